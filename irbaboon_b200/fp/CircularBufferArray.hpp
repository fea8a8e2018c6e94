// fp/CircularBufferArray.hpp -- drop-in for fp/CircularBufferArray.hpp:18-64: a ring of AudioBuffers with
// wrapping read/write cursors.  Host-side container only (the plugin uses it for re-blocking and capture);
// the device-resident analogue -- the FDL ring of spectra -- lives inside irb_engine.
#pragma once
#ifdef IRB_USE_REAL_JUCE
#include <JuceHeader.h>
#else
#include "juce_stub/JuceHeader.h"
#endif
#include <vector>

namespace fp {

class CircularBufferArray {
public:
    CircularBufferArray();
    CircularBufferArray(int amountOfBuffers, int bufferChannelSize, int bufferSampleSize);
    ~CircularBufferArray();

    void clearAndResize(int amountOfBuffers, int bufferChannelSize, int bufferSampleSize);
    // growing keeps all data; shrinking keeps the most recently written buffers in ascending order
    void changeArraySize(int amountOfBuffers);

    AudioBuffer<float>* getReadBufferPtr();
    AudioBuffer<float>* getWriteBufferPtr();              // also remembers this slot as "last written"
    AudioBuffer<float>* getBufferPtrAtIndex(int index);

    void incrReadIndex();
    void decrReadIndex();
    void incrWriteIndex();

    AudioBuffer<float> consolidate(int bufOffset = 0);   // all slots appended, starting at slot bufOffset

    int getReadIndex();
    void setReadIndex(int index);
    int getWriteIndex();
    void setWriteIndex(int index);
    int getArraySize();
    int getChannelsPerBuffer();
    int getSamplesPerBuffer();

private:
    void initBuffers(int amountOfBuffers, int bufferChannelSize, int bufferSampleSize);
    std::vector<AudioBuffer<float>> bufferArray;
    int channelsPerBuffer = 0, samplesPerBuffer = 0, readIndex = 0, writeIndex = 0, arraySize = 0;
    int lastWrittenIndex = -1;
};

}  // namespace fp
