// fp/StreamingConvolver.hpp -- C++ RAII handle over irb_engine (include/irb_b200.h): the batched,
// device-resident replacement of the real-time UPOLA engine the reference inlines in
// IRBaboonAudioProcessor::processBlock (Source/PluginProcessor.cpp:403-562).  The reference has no
// class for this; this is the object a processBlock replacement or a headless harness owns.
#pragma once
#ifdef IRB_USE_REAL_JUCE
#include <JuceHeader.h>
#else
#include "juce_stub/JuceHeader.h"
#endif
#include <vector>

struct irb_engine;
struct irb_group;

namespace fp {
namespace b200 {

class StreamingConvolver {
public:
    // channels = independent stream-channels processed per block; maxPartitions = ceil(longest IR / blockSize)
    StreamingConvolver(int blockSize, int maxPartitions, int channels, int numIRs = 1, int device = 0);
    ~StreamingConvolver();
    StreamingConvolver(const StreamingConvolver&) = delete;
    StreamingConvolver& operator=(const StreamingConvolver&) = delete;

    // mono IR (channel 0) or stereo IR folded to (L+R)/2, as convolvePeriodic does for mono audio
    void setIR(int irId, const AudioBuffer<float>& ir, bool foldStereo = false);
    void bind(int channelBegin, int channelEnd, int irId);
    void reset();
    // one block per call: buffer holds `channels` channels of exactly blockSize samples, processed in place
    void processBlock(AudioBuffer<float>& buffer);
    // dense [nBlocks][channels][blockSize] host arrays (pinned memory from irb_host_alloc is fastest)
    void process(const float* in, float* out, int nBlocks);

    int getBlockSize() const { return blockSize; }
    int getNumChannels() const { return channels; }
    int getLatencySamples() const { return blockSize; }      // PluginProcessor.cpp:59,169-174 with hostBlock <= B
    irb_engine* handle() const { return engine; }

private:
    irb_engine* engine = nullptr;
    int blockSize, channels;
    std::vector<float> stageIn, stageOut;
};

// The same over several GPUs driven from this process (irb_group_*): channels are sharded by contiguous ranges, shared IRs are
// replicated on every device (numIRs > 0) or every channel owns the IR with its own number (numIRs == 0).
class MultiGpuConvolver {
public:
    MultiGpuConvolver(const std::vector<int>& devices, int blockSize, int maxPartitions, int channels, int numIRs = 1);
    ~MultiGpuConvolver();
    MultiGpuConvolver(const MultiGpuConvolver&) = delete;
    MultiGpuConvolver& operator=(const MultiGpuConvolver&) = delete;
    void setIR(int irId, const AudioBuffer<float>& ir, bool foldStereo = false);
    void bind(int channelBegin, int channelEnd, int irId);
    void reset();
    void process(const float* in, float* out, int nBlocks);   // dense [nBlocks][channels][blockSize] host arrays
    irb_group* handle() const { return group; }

private:
    irb_group* group = nullptr;
};

}  // namespace b200
}  // namespace fp
