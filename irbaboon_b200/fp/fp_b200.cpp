// fp_b200.cpp -- implementation of the fp:: facade over the C ABI (include/irb_b200.h).
// Host code only marshals buffers; every transform / multiply-accumulate runs in libirb_b200.so on the GPU.
// A failing CUDA call throws std::runtime_error: there is no host fallback to hide behind.
#include <cmath>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "../../include/irb_b200.h"
#include "CircularBufferArray.hpp"
#include "StreamingConvolver.hpp"
#include "convolution.hpp"

namespace {

[[noreturn]] void raise(const char* what) {
    throw std::runtime_error(std::string(what) + ": " + irb_last_error());
}

// planar AudioBuffer -> one dense [channels][samples] array (JUCE channels need not be adjacent in memory)
std::vector<float> pack(const AudioBuffer<float>& b) {
    const int ch = b.getNumChannels(), n = b.getNumSamples();
    std::vector<float> v((size_t) ch * (size_t) n);
    for (int c = 0; c < ch; ++c) std::copy_n(b.getReadPointer(c), n, v.begin() + (size_t) c * n);
    return v;
}

}  // namespace

namespace fp {

// ---- convolution ---------------------------------------------------------------------------------
AudioBuffer<float> convolution::convolvePeriodic(AudioBuffer<float>& buffer1, AudioBuffer<float>& buffer2, int processBlockSize) {
    const int chx = buffer1.getNumChannels(), lx = buffer1.getNumSamples();
    const int chh = buffer2.getNumChannels(), lh = buffer2.getNumSamples();
    AudioBuffer<float> out(chx, lx + lh - 1);
    out.clear();
    if (out.getNumSamples() <= 0 || chx <= 0) return out;
    std::vector<float> x = pack(buffer1), h = pack(buffer2), y((size_t) chx * out.getNumSamples());
    const int rc = irb_convolve_periodic(x.data(), chx, lx, h.data(), chh, lh, processBlockSize, y.data());
    if (rc == IRB_ERR_LAYOUT) {
        DBG("Either buffer1 or buffer2 is not mono nor stereo. Abort abort \n");
        return out;
    }
    if (rc != IRB_OK) raise("fp::convolution::convolvePeriodic");
    for (int c = 0; c < chx; ++c) out.copyFrom(c, 0, y.data() + (size_t) c * out.getNumSamples(), out.getNumSamples());
    return out;
}

// ---- CircularBufferArray (fp/CircularBufferArray.cpp:12-205) ---------------------------------------
CircularBufferArray::CircularBufferArray() {}
CircularBufferArray::CircularBufferArray(int amountOfBuffers, int bufferChannelSize, int bufferSampleSize) {
    initBuffers(amountOfBuffers, bufferChannelSize, bufferSampleSize);
}
CircularBufferArray::~CircularBufferArray() {}

void CircularBufferArray::initBuffers(int amountOfBuffers, int bufferChannelSize, int bufferSampleSize) {
    channelsPerBuffer = bufferChannelSize;
    samplesPerBuffer = bufferSampleSize;
    arraySize = amountOfBuffers;
    readIndex = writeIndex = 0;
    bufferArray.assign((size_t) std::max(0, amountOfBuffers), AudioBuffer<float>());
    for (auto& b : bufferArray) { b.setSize(bufferChannelSize, bufferSampleSize); b.clear(); }
}
void CircularBufferArray::clearAndResize(int amountOfBuffers, int bufferChannelSize, int bufferSampleSize) {
    initBuffers(amountOfBuffers, bufferChannelSize, bufferSampleSize);
}
void CircularBufferArray::changeArraySize(int amountOfBuffers) {
    if (amountOfBuffers == arraySize) return;
    if (amountOfBuffers == 0) {
        bufferArray.clear();
        readIndex = writeIndex = arraySize = 0;
        return;
    }
    if (amountOfBuffers > arraySize) {
        const int extra = amountOfBuffers - arraySize;
        for (int i = 0; i < extra; ++i) {
            AudioBuffer<float> b(channelsPerBuffer, samplesPerBuffer);
            b.clear();
            bufferArray.push_back(b);
            // the reference clears slot i here -- an OLD front slot, not the new one (CircularBufferArray.cpp:49-52);
            // observable, so kept
            if (i < (int) bufferArray.size()) bufferArray[(size_t) i].clear();
        }
        arraySize = (int) bufferArray.size();
        return;
    }
    // shrinking
    if (lastWrittenIndex == -1) {          // never written: plain truncation
        bufferArray.resize((size_t) amountOfBuffers);
        return;                            // (arraySize is left untouched by the reference on this path, :84-86)
    }
    const int readBack = lastWrittenIndex - readIndex, writeBack = lastWrittenIndex - writeIndex;
    std::vector<AudioBuffer<float>> kept((size_t) amountOfBuffers);
    int src = lastWrittenIndex;
    for (int i = amountOfBuffers - 1; i >= 0; --i) {      // newest ends up last
        kept[(size_t) i] = bufferArray[(size_t) src];
        src = src == 0 ? arraySize - 1 : src - 1;
    }
    bufferArray.swap(kept);
    arraySize = amountOfBuffers;
    const int last = arraySize - 1;
    auto remap = [&](int back) { return back > last ? 0 : (back < 0 ? last - (arraySize + back) : last - back); };
    readIndex = remap(readBack);
    writeIndex = remap(writeBack);
}
AudioBuffer<float>* CircularBufferArray::getReadBufferPtr() { return &bufferArray[(size_t) readIndex]; }
AudioBuffer<float>* CircularBufferArray::getWriteBufferPtr() { lastWrittenIndex = writeIndex; return &bufferArray[(size_t) writeIndex]; }
AudioBuffer<float>* CircularBufferArray::getBufferPtrAtIndex(int index) { return &bufferArray[(size_t) index]; }
void CircularBufferArray::incrReadIndex() { if (++readIndex >= arraySize) readIndex = 0; }
void CircularBufferArray::decrReadIndex() { if (--readIndex < 0) readIndex = arraySize - 1; }
void CircularBufferArray::incrWriteIndex() { if (++writeIndex >= arraySize) writeIndex = 0; }
AudioBuffer<float> CircularBufferArray::consolidate(int bufOffset) {
    AudioBuffer<float> all(channelsPerBuffer, samplesPerBuffer * arraySize);
    int start = bufOffset < 0 ? arraySize - 1 - std::abs(bufOffset) : bufOffset;
    for (int k = 0; k < arraySize; ++k) {
        const int slot = (start + k) % std::max(1, arraySize);
        for (int c = 0; c < channelsPerBuffer; ++c)
            all.copyFrom(c, k * samplesPerBuffer, bufferArray[(size_t) slot], c, 0, samplesPerBuffer);
    }
    return all;
}
int CircularBufferArray::getReadIndex() { return readIndex; }
void CircularBufferArray::setReadIndex(int index) { readIndex = index; }
int CircularBufferArray::getWriteIndex() { return writeIndex; }
void CircularBufferArray::setWriteIndex(int index) { writeIndex = index; }
int CircularBufferArray::getArraySize() { return arraySize; }
int CircularBufferArray::getChannelsPerBuffer() { return channelsPerBuffer; }
int CircularBufferArray::getSamplesPerBuffer() { return samplesPerBuffer; }

// ---- StreamingConvolver ----------------------------------------------------------------------------
namespace b200 {

StreamingConvolver::StreamingConvolver(int blockSize_, int maxPartitions, int channels_, int numIRs, int device)
    : blockSize(blockSize_), channels(channels_) {
    if (irb_engine_create(&engine, device, blockSize, maxPartitions, channels, numIRs) != IRB_OK) raise("irb_engine_create");
    stageIn.resize((size_t) blockSize * channels);
    stageOut.resize((size_t) blockSize * channels);
}
StreamingConvolver::~StreamingConvolver() { irb_engine_destroy(engine); }
void StreamingConvolver::setIR(int irId, const AudioBuffer<float>& ir, bool foldStereo) {
    const float* r = (foldStereo && ir.getNumChannels() == 2) ? ir.getReadPointer(1) : nullptr;
    if (irb_engine_set_ir(engine, irId, ir.getReadPointer(0), r, ir.getNumSamples()) != IRB_OK) raise("irb_engine_set_ir");
}
void StreamingConvolver::bind(int channelBegin, int channelEnd, int irId) {
    if (irb_engine_bind(engine, channelBegin, channelEnd, irId) != IRB_OK) raise("irb_engine_bind");
}
void StreamingConvolver::reset() {
    if (irb_engine_reset(engine) != IRB_OK) raise("irb_engine_reset");
}
void StreamingConvolver::processBlock(AudioBuffer<float>& buffer) {
    if (buffer.getNumChannels() != channels || buffer.getNumSamples() != blockSize)
        throw std::invalid_argument("StreamingConvolver::processBlock: buffer must be channels x blockSize");
    for (int c = 0; c < channels; ++c) std::copy_n(buffer.getReadPointer(c), blockSize, stageIn.begin() + (size_t) c * blockSize);
    if (irb_engine_process(engine, stageIn.data(), stageOut.data(), 1) != IRB_OK) raise("irb_engine_process");
    for (int c = 0; c < channels; ++c) buffer.copyFrom(c, 0, stageOut.data() + (size_t) c * blockSize, blockSize);
}
void StreamingConvolver::process(const float* in, float* out, int nBlocks) {
    if (irb_engine_process(engine, in, out, nBlocks) != IRB_OK) raise("irb_engine_process");
}

}  // namespace b200
}  // namespace fp
