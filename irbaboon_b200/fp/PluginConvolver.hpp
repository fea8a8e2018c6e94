// fp/PluginConvolver.hpp -- the convolution half of IRBaboonAudioProcessor with processBlock semantics
// (Source/PluginProcessor.cpp:161-234 prepareToPlay, :403-574 processBlock in IRCAP_IDLE, :584-590 bypass), over the
// CUDA engine.  What the reference does with six CircularBufferArrays and three host FFTs becomes: two host rings that
// re-block between the host's buffer size and processBlockSize (sample shuffling only -- no arithmetic), and one
// irb_engine_process_callback() per host callback for everything else (forward FFTs, FDL, round-robin IR refresh, MAC,
// inverse FFT, overlap-add -- all on the GPU).
//
// Kept from the reference: latency max(hostBlock, processBlockSize) reported, output ring one buffer ahead (:218), the
// IR switched partition by partition (one per processed block, :455-461), IR channel 0 for every audio channel (:485),
// output gain then the "makeshift limiter" (:567-574), the bypass delay line (:584-590), and -- with
// exactReferenceOrder (default) -- the order "transform all blocks of the callback, then convolve them" together with
// the reference's FDL ring of max(partitions, blocksPerCallback) slots, which makes the oldest partitions read the
// callback's later blocks when hostBlock > processBlockSize.  exactReferenceOrder = false sizes the ring for
// partitions + blocksPerCallback - 1 spectra instead and gives the causal convolution.
// Differences: processBlockSize and the channel count are constructor parameters (the reference pins 256 and 2), and the
// partition count follows the IR given to prepareToPlay (the reference pins it to IRpulse = 2048 taps).
#pragma once
#ifdef IRB_USE_REAL_JUCE
#include <JuceHeader.h>
#else
#include "juce_stub/JuceHeader.h"
#endif
#include <vector>

struct irb_engine;

namespace fp {
namespace b200 {

class PluginConvolver {
public:
    explicit PluginConvolver(int processBlockSize = 256, int channels = 2, int device = 0);
    ~PluginConvolver();
    PluginConvolver(const PluginConvolver&) = delete;
    PluginConvolver& operator=(const PluginConvolver&) = delete;

    void prepareToPlay(double sampleRate, int samplesPerBlock, const AudioBuffer<float>& ir);
    void setIR(const AudioBuffer<float>& ir);                 // IRtoConvolve = &ir: picked up one partition per processed block
    void processBlock(AudioBuffer<float>& buffer);            // buffer: >= channels channels, <= samplesPerBlock samples, in place
    void processBlockBypassed(AudioBuffer<float>& buffer);    // keeps the latency while bypassed
    int getLatencySamples() const { return latency; }
    void setOutputVolumedB(float dB) { outputVolumedB = dB; } // PluginProcessor.h:168 default -30 dB
    void setExactReferenceOrder(bool exact) { exactOrder = exact; }   // takes effect at the next prepareToPlay
    int getNumPartitions() const { return partitions; }
    irb_engine* handle() const { return engine; }

private:
    void destroyEngine();
    irb_engine* engine = nullptr;
    int B, channels, device;
    int hostBlock = 0, latency = 0, partitions = 0;
    float outputVolumedB = -30.0f;
    bool exactOrder = true;
    // input side: samples collected into blocks, laid out [block][channel][B] for the engine
    std::vector<float> inBlocks, outBlocks;
    int inArray = 0, inSample = 0, blocksToProcess = 0;
    // output side: outArray host-sized buffers, written block by block, read one buffer behind
    std::vector<float> outRing, bypassRing;
    int outArray = 0, outWrite = 0, outRead = 1, outWriteSample = 0, outReadSample = 0;
    int bypassWrite = 0, bypassRead = 0;
};

}  // namespace b200
}  // namespace fp
