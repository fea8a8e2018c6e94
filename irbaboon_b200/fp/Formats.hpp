// fp/Formats.hpp -- the wire / on-disk formats around the convolution path.  Host I/O only.
//
//  * ".sweepandir": what IRBaboonAudioProcessor::saveCustomExt / loadTarget / loadBase exchange
//    (Source/PluginProcessor.cpp:661-695,897-960): a 2-channel 24-bit PCM WAV renamed to the custom extension, channel 0 the
//    recorded sweep, channel 1 the deconvolved IR, totalSweepBreakSamples (65 536) samples, written through
//    ParallelBufferPrinter::printToWav (fp/ParallelBufferPrinter.cpp:205-262: WavAudioFormat, 24 bits).
//  * the TSV spectrum dump of ParallelBufferPrinter::printFreqToCsv (fp/ParallelBufferPrinter.cpp:271-334):
//    "freq  <name>[lin]  <name>[dB]  bin  <name> phase[rad]", one row per bin 0..N/2 of an interleaved spectrum.
//  * the text dump fp::ir::IRtoRealFFTRaw leaves behind (fp/ir.cpp:136-143): the packed partition spectra as
//    comma-separated floats, a line break every 8 values and a blank line every partition.
#pragma once
#ifdef IRB_USE_REAL_JUCE
#include <JuceHeader.h>
#else
#include "juce_stub/JuceHeader.h"
#endif
#include <string>

namespace fp {
namespace b200 {
namespace formats {

// PCM WAV (RIFF/WAVE, format tag 1), bitsPerSample 16 or 24, all channels of the buffer.  Samples are clipped to [-1, 1]
// and quantised as JUCE's integer writers do: round(v * (2^31 - 1)) keeping the top `bits` bits.  false on I/O failure.
bool writeWav(const std::string& path, const AudioBuffer<float>& buffer, int sampleRate, int bitsPerSample = 24);
// reads 16/24/32-bit PCM and 32-bit float WAV; integer samples become value / 2^(bits-1).  Empty buffer on failure.
AudioBuffer<float> readWav(const std::string& path, int* sampleRate = nullptr);

// sweepRecording and ir: channel 0 of each, numSamples samples (zero-padded / truncated), -> <path> as 2-channel 24-bit WAV
bool writeSweepAndIR(const std::string& path, const AudioBuffer<float>& sweepRecording, const AudioBuffer<float>& ir, int sampleRate, int numSamples = 65536);
// -> channel 0 into sweepRecording, channel 1 into ir (each [1][numSamples], cleared first as loadTarget does)
bool readSweepAndIR(const std::string& path, AudioBuffer<float>& sweepRecording, AudioBuffer<float>& ir, int numSamples = 65536);

// interleaved spectrum (channel 0, [2N] floats as tools::fftTransform returns it) -> TSV; non-power-of-two sizes are refused
bool writeSpectrumTsv(const std::string& path, const std::string& name, const AudioBuffer<float>& spectrum, int sampleRate);

// packed partition spectra (the result of ir::IRtoRealFFTRaw) -> the reference's text dump
bool writeRawSpectraText(const std::string& path, const AudioBuffer<float>& packed, int irPartSize);

}  // namespace formats

}  // namespace b200
}  // namespace fp
