// fp/convolution.hpp -- drop-in for the reference header of the same name (fp/convolution.hpp:14-48):
// same namespace, function names, argument meaning, default arguments and error behaviour; the work is
// done on a B200 through the C ABI in include/irb_b200.h.  No host compute path exists behind these calls.
#pragma once
#ifdef IRB_USE_REAL_JUCE
#include <JuceHeader.h>
#else
#include "juce_stub/JuceHeader.h"
#endif

namespace fp {

// same enumerators, same order as fp/convolution.hpp:18-24
enum ChannelLayout { unknown, IRMonoAudioMono, IRMonoAudioStereo, IRStereoAudioMono, IRStereoAudioStereo };

namespace convolution {

// Uniformly partitioned FFT convolution of whole buffers; buffer1 = audio, buffer2 = impulse response.
// Returns [audio channels][len1 + len2 - 1]; mono/stereo x mono/stereo only -- anything else prints a debug
// line and returns the cleared buffer, as fp/convolution.cpp:39-42 does.
AudioBuffer<float> convolvePeriodic(AudioBuffer<float>& buffer1, AudioBuffer<float>& buffer2, int processBlockSize = 256);
// One-FFT linear convolution (fp/convolution.cpp:246-347); an unsupported layout returns a cleared copy of buffer1.
AudioBuffer<float> convolveNonPeriodic(AudioBuffer<float>& buffer1, AudioBuffer<float>& buffer2);

// Spectral division FFT(numerator)/FFT(denominator) at N = nextPowerOfTwo(longer length) (circular), optional three
// passes of 1/13-octave log smoothing, inverse FFT, half-swap when the phase is dropped (fp/convolution.cpp:351-403).
// Channel 0 of each buffer is used.
AudioBuffer<float> deconvolve(AudioBuffer<float>* numeratorBuffer, AudioBuffer<float>* denominatorBuffer, double sampleRate, bool smoothing = true,
                              bool includePhase = true, bool includeAmplitude = true);

// Fractional-octave moving average of the bin amplitudes of an interleaved spectrum, in place (fp/convolution.cpp:406-546).
// The reference header names the last two parameters nullifyPhase / nullifyAmplitude but its body treats them as
// INCLUDE flags; the defaults (false, false) are kept, and so is that meaning.
void averagingFilter(AudioBuffer<float>* buffer, double octaveFraction, double sampleRate, bool logAvg, bool nullifyPhase = false, bool nullifyAmplitude = false);

}  // namespace convolution
}  // namespace fp
