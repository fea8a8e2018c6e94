// fp/tools.hpp -- drop-in for the hot-path part of the reference's fp/tools.hpp:18-92.  The per-bin scalar
// primitives (complexMul, complexDivCartesian, binAmpl, ...) are kept as host inline functions for source
// compatibility; the loops that called them per bin in the reference (fp/convolution.cpp:184-189,373-381,
// 429-543) run as CUDA kernels behind convolution.hpp.  fftTransform / fftInvTransform run on the GPU.
// Not provided: fileToBuffer / DescribeIosFailure (file I/O, outside the convolution path -- DESIGN.md section 8).
#pragma once
#ifdef IRB_USE_REAL_JUCE
#include <JuceHeader.h>
#else
#include "juce_stub/JuceHeader.h"
#endif
#include <cmath>

namespace fp {
namespace tools {

void sumToMono(AudioBuffer<float>* buffer);                    // in place: ch0 = (ch0 + ch1) / 2, ch1 = 0
void makeStereo(AudioBuffer<float>* buffer);                   // mono -> two identical channels

inline void complexMul(float* a, float* b, float c, float d) { // (a,b) := (a,b) * (c,d)
    const float re = (*a) * c - (*b) * d, im = (*b) * c + (*a) * d;
    *a = re; *b = im;
}
inline void complexDivCartesian(float* a, float* b, float c, float d) {   // (a,b) := (a,b) / (c,d); 0/0 denominators leave it untouched
    if (c == 0.0 && d == 0.0) return;
    const float re = ((*a) * c + (*b) * d) / (c * c + d * d), im = ((*b) * c - (*a) * d) / (c * c + d * d);
    *a = re; *b = im;
}
void complexDivPolar(float* a, float* b, float c, float d);

void normalize(AudioBuffer<float>* buffer, float dBGoalLevel = 0.0f, bool printGain = false);
float dBToLin(float dB);
double dBToLin(double dB);
float linTodB(float lin);
double linTodB(double lin);

bool isPowerOfTwo(int x);
int nextPowerOfTwo(int x, int result = 1);
void roundToZero(float* x, float threshold);
void roundTo1TenQuadrillionth(float* x);
float binAmpl(float* binPtr);
float binPhase(float* binPtr);

AudioBuffer<float> generatePulse(int numSamples, int pulseOffset = 0);
void linearFade(AudioBuffer<float>* buffer, bool fadeIn, int startSample, int numSamples);
void sineFill(AudioBuffer<float>* buffer, float freq, float sampleRate, float ampl = 1.0f);

// real FFT of channel 0 at N = nextPowerOfTwo(length) -> [channels][2N] interleaved {re, im}; inverse -> [channels][N]
AudioBuffer<float> fftTransform(AudioBuffer<float>& buffer, bool formatAmplPhase = false);
AudioBuffer<float> fftInvTransform(AudioBuffer<float>& buffer);

}  // namespace tools
}  // namespace fp
