// fp/ExpSineSweep.hpp -- drop-in for the reference's fp/ExpSineSweep.hpp:20-64 (Farina exponential sine sweep).
// generate / generateInv evaluate the FP64 sweep on the GPU (irb_ess_generate); the fade-outs and index helpers
// are a handful of host operations.
#pragma once
#include "tools.hpp"

namespace fp {

class ExpSineSweep {
public:
    ExpSineSweep();
    ~ExpSineSweep();
    void generate(double durationSecs, double sampleRate, double lowFreq, double highFreq, double dBGain);
    AudioBuffer<double> getSweep();
    AudioBuffer<float> getSweepFloat();
    void generateInv();
    void generateInv(double durationSecs, double sampleRate, double lowFreq, double highFreq, double dBGain);
    AudioBuffer<double> getSweepInv();
    AudioBuffer<float> getSweepInvFloat();
    int getSampleIndexAtFreq(double freq);
    int getSampleIndexAtFreq(double freq, double durationSecs, double sampleRate, double lowFreq, double highFreq);
    double getFreqAtSampleIndex(int index);
    double getFreqAtSampleIndex(int index, double durationSecs, double sampleRate, double lowFreq, double highFreq);
    void linFadeout(double freq);
    void dBFadeout(double freq);
    void brickwallFadeout(double freq);

private:
    void assignParameters(double durationSecs, double sampleRate, double lowFreq, double highFreq);
    double getFreqAtSampleIndexHelper(int index);
    int getSampleHelper(double freq);
    double SR = 0, w1 = 0, w2 = 0, T = 0, K = 0, L = 0, k = 0, kend = 0;
    double genDuration = 0, genLow = 0, genHigh = 0, gendB = 0;
    AudioBuffer<double> sweep;
    AudioBuffer<double> sweepInv;
};

}  // namespace fp
