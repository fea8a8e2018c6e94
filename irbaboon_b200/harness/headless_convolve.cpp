// headless_convolve.cpp -- headless C++ harness: BASELINE config 1 (mono 48 kHz, block 512, 1 s IR, 10 s white
// noise) through the fp:: facade exactly as code written against the reference's fp/convolution.hpp would call
// it, then a streaming run of the same signal through fp::b200::StreamingConvolver.  Prints timings and
// checksums; tests/test_facade.py compares the checksums with the oracle's.
//   usage: headless_convolve [seconds=10] [ir_taps=48000] [block=512]
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../fp/StreamingConvolver.hpp"
#include "../fp/convolution.hpp"

static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static float noise(uint64_t seed, uint64_t stream, uint64_t i) {           // SURVEY 8(d) generator
    const uint64_t u = splitmix64(seed + (stream << 40) + i);
    return (float) ((double) (u >> 40) * (1.0 / 16777216.0) * 2.0 - 1.0);
}

int main(int argc, char** argv) {
    const double seconds = argc > 1 ? atof(argv[1]) : 10.0;
    const int taps = argc > 2 ? atoi(argv[2]) : 48000;
    const int B = argc > 3 ? atoi(argv[3]) : 512;
    const int Lx = (int) (seconds * 48000.0);
    AudioBuffer<float> x(1, Lx), h(1, taps);
    for (int i = 0; i < Lx; ++i) x.setSample(0, i, noise(1001, 0, (uint64_t) i));
    double e = 0.0;
    for (int i = 0; i < taps; ++i) { float v = noise(2000, 0, (uint64_t) i) * (float) std::exp(-6.9078 * i / taps); h.setSample(0, i, v); e += (double) v * v; }
    h.applyGain((float) (1.0 / std::sqrt(e)));

    try {
        fp::convolution::convolvePeriodic(x, h, B);                        // first call pays CUDA context creation
        auto t0 = std::chrono::steady_clock::now();
        AudioBuffer<float> y = fp::convolution::convolvePeriodic(x, h, B);
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        double sum = 0.0, sq = 0.0;
        for (int i = 0; i < y.getNumSamples(); ++i) { sum += y.getSample(0, i); sq += (double) y.getSample(0, i) * y.getSample(0, i); }
        printf("offline  samples=%d seconds=%.6f realtime_factor=%.1f sum=%.9e sumsq=%.9e\n", y.getNumSamples(), dt, seconds / dt, sum, sq);

        const int P = (taps + B - 1) / B, nb = Lx / B;
        fp::b200::StreamingConvolver conv(B, P, 1);
        conv.setIR(0, h);
        AudioBuffer<float> blk(1, B);
        double ssum = 0.0, maxdiff = 0.0;
        t0 = std::chrono::steady_clock::now();
        for (int b = 0; b < nb; ++b) {
            blk.copyFrom(0, 0, x, 0, b * B, B);
            conv.processBlock(blk);
            for (int i = 0; i < B; ++i) { ssum += blk.getSample(0, i); maxdiff = std::max(maxdiff, (double) std::fabs(blk.getSample(0, i) - y.getSample(0, b * B + i))); }
        }
        const double dts = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("streaming blocks=%d seconds=%.6f us_per_block=%.1f sum=%.9e max_abs_vs_offline=%.3e\n", nb, dts, 1e6 * dts / nb, ssum, maxdiff);
    } catch (const std::exception& ex) {
        fprintf(stderr, "headless_convolve: %s\n", ex.what());
        return 1;
    }
    return 0;
}
