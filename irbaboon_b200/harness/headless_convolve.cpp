// headless_convolve.cpp -- headless C++ harness: BASELINE config 1 (mono 48 kHz, block 512, 1 s IR, 10 s white
// noise) through the fp:: facade exactly as code written against the reference's fp/convolution.hpp would call
// it, then a streaming run of the same signal through fp::b200::StreamingConvolver.  Prints timings and
// checksums; tests/test_facade.py compares the checksums with the oracle's.
//   usage: headless_convolve [seconds=10] [ir_taps=48000] [block=512]
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../fp/PluginConvolver.hpp"
#include "../fp/StreamingConvolver.hpp"
#include "../fp/convolution.hpp"
#include "../fp/tools.hpp"

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

        // the plug-in's own situation: stereo, processBlockSize 256, a host that delivers 480-sample buffers, the default IR
        // generatePulse(2048, 100) -> the output is the input delayed by 100 samples plus the re-blocking latency, times -30 dB
        const int hostBlock = 480, nhost = 200;
        AudioBuffer<float> pulse = fp::tools::generatePulse(2048, 100);
        fp::b200::PluginConvolver plug(256, 2);
        plug.prepareToPlay(48000.0, hostBlock, pulse);
        std::vector<float> in2((size_t) 2 * hostBlock * nhost), out2(in2.size());
        for (int c = 0; c < 2; ++c)
            for (int i = 0; i < hostBlock * nhost; ++i) in2[(size_t) c * hostBlock * nhost + i] = noise(1002, (uint64_t) c, (uint64_t) i);
        AudioBuffer<float> hb(2, hostBlock);
        t0 = std::chrono::steady_clock::now();
        for (int k = 0; k < nhost; ++k) {
            for (int c = 0; c < 2; ++c) hb.copyFrom(c, 0, &in2[(size_t) c * hostBlock * nhost + (size_t) k * hostBlock], hostBlock);
            plug.processBlock(hb);
            for (int c = 0; c < 2; ++c) std::copy_n(hb.getReadPointer(c), hostBlock, &out2[(size_t) c * hostBlock * nhost + (size_t) k * hostBlock]);
        }
        const double dtp = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        // find the delay from the data (the reference's output ring decides it), then measure the deviation from gain * delayed input
        const float g30 = fp::tools::dBToLin(-30.0f);
        int delay = -1;
        for (int d = 0; d < 4 * hostBlock && delay < 0; ++d) {
            double err = 0.0;
            for (int i = 2000; i < 4000; ++i) err = std::max(err, (double) std::fabs(out2[(size_t) i + d] - g30 * in2[(size_t) i]));
            if (err < 1e-6) delay = d;
        }
        double dev = 0.0;
        if (delay >= 0)
            for (int c = 0; c < 2; ++c)
                for (int i = 0; i + delay < hostBlock * nhost; ++i)
                    dev = std::max(dev, (double) std::fabs(out2[(size_t) c * hostBlock * nhost + i + delay] - g30 * in2[(size_t) c * hostBlock * nhost + i]));
        printf("plugin hostBlock=%d callbacks=%d us_per_callback=%.1f reported_latency=%d measured_delay=%d max_dev_from_delayed_input=%.3e\n", hostBlock, nhost,
               1e6 * dtp / nhost, plug.getLatencySamples(), delay, dev);
    } catch (const std::exception& ex) {
        fprintf(stderr, "headless_convolve: %s\n", ex.what());
        return 1;
    }
    return 0;
}
