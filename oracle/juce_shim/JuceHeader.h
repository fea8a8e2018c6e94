/*
 * oracle/juce_shim/JuceHeader.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Minimal stand-in for the JUCE 6.0.1 surface that IRBaboon's fp/ library touches, so that
 * the reference's own fp/{convolution,tools,ir,CircularBufferArray,ExpSineSweep}.cpp compile
 * UNMODIFIED from /root/reference (see oracle/Makefile).  JUCE itself is not vendored in the
 * reference (IRBaboonCombined.jucer:24-58 points at ~/JUCE/modules) and is not in this image.
 *
 * What is restated here (our own code, written from the documented behaviour of the JUCE API):
 *   - juce::AudioBuffer<T>      planar sample buffer (subset used by fp/)
 *   - juce::dsp::FFT            real-only forward / inverse transform, same conventions as
 *                               juce_dsp/frequency/juce_FFT.cpp's built-in fallback engine:
 *                               decimation-in-time, radix-4 then radix-2 factors, float32
 *                               butterflies, twiddles generated in double, inverse scaled 1/N,
 *                               forward = full complex FFT of {x[i], 0}, inverse rebuilds bins
 *                               N/2+1..N-1 by conjugate symmetry and leaves the N reals in the
 *                               first half of the 2N-float buffer.
 *   - juce::BigInteger, String, File, AudioFormatManager/Reader stubs, DBG
 *
 * g++ notes: the reference spells the type `dsp::FFT::FFT`
 * (convolution.cpp:75-77, tools.cpp:331) which g++ only accepts when `FFT` is a namespace
 * holding a class `FFT`; unqualified abs()/signbit() on floats need the std overloads.
 */
#pragma once

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cerrno>
#include <complex>
#include <memory>
#include <string>
#include <vector>
#include <algorithm>
#include <iomanip>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

using std::abs;
using std::signbit;

#ifndef DBG
#define DBG(x) do { } while (0)
#endif
#define JUCE_BREAK_IN_DEBUGGER do { } while (0)

namespace juce {

template <typename T>
class AudioBuffer {
public:
    AudioBuffer() : nch(0), ns(0) {}
    AudioBuffer(int channels, int samples) : nch(0), ns(0) { allocate(channels, samples); }
    AudioBuffer(const AudioBuffer& o) : nch(o.nch), ns(o.ns), chans(o.chans) {}
    AudioBuffer& operator=(const AudioBuffer& o) { nch = o.nch; ns = o.ns; chans = o.chans; return *this; }

    int getNumChannels() const { return nch; }
    int getNumSamples() const { return ns; }

    const T* getReadPointer(int ch) const { return chans[(size_t) ch].data(); }
    const T* getReadPointer(int ch, int idx) const { return chans[(size_t) ch].data() + idx; }
    T* getWritePointer(int ch) { return chans[(size_t) ch].data(); }
    T* getWritePointer(int ch, int idx) { return chans[(size_t) ch].data() + idx; }

    /* new space is zeroed: the reference relies on that in convolveNonPeriodic (convolution.cpp:291-293) */
    void setSize(int channels, int samples, bool keepExisting = false, bool /*clearExtra*/ = false, bool /*avoidRealloc*/ = false) {
        if (channels == nch && samples == ns) return;
        if (! keepExisting) { allocate(channels, samples); return; }
        chans.resize((size_t) channels);
        for (auto& c : chans) c.resize((size_t) samples, T(0));
        nch = channels; ns = samples;
    }

    void clear() { for (auto& c : chans) std::fill(c.begin(), c.end(), T(0)); }
    void clear(int start, int n) { for (auto& c : chans) std::fill(c.begin() + start, c.begin() + start + n, T(0)); }
    void clear(int ch, int start, int n) { auto& c = chans[(size_t) ch]; std::fill(c.begin() + start, c.begin() + start + n, T(0)); }

    void copyFrom(int destCh, int destStart, const AudioBuffer& src, int srcCh, int srcStart, int n) {
        if (n > 0) memmove(chans[(size_t) destCh].data() + destStart, src.chans[(size_t) srcCh].data() + srcStart, sizeof(T) * (size_t) n);
    }
    void copyFrom(int destCh, int destStart, const T* src, int n) {
        if (n > 0) memmove(chans[(size_t) destCh].data() + destStart, src, sizeof(T) * (size_t) n);
    }

    template <typename Other>
    void makeCopyOf(const AudioBuffer<Other>& o, bool /*avoidRealloc*/ = false) {
        setSize(o.getNumChannels(), o.getNumSamples());
        for (int c = 0; c < nch; ++c) {
            const Other* s = o.getReadPointer(c);
            T* d = getWritePointer(c);
            for (int i = 0; i < ns; ++i) d[i] = static_cast<T>(s[i]);
        }
    }

    T getSample(int ch, int i) const { return chans[(size_t) ch][(size_t) i]; }
    void setSample(int ch, int i, T v) { chans[(size_t) ch][(size_t) i] = v; }

    void applyGain(T g) { for (auto& c : chans) for (auto& v : c) v *= g; }
    void applyGain(int start, int n, T g) { for (auto& c : chans) for (int i = start; i < start + n; ++i) c[(size_t) i] *= g; }

    T getMagnitude(int ch, int start, int n) const {
        T m = 0;
        const T* p = chans[(size_t) ch].data() + start;
        for (int i = 0; i < n; ++i) m = std::max(m, (T) std::abs(p[i]));
        return m;
    }
    T getMagnitude(int start, int n) const {
        T m = 0;
        for (int c = 0; c < nch; ++c) m = std::max(m, getMagnitude(c, start, n));
        return m;
    }

    void reverse(int ch, int start, int n) { auto& c = chans[(size_t) ch]; std::reverse(c.begin() + start, c.begin() + start + n); }
    void reverse(int start, int n) { for (int c = 0; c < nch; ++c) reverse(c, start, n); }

private:
    void allocate(int channels, int samples) {
        nch = channels; ns = samples;
        chans.assign((size_t) channels, std::vector<T>((size_t) samples, T(0)));
    }
    int nch, ns;
    std::vector<std::vector<T>> chans;
};

using AudioSampleBuffer = AudioBuffer<float>;

class BigInteger {
public:
    BigInteger(int v = 0) : value(v) {}
    int getHighestBit() const { int b = -1; unsigned v = (unsigned) value; while (v) { ++b; v >>= 1; } return b; }
private:
    int value;
};

class String {
public:
    String() {}
    String(const char* c) : s(c) {}
    String(const std::string& c) : s(c) {}
    std::string toStdString() const { return s; }
    String operator+(const String& o) const { return String(s + o.s); }
    String operator+(const char* o) const { return String(s + o); }
    String operator+(const std::string& o) const { return String(s + o); }
private:
    std::string s;
};

class File {
public:
    File() {}
    File(const String& p) : path(p) {}
    bool existsAsFile() const { return false; }   /* file I/O is out of scope for the oracle */
    bool exists() const { return false; }
    String getFileName() const { return path; }
private:
    String path;
};

struct AudioFormatReader {
    long long lengthInSamples = 0;
    unsigned int numChannels = 0;
    bool read(AudioBuffer<float>*, int, int, long long, bool, bool) { return false; }
};
struct AudioFormatManager {
    void registerBasicFormats() {}
    AudioFormatReader* createReaderFor(const File&) { return nullptr; }
};

namespace dsp {
namespace FFT {

/* Restatement of the algorithmic contract of JUCE 6's built-in FFT engine (see header comment). */
class FFT {
public:
    explicit FFT(int order) : n(1 << order) {
        build(fwd, false);
        build(inv, true);
    }
    int getSize() const { return n; }

    void performRealOnlyForwardTransform(float* d, bool /*dontCalculateNegativeFrequencies*/ = false) const {
        if (n == 1) return;
        std::vector<std::complex<float>> scratch((size_t) n);
        for (int i = 0; i < n; ++i) scratch[(size_t) i] = std::complex<float>(d[i], 0.0f);
        run(fwd, scratch.data(), reinterpret_cast<std::complex<float>*>(d));
    }

    void performRealOnlyInverseTransform(float* d) const {
        if (n == 1) return;
        auto* in = reinterpret_cast<std::complex<float>*>(d);
        for (int i = n >> 1; i < n; ++i) in[i] = std::conj(in[n - i]);
        std::vector<std::complex<float>> scratch((size_t) n);
        run(inv, in, scratch.data());
        const float scale = 1.0f / (float) n;
        for (int i = 0; i < n; ++i) scratch[(size_t) i] *= scale;
        for (int i = 0; i < n; ++i) { d[i] = scratch[(size_t) i].real(); d[i + n] = scratch[(size_t) i].imag(); }
    }

private:
    struct Stage { int radix, length; };
    struct Plan { bool inverse; std::vector<std::complex<float>> tw; std::vector<Stage> stages; };

    void build(Plan& p, bool inverse) const {
        p.inverse = inverse;
        p.tw.resize((size_t) n);
        const double f = (inverse ? 2.0 : -2.0) * M_PI / (double) n;
        for (int i = 0; i < n; ++i) p.tw[(size_t) i] = std::complex<float>((float) std::cos(i * f), (float) std::sin(i * f));
        int rem = n;
        while (rem > 1) {                       /* radix-4 while possible, then one radix-2 */
            int r = (rem % 4 == 0) ? 4 : 2;
            rem /= r;
            p.stages.push_back({ r, rem });
        }
    }

    void run(const Plan& p, const std::complex<float>* in, std::complex<float>* out) const { rec(p, in, out, 1, 0); }

    /* decimation in time: split into `radix` interleaved sub-sequences, transform each, combine */
    void rec(const Plan& p, const std::complex<float>* in, std::complex<float>* out, int stride, size_t level) const {
        const Stage st = p.stages[level];
        if (st.length == 1) {
            for (int i = 0; i < st.radix; ++i) out[i] = in[(size_t) i * (size_t) stride];
        } else {
            for (int i = 0; i < st.radix; ++i)
                rec(p, in + (size_t) i * (size_t) stride, out + (size_t) i * (size_t) st.length, stride * st.radix, level + 1);
        }
        if (st.radix == 2) bfly2(p, out, stride, st.length); else bfly4(p, out, stride, st.length);
    }

    void bfly2(const Plan& p, std::complex<float>* d, int stride, int len) const {
        const std::complex<float>* tw = p.tw.data();
        for (int i = 0; i < len; ++i) {
            std::complex<float> s = cmul(d[i + len], tw[(size_t) i * (size_t) stride]);
            d[i + len] = d[i] - s;
            d[i] += s;
        }
    }

    void bfly4(const Plan& p, std::complex<float>* d, int stride, int len) const {
        const std::complex<float>* tw = p.tw.data();
        for (int i = 0; i < len; ++i) {
            std::complex<float> s0 = cmul(d[i + len], tw[(size_t) i * (size_t) stride]);
            std::complex<float> s1 = cmul(d[i + 2 * len], tw[(size_t) i * (size_t) stride * 2]);
            std::complex<float> s2 = cmul(d[i + 3 * len], tw[(size_t) i * (size_t) stride * 3]);
            std::complex<float> s3 = s0 + s2;
            std::complex<float> s4 = s0 - s2;
            std::complex<float> s5 = d[i] - s1;
            d[i] += s1;
            d[i + 2 * len] = d[i] - s3;
            d[i] += s3;
            if (p.inverse) {
                d[i + len]     = std::complex<float>(s5.real() - s4.imag(), s5.imag() + s4.real());
                d[i + 3 * len] = std::complex<float>(s5.real() + s4.imag(), s5.imag() - s4.real());
            } else {
                d[i + len]     = std::complex<float>(s5.real() + s4.imag(), s5.imag() - s4.real());
                d[i + 3 * len] = std::complex<float>(s5.real() - s4.imag(), s5.imag() + s4.real());
            }
        }
    }

    static std::complex<float> cmul(std::complex<float> a, std::complex<float> b) {
        return std::complex<float>(a.real() * b.real() - a.imag() * b.imag(), a.real() * b.imag() + a.imag() * b.real());
    }

    int n;
    Plan fwd, inv;
};

} // namespace FFT
} // namespace dsp

} // namespace juce

using namespace juce;
