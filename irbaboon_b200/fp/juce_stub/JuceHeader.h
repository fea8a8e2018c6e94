// JuceHeader.h -- minimal stand-in for the Projucer-generated JuceHeader.h, used ONLY when real JUCE is
// not on the include path (the reference does not vendor JUCE; IRBaboonCombined.jucer:24-58 points at the
// author's ~/JUCE/modules).  It provides the subset of juce::AudioBuffer<T> the fp:: facade exchanges at
// the boundary (planar, per-channel contiguous storage).  It deliberately has NO dsp::FFT: the product
// path never transforms on the host.  With real JUCE available, build with -DIRB_USE_REAL_JUCE and put the
// project's JuceLibraryCode on the include path instead of this directory.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace juce {

template <typename T>
class AudioBuffer {
public:
    AudioBuffer() = default;
    AudioBuffer(int channels, int samples) { allocate(channels, samples, false); }   // contents unspecified, as in JUCE
    int getNumChannels() const noexcept { return nch; }
    int getNumSamples() const noexcept { return nsm; }
    const T* getReadPointer(int ch) const noexcept { return data.data() + (size_t) ch * stride; }
    const T* getReadPointer(int ch, int i) const noexcept { return getReadPointer(ch) + i; }
    T* getWritePointer(int ch) noexcept { return data.data() + (size_t) ch * stride; }
    T* getWritePointer(int ch, int i) noexcept { return getWritePointer(ch) + i; }
    T getSample(int ch, int i) const noexcept { return getReadPointer(ch)[i]; }
    void setSample(int ch, int i, T v) noexcept { getWritePointer(ch)[i] = v; }
    void clear() noexcept { std::fill(data.begin(), data.end(), T(0)); }
    void clear(int start, int n) noexcept { for (int c = 0; c < nch; ++c) clear(c, start, n); }
    void clear(int ch, int start, int n) noexcept { std::fill_n(getWritePointer(ch, start), n, T(0)); }
    // keepExisting keeps the overlapping region; new space is zeroed (the reference relies on that)
    void setSize(int channels, int samples, bool keepExisting = false, bool clearExtra = false, bool avoidRealloc = false) {
        (void) clearExtra; (void) avoidRealloc;
        if (channels == nch && samples == nsm) return;
        AudioBuffer old;
        if (keepExisting) old = *this;
        allocate(channels, samples, true);
        if (keepExisting)
            for (int c = 0; c < std::min(channels, old.nch); ++c)
                std::copy_n(old.getReadPointer(c), std::min(samples, old.nsm), getWritePointer(c));
    }
    void copyFrom(int dstCh, int dstStart, const AudioBuffer& src, int srcCh, int srcStart, int n) noexcept {
        if (n > 0) std::memmove(getWritePointer(dstCh, dstStart), src.getReadPointer(srcCh, srcStart), sizeof(T) * (size_t) n);
    }
    void copyFrom(int dstCh, int dstStart, const T* src, int n) noexcept {
        if (n > 0) std::memmove(getWritePointer(dstCh, dstStart), src, sizeof(T) * (size_t) n);
    }
    template <typename U> void makeCopyOf(const AudioBuffer<U>& o) {
        allocate(o.getNumChannels(), o.getNumSamples(), false);
        for (int c = 0; c < nch; ++c) { const U* s = o.getReadPointer(c); T* d = getWritePointer(c); for (int i = 0; i < nsm; ++i) d[i] = (T) s[i]; }
    }
    void applyGain(T g) noexcept { for (auto& v : data) v *= g; }
    void applyGain(int ch, int start, int n, T g) noexcept { T* p = getWritePointer(ch, start); for (int i = 0; i < n; ++i) p[i] *= g; }
    T getMagnitude(int ch, int start, int n) const noexcept { T m = 0; const T* p = getReadPointer(ch, start); for (int i = 0; i < n; ++i) m = std::max(m, (T) std::fabs(p[i])); return m; }
    T getMagnitude(int start, int n) const noexcept { T m = 0; for (int c = 0; c < nch; ++c) m = std::max(m, getMagnitude(c, start, n)); return m; }
    void reverse(int ch, int start, int n) noexcept { std::reverse(getWritePointer(ch, start), getWritePointer(ch, start) + n); }
    void reverse(int start, int n) noexcept { for (int c = 0; c < nch; ++c) reverse(c, start, n); }

private:
    void allocate(int channels, int samples, bool zero) {
        nch = std::max(0, channels); nsm = std::max(0, samples); stride = (size_t) nsm;
        if (zero) data.assign((size_t) nch * stride, T(0)); else data.resize((size_t) nch * stride);
    }
    int nch = 0, nsm = 0;
    size_t stride = 0;
    std::vector<T> data;
};

using AudioSampleBuffer = AudioBuffer<float>;
using String = std::string;

}  // namespace juce

#ifndef DBG
#define DBG(msg) do { std::fprintf(stderr, "%s", std::string(msg).c_str()); } while (0)
#endif
using namespace juce;      // JuceHeader.h:42 of the reference does the same
