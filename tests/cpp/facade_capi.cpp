// tests/cpp/facade_capi.cpp -- TEST INFRASTRUCTURE: a flat C view of the C++ facade (irbaboon_b200/fp) so that
// python tests can drive it next to the reference's own classes (oracle/_ref exposes the same view as ref_cba_*).
#include "../../irbaboon_b200/fp/CircularBufferArray.hpp"
#include "../../irbaboon_b200/fp/convolution.hpp"

static void fill(AudioBuffer<float>& b, const float* data) {
    for (int c = 0; c < b.getNumChannels(); ++c) b.copyFrom(c, 0, data + (size_t) c * b.getNumSamples(), b.getNumSamples());
}
static void dump(const AudioBuffer<float>& b, float* out) {
    for (int c = 0; c < b.getNumChannels(); ++c) std::copy_n(b.getReadPointer(c), b.getNumSamples(), out + (size_t) c * b.getNumSamples());
}
extern "C" {
void* fac_cba_create(int buffers, int ch, int n) { return new fp::CircularBufferArray(buffers, ch, n); }
void fac_cba_destroy(void* h) { delete (fp::CircularBufferArray*) h; }
void fac_cba_clear_and_resize(void* h, int buffers, int ch, int n) { ((fp::CircularBufferArray*) h)->clearAndResize(buffers, ch, n); }
void fac_cba_change_array_size(void* h, int buffers) { ((fp::CircularBufferArray*) h)->changeArraySize(buffers); }
void fac_cba_write(void* h, const float* data) { fill(*((fp::CircularBufferArray*) h)->getWriteBufferPtr(), data); }
void fac_cba_read(void* h, float* out) { dump(*((fp::CircularBufferArray*) h)->getReadBufferPtr(), out); }
void fac_cba_read_at(void* h, int idx, float* out) { dump(*((fp::CircularBufferArray*) h)->getBufferPtrAtIndex(idx), out); }
void fac_cba_incr_read(void* h) { ((fp::CircularBufferArray*) h)->incrReadIndex(); }
void fac_cba_decr_read(void* h) { ((fp::CircularBufferArray*) h)->decrReadIndex(); }
void fac_cba_incr_write(void* h) { ((fp::CircularBufferArray*) h)->incrWriteIndex(); }
int fac_cba_get_read_index(void* h) { return ((fp::CircularBufferArray*) h)->getReadIndex(); }
int fac_cba_get_write_index(void* h) { return ((fp::CircularBufferArray*) h)->getWriteIndex(); }
void fac_cba_set_read_index(void* h, int i) { ((fp::CircularBufferArray*) h)->setReadIndex(i); }
void fac_cba_set_write_index(void* h, int i) { ((fp::CircularBufferArray*) h)->setWriteIndex(i); }
int fac_cba_get_array_size(void* h) { return ((fp::CircularBufferArray*) h)->getArraySize(); }
int fac_cba_consolidate(void* h, int offset, float* out) {
    AudioBuffer<float> r = ((fp::CircularBufferArray*) h)->consolidate(offset);
    dump(r, out);
    return r.getNumSamples();
}
// fp::convolution::convolvePeriodic through the facade; returns the output length, -1 if the facade threw
int fac_convolve_periodic(const float* x, int chx, int Lx, const float* h, int chh, int Lh, int B, float* out) {
    AudioBuffer<float> bx(chx, Lx), bh(chh, Lh);
    fill(bx, x);
    fill(bh, h);
    try {
        AudioBuffer<float> r = fp::convolution::convolvePeriodic(bx, bh, B);
        dump(r, out);
        return r.getNumSamples();
    } catch (const std::exception&) {
        return -1;
    }
}
}

// ---- the rest of the fp:: surface ---------------------------------------------------------------------------
#include "../../irbaboon_b200/fp/ExpSineSweep.hpp"
#include "../../irbaboon_b200/fp/ir.hpp"
#include "../../irbaboon_b200/fp/tools.hpp"
extern "C" {
int fac_convolve_nonperiodic(const float* x, int chx, int Lx, const float* h, int chh, int Lh, float* out) {
    AudioBuffer<float> bx(chx, Lx), bh(chh, Lh);
    fill(bx, x); fill(bh, h);
    try { AudioBuffer<float> r = fp::convolution::convolveNonPeriodic(bx, bh); dump(r, out); return r.getNumSamples(); } catch (const std::exception&) { return -1; }
}
int fac_deconvolve(const float* num, int Ln, const float* den, int Ld, double sr, int smoothing, int phase, int ampl, float* out) {
    AudioBuffer<float> bn(1, Ln), bd(1, Ld);
    fill(bn, num); fill(bd, den);
    try { AudioBuffer<float> r = fp::convolution::deconvolve(&bn, &bd, sr, smoothing, phase, ampl); dump(r, out); return r.getNumSamples(); } catch (const std::exception&) { return -1; }
}
int fac_averaging_filter(float* spec, int ch, int fftSize, double oct, double sr, int logAvg, int phase, int ampl) {
    AudioBuffer<float> b(ch, fftSize);
    fill(b, spec);
    try { fp::convolution::averagingFilter(&b, oct, sr, logAvg, phase, ampl); dump(b, spec); return 0; } catch (const std::exception&) { return -1; }
}
int fac_fft_roundtrip(const float* x, int ch, int L, float* spec, float* back) {
    AudioBuffer<float> b(ch, L);
    fill(b, x);
    try {
        AudioBuffer<float> s = fp::tools::fftTransform(b);
        dump(s, spec);
        AudioBuffer<float> r = fp::tools::fftInvTransform(s);
        dump(r, back);
        return s.getNumSamples();
    } catch (const std::exception&) { return -1; }
}
int fac_invert_filter(const float* x, int L, int sr, float* out) {
    AudioBuffer<float> b(1, L);
    fill(b, x);
    try { AudioBuffer<float> r = fp::ir::invertFilter(b, sr); dump(r, out); return r.getNumSamples(); } catch (const std::exception&) { return -1; }
}
int fac_ir_to_real_fft_raw(const float* x, int L, int part, float* out) {
    AudioBuffer<float> b(1, L);
    fill(b, x);
    try { AudioBuffer<float> r = fp::ir::IRtoRealFFTRaw(b, part); dump(r, out); return r.getNumSamples(); } catch (const std::exception&) { return -1; }
}
int fac_ir_chop(const float* x, int L, int IRlength, float thr, int consecutive, float* out) {
    AudioBuffer<float> b(1, L);
    fill(b, x);
    AudioBuffer<float> r = fp::ir::IRchop(b, IRlength, thr, consecutive);
    dump(r, out);
    return r.getNumSamples();
}
void fac_shifteroo(float* buf, int ch, int n) { AudioBuffer<float> b(ch, n); fill(b, buf); fp::ir::shifteroo(&b); dump(b, buf); }
void fac_sum_to_mono(float* buf, int ch, int n) { AudioBuffer<float> b(ch, n); fill(b, buf); fp::tools::sumToMono(&b); dump(b, buf); }
void fac_linear_fade(float* buf, int ch, int n, int fadeIn, int start, int count) { AudioBuffer<float> b(ch, n); fill(b, buf); fp::tools::linearFade(&b, fadeIn, start, count); dump(b, buf); }
void fac_normalize(float* buf, int ch, int n, float dB) { AudioBuffer<float> b(ch, n); fill(b, buf); fp::tools::normalize(&b, dB); dump(b, buf); }
void fac_sine_fill(float* buf, int ch, int n, float f, float sr, float a) { AudioBuffer<float> b(ch, n); fp::tools::sineFill(&b, f, sr, a); dump(b, buf); }
void fac_generate_pulse(int n, int off, float* out) { AudioBuffer<float> r = fp::tools::generatePulse(n, off); dump(r, out); }
void fac_complex_ops(float* v) {   // v = {a, b, c, d} -> {mul.re, mul.im, div.re, div.im, polar.re, polar.im, ampl, phase}
    float a = v[0], b = v[1]; fp::tools::complexMul(&a, &b, v[2], v[3]); float m0 = a, m1 = b;
    a = v[0]; b = v[1]; fp::tools::complexDivCartesian(&a, &b, v[2], v[3]); float d0 = a, d1 = b;
    a = v[0]; b = v[1]; fp::tools::complexDivPolar(&a, &b, v[2], v[3]);
    float bin[2] = {v[0], v[1]};
    float am = fp::tools::binAmpl(bin), ph = fp::tools::binPhase(bin);
    v[0] = m0; v[1] = m1; v[2] = d0; v[3] = d1; v[4] = a; v[5] = b; v[6] = am; v[7] = ph;
}
float fac_db_to_lin(float dB) { return fp::tools::dBToLin(dB); }
float fac_lin_to_db(float lin) { return fp::tools::linTodB(lin); }
int fac_next_pow2(int x) { return fp::tools::nextPowerOfTwo(x); }
void fac_round(float* v) { fp::tools::roundToZero(v, 1e-11f); fp::tools::roundTo1TenQuadrillionth(v + 1); }
// mode 0 sweep, 1 inverse; fadeKind 0 none, 1 lin, 2 dB, 3 brickwall (applied before the inverse is built)
int fac_ess(double dur, double sr, double f1, double f2, double dB, int mode, int fadeKind, double fadeFreq, double* out) {
    fp::ExpSineSweep s;
    try { s.generate(dur, sr, f1, f2, dB); } catch (const std::exception&) { return -1; }
    if (fadeKind == 1) s.linFadeout(fadeFreq);
    if (fadeKind == 2) s.dBFadeout(fadeFreq);
    if (fadeKind == 3) s.brickwallFadeout(fadeFreq);
    AudioBuffer<double> r;
    if (mode == 1) { s.generateInv(); r = s.getSweepInv(); } else r = s.getSweep();
    if (out) std::copy_n(r.getReadPointer(0), r.getNumSamples(), out);
    return r.getNumSamples();
}
int fac_ess_index_at_freq(double freq, double dur, double sr, double f1, double f2) { fp::ExpSineSweep s; return s.getSampleIndexAtFreq(freq, dur, sr, f1, f2); }
double fac_ess_freq_at_index(int idx, double dur, double sr, double f1, double f2) { fp::ExpSineSweep s; return s.getFreqAtSampleIndex(idx, dur, sr, f1, f2); }
}

// ---- fp::b200::PluginConvolver: processBlock semantics over the engine ----------------------------------------
#include "../../irbaboon_b200/fp/PluginConvolver.hpp"
extern "C" {
void* fac_plugin_create(int B, int channels, int hostBlock, const float* ir, int irLen, int exactOrder, float volumedB) {
    try {
        auto* p = new fp::b200::PluginConvolver(B, channels, 0);
        AudioBuffer<float> h(1, irLen);
        fill(h, ir);
        p->setExactReferenceOrder(exactOrder != 0);
        p->setOutputVolumedB(volumedB);
        p->prepareToPlay(48000.0, hostBlock, h);
        return p;
    } catch (const std::exception& e) { std::fprintf(stderr, "fac_plugin_create: %s\n", e.what()); return nullptr; }
}
void fac_plugin_destroy(void* h) { delete (fp::b200::PluginConvolver*) h; }
int fac_plugin_latency(void* h) { return ((fp::b200::PluginConvolver*) h)->getLatencySamples(); }
int fac_plugin_set_ir(void* h, const float* ir, int irLen) {
    AudioBuffer<float> b(1, irLen);
    fill(b, ir);
    try { ((fp::b200::PluginConvolver*) h)->setIR(b); return 0; } catch (const std::exception&) { return -1; }
}
// buffer: [channels][n] planar, in place
int fac_plugin_process(void* h, float* buffer, int channels, int n, int bypassed) {
    AudioBuffer<float> b(channels, n);
    fill(b, buffer);
    try {
        if (bypassed) ((fp::b200::PluginConvolver*) h)->processBlockBypassed(b);
        else ((fp::b200::PluginConvolver*) h)->processBlock(b);
        dump(b, buffer);
        return 0;
    } catch (const std::exception& e) { std::fprintf(stderr, "fac_plugin_process: %s\n", e.what()); return -1; }
}
}

// ---- formats (N4) and the capture -> filter chain (N3) --------------------------------------------------------
#include "../../irbaboon_b200/fp/CaptureChain.hpp"
#include "../../irbaboon_b200/fp/Formats.hpp"
extern "C" {
int fac_write_wav(const char* path, const float* data, int ch, int n, int sr, int bits) {
    AudioBuffer<float> b(ch, n);
    fill(b, data);
    return fp::b200::formats::writeWav(path, b, sr, bits) ? 0 : -1;
}
int fac_read_wav(const char* path, float* out, int capacity, int* ch, int* n, int* sr) {
    AudioBuffer<float> b = fp::b200::formats::readWav(path, sr);
    *ch = b.getNumChannels(); *n = b.getNumSamples();
    if (*ch == 0) return -1;
    if ((long long) *ch * *n > capacity) return -2;
    dump(b, out);
    return 0;
}
int fac_write_sweep_and_ir(const char* path, const float* sweep, int ns, const float* ir, int ni, int sr, int num) {
    AudioBuffer<float> s(1, ns), i(1, ni);
    fill(s, sweep); fill(i, ir);
    return fp::b200::formats::writeSweepAndIR(path, s, i, sr, num) ? 0 : -1;
}
int fac_read_sweep_and_ir(const char* path, float* sweep, float* ir, int num) {
    AudioBuffer<float> s, i;
    if (!fp::b200::formats::readSweepAndIR(path, s, i, num)) return -1;
    dump(s, sweep); dump(i, ir);
    return 0;
}
int fac_write_spectrum_tsv(const char* path, const char* name, const float* spec, int fftSize, int sr) {
    AudioBuffer<float> b(1, fftSize);
    fill(b, spec);
    return fp::b200::formats::writeSpectrumTsv(path, name, b, sr) ? 0 : -1;
}
int fac_write_raw_text(const char* path, const float* packed, int n, int part) {
    AudioBuffer<float> b(1, n);
    fill(b, packed);
    return fp::b200::formats::writeRawSpectraText(path, b, part) ? 0 : -1;
}
// blocks: [nb][H] captured host buffers (mono) -> ring -> consolidate -> deconvolve against the sweep; returns N
int fac_capture_to_ir(const float* blocks, int nb, int H, const float* sweep, int ns, double sr, float* recording, float* ir) {
    try {
        fp::CircularBufferArray ring(nb, 1, H);
        for (int b = 0; b < nb; ++b) {                                    // PluginProcessor.cpp:291-292
            ring.getWriteBufferPtr()->copyFrom(0, 0, blocks + (size_t) b * H, H);
            ring.incrWriteIndex();
        }
        AudioBuffer<float> sw(1, ns), rec;
        fill(sw, sweep);
        AudioBuffer<float> r = fp::b200::captureToIR(ring, sw, sr, &rec);
        dump(rec, recording); dump(r, ir);
        return r.getNumSamples();
    } catch (const std::exception& e) { std::fprintf(stderr, "fac_capture_to_ir: %s\n", e.what()); return -1; }
}
int fac_create_ir_filt(const float* targ, int nt, const float* base, int nbse, double sr, int phase, int ampl, float* out) {
    try {
        AudioBuffer<float> t(1, nt), b(1, nbse);
        fill(t, targ); fill(b, base);
        AudioBuffer<float> r = fp::b200::createIRFilt(t, b, sr, phase != 0, ampl != 0);
        dump(r, out);
        return r.getNumSamples();
    } catch (const std::exception&) { return -1; }
}
int fac_chop_and_normalize(const float* x, int n, int len, float thr, int cons, float* out) {
    AudioBuffer<float> b(1, n);
    fill(b, x);
    AudioBuffer<float> r = fp::b200::chopAndNormalize(b, len, thr, cons);
    dump(r, out);
    return r.getNumSamples();
}
}
