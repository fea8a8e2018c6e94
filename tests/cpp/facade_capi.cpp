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
