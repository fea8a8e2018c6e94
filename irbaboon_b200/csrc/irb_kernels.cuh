// irb_kernels.cuh -- sm_100a kernels of the partitioned-convolution (UPOLA) block step.
//
// One CTA owns a TILE of 2048 packed complex bins = ROWS = 2048/M whole spectra ("rows": one stream-channel
// in the streaming engine, one output block in the offline functions).  256 compute threads work in two
// layouts over the same 16 KB of shared memory:
//   * FFT layout : M/8 threads per row, 8 complex points each (irb_fft.cuh);
//   * MAC layout : each thread owns V float4 (= 2V bins) of K rows, K*V = 4, so that one warp instruction
//                  reads 512 contiguous bytes of a frequency-domain delay line (FDL) row.
// A 9th warp is the TMA producer that streams impulse-response partition spectra into a shared-memory
// ring (cp.async.bulk + mbarrier), so every row of the tile -- every stream sharing that IR -- reuses them.
//
// Reference loops replaced (paths relative to /root/reference):
//   k_fwd      fp/convolution.cpp:106-125 (IR partition load + FFT), :128-149 (audio block load + FFT),
//              Source/PluginProcessor.cpp:430-436,455-461
//   k_mac      fp/convolution.cpp:160-215 (MAC over partitions, inverse FFT, overlap-add),
//              Source/PluginProcessor.cpp:480-510
//   k_ola_tail fp/convolution.cpp:210-213 for the offline (all blocks at once) formulation
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "irb_fft.cuh"

namespace irb {

constexpr int kThreads = 256;          // compute threads per CTA
constexpr int kTile = 2048;            // packed complex bins per CTA tile (16 KB)
constexpr int kStages = 4;             // IR ring stages

template <int M> struct Tile {
    static constexpr int V = M > 512 ? M / 512 : 1;      // float4 per thread per row
    static constexpr int TPR = M / (2 * V);              // MAC-layout threads per row
    static constexpr int G = kThreads / TPR;             // row groups
    static constexpr int K = 4 / V;                      // rows per MAC thread
    static constexpr int ROWS = kTile / M;               // rows per CTA (= G*K)
    static constexpr int TPF = M / kPts;                 // FFT-layout threads per row
};

// ---- PTX wrappers: named barrier, mbarrier, 1-D bulk TMA, streaming 128-bit load -----------------
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory"); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    while (!mbar_try_wait(b, parity)) {}
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// ---- full M-point transform of the 8 registers of an FFT-layout thread ----------------------------
template <int M, bool INV, int PASS = 0, int PS = 1>
__device__ __forceinline__ void fft_run(float2 (&v)[kPts], int t, float2* srow, const float2* __restrict__ W) {
    constexpr int R = pass_radix(M, PASS);
    fft_pass<M, R, PS, INV>(v, t, W);
    if constexpr (PS * R < M) {
        bar_compute();                       // every earlier read of the tile has completed
        fft_scatter<M, R, PS>(v, t, srow);
        bar_compute();
        fft_gather<M>(v, t, srow);
        fft_run<M, INV, PASS + 1, PS * R>(v, t, srow, W);
    }
}

// ===================================================================================================
// k_fwd: rows of <= B real samples -> zero-padded 2M-point real FFT -> packed spectrum rows.
// row = chan * blocks_per_chan + blk; the samples are src[chan*src_chan_stride + blk*B + i], i < clamp(L - blk*B, 0, B).
// Destination slot: blk (offline / IR preparation) or, when head != nullptr, the next ring slot of chan
// (streaming: the kernel also advances head[chan]).
struct FwdArgs {
    const float* src;          // device samples
    const float* src2;         // optional second channel folded in as (a + b) / 2  (tools::sumToMono, fp/tools.cpp:25-29)
    long long src_chan_stride; // floats
    int L;                     // samples per channel
    int B;                     // block size (<= M)
    int blocks_per_chan;
    int n_rows;
    float2* dst;               // packed spectra
    long long dst_chan_stride; // float2 units
    int* head;                 // per-chan ring head (nullable)
    int ring;                  // ring slots per chan
    const float2* W;           // N = 2M roots of unity
};

template <int M>
__global__ void __launch_bounds__(kThreads) k_fwd(const FwdArgs a) {
    using T = Tile<M>;
    __shared__ __align__(16) float2 s_spec[kTile];
    const int tid = threadIdx.x;
    {   // FFT layout
        const int rf = tid / T::TPF, t = tid % T::TPF;
        const int row = blockIdx.x * T::ROWS + rf;
        float2 v[kPts];
#pragma unroll
        for (int j = 0; j < kPts; ++j) v[j] = make_float2(0.f, 0.f);
        if (row < a.n_rows) {
            const int chan = row / a.blocks_per_chan, blk = row % a.blocks_per_chan;
            int len = a.L - blk * a.B;
            len = len < 0 ? 0 : (len > a.B ? a.B : len);
            const long long off = chan * a.src_chan_stride + (long long) blk * a.B;
            const float* p = a.src + off;
            const float* q = a.src2 ? a.src2 + off : nullptr;
#pragma unroll
            for (int j = 0; j < kPts; ++j) {
                const int m = 2 * (t + j * T::TPF);
                float x0 = 0.f, x1 = 0.f;
                if (m < len) x0 = p[m];
                if (m + 1 < len) x1 = p[m + 1];
                if (q) {
                    if (m < len) { x0 += q[m]; x0 /= 2.0f; }
                    if (m + 1 < len) { x1 += q[m + 1]; x1 /= 2.0f; }
                }
                v[j] = make_float2(x0, x1);
            }
        }
        float2* srow = s_spec + rf * M;
        fft_run<M, false>(v, t, srow, a.W);
        bar_compute();
#pragma unroll
        for (int j = 0; j < kPts; ++j) srow[t + j * T::TPF] = v[j];
    }
    bar_compute();
    {   // MAC layout: split into the packed real spectrum and store 16 B per lane
        const int g = tid / T::TPR, c0 = tid % T::TPR;
#pragma unroll
        for (int s = 0; s < T::K; ++s) {
            const int rl = s * T::G + g;
            const int row = blockIdx.x * T::ROWS + rl;
            if (row >= a.n_rows) continue;
            const int chan = row / a.blocks_per_chan, blk = row % a.blocks_per_chan;
            int slot = blk;
            if (a.head) { slot = a.head[chan] + 1; if (slot >= a.ring) slot = 0; }
            const float2* z = s_spec + rl * M;
            float2* d = a.dst + chan * a.dst_chan_stride + (long long) slot * M;
#pragma unroll
            for (int vv = 0; vv < T::V; ++vv) {
                const int k = 2 * (c0 + vv * T::TPR);
                const float2 x0 = real_split(z[k], z[(M - k) & (M - 1)], root<false>(a.W, k), k);
                const float2 x1 = real_split(z[k + 1], z[M - k - 1], root<false>(a.W, k + 1), k + 1);
                *reinterpret_cast<float4*>(d + k) = make_float4(x0.x, x0.y, x1.x, x1.y);
            }
        }
    }
    if (a.head) {   // advance the ring heads once every read of the old value is done
        bar_compute();
        if (tid < T::ROWS) {
            const int row = blockIdx.x * T::ROWS + tid;
            if (row < a.n_rows) { int h = a.head[row] + 1; a.head[row] = h >= a.ring ? 0 : h; }
        }
    }
}

// ===================================================================================================
// k_mac: Y[row][bin] = sum_{p < nvalid} FDL[row][(head - p) mod ring][bin] * H[ir][p][bin], ascending p
// (the summation order of fp/convolution.cpp:171-202), then -- when INV -- inverse FFT and overlap-add.
struct MacArgs {
    const float2* fdl;         // packed spectra rows: fdl + chan*fdl_chan_stride + slot*M
    long long fdl_chan_stride; // float2 units
    const int* head;           // per-chan newest slot; nullptr => head = blk (offline: block index)
    int ring;                  // slots per chan (wrap); offline: unused because nvalid <= blk+1
    int blocks_per_chan;
    int n_rows;
    const float2* H;           // IR partition spectra: H + ir*ir_stride + p*M
    long long ir_stride;       // float2 units
    const int* ir_of_chan;     // nullptr => IR 0 for every row
    const int* nparts;         // partitions per IR
    const float2* W;
    // !INV: accumulator spectra out
    float2* Y;                 // [row][M]
    // INV: time-domain epilogue
    int B;
    float* out;                // out + chan*out_chan_stride + blk*B + m, m < clamp(Lout - blk*B, 0, B)
    long long out_chan_stride;
    int Lout;
    float* ov;                 // streaming: overlap state [chan][B] (read, then replaced); offline: nullptr
    float* tail;               // offline: second halves [row][B]
};

// PRI = per-row IR: every row of the tile stages its OWN partition spectrum (per-stream IRs, BASELINE config 4);
// otherwise the whole tile shares one IR and a ring stage holds U consecutive partitions of it.
template <int M, int U, bool PRI>
struct MacSmem {
    float2 spec[kTile];                                  // tile in both layouts (inverse path)
    float2 h[kStages][PRI ? kTile / M : U][M];           // IR ring
    uint64_t full[kStages], empty[kStages];
};

template <int M, int U, bool INV, bool PRI>
__global__ void __launch_bounds__(kThreads + 32, (U == 1 && !PRI) ? 3 : 2) k_mac(const MacArgs a) {
    static_assert(!PRI || U == 1, "per-row IR stages one partition per row");
    using T = Tile<M>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    MacSmem<M, U, PRI>& sm = *reinterpret_cast<MacSmem<M, U, PRI>*>(smem_raw);
    const int tid = threadIdx.x;
    const int row0 = blockIdx.x * T::ROWS;

    // shared mode: by contract every row of a tile is bound to the same IR (the host checks / pads)
    const int chan_first = row0 / a.blocks_per_chan;
    const int ir = a.ir_of_chan ? a.ir_of_chan[chan_first] : 0;
    const int np = a.nparts[ir];
    // number of partitions any row of this tile needs
    int pmax = np;
    if (PRI) {
        pmax = 0;
        for (int r = 0; r < T::ROWS && row0 + r < a.n_rows; ++r) {
            const int chan = (row0 + r) / a.blocks_per_chan;
            const int n = a.nparts[a.ir_of_chan ? a.ir_of_chan[chan] : 0];
            pmax = n > pmax ? n : pmax;
        }
    }
    if (!a.head) {          // offline: row = chan*blocks_per_chan + blk, blocks_per_chan % ROWS == 0 (host pads)
        int last = row0 + T::ROWS - 1;
        if (last >= a.n_rows) last = a.n_rows - 1;
        const int kmax = last % a.blocks_per_chan;
        if (kmax + 1 < pmax) pmax = kmax + 1;
    }
    const int ngroups = (pmax + U - 1) / U;

    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&sm.full[i], 1); mbar_init(&sm.empty[i], kThreads / 32); }
        mbar_fence_init();
    }
    __syncthreads();

    if (tid >= kThreads) {
        // ===== TMA producer warp: stream the IR partition spectra through the ring =====
        if (PRI) {
            // one bulk copy per (row, partition); lanes split the rows of the tile
            const int lane = tid - kThreads;
            for (int g = 0; g < ngroups; ++g) {
                const int st = g % kStages;
                if (g >= kStages) mbar_wait(&sm.empty[st], ((g / kStages) - 1) & 1);
                uint32_t mine = 0;
                for (int r = lane; r < T::ROWS && row0 + r < a.n_rows; r += 32) {
                    const int chan = (row0 + r) / a.blocks_per_chan;
                    if (g < a.nparts[a.ir_of_chan ? a.ir_of_chan[chan] : 0]) mine += M * sizeof(float2);
                }
                uint32_t total = mine;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
                if (lane == 0) mbar_expect_tx(&sm.full[st], total);
                __syncwarp();
                for (int r = lane; r < T::ROWS && row0 + r < a.n_rows; r += 32) {
                    const int chan = (row0 + r) / a.blocks_per_chan;
                    const int irr = a.ir_of_chan ? a.ir_of_chan[chan] : 0;
                    if (g < a.nparts[irr]) tma_bulk_g2s(&sm.h[st][r][0], a.H + irr * a.ir_stride + (long long) g * M, M * sizeof(float2), &sm.full[st]);
                }
            }
        } else if (tid == kThreads) {
            const float2* hsrc = a.H + ir * a.ir_stride;
            for (int g = 0; g < ngroups; ++g) {
                const int st = g % kStages;
                if (g >= kStages) mbar_wait(&sm.empty[st], ((g / kStages) - 1) & 1);
                int cnt = pmax - g * U;
                if (cnt > U) cnt = U;
                const uint32_t bytes = (uint32_t) cnt * M * sizeof(float2);
                mbar_expect_tx(&sm.full[st], bytes);
                tma_bulk_g2s(&sm.h[st][0][0], hsrc + (long long) g * U * M, bytes, &sm.full[st]);
            }
        }
        return;
    }

    // ===== compute threads, MAC layout =====
    const int g_ = tid / T::TPR, c0 = tid % T::TPR;
    const float4* xptr[T::K];      // points at the float4 of partition 0 (the newest slot)
    int slot[T::K], nvalid[T::K];
#pragma unroll
    for (int s = 0; s < T::K; ++s) {
        const int row = row0 + s * T::G + g_;
        nvalid[s] = 0; slot[s] = 0; xptr[s] = nullptr;
        if (row < a.n_rows) {
            const int chan = row / a.blocks_per_chan, blk = row % a.blocks_per_chan;
            const int hd = a.head ? a.head[chan] : blk;
            const int npr = PRI ? a.nparts[a.ir_of_chan ? a.ir_of_chan[chan] : 0] : np;
            nvalid[s] = a.head ? npr : (blk + 1 < npr ? blk + 1 : npr);
            slot[s] = hd;
            xptr[s] = reinterpret_cast<const float4*>(a.fdl + chan * a.fdl_chan_stride + (long long) hd * M) + c0;
        }
    }
    float4 acc[T::K][T::V];
#pragma unroll
    for (int s = 0; s < T::K; ++s)
#pragma unroll
        for (int vv = 0; vv < T::V; ++vv) acc[s][vv] = make_float4(0.f, 0.f, 0.f, 0.f);

    float4 xa[U][T::K][T::V], xb[U][T::K][T::V];
    auto load_group = [&](float4 (&x)[U][T::K][T::V], int g) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = g * U + u;
#pragma unroll
            for (int s = 0; s < T::K; ++s) {
                const bool ok = p < nvalid[s];
#pragma unroll
                for (int vv = 0; vv < T::V; ++vv)
                    x[u][s][vv] = ok ? ldg_stream(xptr[s] + vv * T::TPR) : make_float4(0.f, 0.f, 0.f, 0.f);
                // step one slot back in the ring (wrap to the top)
                if (ok) {
                    if (slot[s] == 0) { slot[s] = a.ring - 1; xptr[s] += (long long) (a.ring - 1) * (M / 2); }
                    else { --slot[s]; xptr[s] -= M / 2; }
                }
            }
        }
    };
    auto consume_group = [&](float4 (&x)[U][T::K][T::V], int g) {
        const int st = g % kStages;
        mbar_wait(&sm.full[st], (g / kStages) & 1);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (g * U + u < pmax) {
#pragma unroll
                for (int vv = 0; vv < T::V; ++vv) {
                    const int c = c0 + vv * T::TPR;
                    float4 h;
                    float h0i, h0q;
                    if (!PRI) {
                        h = *reinterpret_cast<const float4*>(&sm.h[st][u][2 * c]);
                        // bin 0 is the packed {DC, Nyquist} pair: two real products instead of a complex one
                        h0i = c == 0 ? 0.f : h.y; h0q = c == 0 ? h.y : h.x;
                    }
#pragma unroll
                    for (int s = 0; s < T::K; ++s) {
                        if (PRI) {
                            if (g >= nvalid[s]) continue;      // this row's IR is shorter: its stage slot was not filled
                            h = *reinterpret_cast<const float4*>(&sm.h[st][s * T::G + g_][2 * c]);
                            h0i = c == 0 ? 0.f : h.y; h0q = c == 0 ? h.y : h.x;
                        }
                        const float4 xv = x[u][s][vv];
                        float4& ac = acc[s][vv];
                        ac.x = fmaf(xv.x, h.x, fmaf(-xv.y, h0i, ac.x));
                        ac.y = fmaf(xv.y, h0q, fmaf(xv.x, h0i, ac.y));
                        ac.z = fmaf(xv.z, h.z, fmaf(-xv.w, h.w, ac.z));
                        ac.w = fmaf(xv.w, h.z, fmaf(xv.z, h.w, ac.w));
                    }
                }
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&sm.empty[st]);
    };

    if (ngroups > 0) load_group(xa, 0);
    for (int g = 0; g < ngroups; g += 2) {
        if (g + 1 < ngroups) load_group(xb, g + 1);
        consume_group(xa, g);
        if (g + 1 < ngroups) {
            if (g + 2 < ngroups) load_group(xa, g + 2);
            consume_group(xb, g + 1);
        }
    }

    if constexpr (!INV) {
#pragma unroll
        for (int s = 0; s < T::K; ++s) {
            const int row = row0 + s * T::G + g_;
            if (row >= a.n_rows) continue;
#pragma unroll
            for (int vv = 0; vv < T::V; ++vv)
                reinterpret_cast<float4*>(a.Y + (long long) row * M)[c0 + vv * T::TPR] = acc[s][vv];
        }
    } else {
        // ---- accumulators -> shared tile (MAC layout), then inverse real FFT in FFT layout ----
#pragma unroll
        for (int s = 0; s < T::K; ++s)
#pragma unroll
            for (int vv = 0; vv < T::V; ++vv)
                reinterpret_cast<float4*>(sm.spec + (s * T::G + g_) * M)[c0 + vv * T::TPR] = acc[s][vv];
        bar_compute();
        const int rf = tid / T::TPF, t = tid % T::TPF;
        float2* srow = sm.spec + rf * M;
        float2 v[kPts];
#pragma unroll
        for (int j = 0; j < kPts; ++j) {
            const int k = t + j * T::TPF;
            v[j] = real_merge(srow[k], srow[(M - k) & (M - 1)], root<false>(a.W, k), k);
        }
        fft_run<M, true>(v, t, srow, a.W);
        const float scale = 1.0f / (float) (2 * M);      // the 1/N of performRealOnlyInverseTransform
        const int row = row0 + rf;
        const bool live = row < a.n_rows;
        const int chan = live ? row / a.blocks_per_chan : 0, blk = live ? row % a.blocks_per_chan : 0;
        int olen = a.Lout - blk * a.B;
        olen = olen < 0 ? 0 : (olen > a.B ? a.B : olen);
        float* outp = a.out + chan * a.out_chan_stride + (long long) blk * a.B;
        float* ovp = a.ov ? a.ov + (long long) chan * a.B : nullptr;
        float* tailp = a.tail ? a.tail + (long long) row * a.B : nullptr;
        // fp/convolution.cpp:210-213: y[i] += overlap[i]; overlap[i] = y[B + i]
#pragma unroll
        for (int j = 0; j < kPts; ++j) {
            v[j].x *= scale; v[j].y *= scale;
            const int m = 2 * (t + j * T::TPF);
            if (live && ovp) {
                if (m < a.B) v[j].x += ovp[m];
                if (m + 1 < a.B) v[j].y += ovp[m + 1];
            }
        }
        if (a.ov) bar_compute();                           // all overlap reads precede the overlap writes
        if (live) {
#pragma unroll
            for (int j = 0; j < kPts; ++j) {
                const int m = 2 * (t + j * T::TPF);
                const float val[2] = {v[j].x, v[j].y};
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int mm = m + e;
                    if (mm < a.B) { if (mm < olen) outp[mm] = val[e]; }
                    else if (mm < 2 * a.B) { if (ovp) ovp[mm - a.B] = val[e]; else tailp[mm - a.B] = val[e]; }
                }
            }
        }
    }
}

// offline: out[chan][blk*B + m] += tail[chan][blk-1][m]   (the overlap of the previous block)
static __global__ void k_ola_tail(float* out, long long out_chan_stride, int Lout, const float* tail, int B, int blocks_per_chan, int n_chans) {
    const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long per_chan = (long long) blocks_per_chan * B;
    if (i >= per_chan * n_chans) return;
    const int chan = (int) (i / per_chan);
    const long long r = i % per_chan;
    const int blk = (int) (r / B), m = (int) (r % B);
    if (blk == 0) return;
    const long long o = (long long) blk * B + m;
    if (o >= Lout) return;
    out[chan * out_chan_stride + o] += tail[((long long) chan * blocks_per_chan + (blk - 1)) * B + m];
}

}  // namespace irb
