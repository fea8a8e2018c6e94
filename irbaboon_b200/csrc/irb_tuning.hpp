// irb_tuning.hpp -- launch-policy knobs of libirb_b200.so.  The product path reads these plain globals with the defaults
// below and nothing else (no environment variable is consulted anywhere in the library); the only writer is
// irbx_set_tuning() in irb_benchaids.cu, the bench-only translation unit behind include/irb_b200_bench.h, which
// A/B measurements and tests use to select the other form of a kernel.
#pragma once

namespace irbh {

struct Tuning {
    int mac_persistent = 1;      // shared-IR / per-stream-IR block step as the persistent TMA kernel k_mac_p (0: k_mac_tma / k_mac_slots)
    int fuse_split = 1;          // few rows (partitions split over a cluster): forward transform inside the cluster kernel (0: k_fwd launch first)
    int mac_tma = 1;             // non-persistent fused step: FDL through TMA (k_mac_tma); 0: register-staged k_mac<FUSE>
    int mac_wide = 1;            // register-staged MAC: 32-byte FDL loads (0: 16-byte)
    int mac_u = 1;               // register-staged MAC: IR partitions per ring stage (1 or 2)
    int fdl_plain = 0;           // engines created from now on store the FDL [chan][slot][M] instead of tile-interleaved
    int producer_sleep_ns = 200; // the TMA producer sleeps this long between polls of a busy stage (0: spin)
    int no_graph = 0;            // single small blocks: plain launches instead of the captured CUDA graph
    int deconv_sub = 0;          // captures per sub-batch of irb_deconvolve_batch (0: about 48 MB of spectra)
    int avg_fused = 1;           // log-average smoothing: all passes in one wavefront launch (0: one scan + one rebuild kernel per pass)
    int deconv_groups = 0;       // smoothed batch deconvolution: groups of captures smoothed side by side (0: four)
    int deconv_group_cap = 0;    // smoothed batch deconvolution: most captures per group (0: about 1.5 GB of spectra and sums)
    int deconv_streams = 0;      // irb_deconvolve_batch_device: compute streams the sub-batches alternate between (0: two)
    int release_fence = 1;       // fence.proxy.async between the last ld.shared of a ring stage and its release
    int release_dep = 1;         // the release also carries a data dependency on the values read (0 + 0 = the unguarded round-1 form: sanitizer experiments only)
    int persistent_ctas = 0;     // k_mac_p: CTAs per SM (0: the kernel's own choice)
    int ir_replicas = 1;         // copies of the shared IR spectra an engine created from now on keeps (0: as many as fit 24 MB, at most 32); measured: no effect
    int ring_stages = 0;         // k_mac_p: ring stages in use (0: all the kernel has; measurement of the in-flight depth)
    int stagger_ns = 0;          // k_mac_p: CTA starts spread over this many nanoseconds
    int unit_narrowing = 0;      // k_mac_p: run the last partial wave of a launch on tiles of fewer rows (measured slower: a narrow unit has less in flight)
};
extern Tuning g_tuning;

}  // namespace irbh
