"""CPU: the dependency rule of the wavefront smoothing launch (k_avg_passes, irb_spectral.cuh).

The three log-average passes of fp::convolution::averagingFilter (fp/convolution.cpp:406-546, called three times at :389-394) run
inside one CTA per spectrum; pass p + 1 gathers the operands of its chunk c only after pass p has rebuilt every bin up to
need(c) = min(M, hi[min(M, kstart[c + 1])]).  This restates the host-independent plan (window edges, operation list, chunk starts)
with numpy exactly as k_avg_windows / k_avg_oplist / k_avg_chunk_starts lay it out and checks, for a range of sizes, sample rates
and octave fractions, that no operation of chunk c ever touches a bin above need(c) -- i.e. that the wait is sufficient -- and that
the wait can always be satisfied (no deadlock)."""
import numpy as np
import pytest

CHUNK = 2048                                                 # irb::kAvgChunk


def _plan(M, sample_rate, octave_fraction):
    freq_per_bin = (sample_rate / 2) / M                     # fp/convolution.cpp:425
    c_side = 2.0 ** (octave_fraction / 2.0)
    k = np.arange(M + 1, dtype=np.float64)
    f = k * freq_per_bin
    lo = np.round((f / c_side) / freq_per_bin).astype(np.int64)     # fp/convolution.cpp:451-458 (round half away from zero; values >= 0)
    hi = np.round((f * c_side) / freq_per_bin).astype(np.int64)
    # np.round is half-to-even, C's round() half-away-from-zero: fix the exact .5 cases
    for arr, val in ((lo, (f / c_side) / freq_per_bin), (hi, (f * c_side) / freq_per_bin)):
        half = (val - np.floor(val)) == 0.5
        arr[half] = np.floor(val[half]).astype(np.int64) + 1
    endq = lo + hi
    n_ops = int(endq[M]) + 1
    ops = np.empty(n_ops, np.int64)
    l0 = np.concatenate(([0], lo[:-1])); h0 = np.concatenate(([-1], hi[:-1]))
    for kk in range(M + 1):                                  # k_avg_oplist
        j = l0[kk] + h0[kk] + 1
        nsub = lo[kk] - l0[kk]
        ops[j:j + nsub] = ~np.arange(l0[kk], lo[kk])
        nadd = hi[kk] - h0[kk]
        ops[j + nsub:j + nsub + nadd] = np.arange(h0[kk] + 1, hi[kk] + 1)
        assert j + nsub + nadd - 1 == endq[kk]
    nchunks = (n_ops + CHUNK - 1) // CHUNK
    kstart = np.searchsorted(endq // CHUNK, np.arange(nchunks + 1), side="left")      # first bin recorded in chunk c; M + 1 past the end
    return lo, hi, endq, ops, kstart, nchunks, n_ops


@pytest.mark.parametrize("M,sr,frac", [(256, 48000.0, 1 / 13), (2048, 48000.0, 1 / 13), (8192, 44100.0, 1 / 13), (1 << 15, 48000.0, 1 / 13),
                                        (1 << 15, 96000.0, 1 / 3), (4096, 48000.0, 1.0), (5000, 48000.0, 1 / 24), (1 << 17, 48000.0, 1 / 13)])
def test_wait_rule_covers_every_operand_and_cannot_deadlock(M, sr, frac):
    lo, hi, endq, ops, kstart, nchunks, n_ops = _plan(M, sr, frac)
    assert np.all(np.diff(hi) >= 0) and np.all(np.diff(lo) >= 0) and np.all(lo <= hi)
    assert kstart[nchunks] == M + 1 and kstart[0] == 0 and np.all(np.diff(kstart) >= 0)
    # the host's closed form for the list length (Smoother::init)
    f = M * ((sr / 2) / M); c = 2.0 ** (frac / 2.0); fpb = (sr / 2) / M
    assert n_ops == int(round((f / c) / fpb)) + int(round((f * c) / fpb)) + 1
    bins = np.where(ops < 0, ~ops, ops)
    for cidx in range(nchunks):
        seg = bins[cidx * CHUNK:(cidx + 1) * CHUNK]
        seg = seg[seg <= M]                                  # bins past Nyquist hold a constant, not the previous pass's output
        need = min(M, int(hi[min(M, int(kstart[cidx + 1]))]))
        assert seg.size == 0 or int(seg.max()) <= need, (cidx, int(seg.max()), need)
        assert need <= M                                     # pass p publishes M + 1 when it is done: the wait always ends
    # every bin is recorded inside the chunk kstart says it is, so the rebuild of chunk c covers [kstart[c], kstart[c + 1])
    rec_chunk = endq // CHUNK
    for cidx in range(nchunks):
        ks = np.arange(kstart[cidx], kstart[cidx + 1])
        assert np.all(rec_chunk[ks] == cidx)


def test_wavefront_lag_is_a_few_percent():
    """How far pass p + 1 trails pass p: the bins it waits for are at most hi[k] ~ 1.027 k ahead, i.e. three passes cost about one."""
    M = 1 << 16
    lo, hi, endq, ops, kstart, nchunks, n_ops = _plan(M, 48000.0, 1 / 13)
    # chunk of pass p that records bin need(c): pass p + 1 can start chunk c once pass p has finished that chunk (+ 1 for the rebuild)
    need = np.array([min(M, int(hi[min(M, int(kstart[c + 1]))])) for c in range(nchunks)])
    dep_chunk = endq[need] // CHUNK + 1
    lag = dep_chunk - np.arange(nchunks)
    assert lag.max() <= 0.035 * nchunks + 3, (int(lag.max()), nchunks)
