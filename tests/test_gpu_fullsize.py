"""GPU parity at BASELINE.json's FULL sizes, sampled the way SURVEY 8(d) asks ("parity on a random 1 % of streams"):

* configs[2]: 1024 streams sharing one 2 s IR, B = 512 -- a seeded random 1 % of the streams against the reference's own
  convolvePeriodic over 2 P + 4 blocks of white noise (steady state: the FDL ring has wrapped);
* configs[3]: ALL 8192 streams with per-stream 10 s IRs (480 000 taps, 469 partitions), B = 1024 -- a unit impulse returns each
  stream's own IR: every stream over the first blocks, a random 1 % of the streams over the whole IR and past its end;
* configs[4]: deconvolve(smoothing = true, the plug-in default) at N = 2^20 against the reference, and a batch of captures that
  spans several sub-batches of the pipeline, first / middle / last capture against the reference.
"""
import numpy as np
import pytest

from conftest import TOL, conditioned_bound, parity
from irbaboon_b200 import synth

pytestmark = pytest.mark.gpu


def test_config3_random_one_percent_of_streams_in_steady_state(eng, orc, request):
    B, Lh, S = 512, 96000, 1024
    P = 188
    nb = 2 * P + 4
    h = synth.decaying_ir(2000, Lh)
    chans = sorted(set(int(c) for c in np.random.default_rng(2024).choice(S, 11, replace=False)) | {0, S - 1})
    xs = {c: synth.white_noise(1003, 100 + c, nb * B) for c in chans}
    ys = {c: np.zeros(nb * B, np.float32) for c in chans}
    CH = 16                                                                # blocks per host call (33.5 MB per direction)
    with eng.Engine(B, P, S, 1) as e:
        e.set_ir(0, h)
        assert e.mac_plan() == (False, 1, 1)
        rng = np.random.default_rng(7)
        for b0 in range(0, nb, CH):
            k = min(CH, nb - b0)
            x = (rng.random((k, S, B), dtype=np.float32) * 2 - 1).astype(np.float32)      # every other stream: its own noise
            for c in chans:
                x[:, c, :] = xs[c][b0 * B:(b0 + k) * B].reshape(k, B)
            y = e.process(x)
            for c in chans:
                ys[c][b0 * B:(b0 + k) * B] = y[:, c, :].reshape(-1)
    try:
        ref = request.getfixturevalue("ref")                               # oracle/_ref: the reference's own object code
    except pytest.skip.Exception:
        ref = orc
    worst = 0.0
    for c in chans:
        want = ref.convolve_periodic(xs[c], h, B)[0, :nb * B]
        e_, l2 = parity(ys[c], want)
        worst = max(worst, e_, l2)
        assert e_ <= TOL and l2 <= TOL, (c, e_, l2)
        tail = slice((nb - 8) * B, nb * B)                                 # the last blocks alone (FDL wrapped twice)
        e_, l2 = parity(ys[c][tail], want[tail])
        assert e_ <= TOL and l2 <= TOL, (c, e_, l2)
    print("config3 steady state, %d of %d streams, %d blocks: worst error %.3g" % (len(chans), S, nb, worst))


def test_config4_all_8192_streams_return_their_own_ir(eng):
    B, Lh, S, P = 1024, 480000, 8192, 469
    base = [synth.decaying_ir(2100 + j, Lh, j) for j in range(8)]
    rng = np.random.default_rng(4)
    gain = (0.5 + rng.random(S)).astype(np.float32)                        # distinct per stream
    delay = (rng.integers(0, B, S)).astype(np.int64)                       # the impulse sits somewhere in block 0
    sample = sorted(set(int(c) for c in rng.choice(S, 82, replace=False)) | {0, 1, S - 1})
    nb_all, nb = 12, P + 3
    CH = 8
    kept = {c: np.zeros(nb * B, np.float32) for c in sample}
    with eng.Engine(B, P, S, S) as e:
        for s in range(S):
            e.set_ir(s, base[s % 8] * gain[s])
            e.bind(s, s + 1, s)
        assert e.mac_plan() == (True, 1, 1)                                # every row stages its own IR, one launch per block step
        l0 = e.launches
        x = eng.pinned_empty((CH, S, B))
        y = eng.pinned_empty((CH, S, B))
        first = None
        for b0 in range(0, nb, CH):
            k = min(CH, nb - b0)
            x[:] = 0.0
            if b0 == 0:
                x[0, np.arange(S), delay] = 1.0
            e.process(x[:k], y[:k])
            if b0 == 0:
                first = np.ascontiguousarray(np.array(y[:CH]).transpose(1, 0, 2)).reshape(S, CH * B)
            for c in sample:
                kept[c][b0 * B:(b0 + k) * B] = y[:k, c, :].reshape(-1)
        assert e.launches - l0 == nb                                       # fused: exactly one kernel per block step
        eng.pinned_free(x); eng.pinned_free(y)
    # every stream, first blocks: its OWN IR, delayed (also proves no stream reads a neighbour's spectra)
    n0 = min(nb_all, CH) * B
    for s in range(S):
        d = int(delay[s])
        want = np.zeros(n0, np.float32)
        want[d:] = (base[s % 8] * gain[s])[:n0 - d]
        assert np.abs(first[s, :n0] - want).max() <= 2e-6 * max(1.0, float(np.abs(want).max())), s
    # 1 % of the streams: the whole IR and the silence after it
    for c in sample:
        d = int(delay[c])
        want = np.zeros(nb * B, np.float32)
        hc = base[c % 8] * gain[c]
        want[d:d + Lh] = hc[:nb * B - d]
        e_, l2 = parity(kept[c], want)
        assert e_ <= 2e-6 and l2 <= TOL, (c, e_, l2)


def _capture20(ref, seed):
    n = 1 << 20
    sweep = synth.exp_sine_sweep(n / 48000.0, 48000.0, 20.0, 20000.0).astype(np.float32)[:n]
    h = synth.decaying_ir(3000 + seed, 48000)
    cap = ref.convolve_nonperiodic(sweep, h)[0, :n].copy()
    cap += synth.white_noise(4000 + seed, 0, n) * np.float32(1e-3)
    return sweep, cap


def test_deconvolve_smoothed_at_two_to_the_twenty(eng, orc, request):
    """The plug-in default (smoothing = true) at the full capture length of configs[4]: the float32 running sum of
    averagingFilter drifts with N, so the device must execute the reference's exact addition sequence."""
    try:
        ref = request.getfixturevalue("ref")
    except pytest.skip.Exception:
        ref = orc
    sweep, cap = _capture20(ref, 0)
    want = ref.deconvolve(cap, sweep, 48000.0, True)
    got = eng.deconvolve(cap, sweep, 48000.0, True)
    e_, l2 = parity(got, want)
    bs = conditioned_bound(lambda c, s: ref.deconvolve(c, s, 48000.0, True), [cap, sweep], want)
    print("deconvolve(smoothing=true), N = 2^20: max-abs/FS %.3g, relative L2 %.3g (reference's own half-ulp response x 2: %.3g, %.3g)" % (e_, l2, bs[0], bs[1]))
    assert e_ <= TOL and l2 <= bs[1], (e_, l2, bs)
    wantp = ref.deconvolve(cap, sweep, 48000.0, False)
    gotp = eng.deconvolve(cap, sweep, 48000.0, False)
    e_, l2 = parity(gotp, wantp)
    # the sweep's spectrum falls to 2e-4 of its peak: the division amplifies float32 rounding, the reference's included
    be, bl2 = conditioned_bound(lambda c, s: ref.deconvolve(c, s, 48000.0, False), [cap, sweep], wantp)
    print("deconvolve(smoothing=false), N = 2^20: max-abs/FS %.3g, relative L2 %.3g (reference's own half-ulp response x 2: %.3g, %.3g)" % (e_, l2, be, bl2))
    assert e_ <= TOL and l2 <= bl2, (e_, l2, be, bl2)


def test_deconvolve_batch_over_several_sub_batches(eng, orc, request):
    try:
        ref = request.getfixturevalue("ref")
    except pytest.skip.Exception:
        ref = orc
    nb = 30                                                                # sub-batches of 12, 12 and 6 captures
    sweep, cap0 = _capture20(ref, 1)
    caps = eng.pinned_empty((nb, 1 << 20))
    res = eng.pinned_empty((nb, 1 << 20))
    for j in range(nb):
        caps[j] = cap0 * np.float32(0.5 + j / nb) + synth.white_noise(5000 + j, 0, 1 << 20) * np.float32(1e-3)
    eng.deconvolve_batch(caps, sweep, 48000.0, False, out=res)
    for j in (0, 11, 12, nb // 2, nb - 1):
        cj = np.array(caps[j])
        want = ref.deconvolve(cj, sweep, 48000.0, False)[0]
        e_, l2 = parity(res[j], want)
        be, bl2 = conditioned_bound(lambda c, s: ref.deconvolve(c, s, 48000.0, False), [cj, sweep], want[None, :], seed=j)
        assert e_ <= TOL and l2 <= bl2, (j, e_, l2, bl2)
        assert np.array_equal(res[j], eng.deconvolve(np.array(caps[j]), sweep, 48000.0, False)[0])
    eng.pinned_free(caps); eng.pinned_free(res)
