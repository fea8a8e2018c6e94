"""CPU: pin the oracle.  The plain-C restatement (oracle/irb_oracle.c) must agree with the reference's own
object code (oracle/_ref, when present here) and with the committed golden vectors generated from it
(tests/golden/make_golden.py); plus the known-answer identities SURVEY.md section 4 lists (KA1-KA9)."""
import os
import sys

import numpy as np
import pytest

from conftest import parity
from irbaboon_b200 import synth

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))


# ---- golden vectors (travel to the GPU box) ------------------------------------------------------------
@pytest.mark.parametrize("case", sorted(make_golden.PERIODIC_CASES))
def test_oracle_convolve_periodic_matches_golden(orc, case):
    x, h, B = make_golden.periodic_inputs(case)
    got = orc.convolve_periodic(x, h, B)
    want = GOLD["periodic/" + case]
    assert got.shape == want.shape
    assert np.array_equal(got, want)          # same algorithm, same FFT restatement, no FMA: bit-exact


def test_oracle_config1_matches_golden(orc):
    x, h, B = make_golden.config1_inputs()
    y = orc.convolve_periodic(x, h, B)[0]
    assert y.shape == (527999,)
    assert np.array_equal(y[:2048], GOLD["config1/head"])
    assert np.array_equal(y[-2048:], GOLD["config1/tail"])
    assert np.array_equal(y[::257], GOLD["config1/strided"])
    s = GOLD["config1/sums"]
    assert y.astype(np.float64).sum() == s[0] and (y.astype(np.float64) ** 2).sum() == s[1]
    assert not y[527872:].any() and y[527871] != 0          # D6: 127 unflushed samples


def test_oracle_other_functions_match_golden(orc):
    xs = synth.white_noise(1005, 0, 3000)
    hs = synth.decaying_ir(2005, 1000)
    conv = orc.convolve_nonperiodic(xs, hs)
    assert np.array_equal(conv, GOLD["nonperiodic/mono"])
    num, den = conv[0, :4096], np.pad(xs, (0, 1096))
    assert np.array_equal(orc.deconvolve(num, den, 48000.0, False), GOLD["deconvolve/plain"])
    e, l2 = parity(orc.deconvolve(num, den, 48000.0, True), GOLD["deconvolve/smoothed"])
    assert e <= 1e-6 and l2 <= 1e-5            # libm log/exp/atan2 may differ in the last ulp between builds
    assert np.array_equal(orc.deconvolve(num, den, 48000.0, False, False, True), GOLD["deconvolve/nophase"])
    assert np.array_equal(orc.fft_transform(xs[:1000]), GOLD["fft_transform"])
    assert np.array_equal(orc.shifteroo(np.arange(9, dtype=np.float32)), GOLD["shifteroo_odd"])
    assert np.allclose(orc.ess(0.25, 48000.0, 20.0, 20000.0)[::7], GOLD["ess/sweep"], rtol=0, atol=1e-12)
    assert np.allclose(orc.ess(0.25, 48000.0, 20.0, 20000.0, 0.0, True)[::7], GOLD["ess/inverse"], rtol=0, atol=1e-12)
    e, l2 = parity(orc.invert_filter(hs, 48000), GOLD["invert_filter"])
    assert e <= 1e-6 and l2 <= 1e-5


# ---- live comparison with the reference's object code (this container only) ---------------------------
@pytest.mark.parametrize("Lx,Lh,B,chx,chh", [(1, 1, 16, 1, 1), (100, 7, 16, 1, 1), (4096, 1000, 64, 2, 1), (2000, 2000, 256, 1, 2),
                                               (3333, 777, 100, 2, 2), (512, 512, 512, 1, 1), (5000, 100, 1024, 1, 1)])
def test_oracle_equals_reference_convolve_periodic(orc, ref, Lx, Lh, B, chx, chh):
    x = np.stack([synth.white_noise(1001, c, Lx) for c in range(chx)])
    h = np.stack([synth.decaying_ir(2000 + c, Lh, c) for c in range(chh)])
    assert np.array_equal(orc.convolve_periodic(x, h, B), ref.convolve_periodic(x, h, B))


def test_oracle_equals_reference_layout_rejection(orc, ref):
    x = np.zeros((3, 100), np.float32) + 1
    h = np.ones((1, 10), np.float32)
    a, b = orc.convolve_periodic(x, h, 16), ref.convolve_periodic(x, h, 16)
    assert a.shape == b.shape == (3, 109) and not a.any() and not b.any()


def test_oracle_equals_reference_nonperiodic_and_deconvolve(orc, ref):
    x = synth.white_noise(1001, 0, 1500)
    h = synth.decaying_ir(2000, 600)
    assert np.array_equal(orc.convolve_nonperiodic(x, h), ref.convolve_nonperiodic(x, h))
    y = ref.convolve_nonperiodic(x, h)[0]
    for smoothing in (False, True):
        a = orc.deconvolve(y[:2048], np.pad(x, (0, 548)), 48000.0, smoothing)
        b = ref.deconvolve(y[:2048], np.pad(x, (0, 548)), 48000.0, smoothing)
        e, l2 = parity(a, b)
        assert e <= 1e-6 and l2 <= 1e-5


# ---- known-answer identities -------------------------------------------------------
def test_ka1_pulse_ir_is_a_delay(orc):
    x = synth.white_noise(1001, 0, 3000)
    h = np.zeros(2048, np.float32)
    h[100] = 1.0
    y = orc.convolve_periodic(x, h, 256)[0]
    assert np.abs(y[100:3100] - x).max() <= 1e-6 and np.abs(y[:100]).max() <= 1e-6


def test_ka2_ka3_periodic_equals_nonperiodic_equals_direct(orc):
    x = synth.white_noise(1001, 0, 4096)
    h = synth.decaying_ir(2000, 1000)
    a = orc.convolve_periodic(x, h, 64)[0]
    b = orc.convolve_nonperiodic(x, h)[0]
    d = np.convolve(x.astype(np.float64), h.astype(np.float64))
    n = orc.periodic_iterations(4096, 1000, 64) * 64
    n = min(n, len(d))
    assert np.abs(a[:n] - b[:n]).max() <= 1e-5
    assert np.abs(a[:n] - d[:n]).max() <= 1e-5


def test_ka4_result_independent_of_block_size(orc):
    x = synth.white_noise(1001, 0, 5000)
    h = synth.decaying_ir(2000, 1300)
    base = orc.convolve_periodic(x, h, 64)[0]
    for B in (128, 256, 512, 1024):
        y = orc.convolve_periodic(x, h, B)[0]
        n = min(orc.periodic_iterations(5000, 1300, 64) * 64, orc.periodic_iterations(5000, 1300, B) * B, len(y))
        assert np.abs(y[:n] - base[:n]).max() <= 1e-5


@pytest.mark.parametrize("Lx,Lh,B", [(480000, 48000, 512), (1000, 100, 64), (64, 64, 64), (63, 65, 64)])
def test_ka5_unflushed_tail_length(orc, Lx, Lh, B):
    iters = orc.periodic_iterations(Lx, Lh, B)
    P = int(np.ceil(np.float32(Lh) / np.float32(B)))
    assert iters == Lx // B + P
    if (Lx, Lh, B) == (480000, 48000, 512):
        assert iters == 1031 and Lx + Lh - 1 - iters * B == 127


def test_ka6_deconvolve_recovers_the_ir(orc):
    x = synth.white_noise(1001, 0, 2048)
    h = synth.decaying_ir(2000, 500)
    y = orc.convolve_nonperiodic(x, h)[0]            # 2547 samples <= 4096: no circular wrap
    got = orc.deconvolve(np.pad(y, (0, 4096 - len(y))), np.pad(x, (0, 2048)), 48000.0, False)[0]
    assert np.abs(got[:500] - h).max() <= 1e-4 and np.abs(got[500:]).max() <= 1e-4


def test_ka7_shifteroo_involution(orc):
    a = np.arange(10, dtype=np.float32)
    assert np.array_equal(orc.shifteroo(orc.shifteroo(a)), a[None, :])


def test_ka8_sweep_properties(orc):
    s = orc.ess(0.5, 48000.0, 20.0, 20000.0)
    assert s[0] == 0.0 and len(s) == 24000
    assert np.allclose(s, synth.exp_sine_sweep(0.5, 48000.0, 20.0, 20000.0), atol=1e-12)


def test_ka9_stereo_semantics(orc):
    x = np.stack([synth.white_noise(1001, c, 1000) for c in range(2)])
    h = np.stack([synth.decaying_ir(2000 + c, 300, c) for c in range(2)])
    ss = orc.convolve_periodic(x, h, 64)
    assert np.array_equal(ss[0], orc.convolve_periodic(x[0], h[0], 64)[0])       # channel-wise, not 2x2
    assert np.array_equal(ss[1], orc.convolve_periodic(x[1], h[1], 64)[0])
    ms = orc.convolve_periodic(x[0], h, 64)
    fold = ((h[0] + h[1]) / np.float32(2.0)).astype(np.float32)
    assert np.array_equal(ms, orc.convolve_periodic(x[0], fold, 64))


def test_oracle_fft_against_float64(orc):
    for n in (32, 512, 1024, 4096):
        x = synth.white_noise(7, n, n)
        buf = np.zeros(2 * n, np.float32)
        buf[:n] = x
        got = orc.real_forward(buf, n)
        want = np.fft.fft(x.astype(np.float64))
        g = got[0::2] + 1j * got[1::2]
        assert np.abs(g - want).max() / np.abs(want).max() <= 2e-6
        back = orc.real_inverse(got, n)[:n]
        assert np.abs(back - x).max() <= 2e-6


def test_synth_generators_match_c(orc):
    assert np.array_equal(orc.white_noise(1001, 3, 1000), synth.white_noise(1001, 3, 1000))
    h = synth.decaying_ir(2000, 4800)
    assert abs(float(np.sqrt((h.astype(np.float64) ** 2).sum())) - 1.0) < 1e-6


# ---- the processBlock restatement (orc_rt_*) ---------------------------------------------------------------------
@pytest.mark.parametrize("B,H,delay", [(256, 256, 256), (256, 64, 448), (256, 100, 300), (128, 128, 128), (64, 32, 96)])
def test_rt_restatement_is_the_offline_convolution_delayed(orc, B, H, delay):
    """Host blocks <= processBlockSize: the streaming engine of PluginProcessor.cpp:403-562 is convolvePeriodic delayed
    by (outputArraySize - 1) host buffers (>= the latency the plug-in reports, max(B, H))."""
    n = H * 40
    x = np.stack([synth.white_noise(1, c, n) for c in range(2)])
    h = synth.decaying_ir(2000, 3 * B + 5)
    e = orc.rt_engine(B, H, 2, h)
    y = np.concatenate([e.process(x[:, i:i + H]) for i in range(0, n, H)], axis=1)
    e.close()
    want = orc.convolve_periodic(x, np.stack([h, h]), B)[:, :n - delay]
    assert not y[:, :delay].any()
    assert np.abs(y[:, delay:] - want).max() <= 1e-6 * max(1.0, np.abs(want).max())


def test_rt_restatement_large_host_blocks_read_future_blocks(orc):
    """hostBlock > processBlockSize with the reference's ring of max(P, hostBlock/B) spectra: all blocks of a callback
    are transformed before any is convolved (PluginProcessor.cpp:421-445 then :452-518), so the oldest partitions of
    the callback's first blocks multiply spectra of its LATER blocks -- the output is not a delayed convolution."""
    B, H, P = 256, 512, 8
    n = H * 30
    x = synth.white_noise(1, 0, n)[None, :]
    h = synth.decaying_ir(2000, P * B)
    e = orc.rt_engine(B, H, 1, h)
    y = np.concatenate([e.process(x[:, i:i + H]) for i in range(0, n, H)], axis=1)
    e.close()
    want = orc.convolve_periodic(x, h, B)[:, :n - H]
    assert np.abs(y[:, H:] - want).max() > 1e-3
    # with an IR one partition shorter every partition still meets the right slot
    h2 = h[:(P - 1) * B]
    e = orc.rt_engine(B, H, 1, np.concatenate([h2, np.zeros(B, np.float32)]))
    y = np.concatenate([e.process(x[:, i:i + H]) for i in range(0, n, H)], axis=1)
    e.close()
    # partition P-1 is all zeros, so its acausal product vanishes
    assert np.abs(y[:, H:] - orc.convolve_periodic(x, h2, B)[:, :n - H]).max() <= 1e-6 * 10


def test_rt_restatement_round_robin_ir_switch(orc):
    """After orc_rt_set_ir the partitions are replaced one per processed block: P blocks later the output is the new
    IR's convolution of the recent input; before that it is a partition-wise mix of both."""
    B, P = 128, 6
    n = B * 40
    x = synth.white_noise(1, 0, n)[None, :]
    h0, h1 = synth.decaying_ir(2000, P * B), synth.decaying_ir(2001, P * B, 1)
    e = orc.rt_engine(B, B, 1, h0)
    out = []
    for k in range(40):
        if k == 15:
            e.set_ir(h1)
        out.append(e.process(x[:, k * B:(k + 1) * B]))
    e.close()
    y = np.concatenate(out, axis=1)[:, B:]
    y0 = orc.convolve_periodic(x, h0, B)[:, :n - B]
    y1 = orc.convolve_periodic(x, h1, B)[:, :n - B]
    assert np.abs(y[:, :15 * B] - y0[:, :15 * B]).max() <= 1e-5
    assert np.abs(y[:, 15 * B:16 * B] - y0[:, 15 * B:16 * B]).max() > 1e-3          # the first partition is already new
    # the write position was 15 % 6 = 3 at the switch: all six partitions are new after 6 more blocks, and the FDL
    # only holds blocks convolved with ... the new spectra from then on
    assert np.abs(y[:, 21 * B:] - y1[:, 21 * B:]).max() <= 1e-5
