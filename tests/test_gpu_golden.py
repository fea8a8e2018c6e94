"""GPU: the CUDA path against the COMMITTED golden vectors (tests/golden/reference_vectors.npz = outputs of the reference's
own object code, generated in the build container by tests/golden/make_golden.py), so parity on the GPU box does not
depend on any CPU library being present there; plus size-independent properties at BASELINE.json's full sizes for the
configurations whose full-size oracle run would take minutes (c3, c4)."""
import os
import sys

import numpy as np
import pytest

from conftest import TOL, parity
from irbaboon_b200 import synth

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))


def _check(got, want, tol=TOL, l2tol=None):
    assert got.shape == want.shape, (got.shape, want.shape)
    e, l2 = parity(got, want)
    assert e <= tol and l2 <= (l2tol or tol), (e, l2)


@pytest.mark.parametrize("case", sorted(make_golden.PERIODIC_CASES))
def test_convolve_periodic_matches_golden(eng, case):
    x, h, B = make_golden.periodic_inputs(case)
    _check(eng.convolve_periodic(x, h, B), GOLD["periodic/" + case])


@pytest.mark.parametrize("case", sorted(make_golden.PERIODIC_CASES))
def test_streaming_engine_matches_golden(eng, case):
    """The same cases block by block through the streaming engine (mono IR cases; every audio channel uses IR 0)."""
    x, h, B = make_golden.periodic_inputs(case)
    if h.shape[0] != 1:
        pytest.skip("the engine binds one mono IR per channel; stereo-IR layouts are covered by the offline function")
    P = -(-h.shape[1] // B)
    with eng.Engine(B, P, x.shape[0], 1) as e:
        e.set_ir(0, h[0])
        y = e.process_stream(x)
    _check(y, GOLD["periodic/" + case][:, :x.shape[1]])


def test_config1_matches_golden(eng):
    x, h, B = make_golden.config1_inputs()
    y = eng.convolve_periodic(x, h, B)[0]
    assert y.shape == (527999,)
    fs = max(1.0, float(GOLD["config1/sums"][2]))
    for got, key in ((y[:2048], "config1/head"), (y[-2048:], "config1/tail"), (y[::257], "config1/strided")):
        assert np.abs(got - GOLD[key]).max() / fs <= TOL
    s = GOLD["config1/sums"]
    assert abs(y.astype(np.float64).sum() - s[0]) <= 1e-5 * np.sqrt(s[1]) and abs((y.astype(np.float64) ** 2).sum() / s[1] - 1.0) <= 1e-5
    assert not y[527872:].any()


def test_other_functions_match_golden(eng):
    xs = synth.white_noise(1005, 0, 3000)
    hs = synth.decaying_ir(2005, 1000)
    conv = eng.convolve_nonperiodic(xs, hs)
    _check(conv, GOLD["nonperiodic/mono"])
    num, den = GOLD["nonperiodic/mono"][0, :4096], np.pad(xs, (0, 1096))
    _check(eng.deconvolve(num, den, 48000.0, False), GOLD["deconvolve/plain"])
    _check(eng.deconvolve(num, den, 48000.0, True), GOLD["deconvolve/smoothed"])
    _check(eng.deconvolve(num, den, 48000.0, False, False, True), GOLD["deconvolve/nophase"])
    _check(eng.invert_filter(hs, 48000), GOLD["invert_filter"])
    got = eng.fft_transform(xs[:1000])
    assert np.abs(got - GOLD["fft_transform"]).max() <= 1e-5 * np.abs(GOLD["fft_transform"]).max()
    assert np.abs(eng.ess(0.25, 48000.0, 20.0, 20000.0)[::7] - GOLD["ess/sweep"]).max() <= 1e-9
    assert np.abs(eng.ess(0.25, 48000.0, 20.0, 20000.0, 0.0, True)[::7] - GOLD["ess/inverse"]).max() <= 1e-9


# ---- full-size properties ------------------------------------------------------------------------------------------
def test_config3_full_size_impulse_response_and_linearity(eng, orc):
    """BASELINE configs[2] at full size: 1024 streams sharing one 2 s IR (96 000 taps, 188 partitions), B=512.
    (a) a unit impulse in stream s at sample 17*s%512 returns the IR itself, delayed -- for every stream;
    (b) superposition: y(x1 + x2) == y(x1) + y(x2) within float32 rounding on a noise input;
    (c) stream 0 against the oracle over 20 blocks."""
    B, Lh, S, nb = 512, 96000, 1024, 20
    h = synth.decaying_ir(2000, Lh)
    n = nb * B
    imp = np.zeros((S, n), np.float32)
    d = (17 * np.arange(S)) % 512
    imp[np.arange(S), d] = 1.0
    x1 = np.stack([synth.white_noise(1003, s % 7, n) for s in range(S)])
    with eng.Engine(B, 188, S, 1) as e:
        e.set_ir(0, h)
        assert e.mac_plan() == (False, 1, 1)
        yi = e.process_stream(imp)
        e.reset()
        y1 = e.process_stream(x1)
        e.reset()
        y2 = e.process_stream(imp + x1)
    for s in (0, 1, 511, 1023) + tuple(range(5, S, 97)):
        assert np.abs(yi[s, d[s]:] - h[:n - d[s]]).max() <= 2e-6 and (d[s] == 0 or np.abs(yi[s, :d[s]]).max() <= 2e-6)
    assert np.abs(y2 - (y1 + yi)).max() <= 2e-5
    _check(y1[:1], orc.convolve_periodic(x1[0], h, B)[:, :n])


def test_config4_full_size_per_stream_irs_return_their_own_ir(eng):
    """BASELINE configs[3] shape at full IR length and a tenth of the stream count (819 of 8192 streams, 6.3 GB): every
    stream owns a 480 000-tap IR (469 partitions, B=1024).  A unit impulse returns each stream's OWN IR: checked for every
    stream on the blocks processed, which also proves no stream reads a neighbour's spectra."""
    B, Lh, S, nb = 1024, 480000, 819, 6
    base = [synth.decaying_ir(2100 + j, Lh, j) for j in range(8)]
    rng = np.random.default_rng(4)
    gain = (0.5 + rng.random(S)).astype(np.float32)                        # distinct per stream
    with eng.Engine(B, 469, S, S) as e:
        for s in range(S):
            e.set_ir(s, base[s % 8] * gain[s])
            e.bind(s, s + 1, s)
        assert e.mac_plan()[0]                                              # mixed IRs per tile: the slot kernel
        x = np.zeros((nb, S, B), np.float32)
        x[0, :, 0] = 1.0
        y = e.process(x)
    y = np.ascontiguousarray(y.transpose(1, 0, 2)).reshape(S, nb * B)
    for s in range(S):
        want = (base[s % 8] * gain[s])[:nb * B]
        assert np.abs(y[s] - want).max() <= 2e-6 * max(1.0, np.abs(want).max()), s
