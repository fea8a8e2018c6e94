"""GPU: randomised parity (hypothesis) of the C-ABI entry points against the CPU oracle -- arbitrary lengths, block sizes,
channel layouts, IR bindings and launch plans, the ragged and degenerate cases included.  Tolerance as everywhere:
max-abs <= 1e-5 of full scale and relative L2 <= 1e-5 (FP32)."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from conftest import TOL, parity
from irbaboon_b200 import synth

pytestmark = pytest.mark.gpu
COMMON = dict(deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow], derandomize=True)


def _check(got, want, tol=TOL):
    assert got.shape == want.shape, (got.shape, want.shape)
    e, l2 = parity(got, want)
    assert e <= tol and l2 <= tol, (e, l2)


block_sizes = st.one_of(st.integers(1, 300), st.sampled_from([16, 32, 64, 128, 256, 512, 1024, 2048, 3000, 5000]))


@settings(max_examples=200, **COMMON)
@given(Lx=st.integers(1, 3000), Lh=st.integers(1, 2000), B=block_sizes, chx=st.integers(1, 2), chh=st.integers(1, 2), seed=st.integers(0, 1000))
def test_convolve_periodic_any_shape(eng, orc, Lx, Lh, B, chx, chh, seed):
    x = np.stack([synth.white_noise(seed, c, Lx) for c in range(chx)])
    h = np.stack([synth.decaying_ir(2000 + seed, Lh, c) for c in range(chh)])
    want = orc.convolve_periodic(x, h, B)
    got = eng.convolve_periodic(x, h, B)
    _check(got, want)
    written = min(Lx + Lh - 1, orc.periodic_iterations(Lx, Lh, B) * B)
    assert not got[:, written:].any()                      # the unflushed tail stays zero exactly as in the reference


@settings(max_examples=120, **COMMON)
@given(B=st.one_of(st.integers(1, 200), st.sampled_from([256, 512, 1024])), C=st.integers(1, 24), P=st.integers(1, 9), n_irs=st.integers(1, 3),
       nb=st.integers(1, 12), plan=st.sampled_from([(0, 0), (1, 1), (2, 1), (1, 2), (4, 4), (8, 16)]), fused=st.booleans(), seed=st.integers(0, 1000))
def test_streaming_engine_any_shape(eng, orc, B, C, P, n_irs, nb, plan, fused, seed):
    rng = np.random.default_rng(seed)
    n = nb * B
    x = np.stack([synth.white_noise(seed, c, n) for c in range(C)])
    lens = [int(rng.integers(1, P * B + 1)) for _ in range(n_irs)]
    irs = [synth.decaying_ir(2100 + seed + j, lens[j], j) for j in range(n_irs)]
    which = rng.integers(0, n_irs, C)
    with eng.Engine(B, P, C, n_irs) as e:
        for j in range(n_irs):
            e.set_ir(j, irs[j])
        for c in range(C):
            e.bind(c, c + 1, int(which[c]))
        e.set_mac_split(*plan)
        e.set_fused_step(fused)
        if nb % 2:
            y = e.process_stream(x)                         # one multi-block call
        else:                                               # block by block (plain, then graph replays)
            blocks = np.ascontiguousarray(x.reshape(C, nb, B).transpose(1, 0, 2))
            y = np.ascontiguousarray(np.stack([e.process(blocks[k]) for k in range(nb)]).transpose(1, 0, 2)).reshape(C, n)
    for c in range(C):
        _check(y[c:c + 1], orc.convolve_periodic(x[c], irs[which[c]], B)[:, :n])


@settings(max_examples=80, **COMMON)
@given(Lx=st.integers(1, 6000), Lh=st.integers(1, 6000), chx=st.integers(1, 2), chh=st.integers(1, 2), seed=st.integers(0, 1000))
def test_convolve_nonperiodic_any_shape(eng, orc, Lx, Lh, chx, chh, seed):
    x = np.stack([synth.white_noise(seed, c, Lx) for c in range(chx)])
    h = np.stack([synth.decaying_ir(2000 + seed, Lh, c) for c in range(chh)])
    _check(eng.convolve_nonperiodic(x, h), orc.convolve_nonperiodic(x, h))


@settings(max_examples=60, **COMMON)
@given(Ln=st.integers(16, 5000), Ld=st.integers(16, 5000), smoothing=st.booleans(), phase=st.booleans(), ampl=st.booleans(), seed=st.integers(0, 1000))
def test_deconvolve_any_shape(eng, orc, Ln, Ld, smoothing, phase, ampl, seed):
    den = synth.white_noise(seed, 1, Ld) + np.float32(0.05)         # broadband: no vanishing bins
    num = synth.white_noise(seed, 0, Ln)
    want = orc.deconvolve(num, den, 48000.0, smoothing, phase, ampl)
    got = eng.deconvolve(num, den, 48000.0, smoothing, phase, ampl)
    assert got.shape == want.shape
    e, l2 = parity(got, want)
    # the spectral division amplifies float32 rounding where |den| is small; smoothing adds libm last-ulp differences
    assert e <= (2e-4 if smoothing else 5e-5) and l2 <= (5e-4 if smoothing else 1e-4), (e, l2)
