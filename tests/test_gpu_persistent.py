"""GPU: the persistent block-step kernel k_mac_p (irb_mac_p.cuh) -- one launch per block step, units fetched from a device
counter, the last partial wave on tiles of fewer rows, shared-IR and per-stream-IR (PERROW) forms.

Checked against (a) the CPU oracle (fp/convolution.cpp:14-242 restated), (b) the other forms of the same step, bit for bit:
the one-CTA-per-tile kernel k_mac_tma, and the two-launch form (k_fwd, then k_mac / k_mac_slots)."""
import numpy as np
import pytest

from conftest import TOL, parity
from irbaboon_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _default_tuning(eng):
    yield
    for k, v in (("mac_persistent", 1), ("unit_narrowing", 0), ("release_fence", 1), ("persistent_ctas", 0), ("mac_tma", 1)):
        eng.set_tuning(k, v)


def _blocks(C, nb, B, seed=1002):
    rng = np.random.default_rng(seed)
    return (rng.random((nb, C, B), dtype=np.float32) * 2 - 1).astype(np.float32)


def _run(eng, B, ring, C, x, irs, bind=None, fused=True, tune=()):
    for k, v in tune:
        eng.set_tuning(k, v)
    try:
        with eng.Engine(B, ring, C, len(irs)) as e:
            for i, h in enumerate(irs):
                e.set_ir(i, h)
            if bind is not None:
                for c in range(C):
                    e.bind(c, c + 1, bind(c))
            e.set_fused_step(fused)
            l0 = e.launches
            y = e.process(x)
            return y, e.launches - l0, e.mac_plan()
    finally:
        for k, _ in tune:
            eng.set_tuning(k, {"mac_persistent": 1, "unit_narrowing": 0, "release_fence": 1, "persistent_ctas": 0, "mac_tma": 1}[k])


# stream counts chosen against 2 x 148 resident CTAs so that the last wave runs on half / quarter / single-row units,
# a clipped last unit, fewer units than CTAs, and every FFT size of the kernel
@pytest.mark.parametrize("B,C,P,nb", [
    (512, 1879, 5, 9),            # 296 tiles of 4 rows, then 296 units of 2 rows, then 103 units of 1 row
    (512, 1186, 3, 7),            # one full wave + two single-row units
    (512, 4 * 296 * 2 + 4 * 100 + 3, 4, 6),   # two full waves, a partial wave of whole tiles, last tile clipped
    (256, 8 * 296 + 4 * 296 + 2 * 100 + 1, 4, 6),   # ROWS = 8: levels 8, 4, 2/1
    (1024, 2 * 296 + 150, 6, 8),  # ROWS = 2
    (2048, 300, 3, 5),            # ROWS = 1 (no narrowing possible)
    (300, 1000, 4, 6),            # block size that is not a power of two (M = 512)
    (512, 640, 9, 12),            # fewer rows than one wave
])
def test_persistent_step_equals_every_other_form_and_the_oracle(eng, orc, B, C, P, nb):
    h = synth.decaying_ir(2001, P * B - 7)
    x = _blocks(C, nb, B)
    y, launches, plan = _run(eng, B, P, C, x, [h])
    assert plan == (False, 1, 1) and launches == nb                       # one launch per block step
    y_plain, _, _ = _run(eng, B, P, C, x, [h], tune=[("unit_narrowing", 1)])             # last partial wave on units of fewer rows
    y_tile, _, _ = _run(eng, B, P, C, x, [h], tune=[("mac_persistent", 0)])             # k_mac_tma, one CTA per tile
    y_two, l2, _ = _run(eng, B, P, C, x, [h], fused=False)                               # k_fwd + k_mac
    assert l2 == 2 * nb
    assert np.array_equal(y, y_plain) and np.array_equal(y, y_tile) and np.array_equal(y, y_two)
    for c in (0, 1, C // 2, C - 2, C - 1):
        want = orc.convolve_periodic(np.ascontiguousarray(x[:, c, :]).reshape(-1), h, B)[0, :nb * B]
        e, l2n = parity(y[:, c, :].reshape(-1), want)
        assert e <= TOL and l2n <= TOL, (c, e, l2n)


def test_one_cta_per_sm_and_no_release_fence_give_the_same_bits(eng):
    B, C, P, nb = 512, 2500, 6, 8
    h = synth.decaying_ir(2002, P * B)
    x = _blocks(C, nb, B, 5)
    y, _, _ = _run(eng, B, P, C, x, [h])
    y1, _, _ = _run(eng, B, P, C, x, [h], tune=[("persistent_ctas", 1)])
    y2, _, _ = _run(eng, B, P, C, x, [h], tune=[("release_fence", 0)])
    assert np.array_equal(y, y1) and np.array_equal(y, y2)


@pytest.mark.parametrize("B,C,ring", [(1024, 2 * 296 + 151, 9), (512, 4 * 296 + 2 * 200 + 3, 7), (256, 1203, 6), (2048, 310, 4)])
def test_per_stream_irs_persistent_equals_slot_kernel_and_oracle(eng, orc, B, C, ring):
    """BASELINE configs[3] shape: every stream owns an IR (here of 8 different lengths, 1 .. ring partitions)."""
    irs = [synth.decaying_ir(2100 + j, max(1, (1 + j * (ring - 1) // 7) * B - 3 * j), j) for j in range(8)]
    nb = ring + 4
    x = _blocks(C, nb, B, 7)
    y, launches, plan = _run(eng, B, ring, C, x, irs, bind=lambda c: (c * 5 + c // 3) % 8)
    assert plan == (B < 2048, 1, 1) and launches == nb                    # per-row IR staging (a one-row tile cannot mix IRs), still ONE launch per block step
    y_two, l2, _ = _run(eng, B, ring, C, x, irs, bind=lambda c: (c * 5 + c // 3) % 8, fused=False)      # k_fwd + k_mac_slots
    assert l2 == 2 * nb and np.array_equal(y, y_two)
    y_old, l3, _ = _run(eng, B, ring, C, x, irs, bind=lambda c: (c * 5 + c // 3) % 8, tune=[("mac_persistent", 0)])
    assert l3 == (2 * nb if B < 2048 else nb) and np.array_equal(y, y_old)     # (one-row tiles never mix IRs: the tile kernel, fused)
    for c in (0, 1, 2, 3, C // 2, C - 1):
        hc = irs[(c * 5 + c // 3) % 8]
        want = orc.convolve_periodic(np.ascontiguousarray(x[:, c, :]).reshape(-1), hc, B)[0, :nb * B]
        e, l2n = parity(y[:, c, :].reshape(-1), want)
        assert e <= TOL and l2n <= TOL, (c, e, l2n)


def test_active_channels_freeze_the_rest(eng, orc):
    """irb_engine_set_active_channels: the inactive channels' FDL rings, heads and overlaps stay untouched."""
    B, C, P, A = 512, 1200, 5, 800
    h = synth.decaying_ir(2003, P * B - 1)
    x = _blocks(C, 9, B, 11)
    with eng.Engine(B, P, C, 1) as e:
        e.set_ir(0, h)
        y0 = e.process(x[0:3])
        with pytest.raises(eng.IrbError):
            e.set_active_channels(A + 1)                                   # not a whole number of tiles
        e.set_active_channels(A)
        y1 = e.process(np.ascontiguousarray(x[3:6, :A]))
        with pytest.raises(ValueError):
            e.process(x[3:6])                                              # I/O arrays are dense over the active channels
        e.set_active_channels(C)
        y2 = e.process(x[6:9])
    for c in (0, A - 1):                                                   # always active: nine consecutive blocks
        want = orc.convolve_periodic(np.ascontiguousarray(x[:, c, :]).reshape(-1), h, B)[0, :9 * B]
        got = np.concatenate([y0[:, c], y1[:, c], y2[:, c]]).reshape(-1)
        e_, l2 = parity(got, want)
        assert e_ <= TOL and l2 <= TOL
    for c in (A, C - 1):                                                   # frozen during blocks 3..5: their stream is blocks 0-2, 6-8
        xs = np.concatenate([x[0:3, c], x[6:9, c]]).reshape(-1)
        want = orc.convolve_periodic(xs, h, B)[0, :6 * B]
        got = np.concatenate([y0[:, c], y2[:, c]]).reshape(-1)
        e_, l2 = parity(got, want)
        assert e_ <= TOL and l2 <= TOL


def test_full_gpu_many_laps_stays_in_lockstep_with_the_two_launch_form(eng):
    """More units than resident CTAs, several laps of the FDL ring, the shared-memory ring lapped hundreds of times per CTA."""
    B, C, P, nb = 512, 6000, 23, 60
    h = synth.decaying_ir(2004, P * B)
    x = _blocks(C, nb, B, 13)
    y, _, _ = _run(eng, B, P, C, x, [h])
    y_two, _, _ = _run(eng, B, P, C, x, [h], fused=False)
    assert np.array_equal(y, y_two)


# ---- randomised shapes through the persistent kernel (enough channels that the launch plan never splits rows) ----------
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st


@settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(B=st.sampled_from([129, 200, 256, 300, 512, 777, 1024, 1500, 2048]), tiles=st.integers(160, 700), extra=st.integers(0, 7), P=st.integers(1, 11), nb=st.integers(1, 9),
       n_irs=st.integers(1, 4), mixed=st.booleans(), active_tiles=st.integers(0, 150), seed=st.integers(0, 1000))
def test_persistent_step_any_shape(eng, orc, B, tiles, extra, P, nb, n_irs, mixed, active_tiles, seed):
    """Random block sizes (not powers of two included), stream counts with a ragged last tile, IR lengths, shared IRs bound per
    tile or mixed inside tiles (PERROW), and a change of the active channel count half way: fused == two-launch bit for bit, and
    three random channels against the oracle."""
    rng = np.random.default_rng(seed)
    M = 16
    while M < B:
        M *= 2
    rows = 2048 // max(M, 256) if M >= 256 else 2048 // M
    C = tiles * rows // 4 + (extra % rows)
    C = max(C, 2 * 148 * rows // 2 + 1)                                    # never few enough rows for the split plan
    lens = [int(rng.integers(1, P * B + 1)) for _ in range(n_irs)]
    irs = [synth.decaying_ir(2100 + seed + j, lens[j], j) for j in range(n_irs)]
    bind = (lambda c: int((c * 7 + c // 3) % n_irs)) if mixed else (lambda c: int((c // rows) % n_irs))
    x = _blocks(C, nb, B, seed)
    A = min(C, max(rows, (C // rows - active_tiles) * rows))               # active channels for the second half (whole tiles)
    k = nb // 2
    outs = []
    for fused in (True, False):
        with eng.Engine(B, P, C, n_irs) as e:
            for j, h in enumerate(irs):
                e.set_ir(j, h)
            for c in range(C):
                e.bind(c, c + 1, bind(c))
            e.set_fused_step(fused)
            plan = e.mac_plan()
            assert plan[1:] == (1, 1)
            y = np.zeros((nb, C, B), np.float32)
            if k:
                y[:k] = e.process(x[:k])
            e.set_active_channels(A)
            if nb - k:
                y[k:, :A] = e.process(np.ascontiguousarray(x[k:, :A]))
            outs.append(y)
    assert np.array_equal(outs[0], outs[1])
    for c in [int(v) for v in rng.integers(0, A, 3)]:
        want = orc.convolve_periodic(np.ascontiguousarray(x[:, c, :]).reshape(-1), irs[bind(c)], B)[0, :nb * B]
        e_, l2 = parity(outs[0][:, c, :].reshape(-1), want)
        assert e_ <= TOL and l2 <= TOL, (c, e_, l2)
