"""GPU: several devices driven from one process (irb_group_*).  The sharding logic is exercised on whatever
devices the box has -- with one GPU the group's engines all live on device 0, which still checks the channel ranges, the
strided host layout and the routing of shared / private IRs; with more GPUs each engine gets its own device."""
import numpy as np
import pytest

from irbaboon_b200 import synth

pytestmark = pytest.mark.gpu


def _devices(eng, n):
    have = max(1, eng.device_count())
    return [i % have for i in range(n)]


@pytest.mark.parametrize("B,C,n_dev", [(512, 64, 3), (256, 100, 2), (64, 1000, 4), (1024, 9, 2)])
def test_group_equals_single_engine_shared_irs(eng, B, C, n_dev):
    nb, P = 6, 4
    rng = np.random.default_rng(C)
    x = (rng.random((nb, C, B), dtype=np.float32) * 2 - 1).astype(np.float32)
    h0, h1 = synth.decaying_ir(2000, P * B - 3), synth.decaying_ir(2001, 2 * B + 1, 1)
    with eng.Engine(B, P, C, 2) as e:
        T = e.tile_channels
        e.set_ir(0, h0); e.set_ir(1, h1)
        e.bind(T, min(C, 3 * T), 1)
        want = e.process(x)
        e.reset()
        want1 = np.stack([e.process(x[k]) for k in range(nb)])
    with eng.Group(_devices(eng, n_dev), B, P, C, 2) as g:
        r = g.ranges
        assert r[0][0] == 0 and r[-1][1] == C and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert all(b % T == 0 for b, _ in r) and all(e_ > b for b, e_ in r)            # whole tiles, nobody empty
        g.set_ir(0, h0); g.set_ir(1, h1)
        g.bind(T, min(C, 3 * T), 1)
        got = g.process(x)
        g.reset()
        got1 = np.stack([g.process(x[k]) for k in range(nb)])
    assert np.array_equal(got, want) and np.array_equal(got1, want) and np.array_equal(want1, want)


def test_group_private_irs_live_with_their_channel(eng, orc):
    B, C, P, nb = 256, 37, 5, 8
    x = np.stack([synth.white_noise(1011, c, nb * B) for c in range(C)])
    irs = [synth.decaying_ir(2200 + c, P * B - 11 * c, c) for c in range(C)]
    blocks = np.ascontiguousarray(x.reshape(C, nb, B).transpose(1, 0, 2))
    with eng.Group(_devices(eng, 3), B, P, C, 0) as g:
        for c in range(C):
            g.set_ir(c, irs[c])
        with pytest.raises(eng.IrbError):
            g.bind(0, C, 0)
        y = g.process(blocks)
    y = np.ascontiguousarray(y.transpose(1, 0, 2)).reshape(C, nb * B)
    for c in (0, 1, 12, 13, 24, 25, C - 1):
        want = orc.convolve_periodic(x[c], irs[c], B)[0, :nb * B]
        assert np.abs(y[c] - want).max() <= 1e-5 * max(1.0, np.abs(want).max())


def test_group_argument_checks(eng):
    L = eng.lib()
    with pytest.raises(eng.IrbError):
        eng.Group([0, 0, 0, 0], 512, 4, 9, 1)                       # fewer tiles than devices
    with pytest.raises(eng.IrbError):
        eng.Group([], 512, 4, 64, 1)
    assert L.irb_group_destroy(None) == 0 and L.irb_group_reset(None) == eng.IRB_ERR_ARG
