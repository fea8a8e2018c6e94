"""GPU parity of the streaming engine's launch modes against the CPU oracle:

* the few-row MAC (k_mac_slots: a row's partitions split over tile slots and the CTAs of a cluster, partial sums
  reduced through distributed shared memory) -- BASELINE config 2, the single-stream latency path;
* the round-robin IR switch (irb_engine_stage_ir; Source/PluginProcessor.cpp:411-414,455-461) and the block order of a
  plug-in callback (irb_engine_process_callback; :421-518) against the oracle's restatement of processBlock.
"""
import numpy as np
import pytest

from conftest import TOL, parity
from irbaboon_b200 import synth

pytestmark = pytest.mark.gpu


def _check(got, want, tol=TOL):
    assert got.shape == want.shape
    e, l2 = parity(got, want)
    assert e <= tol and l2 <= tol, (e, l2)


# ---- few rows: partitions split inside the tile and across a cluster -------------------------------------------
@pytest.mark.parametrize("B,Lh,C,split,cluster", [
    (256, 256 * 37 + 11, 2, 1, 1), (256, 256 * 37 + 11, 2, 8, 1), (256, 256 * 37 + 11, 2, 1, 8), (256, 256 * 37 + 11, 2, 8, 16),
    (256, 256 * 37 + 11, 2, 2, 4), (512, 512 * 21, 1, 4, 16), (512, 512 * 21, 5, 2, 2), (1024, 1024 * 9 + 1, 3, 2, 8),
    (2048, 2048 * 6, 2, 1, 16), (64, 64 * 50 + 3, 3, 32, 8), (64, 64 * 50 + 3, 40, 4, 2), (16, 16 * 70, 2, 128, 8),
    (128, 128 * 3, 2, 16, 16),          # more ranges than partitions: most slots are empty
])
def test_split_mac_matches_oracle(eng, orc, B, Lh, C, split, cluster):
    n = 14 * B
    x = np.stack([synth.white_noise(1002, c, n) for c in range(C)])
    h = synth.decaying_ir(2000, Lh)
    P = -(-Lh // B)
    with eng.Engine(B, P + 2, C, 1) as e:
        e.set_ir(0, h)
        e.set_mac_split(split, cluster)
        slots, s_in, cl = e.mac_plan()
        assert slots == (s_in > 1 or cl > 1)
        y = e.process_stream(x)
        # one CTA per tile, the bandwidth path, on the same state
        e.reset()
        e.set_mac_split(1, 1)
        assert e.mac_plan() == (False, 1, 1)
        y1 = e.process_stream(x)
    for c in range(C):
        want = orc.convolve_periodic(x[c], h, B)[:, :n]
        _check(y[c:c + 1], want)
        _check(y1[c:c + 1], want)


@pytest.mark.parametrize("B,Lh,C,split,cluster", [(256, 256 * 37 + 11, 2, 8, 16), (256, 256 * 37 + 11, 2, 2, 4), (512, 512 * 21, 5, 2, 2), (1024, 1024 * 9 + 1, 3, 2, 8),
                                                   (2048, 2048 * 6, 2, 1, 16), (64, 64 * 50 + 3, 3, 32, 8), (16, 16 * 70, 2, 128, 8), (300, 300 * 20 + 5, 3, 4, 4), (512, 512 * 21, 5, 4, 1), (256, 256 * 3, 2, 2, 8)])
def test_split_mac_with_fused_forward_equals_its_two_launch_form(eng, orc, B, Lh, C, split, cluster):
    """The latency path in ONE launch: cluster rank 0 transforms the new blocks itself (partition 0's operand comes out of
    shared memory), the heads move after the cluster's first barrier.  Bit-identical to k_fwd + the cluster kernel; ring wrap."""
    P = -(-Lh // B)
    nb = P + 6
    rng = np.random.default_rng(B + C)
    x = (rng.random((nb, C, B), dtype=np.float32) * 2 - 1).astype(np.float32)
    h = synth.decaying_ir(2000, Lh)
    outs = []
    for fuse in (1, 0):
        eng.set_tuning("fuse_split", fuse)
        try:
            with eng.Engine(B, P, C, 1) as e:
                e.set_ir(0, h)
                e.set_mac_split(split, cluster)
                l0 = e.launches
                outs.append(np.stack([e.process(x[k]) for k in range(nb)]))           # one block per call: plain, then graph replays
                assert e.launches - l0 == (nb if fuse else 2 * nb)
        finally:
            eng.set_tuning("fuse_split", 1)
    # on a cluster the fused form gives rank 0 partition 0 alone (it also runs the transform), so the ranges -- the blocking of
    # the sum -- differ from the two-launch form's: equal within rounding there, bit for bit without a cluster
    if cluster == 1:
        assert np.array_equal(outs[0], outs[1])
    else:
        assert np.abs(outs[0] - outs[1]).max() <= 2e-6 * max(1.0, float(np.abs(outs[1]).max()))
    for c in range(C):
        want = orc.convolve_periodic(np.ascontiguousarray(x[:, c, :]).reshape(-1), h, B)[:, :nb * B]
        _check(outs[0][:, c, :].reshape(1, -1), want)
        _check(outs[1][:, c, :].reshape(1, -1), want)


def test_split_is_chosen_automatically_for_few_rows_only(eng):
    with eng.Engine(256, 750, 2, 2) as e:                       # BASELINE config 2: stereo, 4 s IR, one stream
        h = synth.decaying_ir(2000, 192000)
        e.set_ir(0, h)
        e.set_ir(1, h)
        e.bind(1, 2, 1)
        slots, s_in, cl = e.mac_plan()
        assert slots and s_in == 8 and cl == 16
    with eng.Engine(512, 4, 4096, 1) as e:                      # throughput shape: one CTA per tile, shared-IR kernel
        e.set_ir(0, synth.decaying_ir(2000, 2000))
        assert e.mac_plan() == (False, 1, 1)


def test_split_mac_per_stream_irs_of_different_lengths(eng, orc):
    B, C = 256, 3
    n = 12 * B
    lens = [B * 20 + 5, B * 3, B * 11 + 100]
    irs = [synth.decaying_ir(2100 + c, lens[c], c) for c in range(C)]
    x = np.stack([synth.white_noise(1004, c, n) for c in range(C)])
    with eng.Engine(B, 21, C, C) as e:
        for c in range(C):
            e.set_ir(c, irs[c])
            e.bind(c, c + 1, c)
        for split, cluster in [(0, 0), (2, 1), (1, 4), (4, 4)]:
            e.reset()
            e.set_mac_split(split, cluster)
            y = e.process_stream(x)
            for c in range(C):
                _check(y[c:c + 1], orc.convolve_periodic(x[c], irs[c], B)[:, :n])


def test_config2_shape_against_reference(eng, orc):
    """BASELINE config 2 at full IR size: stereo, B=256, 4 s stereo IR (channel-wise), one stream, block by block."""
    B, Lh, nb = 256, 192000, 40
    n = nb * B
    x = np.stack([synth.white_noise(1002, c, n) for c in range(2)])
    h = np.stack([synth.decaying_ir(2000 + c, Lh, c) for c in range(2)])
    with eng.Engine(B, 750, 2, 2) as e:
        e.set_ir(0, h[0])
        e.set_ir(1, h[1])
        e.bind(1, 2, 1)
        blocks = np.ascontiguousarray(x.reshape(2, nb, B).transpose(1, 0, 2))
        y = np.stack([e.process(blocks[b]) for b in range(nb)])           # one call per block: the latency path
    y = np.ascontiguousarray(y.transpose(1, 0, 2)).reshape(2, n)
    _check(y, orc.convolve_periodic(x, h, B)[:, :n])


# ---- round-robin IR switch and callback order against the oracle's processBlock restatement --------------------
def _rt_oracle_run(orc, B, H, x, h0, switch_at=None, h1=None):
    C, n = x.shape
    e = orc.rt_engine(B, H, C, h0)
    out = []
    for k, i in enumerate(range(0, n, H)):
        if switch_at is not None and k == switch_at:
            e.set_ir(h1)
        out.append(e.process(x[:, i:i + H]))
    e.close()
    return np.concatenate(out, axis=1)


@pytest.mark.parametrize("B,C,P", [(256, 2, 8), (64, 3, 5), (512, 1, 12)])
def test_staged_ir_round_robin_matches_oracle_processblock(eng, orc, B, C, P):
    """Host block == B: the engine's block k is the oracle's output delayed by its reported latency (B samples)."""
    nb = 5 * P
    n = nb * B
    x = np.stack([synth.white_noise(1005, c, n) for c in range(C)])
    h0 = synth.decaying_ir(2000, P * B - 7)
    h1 = synth.decaying_ir(2001, P * B - 40, 1)
    sw = 2 * P + 3
    want = _rt_oracle_run(orc, B, B, x, h0, sw, h1)
    with eng.Engine(B, P, C, 1) as e:
        e.stage_ir(0, h0)                                     # nothing transformed yet: partitions arrive one per block
        assert e.partitions(0) == P
        blocks = np.ascontiguousarray(x.reshape(C, nb, B).transpose(1, 0, 2))
        ys = []
        for k in range(nb):
            if k == sw:
                e.stage_ir(0, h1)
            ys.append(e.process(blocks[k]))
    y = np.ascontiguousarray(np.stack(ys).transpose(1, 0, 2)).reshape(C, n)
    assert not want[:, :B].any()
    _check(y[:, :n - B], want[:, B:])
    # while the new IR fades in the output is neither the old nor the new convolution: the test above pins the mix


def test_staged_ir_equals_preloaded_ir_from_a_cold_start(eng):
    """Round-robin loading from block 0 is indistinguishable from a preloaded IR (partition p is first needed at block p)."""
    B, C, P = 128, 2, 6
    x = np.stack([synth.white_noise(1005, c, 20 * B) for c in range(C)])
    h = synth.decaying_ir(2000, P * B)
    with eng.Engine(B, P, C, 1) as e:
        e.set_ir(0, h)
        a = e.process_stream(x)
    with eng.Engine(B, P, C, 1) as e:
        e.stage_ir(0, h)
        b = e.process_stream(x)
        e.reset()                                             # prepareToPlay: spectra cleared, refresh restarts at 0
        c = e.process_stream(x)
    assert np.array_equal(a, b) and np.array_equal(a, c)


@pytest.mark.parametrize("B,H,C,P,ring", [(256, 512, 2, 8, 8), (256, 1024, 2, 8, 8), (64, 256, 1, 3, 4), (128, 256, 2, 4, 16)])
def test_callback_order_matches_oracle_when_host_block_exceeds_B(eng, orc, B, H, C, P, ring):
    """hostBlock = m*B: the reference transforms the callback's m blocks before convolving any of them, so with its
    ring of max(P, m) slots the oldest partitions see the callback's later blocks.  irb_engine_process_callback keeps
    that order; the oracle restates PluginProcessor.cpp:403-562."""
    m = H // B
    ncb = 12
    n = ncb * H
    x = np.stack([synth.white_noise(1006, c, n) for c in range(C)])
    h = synth.decaying_ir(2000, P * B - 3)
    want = _rt_oracle_run(orc, B, H, x, h)
    with eng.Engine(B, ring, C, 1) as e:
        e.stage_ir(0, h, P)
        ys = []
        for k in range(ncb):
            blk = x[:, k * H:(k + 1) * H].reshape(C, m, B).transpose(1, 0, 2)
            ys.append(e.process_callback(blk))
    y = np.ascontiguousarray(np.concatenate(ys).transpose(1, 0, 2)).reshape(C, n)
    if ring == max(P, m):
        assert not want[:, :H].any()
        _check(y[:, :n - H], want[:, H:])                     # latency = hostBlock (PluginProcessor.cpp:166-171)
    else:
        # a ring with room for partitions + m - 1 spectra gives the causal result instead
        causal = orc.convolve_periodic(x, h[None, :], B)[:, :n]
        _check(y, causal)


# ---- one launch per block step: forward transform fused into the MAC kernel --------------------------------------
@pytest.mark.parametrize("B,C,Lh", [(512, 300, 512 * 9 + 5), (256, 1200, 256 * 3), (1024, 150, 1024 * 5), (2048, 160, 2048 * 2 + 1), (64, 5000, 64 * 7), (100, 2000, 900), (101, 1500, 700), (33, 9000, 100)])
def test_fused_step_is_bit_identical_to_two_launches_and_matches_oracle(eng, orc, B, C, Lh):
    n = 7 * B
    x = np.stack([synth.white_noise(1009, c % 5, n) for c in range(C)])
    h = synth.decaying_ir(2000, Lh)
    P = -(-Lh // B)
    with eng.Engine(B, P, C, 1) as e:
        e.set_ir(0, h)
        assert e.mac_plan() == (False, 1, 1)                   # enough tiles: the shared-IR kernel
        l0 = e.launches
        a = e.process_stream(x)
        fused_launches = e.launches - l0
        e.reset()
        e.set_fused_step(False)
        l0 = e.launches
        b = e.process_stream(x)
        assert e.launches - l0 == 2 * fused_launches == 2 * 7   # one launch per block instead of two
        fdl_b = [e.fdl_spectrum(C // 2, age) for age in range(min(P, 4))]
        e.reset()
        e.set_fused_step(True)
        e.process_stream(x)
        fdl_a = [e.fdl_spectrum(C // 2, age) for age in range(min(P, 4))]
    assert np.array_equal(a, b)
    assert all(np.array_equal(u, v) for u, v in zip(fdl_a, fdl_b))      # the fused kernel writes the same spectra into the FDL
    for c in (0, C - 1):
        _check(a[c:c + 1], orc.convolve_periodic(x[c], h, B)[:, :n])


# ---- the IR ring under load: more tiles than fit on the GPU at once, several laps of the ring -------------------
def test_ir_ring_release_under_load_fused_equals_two_launches_on_every_channel(eng, orc):
    """Regression (profiles/r01_ring_release_race.md): with every SM holding three CTAs, a warp's last ld.shared of an IR
    stage could still be in flight when its release of the stage was seen by the TMA producer; about 1 % of the (tile,
    block) pairs then used partition g + kStages for half of a warp's bins.  Small cases never showed it, so this one
    fills the GPU (513 tiles of 4 channels, 444 fit at once) and compares EVERY channel of every block."""
    B, C, P, nb = 512, 2050, 48, 24
    rng = np.random.default_rng(11)
    h = synth.decaying_ir(2000, P * B - 3)
    x = (rng.random((nb, C, B), dtype=np.float32) * 2 - 1).astype(np.float32)
    outs = []
    for fused in (True, False, True):
        with eng.Engine(B, P, C, 1) as e:
            e.set_ir(0, h)
            e.set_fused_step(fused)
            assert e.mac_plan() == (False, 1, 1)
            outs.append(np.concatenate([e.process(x[k:k + 8]) for k in range(0, nb, 8)]))
    assert np.array_equal(outs[0], outs[1])
    assert np.array_equal(outs[0], outs[2])
    for c in (0, 1027, C - 1):
        xc = np.ascontiguousarray(outs[0][:, c, :]).reshape(1, -1)
        _check(xc, orc.convolve_periodic(np.ascontiguousarray(x[:, c, :]).reshape(-1), h, B)[:, :nb * B])


@pytest.mark.parametrize("B,C,P,nb", [(256, 4100, 24, 16), (1024, 1100, 14, 16), (2048, 620, 11, 8), (100, 8200, 30, 16), (64, 20000, 40, 16)])
def test_full_gpu_block_steps_other_block_sizes(eng, orc, B, C, P, nb):
    """Every FFT size of the streaming kernels with more tiles than the GPU holds at once and several ring laps."""
    rng = np.random.default_rng(13)
    h = synth.decaying_ir(2000, P * B - 1)
    x = (rng.random((nb, C, B), dtype=np.float32) * 2 - 1).astype(np.float32)
    outs = []
    for fused in (True, False):
        with eng.Engine(B, P, C, 1) as e:
            e.set_ir(0, h)
            e.set_fused_step(fused)
            assert e.mac_plan() == (False, 1, 1)
            outs.append(np.concatenate([e.process(x[k:k + 8]) for k in range(0, nb, 8)]))
    assert np.array_equal(outs[0], outs[1])
    for c in (0, C // 2, C - 1):
        xc = np.ascontiguousarray(outs[0][:, c, :]).reshape(1, -1)
        _check(xc, orc.convolve_periodic(np.ascontiguousarray(x[:, c, :]).reshape(-1), h, B)[:, :nb * B])


@pytest.mark.parametrize("B,np_ir", [(512, 0), (512, 1), (512, 2), (512, 3), (512, 4), (256, 2), (1024, 1), (2048, 2)])
def test_fused_step_with_fewer_partitions_than_ring_stages(eng, orc, B, np_ir):
    """IRs shorter than the TMA ring (0 = no IR loaded at all: silence, no hang), inside a longer FDL ring, GPU full."""
    C, ring, nb = 2050 * 512 // B, 8, 10
    rng = np.random.default_rng(17)
    x = (rng.random((nb, C, B), dtype=np.float32) * 2 - 1).astype(np.float32)
    h = synth.decaying_ir(2000, np_ir * B - 1) if np_ir else None
    outs = []
    for fused in (True, False):
        with eng.Engine(B, ring, C, 1) as e:
            if h is not None:
                e.set_ir(0, h)
                assert e.partitions(0) == np_ir
            e.set_fused_step(fused)
            outs.append(e.process(x))
    assert np.array_equal(outs[0], outs[1])
    if h is None:
        assert not outs[0].any()
    else:
        for c in (0, C - 1):
            xc = np.ascontiguousarray(outs[0][:, c, :]).reshape(1, -1)
            _check(xc, orc.convolve_periodic(np.ascontiguousarray(x[:, c, :]).reshape(-1), h, B)[:, :nb * B])


def test_ir_ring_release_under_load_per_stream_irs(eng, orc):
    """The same for the slot kernel (every tile slot stages its own IR partitions): repeatable and equal to the oracle."""
    B, C, P, nb = 512, 1640, 20, 16
    rng = np.random.default_rng(12)
    hs = [synth.decaying_ir(3000 + c % 7, P * B - (c % 5), c % 2) for c in range(C)]
    x = (rng.random((nb, C, B), dtype=np.float32) * 2 - 1).astype(np.float32)
    outs = []
    for rep in range(2):
        with eng.Engine(B, P, C, C) as e:
            for c in range(C):
                e.set_ir(c, hs[c])
                e.bind(c, c + 1, c)
            assert e.mac_plan()[0]
            outs.append(np.concatenate([e.process(x[k:k + 8]) for k in range(0, nb, 8)]))
    assert np.array_equal(outs[0], outs[1])
    for c in (0, 821, C - 1):
        xc = np.ascontiguousarray(outs[0][:, c, :]).reshape(1, -1)
        _check(xc, orc.convolve_periodic(np.ascontiguousarray(x[:, c, :]).reshape(-1), hs[c], B)[:, :nb * B])


def test_fused_step_with_staged_ir_refresh_rows(eng, orc):
    """Round-robin IR refresh rows still run (their own k_fwd launch) next to the fused block step."""
    B, C, P = 256, 600, 6
    nb = 4 * P
    n = nb * B
    x = np.stack([synth.white_noise(1005, c % 3, n) for c in range(C)])
    h0, h1 = synth.decaying_ir(2000, P * B), synth.decaying_ir(2001, P * B, 1)
    want = _rt_oracle_run(orc, B, B, x[:3], h0, P + 2, h1)
    with eng.Engine(B, P, C, 1) as e:
        e.stage_ir(0, h0)
        assert e.mac_plan() == (False, 1, 1)
        blocks = np.ascontiguousarray(x.reshape(C, nb, B).transpose(1, 0, 2))
        ys = []
        for k in range(nb):
            if k == P + 2:
                e.stage_ir(0, h1)
            ys.append(e.process(blocks[k]))
    y = np.ascontiguousarray(np.stack(ys).transpose(1, 0, 2)).reshape(C, n)
    _check(y[:3, :n - B], want[:, B:])
    assert np.array_equal(y[:3], y[3:6])                       # channels repeat every 3: same input, same output


def test_single_block_calls_replay_a_graph_with_identical_results(eng, orc):
    """One small block per call (the live-callback shape): from the second call on, copy-in, the step's kernels and copy-out
    are one captured CUDA graph.  Results equal the multi-block pipelined call bit for bit, across an IR switch, a reset and
    a change of the launch plan (each of which re-captures)."""
    B, C, P = 256, 2, 12
    nb = 30
    x = np.stack([synth.white_noise(1010, c, nb * B) for c in range(C)])
    h0, h1 = synth.decaying_ir(2000, P * B), synth.decaying_ir(2001, P * B - 9, 1)
    blocks = np.ascontiguousarray(x.reshape(C, nb, B).transpose(1, 0, 2))
    with eng.Engine(B, P, C, 1) as e:
        e.set_ir(0, h0)
        whole = e.process(blocks)                                  # pipelined multi-block path
        e.reset()
        one = np.stack([e.process(blocks[k]) for k in range(nb)])  # plain first call, graph replays after
        assert np.array_equal(one, whole)
        l0 = e.launches
        e.process(blocks[0])
        assert e.launches - l0 == 1                                # the replayed graph holds the step's ONE kernel (cluster kernel, forward transform fused)
        e.reset()
        e.set_mac_split(1, 1)                                      # different plan -> different kernels -> re-capture
        one2 = np.stack([e.process(blocks[k]) for k in range(nb)])
        e.reset()
        e.set_mac_split(0, 0)
        e.stage_ir(0, h0)                                          # adds a refresh row to the forward launch -> re-capture
        ys = []
        for k in range(nb):
            if k == 10:
                e.stage_ir(0, h1)
            ys.append(e.process(blocks[k]))
    y1 = np.ascontiguousarray(one2.transpose(1, 0, 2)).reshape(C, nb * B)
    _check(y1, orc.convolve_periodic(x, np.stack([h0, h0]), B)[:, :nb * B])
    y = np.ascontiguousarray(np.stack(ys).transpose(1, 0, 2)).reshape(C, nb * B)
    want = _rt_oracle_run(orc, B, B, x, h0, 10, h1)
    _check(y[:, :-B], want[:, B:])


@pytest.mark.parametrize("B,C,ns", [(64, 40000, 1), (512, 9500, 1), (256, 5000, 5000)])
def test_large_single_block_calls_pipeline_channel_groups(eng, B, C, ns):
    """A single-block call with megabytes of audio is cut into channel groups whose upload, kernels and download overlap.
    Same results, bit for bit, as the multi-block call (which pipelines whole blocks) -- shared IR and per-stream IRs."""
    nb, P = 5, 3
    rng = np.random.default_rng(B)
    x = (rng.random((nb, C, B), dtype=np.float32) * 2 - 1).astype(np.float32)
    with eng.Engine(B, P, C, ns) as e:
        if ns == 1:
            e.set_ir(0, synth.decaying_ir(2000, P * B - 5))
        else:
            irs = [synth.decaying_ir(2100 + j, P * B - j, j) for j in range(4)]
            for c in range(C):
                e.set_ir(c, irs[c % 4])
                e.bind(c, c + 1, c)
        whole = e.process(x)
        e.reset()
        single = np.stack([e.process(x[k]) for k in range(nb)])
    assert np.array_equal(single, whole)
    assert np.abs(whole).max() > 0.1


def test_submit_wait_keeps_state_and_results_of_the_synchronous_path(eng):
    """irb_engine_submit / irb_engine_wait: several submissions in flight, then single-block and multi-block synchronous calls
    on the same engine -- the output sequence is bit-identical to one synchronous call over all blocks."""
    B, C, P, nb = 128, 600, 5, 24
    rng = np.random.default_rng(3)
    x = (rng.random((nb, C, B), dtype=np.float32) * 2 - 1).astype(np.float32)
    with eng.Engine(B, P, C, 1) as e:
        e.set_ir(0, synth.decaying_ir(2000, P * B - 1))
        want = e.process(x)
        e.reset()
        hin, hout = eng.pinned_empty((nb, C, B)), eng.pinned_empty((nb, C, B))
        hin[:] = x
        hout[:] = 0
        e.submit(hin[0:6], hout[0:6])
        e.submit(hin[6:8], hout[6:8])
        e.submit(hin[8:13], hout[8:13])
        e.wait()
        hout[13] = e.process(hin[13])                      # a live single block drains nothing it should not
        e.submit(hin[14:20], hout[14:20])
        hout[20] = e.process(hin[20])                      # single block right behind an un-waited submission
        hout[21:24] = e.process(hin[21:24])
        got = np.array(hout)
        with pytest.raises(eng.IrbError):
            e.submit(hin[0:1], hout[0:1])
        eng.pinned_free(hin); eng.pinned_free(hout)
    assert np.array_equal(got, want)


def test_host_pipeline_equals_device_resident_steps_at_scale(eng):
    """The three-stage host pipeline (upload | kernels | download over double-buffered staging, irb_engine_process /
    submit / wait) and the single-large-block path against block steps on device-resident buffers: same bits, with enough
    channels (32 MB per block) that copies and kernels of neighbouring blocks really overlap."""
    torch = pytest.importorskip("torch")
    B, C, P, nb = 512, 16384, 12, 12
    rng = np.random.default_rng(21)
    h = synth.decaying_ir(2000, P * B - 2)
    x = (rng.random((nb, C, B), dtype=np.float32) * 2 - 1).astype(np.float32)
    with eng.Engine(B, P, C, 1) as e:
        e.set_ir(0, h)
        d_in = torch.from_numpy(x).cuda()
        d_out = torch.empty((nb, C, B), device="cuda", dtype=torch.float32)
        torch.cuda.synchronize()
        for k in range(nb):
            e.process_device(d_in[k].data_ptr(), d_out[k].data_ptr(), 1)
        e.synchronize()
        want = d_out.cpu().numpy()
        e.reset()
        got = np.concatenate([e.process(x[0:5]), e.process(x[5:6]), e.process(x[6:12])])      # pipeline, one large block, pipeline
        assert np.array_equal(got, want)
        e.reset()
        hin, hout = eng.pinned_empty((nb, C, B)), eng.pinned_empty((nb, C, B))
        hin[:] = x
        e.submit(hin[:7], hout[:7])
        e.submit(hin[7:], hout[7:])
        e.wait()
        assert np.array_equal(hout, want)
        eng.pinned_free(hin)
        eng.pinned_free(hout)
