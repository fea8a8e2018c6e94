"""Soak run (not collected by pytest): the fused block step (the persistent kernel k_mac_p; with per-stream IRs its PERROW form; with
few rows the cluster kernel with the forward transform on rank 0) against its two-launch form in lockstep, every channel of every
block, with more units than the GPU holds at once and many laps of the FDL ring, for every FFT size.
python tests/soak_block_step.py  (needs a B200)."""
import sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from irbaboon_b200 import engine as eng, synth
def soak(B, C, P, ncalls, seed, per_stream=False, exact=True):
    rng = np.random.default_rng(seed)
    h = synth.decaying_ir(2000, P * B - 3)
    nir = 8 if per_stream else 1
    ew = eng.Engine(B, P, C, nir); et = eng.Engine(B, P, C, nir)
    for e in (ew, et):
        if per_stream:
            for j in range(8): e.set_ir(j, synth.decaying_ir(2100 + j, max(B, (P - j) * B - 5 * j), j))
            for c in range(C): e.bind(c, c + 1, (3 * c + c // 5) % 8)
        else:
            e.set_ir(0, h)
    et.set_fused_step(False)
    bad = 0; t0 = time.time()
    x = (rng.random((8, C, B), dtype=np.float32) * 2 - 1).astype(np.float32)
    for k in range(ncalls):
        x = np.roll(x, 1, axis=0); x[0] = -x[0]            # cheap new input per call
        yw = ew.process(x); yt = et.process(x)
        if exact:
            if not np.array_equal(yw, yt): bad += 1
        elif np.abs(yw - yt).max() > 2e-6 * max(1.0, float(np.abs(yt).max())): bad += 1
    plan = ew.mac_plan()
    ew.close(); et.close()
    print("B=%d C=%d P=%d %s plan %s: %d blocks x %d tiles, mismatching calls: %d  (%.1f s)" % (B, C, P, "per-stream IRs" if per_stream else "shared IR", plan, 8 * ncalls, -(-C // (2048 // max(16, B))), bad, time.time() - t0), flush=True)
    return bad
tot = 0
tot += soak(512, 2050, 188, 150, 1)
tot += soak(512, 8200, 40, 60, 2)
tot += soak(256, 4100, 100, 60, 3)
tot += soak(1024, 1030, 94, 100, 4)
tot += soak(2048, 520, 47, 100, 5)
tot += soak(1024, 1030, 94, 60, 6, per_stream=True)
tot += soak(512, 2050, 60, 60, 7, per_stream=True)
tot += soak(256, 4100, 40, 40, 8, per_stream=True)
tot += soak(256, 2, 750, 300, 9, exact=False)          # the latency path: cluster kernel, blocked summation differs from its two-launch form's
tot += soak(512, 6, 100, 200, 10, exact=False)
print("SOAK", "OK" if tot == 0 else "FAILED", tot)
