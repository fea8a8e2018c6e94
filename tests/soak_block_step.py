"""Soak run (not collected by pytest): the fused block step against its two-launch form in lockstep, every channel of every block,
with more tiles than the GPU holds at once, for every FFT size of k_mac_tma.  python tests/soak_block_step.py  (needs a B200)."""
import sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from irbaboon_b200 import engine as eng, synth
def soak(B, C, P, ncalls, seed):
    rng = np.random.default_rng(seed)
    h = synth.decaying_ir(2000, P * B - 3)
    ew = eng.Engine(B, P, C, 1); et = eng.Engine(B, P, C, 1)
    ew.set_ir(0, h); et.set_ir(0, h); et.set_fused_step(False)
    bad = 0; t0 = time.time()
    x = (rng.random((8, C, B), dtype=np.float32) * 2 - 1).astype(np.float32)
    for k in range(ncalls):
        x = np.roll(x, 1, axis=0); x[0] = -x[0]            # cheap new input per call
        yw = ew.process(x); yt = et.process(x)
        if not np.array_equal(yw, yt): bad += 1
    ew.close(); et.close()
    print("B=%d C=%d P=%d: %d blocks x %d tiles, mismatching calls: %d  (%.1f s)" % (B, C, P, 8 * ncalls, -(-C // (2048 // max(16, B))), bad, time.time() - t0), flush=True)
    return bad
tot = 0
tot += soak(512, 2050, 188, 150, 1)
tot += soak(512, 8200, 40, 60, 2)
tot += soak(256, 4100, 100, 60, 3)
tot += soak(1024, 1030, 94, 100, 4)
tot += soak(2048, 520, 47, 100, 5)
print("SOAK", "OK" if tot == 0 else "FAILED", tot)
