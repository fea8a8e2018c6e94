import sys, time, numpy as np
sys.path.insert(0, '.')
from irbaboon_b200 import engine as eng, synth
n = 1 << 20
sweep = eng.ess(n / 48000.0, 48000.0, 20.0, 24000.0).astype(np.float32)
caps = np.stack([synth.white_noise(4000 + j, 0, n) for j in range(8)])
eng.deconvolve_batch(caps[:2], sweep, 48000.0, True)
t0 = time.perf_counter(); eng.deconvolve_batch(caps, sweep, 48000.0, True); print("device ms", eng.last_compute_ms(), "wall", time.perf_counter() - t0)
