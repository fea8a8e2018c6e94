"""Soak of the wavefront smoothing launch (k_avg_passes): the same smoothed batch over and over, every result compared bit for bit
with the per-pass kernels' (one k_avg_scan + k_avg_apply per pass).  Batches large enough that two CTAs share an SM and several
groups run side by side.  Not collected by pytest (no test_ prefix): python tests/soak_smoothing.py [rounds]"""
import sys
import time

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from irbaboon_b200 import engine as eng, synth  # noqa: E402


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    bad = 0
    t0 = time.time()
    for n, nb in ((1 << 15, 600), (1 << 17, 160), (1 << 19, 40)):
        sweep = synth.exp_sine_sweep(n / 48000.0, 48000.0, 20.0, 20000.0).astype(np.float32)[:n]
        rng = np.random.default_rng(n)
        caps = (rng.standard_normal((nb, n)) * 0.1).astype(np.float32)
        caps += sweep[None, :] * rng.uniform(0.2, 1.0, (nb, 1)).astype(np.float32)
        eng.set_tuning("avg_fused", 0)
        want = eng.deconvolve_batch(caps, sweep, 48000.0, True)
        eng.set_tuning("avg_fused", 1)
        assert np.isfinite(want).all()
        for r in range(rounds):
            got = eng.deconvolve_batch(caps, sweep, 48000.0, True)
            if not np.array_equal(got, want):
                bad += 1
                d = np.abs(got - want).max(axis=1)
                print("MISMATCH n=%d round %d: %d captures differ, max |diff| %.3g" % (n, r, int((d > 0).sum()), float(d.max())), flush=True)
        print("n = 2^%d, %d captures, %d rounds: %s" % (int(np.log2(n)), nb, rounds, "all bit-identical to the per-pass kernels" if not bad else "%d mismatching rounds" % bad), flush=True)
    print("soak %s in %.1f s" % ("OK" if not bad else "FAILED", time.time() - t0))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
