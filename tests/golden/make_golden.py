"""Generate tests/golden/*.npz from the REFERENCE's own object code (oracle/_ref/libirb_ref.so =
/root/reference/fp/*.cpp compiled unmodified against oracle/juce_shim).

Run here, where /root/reference exists:  python tests/golden/make_golden.py
The reference ships no tests or fixtures of its own, so these vectors -- outputs of the
reference run in this container on the deterministic inputs of irbaboon_b200/synth.py -- are the pins
that travel to the GPU box.  Inputs are regenerated from seeds; only outputs (and their float64 checksums) are
stored, float32, small.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from irbaboon_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (Lx, Lh, B, chx, chh, input kind)
PERIODIC_CASES = {
    "mono_b64": (3000, 700, 64, 1, 1, "noise"),
    "mono_b256_sine": (6000, 2048, 256, 1, 1, "sine"),
    "stereo_audio_mono_ir_b128": (2500, 900, 128, 2, 1, "noise"),
    "mono_audio_stereo_ir_b128": (2500, 900, 128, 1, 2, "noise"),
    "stereo_stereo_b512": (5000, 3000, 512, 2, 2, "noise"),
    "ragged_b100": (1234, 321, 100, 1, 1, "noise"),
    "multiple_of_block_b32": (640, 96, 32, 1, 1, "noise"),
}


def periodic_inputs(case):
    Lx, Lh, B, chx, chh, kind = PERIODIC_CASES[case]
    if kind == "noise":
        x = np.stack([synth.white_noise(1001, c, Lx) for c in range(chx)])
    else:
        x = np.stack([synth.sine(Lx) for _ in range(chx)])
    h = np.stack([synth.decaying_ir(2000 + c, Lh, c) for c in range(chh)])
    return x, h, B


def config1_inputs():
    return synth.white_noise(1001, 0, 480000), synth.decaying_ir(2000, 48000), 512


def main():
    ref = oracle.Reference()
    out = {}
    for case in PERIODIC_CASES:
        x, h, B = periodic_inputs(case)
        out["periodic/" + case] = ref.convolve_periodic(x, h, B)
    # BASELINE config 1 at full size: keep a strided sample, the head, the tail and float64 sums
    x, h, B = config1_inputs()
    y = ref.convolve_periodic(x, h, B)[0]
    out["config1/head"] = y[:2048].copy()
    out["config1/tail"] = y[-2048:].copy()
    out["config1/strided"] = y[::257].copy()
    out["config1/sums"] = np.array([y.astype(np.float64).sum(), (y.astype(np.float64) ** 2).sum(), float(np.abs(y).max())])
    # non-periodic + deconvolution + helpers
    xs = synth.white_noise(1005, 0, 3000)
    hs = synth.decaying_ir(2005, 1000)
    out["nonperiodic/mono"] = ref.convolve_nonperiodic(xs, hs)
    conv = ref.convolve_nonperiodic(xs, hs)[0]
    out["deconvolve/plain"] = ref.deconvolve(conv[:4096], np.pad(xs, (0, 1096)), 48000.0, False, True, True)
    out["deconvolve/smoothed"] = ref.deconvolve(conv[:4096], np.pad(xs, (0, 1096)), 48000.0, True, True, True)
    out["deconvolve/nophase"] = ref.deconvolve(conv[:4096], np.pad(xs, (0, 1096)), 48000.0, False, False, True)
    out["invert_filter"] = ref.invert_filter(hs, 48000)
    out["fft_transform"] = ref.fft_transform(xs[:1000])
    out["shifteroo_odd"] = ref.shifteroo(np.arange(9, dtype=np.float32))
    out["ess/sweep"] = ref.ess(0.25, 48000.0, 20.0, 20000.0).astype(np.float64)[::7]
    out["ess/inverse"] = ref.ess(0.25, 48000.0, 20.0, 20000.0, 0.0, True).astype(np.float64)[::7]
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print("wrote", len(out), "vectors,", os.path.getsize(os.path.join(HERE, "reference_vectors.npz")), "bytes")


if __name__ == "__main__":
    main()
