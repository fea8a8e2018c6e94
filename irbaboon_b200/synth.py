"""Deterministic synthetic inputs.

Counter-based so C, CUDA-side tests and numpy agree bit for bit:
``u64 = splitmix64(seed + (stream << 40) + i)``; white noise = ``((u64 >> 40) * 2^-24) * 2 - 1``.
Sine follows tools::sineFill (fp/tools.cpp:270-280); the sweep follows ExpSineSweep::generate
(fp/ExpSineSweep.cpp:26-41, 212-220).  Host-side helpers only -- no device code here.
"""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    x = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15)) & _M64
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return x ^ (x >> np.uint64(31))


def _u64(seed, stream, n):
    with np.errstate(over="ignore"):
        base = np.uint64(seed) + (np.uint64(stream) << np.uint64(40))
        return splitmix64(base + np.arange(n, dtype=np.uint64))


def white_noise(seed, stream, n):
    """float32 white noise in [-1, 1)."""
    u = _u64(seed, stream, n)
    return ((u >> np.uint64(40)).astype(np.float64) * (1.0 / 16777216.0) * 2.0 - 1.0).astype(np.float32)


def sine(n, freq=1000.0, sample_rate=48000.0, ampl=1.0):
    i = np.arange(n, dtype=np.float64)
    return (ampl * np.cos(np.pi * 2 * freq * i / sample_rate)).astype(np.float32)


def normal(seed, stream, n):
    """float64 standard normal via Box-Muller from the same counter RNG."""
    u = _u64(seed, stream, 2 * n)
    u1 = ((u[0::2] >> np.uint64(11)).astype(np.float64) + 1.0) * (1.0 / 9007199254740993.0)
    u2 = (u[1::2] >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def decaying_ir(seed, n, ir_id=0):
    """Exponentially decaying noise IR (-60 dB at the end), unit L2 norm, float32."""
    i = np.arange(n, dtype=np.float64)
    h = normal(seed, ir_id, n) * np.exp(-6.9078 * i / n)
    h /= np.sqrt(np.sum(h * h))
    return h.astype(np.float32)


def exp_sine_sweep(duration_s, sample_rate, f1, f2, gain_db=0.0):
    """float64 Farina sweep, same formula as ExpSineSweep::generate."""
    T = sample_rate * duration_s
    w1 = f1 / sample_rate * 2 * np.pi
    w2 = f2 / sample_rate * 2 * np.pi
    K = T * w1 / np.log(w2 / w1)
    L = T / np.log(w2 / w1)
    i = np.arange(int(T), dtype=np.float64)
    return 10.0 ** (gain_db / 20.0) * np.sin(K * (np.exp(i / L) - 1.0))
