"""CPU oracle for the IRBaboon partitioned-convolution path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  It binds two checkers through ctypes:

* ``libirb_oracle.so``      -- plain-C restatement (oracle/irb_oracle.c), always buildable with gcc;
* ``_ref/libirb_ref.so``    -- the reference's own fp/*.cpp compiled unmodified against oracle/juce_shim
                               (built here where /root/reference exists; travels to the GPU box prebuilt).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)


def build(reference="/root/reference"):
    """Compile the C restatement and, when the reference tree is present, oracle/_ref."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    if os.path.isdir(os.path.join(reference, "fp")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref", "REFERENCE=" + reference])


def _p(a):
    return a.ctypes.data_as(_f32p)


def _planar(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim == 1:
        a = a[None, :]
    return a


def next_pow2(x):
    if x != 0 and (x & (x - 1)) == 0:
        return x
    r = 1
    while r <= x:
        r *= 2
    return r


class _Lib:
    prefix = ""
    path = ""

    def __init__(self):
        if not os.path.exists(self.path):
            raise FileNotFoundError(self.path + " missing: run `make -C oracle` (or __graft_entry__.build())")
        self.lib = ctypes.CDLL(self.path)

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    # ---- shared surface -------------------------------------------------------------------------
    def convolve_periodic(self, x, h, B=256):
        x, h = _planar(x), _planar(h)
        out = np.zeros((x.shape[0], x.shape[1] + h.shape[1] - 1), np.float32)
        self._fn("convolve_periodic")(_p(x), x.shape[0], x.shape[1], _p(h), h.shape[0], h.shape[1], int(B), _p(out))
        return out

    def convolve_nonperiodic(self, x, h):
        x, h = _planar(x), _planar(h)
        Lo = x.shape[1] + h.shape[1] - 1
        out = np.zeros((x.shape[0], max(Lo, x.shape[1])), np.float32)
        n = self._fn("convolve_nonperiodic")(_p(x), x.shape[0], x.shape[1], _p(h), h.shape[0], h.shape[1], _p(out))
        return out.reshape(-1)[: x.shape[0] * n].reshape(x.shape[0], n).copy()

    def averaging_filter(self, spec, octave_fraction, sample_rate, log_avg=True, include_phase=True, include_amplitude=True):
        s = _planar(spec).copy()
        fn = self._fn("averaging_filter")
        fn.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        fn(_p(s), s.shape[0], s.shape[1], float(octave_fraction), float(sample_rate), int(log_avg), int(include_phase), int(include_amplitude))
        return s

    def fft_transform(self, x):
        x = _planar(x)
        N = next_pow2(x.shape[1])
        out = np.zeros((x.shape[0], 2 * N), np.float32)
        self._fft_transform(x, out)
        return out

    def fft_inv_transform(self, spec):
        s = _planar(spec)
        out = np.zeros((s.shape[0], s.shape[1] // 2), np.float32)
        self._fn("fft_inv_transform")(_p(s), s.shape[0], s.shape[1], _p(out))
        return out

    def shifteroo(self, buf):
        b = _planar(buf).copy()
        self._fn("shifteroo")(_p(b), b.shape[0], b.shape[1])
        return b


class Oracle(_Lib):
    """Plain-C restatement (oracle/irb_oracle.c)."""
    prefix = "orc_"
    path = os.path.join(_HERE, "libirb_oracle.so")

    def _fft_transform(self, x, out):
        self.lib.orc_fft_transform(_p(x), x.shape[0], x.shape[1], _p(out))

    def deconvolve(self, num, den, sample_rate=48000.0, smoothing=True, include_phase=True, include_amplitude=True):
        num, den = _planar(num)[0].copy(), _planar(den)[0].copy()
        N = next_pow2(max(len(num), len(den)))
        out = np.zeros(N, np.float32)
        fn = self.lib.orc_deconvolve
        fn.argtypes = [_f32p, ctypes.c_int, _f32p, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p]
        fn(_p(num), len(num), _p(den), len(den), float(sample_rate), int(smoothing), int(include_phase), int(include_amplitude), _p(out))
        return out[None, :]

    def invert_filter(self, x, sample_rate=48000):
        x = _planar(x)[0].copy()
        out = np.zeros(next_pow2(len(x)), np.float32)
        self.lib.orc_invert_filter(_p(x), len(x), int(sample_rate), _p(out))
        return out[None, :]

    def real_forward(self, buf, n):
        b = np.ascontiguousarray(buf, np.float32).copy()
        assert b.size == 2 * n
        self.lib.orc_real_forward(_p(b), n)
        return b

    def real_inverse(self, buf, n):
        b = np.ascontiguousarray(buf, np.float32).copy()
        assert b.size == 2 * n
        self.lib.orc_real_inverse(_p(b), n)
        return b

    def periodic_iterations(self, Lx, Lh, B):
        return self.lib.orc_periodic_iterations(Lx, Lh, B)

    def ess(self, dur, sr, f1, f2, gain_db=0.0, inverse=False):
        fn = self.lib.orc_ess
        fn.argtypes = [ctypes.c_double] * 5 + [ctypes.c_int, _f64p]
        n = fn(dur, sr, f1, f2, gain_db, int(inverse), None)
        out = np.zeros(n, np.float64)
        fn(dur, sr, f1, f2, gain_db, int(inverse), out.ctypes.data_as(_f64p))
        return out

    def white_noise(self, seed, stream, n):
        out = np.zeros(n, np.float32)
        fn = self.lib.orc_white_noise
        fn.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, _f32p]
        fn(seed, stream, n, _p(out))
        return out

    def rt_engine(self, B, host_block, channels, ir):
        return OracleRtEngine(self.lib, B, host_block, channels, ir)

    def rt_post(self, buf, output_volume_db=-30.0):
        b = _planar(buf).copy()
        fn = self.lib.orc_rt_post
        fn.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_float]
        fn(_p(b), b.shape[0], b.shape[1], output_volume_db)
        return b


class OracleRtEngine:
    """Restatement of the plugin's streaming engine (PluginProcessor.cpp:403-562)."""

    def __init__(self, lib, B, host_block, channels, ir):
        self.lib = lib
        ir = np.ascontiguousarray(ir, np.float32).reshape(-1)
        lib.orc_rt_create.restype = ctypes.c_void_p
        self.h = ctypes.c_void_p(lib.orc_rt_create(int(B), int(host_block), int(channels), _p(ir), len(ir)))
        self.channels = channels

    def set_ir(self, ir):
        ir = np.ascontiguousarray(ir, np.float32).reshape(-1)
        self.lib.orc_rt_set_ir(self.h, _p(ir), len(ir))

    def process(self, buf):
        b = _planar(buf).copy()
        assert b.shape[0] == self.channels
        self.lib.orc_rt_process(self.h, _p(b), b.shape[1])
        return b

    def close(self):
        if self.h:
            self.lib.orc_rt_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


class Reference(_Lib):
    """The reference's own object code (oracle/_ref/libirb_ref.so)."""
    prefix = "ref_"
    path = os.path.join(_HERE, "_ref", "libirb_ref.so")

    def _fft_transform(self, x, out):
        self.lib.ref_fft_transform(_p(x), x.shape[0], x.shape[1], 0, _p(out))

    def deconvolve(self, num, den, sample_rate=48000.0, smoothing=True, include_phase=True, include_amplitude=True):
        num, den = _planar(num), _planar(den)
        N = next_pow2(max(num.shape[1], den.shape[1]))
        out = np.zeros(N, np.float32)
        fn = self.lib.ref_deconvolve
        fn.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, _f32p, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                       ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p]
        fn(_p(num), num.shape[0], num.shape[1], _p(den), den.shape[0], den.shape[1], float(sample_rate),
           int(smoothing), int(include_phase), int(include_amplitude), _p(out))
        return out[None, :]

    def invert_filter(self, x, sample_rate=48000):
        x = _planar(x)
        out = np.zeros(next_pow2(x.shape[1]), np.float32)
        self.lib.ref_invert_filter(_p(x), x.shape[0], x.shape[1], int(sample_rate), _p(out))
        return out[None, :]

    def ir_chop(self, x, ir_length, threshold_db, consecutive):
        x = _planar(x)
        out = np.zeros(ir_length, np.float32)
        fn = self.lib.ref_ir_chop
        fn.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int, _f32p]
        fn(_p(x), x.shape[0], x.shape[1], ir_length, threshold_db, consecutive, _p(out))
        return out[None, :]

    def ir_to_real_fft_raw(self, x, part):
        x = _planar(x)[0].copy()
        n = (len(x) // part + 1) * 2 * part
        out = np.zeros(n, np.float32)
        self.lib.ref_ir_to_real_fft_raw(_p(x), len(x), part, _p(out))
        return out

    def ess(self, dur, sr, f1, f2, gain_db=0.0, inverse=False, fade_kind=0, fade_freq=0.0):
        fn = self.lib.ref_ess
        fn.argtypes = [ctypes.c_double] * 5 + [ctypes.c_int, ctypes.c_int, ctypes.c_double, _f64p]
        n = fn(dur, sr, f1, f2, gain_db, int(inverse), fade_kind, fade_freq, None)
        out = np.zeros(n, np.float64)
        fn(dur, sr, f1, f2, gain_db, int(inverse), fade_kind, fade_freq, out.ctypes.data_as(_f64p))
        return out

    def ess_index_at_freq(self, freq, dur, sr, f1, f2):
        fn = self.lib.ref_ess_index_at_freq
        fn.argtypes = [ctypes.c_double] * 5
        return fn(freq, dur, sr, f1, f2)

    def ess_freq_at_index(self, idx, dur, sr, f1, f2):
        fn = self.lib.ref_ess_freq_at_index
        fn.argtypes = [ctypes.c_int] + [ctypes.c_double] * 4
        fn.restype = ctypes.c_double
        return fn(idx, dur, sr, f1, f2)

    def generate_pulse(self, n, offset=0):
        out = np.zeros(n, np.float32)
        self.lib.ref_generate_pulse(n, offset, _p(out))
        return out[None, :]

    def hardware_threads(self):
        return self.lib.ref_hardware_threads()

    def bench_convolve_periodic(self, threads, streams, Lx, h, B, seed):
        h = np.ascontiguousarray(h, np.float32).reshape(-1)
        fn = self.lib.ref_bench_convolve_periodic
        fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, _f64p]
        fn.restype = ctypes.c_double
        cs = ctypes.c_double(0.0)
        secs = fn(threads, streams, Lx, _p(h), len(h), B, seed, ctypes.byref(cs))
        return secs, cs.value

    def bench_deconvolve(self, threads, captures, sweep, sr, smoothing):
        captures = np.ascontiguousarray(captures, np.float32)
        sweep = np.ascontiguousarray(sweep, np.float32).reshape(-1)
        fn = self.lib.ref_bench_deconvolve
        fn.argtypes = [ctypes.c_int, ctypes.c_int, _f32p, ctypes.c_int, _f32p, ctypes.c_double, ctypes.c_int, _f64p]
        fn.restype = ctypes.c_double
        cs = ctypes.c_double(0.0)
        secs = fn(threads, captures.shape[0], _p(captures), captures.shape[1], _p(sweep), sr, int(smoothing), ctypes.byref(cs))
        return secs, cs.value


def have_reference():
    return os.path.exists(Reference.path)
