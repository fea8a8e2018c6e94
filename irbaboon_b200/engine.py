"""ctypes binding of libirb_b200.so (include/irb_b200.h) -- the host-side driver used by tests and bench.py.

This module is plumbing: it loads the C-ABI library, turns its status codes into exceptions and moves numpy
arrays (or raw device pointers from torch) across the boundary.  All arithmetic happens in the CUDA library;
there is no CPU fallback here -- if the library is missing, or no B200 is visible, calls raise.
"""
import ctypes
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IRB_LIB") or os.path.join(_HERE, "libirb_b200.so")      # IRB_LIB: another build of the library (A/B measurements)
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "irb_b200.h")
BENCH_HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "irb_b200_bench.h")      # measurement aids, not the drop-in boundary

_f32p = ctypes.POINTER(ctypes.c_float)
_vp = ctypes.c_void_p
_lib = None


class IrbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("irb status %d: %s" % (code, msg))
        self.code = code


IRB_ERR_ARG, IRB_ERR_LAYOUT, IRB_ERR_CUDA, IRB_ERR_STATE = -1, -2, -3, -4

_SIGS = {
    "irb_last_error": (ctypes.c_char_p, []),
    "irb_version": (ctypes.c_int, []),
    "irb_device_count": (ctypes.c_int, []),
    "irb_set_device": (ctypes.c_int, [ctypes.c_int]),
    "irb_max_block_size": (ctypes.c_int, []),
    "irb_host_alloc": (_vp, [ctypes.c_size_t]),
    "irb_host_alloc_write_combined": (_vp, [ctypes.c_size_t]),
    "irb_host_free": (None, [_vp]),
    "irb_engine_create": (ctypes.c_int, [ctypes.POINTER(_vp)] + [ctypes.c_int] * 5),
    "irb_engine_destroy": (ctypes.c_int, [_vp]),
    "irb_engine_set_stream": (ctypes.c_int, [_vp, _vp]),
    "irb_engine_set_ir": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, ctypes.c_int]),
    "irb_engine_stage_ir": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, ctypes.c_int, ctypes.c_int]),
    "irb_engine_mac_plan": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "irb_engine_set_mac_split": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int]),
    "irb_engine_set_fused_step": (ctypes.c_int, [_vp, ctypes.c_int]),
    "irb_engine_bind": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "irb_engine_tile_channels": (ctypes.c_int, [_vp]),
    "irb_engine_reset": (ctypes.c_int, [_vp]),
    "irb_engine_set_active_channels": (ctypes.c_int, [_vp, ctypes.c_int]),
    "irb_engine_active_channels": (ctypes.c_int, [_vp]),
    "irb_engine_process": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int]),
    "irb_engine_process_callback": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int]),
    "irb_engine_process_device": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int]),
    "irb_engine_synchronize": (ctypes.c_int, [_vp]),
    "irb_engine_submit": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int]),
    "irb_engine_wait": (ctypes.c_int, [_vp]),
    "irb_engine_set_timing": (ctypes.c_int, [_vp, ctypes.c_int]),
    "irb_engine_get_timings": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int]),
    "irb_engine_state_bytes": (ctypes.c_size_t, [_vp]),
    "irb_engine_fft_size": (ctypes.c_int, [_vp]),
    "irb_engine_partitions": (ctypes.c_int, [_vp, ctypes.c_int]),
    "irb_engine_read_ir_spectrum": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp]),
    "irb_engine_read_fdl_spectrum": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp]),
    "irb_engine_launch_count": (ctypes.c_longlong, [_vp]),
    "irb_launch_count": (ctypes.c_longlong, []),
    "irb_release_workspace": (ctypes.c_size_t, []),
    "irb_last_compute_ms": (ctypes.c_double, []),
    "irb_group_create": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.POINTER(ctypes.c_int)] + [ctypes.c_int] * 5),
    "irb_group_destroy": (ctypes.c_int, [_vp]),
    "irb_group_device_count": (ctypes.c_int, [_vp]),
    "irb_group_channel_range": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "irb_group_set_ir": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, ctypes.c_int]),
    "irb_group_stage_ir": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, ctypes.c_int, ctypes.c_int]),
    "irb_group_bind": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "irb_group_reset": (ctypes.c_int, [_vp]),
    "irb_group_process": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int]),
    "irb_group_state_bytes": (ctypes.c_size_t, [_vp]),
    "irb_convolve_periodic": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp]),
    "irb_convolve_nonperiodic": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp, ctypes.c_int, ctypes.c_int, _vp]),
    "irb_deconvolve": (ctypes.c_int, [_vp, ctypes.c_int, _vp, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp]),
    "irb_deconvolve_batch": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp]),
    "irb_deconvolve_batch_device": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp]),
    "irb_invert_filter": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp]),
    "irb_averaging_filter": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "irb_fft_transform": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp]),
    "irb_fft_inv_transform": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp]),
    "irb_ir_to_real_fft_raw": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp]),
    "irb_ess_generate": (ctypes.c_int, [ctypes.c_double] * 5 + [ctypes.c_int, _vp, ctypes.c_int]),
}

# include/irb_b200_bench.h: bench-only entry points (irbaboon_b200/csrc/irb_benchaids.cu)
_BENCH_SIGS = {
    "irbx_set_tuning": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int]),
    "irbx_get_tuning": (ctypes.c_int, [ctypes.c_char_p]),
    "irbx_engine_mac_only_device": (ctypes.c_int, [_vp, _vp]),
    "irbx_engine_set_stamps": (ctypes.c_int, [_vp, _vp]),
    "irbx_hbm_read_probe": (ctypes.c_int, [ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]),
    "irbx_copy_probe_create": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int, ctypes.c_size_t, ctypes.c_int]),
    "irbx_copy_probe_run": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.POINTER(ctypes.c_double)]),
    "irbx_copy_probe_destroy": (ctypes.c_int, [_vp]),
}


def lib():
    """Load libirb_b200.so; raises if it has not been built (python -m irbaboon_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(LIB_PATH + " is missing: run `python -m irbaboon_b200.build` (there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in list(_SIGS.items()) + list(_BENCH_SIGS.items()):
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _ck(rc):
    if rc < 0:
        raise IrbError(rc, lib().irb_last_error().decode("utf-8", "replace"))
    return rc


def _ptr(a):
    return a.ctypes.data_as(_vp)


def _planar(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a[None, :] if a.ndim == 1 else a


def device_count():
    return _ck(lib().irb_device_count())


def set_device(device):
    _ck(lib().irb_set_device(int(device)))


def launch_count():
    return lib().irb_launch_count()


def release_workspace():
    """Return the offline functions' recycled device scratch buffers to the driver; -> bytes released."""
    return lib().irb_release_workspace()


def last_compute_ms():
    return lib().irb_last_compute_ms()


def hbm_read_probe(nbytes, iters=5, write_every=0, store_kind=0):
    """GB/s of a read-only streaming kernel over nbytes of device memory (measurement aid, irb_b200_bench.h)."""
    g = ctypes.c_double(0.0)
    _ck(lib().irbx_hbm_read_probe(int(nbytes), int(iters), int(write_every), int(store_kind), ctypes.byref(g)))
    return g.value


def set_tuning(name, value):
    """Launch-policy knob of the library by name (irb_b200_bench.h; A/B measurements and tests only)."""
    _ck(lib().irbx_set_tuning(name.encode(), int(value)))


def get_tuning(name):
    return _ck(lib().irbx_get_tuning(name.encode()))


class CopyProbe:
    """Copy-only host<->device ceiling (irb_b200_bench.h): `nbytes` per direction, no kernels."""

    def __init__(self, nbytes, host_mode=0, device=0):
        self._h = _vp()
        _ck(lib().irbx_copy_probe_create(ctypes.byref(self._h), int(device), int(nbytes), int(host_mode)))
        self.nbytes = int(nbytes)

    def run(self, iters=8, direction=3, chunk_bytes=0):
        """-> wall seconds of `iters` rounds (direction 1: H2D, 2: D2H, 3: both at once)."""
        s = ctypes.c_double(0.0)
        _ck(lib().irbx_copy_probe_run(self._h, int(iters), int(direction), int(chunk_bytes), ctypes.byref(s)))
        return s.value

    def close(self):
        if self._h:
            lib().irbx_copy_probe_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _PinnedOwner:
    """Owns one cudaMallocHost block: the block is released when the last numpy view of it is collected, or by pinned_free."""

    def __init__(self, ptr):
        self.ptr = ptr
        self._fin = weakref.finalize(self, lib().irb_host_free, ptr)

    def free(self):
        self._fin()


def pinned_empty(shape, dtype=np.float32, write_combined=False):
    """numpy array over cudaMallocHost memory (write_combined: for input the host only writes; never read it back from the CPU).
    The memory goes back when the array (and every view of it) is collected; pinned_free(arr) releases it at once (the array
    must not be touched afterwards)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = lib().irb_host_alloc_write_combined(n) if write_combined else lib().irb_host_alloc(n)
    if not p:
        raise IrbError(IRB_ERR_CUDA, lib().irb_last_error().decode())
    owner = _PinnedOwner(p)
    buf = (ctypes.c_char * max(n, 1)).from_address(p)
    buf._irb_owner = owner                      # the ctypes buffer is the base of every numpy view: it keeps the owner alive
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[p] = weakref.ref(owner)
    return arr


_PINNED = {}


def pinned_free(arr):
    ref = _PINNED.pop(arr.__array_interface__["data"][0], None)
    owner = ref() if ref else None
    if owner is not None:
        owner.free()


def convolve_periodic(x, h, block_size=256):
    """fp::convolution::convolvePeriodic (fp/convolution.cpp:14-242): planar [ch][n] float32 in and out."""
    x, h = _planar(x), _planar(h)
    out = np.zeros((x.shape[0], x.shape[1] + h.shape[1] - 1), np.float32)
    rc = lib().irb_convolve_periodic(_ptr(x), x.shape[0], x.shape[1], _ptr(h), h.shape[0], h.shape[1], int(block_size), _ptr(out))
    if rc == IRB_ERR_LAYOUT:
        return out                      # the reference prints a DBG line and returns the cleared buffer
    _ck(rc)
    return out


def next_pow2(x):
    """tools::nextPowerOfTwo (fp/tools.cpp:189-196): x itself when already a power of two."""
    if x > 0 and (x & (x - 1)) == 0:
        return x
    r = 1
    while r <= x:
        r *= 2
    return r


def convolve_nonperiodic(x, h):
    """fp::convolution::convolveNonPeriodic (fp/convolution.cpp:246-347)."""
    x, h = _planar(x), _planar(h)
    out = np.zeros((x.shape[0], x.shape[1] + h.shape[1] - 1), np.float32)
    rc = lib().irb_convolve_nonperiodic(_ptr(x), x.shape[0], x.shape[1], _ptr(h), h.shape[0], h.shape[1], _ptr(out))
    if rc == IRB_ERR_LAYOUT:
        return np.zeros_like(x)         # the reference returns a cleared copy of the input (fp/convolution.cpp:271-275)
    _ck(rc)
    return out


def deconvolve(num, den, sample_rate=48000.0, smoothing=True, include_phase=True, include_amplitude=True):
    """fp::convolution::deconvolve (fp/convolution.cpp:351-403); channel 0 of each input is used.  -> [1][N]"""
    num, den = np.ascontiguousarray(_planar(num)[0]), np.ascontiguousarray(_planar(den)[0])
    N = next_pow2(max(len(num), len(den)))
    out = np.zeros((1, N), np.float32)
    _ck(lib().irb_deconvolve(_ptr(num), len(num), _ptr(den), len(den), float(sample_rate), int(smoothing), int(include_phase), int(include_amplitude), _ptr(out)))
    return out


def deconvolve_batch(nums, den, sample_rate=48000.0, smoothing=False, include_phase=True, include_amplitude=True, out=None):
    """`batch` captures nums[batch][len] divided by one sweep den[len_den] -> [batch][N]."""
    nums = np.ascontiguousarray(nums, np.float32)
    den = np.ascontiguousarray(den, np.float32).reshape(-1)
    N = next_pow2(max(nums.shape[1], len(den)))
    if out is None:
        out = np.zeros((nums.shape[0], N), np.float32)
    assert out.shape == (nums.shape[0], N) and out.dtype == np.float32 and out.flags.c_contiguous
    _ck(lib().irb_deconvolve_batch(_ptr(nums), nums.shape[0], nums.shape[1], _ptr(den), len(den), float(sample_rate), int(smoothing), int(include_phase),
                                   int(include_amplitude), _ptr(out)))
    return out


def deconvolve_batch_device(nums_ptr, batch, len_num, den, out_ptr, sample_rate=48000.0, smoothing=False, include_phase=True, include_amplitude=True):
    """irb_deconvolve_batch_device: captures and results are raw DEVICE pointers ([batch][len_num] / [batch][N] float32)."""
    den = np.ascontiguousarray(den, np.float32).reshape(-1)
    _ck(lib().irb_deconvolve_batch_device(_vp(int(nums_ptr)), int(batch), int(len_num), _ptr(den), len(den), float(sample_rate), int(smoothing), int(include_phase),
                                          int(include_amplitude), _vp(int(out_ptr))))


def invert_filter(x, sample_rate=48000):
    x = np.ascontiguousarray(_planar(x)[0])
    out = np.zeros((1, next_pow2(len(x))), np.float32)
    _ck(lib().irb_invert_filter(_ptr(x), len(x), int(sample_rate), _ptr(out)))
    return out


def averaging_filter(spec, octave_fraction, sample_rate, log_avg=True, include_phase=True, include_amplitude=True):
    s = _planar(spec).copy()
    _ck(lib().irb_averaging_filter(_ptr(s), s.shape[0], s.shape[1], float(octave_fraction), float(sample_rate), int(log_avg), int(include_phase), int(include_amplitude)))
    return s


def fft_transform(x, format_ampl_phase=False):
    x = _planar(x)
    N = next_pow2(x.shape[1])
    out = np.zeros((x.shape[0], 2 * N), np.float32)
    _ck(lib().irb_fft_transform(_ptr(x), x.shape[0], x.shape[1], int(format_ampl_phase), _ptr(out)))
    return out


def fft_inv_transform(spec):
    s = _planar(spec)
    out = np.zeros((s.shape[0], s.shape[1] // 2), np.float32)
    _ck(lib().irb_fft_inv_transform(_ptr(s), s.shape[0], s.shape[1], _ptr(out)))
    return out


def ir_to_real_fft_raw(x, part_size):
    """fp::ir::IRtoRealFFTRaw (fp/ir.cpp:106-147): packed partition spectra, (len/part + 1) * 2*part floats."""
    x = np.ascontiguousarray(_planar(x)[0])
    out = np.zeros((len(x) // part_size + 1) * 2 * part_size, np.float32)
    _ck(lib().irb_ir_to_real_fft_raw(_ptr(x), len(x), int(part_size), _ptr(out)))
    return out


def ess(duration_s, sample_rate, f1, f2, gain_db=0.0, inverse=False):
    """fp::ExpSineSweep::generate / generateInv in FP64 on the device -> float64 array."""
    n = _ck(lib().irb_ess_generate(duration_s, sample_rate, f1, f2, gain_db, int(inverse), None, 0))
    out = np.zeros(n, np.float64)
    _ck(lib().irb_ess_generate(duration_s, sample_rate, f1, f2, gain_db, int(inverse), _ptr(out), n))
    return out


class Engine:
    """Streaming UPOLA engine over n_channels independent stream-channels on one GPU."""

    def __init__(self, block_size, max_partitions, n_channels, n_irs=1, device=0):
        self._h = _vp()
        _ck(lib().irb_engine_create(ctypes.byref(self._h), int(device), int(block_size), int(max_partitions), int(n_channels), int(n_irs)))
        self.block_size, self.max_partitions, self.n_channels, self.n_irs, self.device = block_size, max_partitions, n_channels, n_irs, device
        self.fft_size = lib().irb_engine_fft_size(self._h)
        self.active_channels = n_channels

    def close(self):
        if self._h:
            lib().irb_engine_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream):
        """cuda_stream: a cudaStream_t handle as int (0 = legacy default stream); None = the engine's own stream."""
        h = _vp(ctypes.c_size_t(-1).value) if cuda_stream is None else _vp(int(cuda_stream))
        _ck(lib().irb_engine_set_stream(self._h, h))

    def set_ir(self, ir_id, taps):
        """taps: [n] mono, or [2][n] stereo folded to (L+R)/2."""
        t = _planar(taps)
        if t.shape[0] == 1:
            _ck(lib().irb_engine_set_ir(self._h, int(ir_id), _ptr(t), None, t.shape[1]))
        elif t.shape[0] == 2:
            l, r = np.ascontiguousarray(t[0]), np.ascontiguousarray(t[1])
            _ck(lib().irb_engine_set_ir(self._h, int(ir_id), _ptr(l), _ptr(r), t.shape[1]))
        else:
            raise ValueError("IR must be mono or stereo")

    def stage_ir(self, ir_id, taps, n_partitions=0):
        """Round-robin IR switch (PluginProcessor.cpp:411-414,455-461): one partition re-transformed per block step."""
        t = _planar(taps)
        if t.shape[0] == 1:
            _ck(lib().irb_engine_stage_ir(self._h, int(ir_id), _ptr(t), None, t.shape[1], int(n_partitions)))
        elif t.shape[0] == 2:
            l, r = np.ascontiguousarray(t[0]), np.ascontiguousarray(t[1])
            _ck(lib().irb_engine_stage_ir(self._h, int(ir_id), _ptr(l), _ptr(r), t.shape[1], int(n_partitions)))
        else:
            raise ValueError("IR must be mono or stereo")

    def mac_plan(self):
        """(slots_kernel, split_in, cluster) of the next block step's MAC launch."""
        a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _ck(lib().irb_engine_mac_plan(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return bool(a.value), b.value, c.value

    def set_mac_split(self, split_in=0, cluster=0):
        """Force how few-row launches split a row's partitions (0, 0 = automatic)."""
        _ck(lib().irb_engine_set_mac_split(self._h, int(split_in), int(cluster)))

    def set_fused_step(self, enable=True):
        """One launch per block step (forward transform inside the MAC kernel) or the two-launch form."""
        _ck(lib().irb_engine_set_fused_step(self._h, int(bool(enable))))

    def bind(self, chan_begin, chan_end, ir_id):
        _ck(lib().irb_engine_bind(self._h, int(chan_begin), int(chan_end), int(ir_id)))

    @property
    def tile_channels(self):
        return lib().irb_engine_tile_channels(self._h)

    def reset(self):
        _ck(lib().irb_engine_reset(self._h))

    def process(self, x, out=None):
        """x: host [n_blocks][n_channels][B] (or [n_channels][B]); returns the same shape."""
        x = np.ascontiguousarray(x, np.float32)
        shp = x.shape
        nb = 1 if x.ndim == 2 else shp[0]
        if x.ndim not in (2, 3) or shp[-2:] != (self.active_channels, self.block_size):
            raise ValueError("input shape %r, expected [n_blocks][%d][%d]" % (shp, self.active_channels, self.block_size))
        if out is None:
            out = np.empty(shp, np.float32)
        elif not (isinstance(out, np.ndarray) and out.dtype == np.float32 and out.shape == shp and out.flags.c_contiguous and out.flags.writeable):
            raise ValueError("out must be a writable C-contiguous float32 array of shape %r" % (shp,))      # its raw pointer goes to the library
        _ck(lib().irb_engine_process(self._h, _ptr(x), _ptr(out), nb))
        return out

    def process_callback(self, x):
        """Blocks of one plug-in callback, reference order (all forward FFTs first): host [n_blocks][n_channels][B]."""
        x = np.ascontiguousarray(x, np.float32)
        assert x.ndim == 3 and x.shape[1:] == (self.active_channels, self.block_size), x.shape
        out = np.empty(x.shape, np.float32)
        _ck(lib().irb_engine_process_callback(self._h, _ptr(x), _ptr(out), x.shape[0]))
        return out

    def submit(self, x, out):
        """Asynchronous multi-block call on pinned host arrays [n_blocks >= 2][n_channels][B]; pair with wait()."""
        assert x.dtype == np.float32 and out.dtype == np.float32 and x.flags.c_contiguous and out.flags.c_contiguous and x.shape == out.shape
        assert x.ndim == 3 and x.shape[1:] == (self.active_channels, self.block_size), x.shape
        _ck(lib().irb_engine_submit(self._h, _ptr(x), _ptr(out), x.shape[0]))

    def wait(self):
        _ck(lib().irb_engine_wait(self._h))

    def process_device(self, in_ptr, out_ptr, n_blocks=1):
        _ck(lib().irb_engine_process_device(self._h, _vp(int(in_ptr)), _vp(int(out_ptr)), int(n_blocks)))

    def mac_only_device(self, acc_ptr):
        """bench-only (irb_b200_bench.h): the bare FDL multiply-accumulate into a device buffer"""
        _ck(lib().irbx_engine_mac_only_device(self._h, _vp(int(acc_ptr))))

    def set_stamps(self, dev_ptr):
        """bench-only: device array of 16 uint64 for the phase time stamps of the latency-path kernel (0 / None: off)"""
        _ck(lib().irbx_engine_set_stamps(self._h, _vp(int(dev_ptr or 0))))

    def set_active_channels(self, n_active):
        """Only channels [0, n_active) take part in the following block steps; I/O arrays are then [n_blocks][n_active][B]."""
        _ck(lib().irb_engine_set_active_channels(self._h, int(n_active)))
        self.active_channels = int(n_active)

    def set_timing(self, enable=True):
        _ck(lib().irb_engine_set_timing(self._h, int(bool(enable))))

    def timings(self, max_steps=16384):
        """(step_ms, mac_ms) device times of the steps recorded since set_timing(True)."""
        a = np.zeros(max_steps, np.float32)
        b = np.zeros(max_steps, np.float32)
        n = _ck(lib().irb_engine_get_timings(self._h, _ptr(a), _ptr(b), max_steps))
        return a[:n].copy(), b[:n].copy()

    def synchronize(self):
        _ck(lib().irb_engine_synchronize(self._h))

    def process_stream(self, x):
        """Convenience for tests: planar audio x[n_channels][n] -> y[n_channels][n] block by block
        (n is padded up to whole blocks with zeros and cut back)."""
        x = _planar(x)
        B = self.block_size
        n = x.shape[1]
        nb = (n + B - 1) // B
        xp = np.zeros((self.n_channels, nb * B), np.float32)
        xp[:, :n] = x
        blocks = np.ascontiguousarray(xp.reshape(self.n_channels, nb, B).transpose(1, 0, 2))
        y = self.process(blocks)
        return np.ascontiguousarray(y.transpose(1, 0, 2)).reshape(self.n_channels, nb * B)[:, :n]

    @property
    def state_bytes(self):
        return lib().irb_engine_state_bytes(self._h)

    def partitions(self, ir_id=0):
        return _ck(lib().irb_engine_partitions(self._h, int(ir_id)))

    @property
    def launches(self):
        return lib().irb_engine_launch_count(self._h)

    def ir_spectrum(self, ir_id, part):
        out = np.zeros(self.fft_size, np.float32)
        _ck(lib().irb_engine_read_ir_spectrum(self._h, int(ir_id), int(part), _ptr(out)))
        return out

    def fdl_spectrum(self, chan, age=0):
        out = np.zeros(self.fft_size, np.float32)
        _ck(lib().irb_engine_read_fdl_spectrum(self._h, int(chan), int(age), _ptr(out)))
        return out


class Group:
    """Several GPUs driven from one process: channels sharded by contiguous tile-aligned ranges (irb_group_*).
    n_irs > 0: shared IRs replicated on every device; n_irs == 0: one private IR per channel, living on its device."""

    def __init__(self, devices, block_size, max_partitions, n_channels, n_irs=1):
        self._h = _vp()
        devs = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
        _ck(lib().irb_group_create(ctypes.byref(self._h), devs, len(devices), int(block_size), int(max_partitions), int(n_channels), int(n_irs)))
        self.block_size, self.n_channels, self.n_irs = block_size, n_channels, n_irs

    def close(self):
        if self._h:
            lib().irb_group_destroy(self._h)
            self._h = _vp()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def ranges(self):
        out = []
        for i in range(_ck(lib().irb_group_device_count(self._h))):
            b, e = ctypes.c_int(), ctypes.c_int()
            _ck(lib().irb_group_channel_range(self._h, i, ctypes.byref(b), ctypes.byref(e)))
            out.append((b.value, e.value))
        return out

    def set_ir(self, ir_id, taps):
        t = np.ascontiguousarray(_planar(taps)[0])
        _ck(lib().irb_group_set_ir(self._h, int(ir_id), _ptr(t), None, len(t)))

    def stage_ir(self, ir_id, taps, n_partitions=0):
        t = np.ascontiguousarray(_planar(taps)[0])
        _ck(lib().irb_group_stage_ir(self._h, int(ir_id), _ptr(t), None, len(t), int(n_partitions)))

    def bind(self, chan_begin, chan_end, ir_id):
        _ck(lib().irb_group_bind(self._h, int(chan_begin), int(chan_end), int(ir_id)))

    def reset(self):
        _ck(lib().irb_group_reset(self._h))

    def process(self, x, out=None):
        x = np.ascontiguousarray(x, np.float32)
        nb = 1 if x.ndim == 2 else x.shape[0]
        assert x.shape[-2:] == (self.n_channels, self.block_size), x.shape
        if out is None:
            out = np.empty(x.shape, np.float32)
        elif not (isinstance(out, np.ndarray) and out.dtype == np.float32 and out.shape == x.shape and out.flags.c_contiguous and out.flags.writeable):
            raise ValueError("out must be a writable C-contiguous float32 array of shape %r" % (x.shape,))
        _ck(lib().irb_group_process(self._h, _ptr(x), _ptr(out), nb))
        return out

    @property
    def state_bytes(self):
        return lib().irb_group_state_bytes(self._h)


def unpack_spectrum(packed):
    """packed M complex (bin 0 = {Re X[0], Re X[M]}) -> M+1 complex bins."""
    p = np.asarray(packed, np.float32)
    M = p.size // 2
    z = (p[0::2] + 1j * p[1::2]).astype(np.complex64)
    out = np.zeros(M + 1, np.complex64)
    out[1:M] = z[1:]
    out[0] = z[0].real
    out[M] = z[0].imag
    return out
