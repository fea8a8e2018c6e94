"""GPU: error behaviour at the C ABI -- status codes instead of exceptions or crashes, the engine stays usable after a
rejected call, and the reference's "debug line + cleared buffer" cases (fp/convolution.cpp:39-42,271-275,412-415)."""
import ctypes

import numpy as np
import pytest

from irbaboon_b200 import synth

pytestmark = pytest.mark.gpu
_vp = ctypes.c_void_p


def _p(a):
    return a.ctypes.data_as(_vp)


def test_engine_rejects_bad_arguments_and_keeps_working(eng):
    L = eng.lib()
    B, P, C = 64, 4, 3
    x = np.stack([synth.white_noise(1002, c, 8 * B) for c in range(C)])
    h = synth.decaying_ir(2000, P * B)
    with eng.Engine(B, P, C, 2) as e:
        e.set_ir(0, h)
        want = e.process_stream(x)
        e.reset()
        hh = e._h
        buf = np.zeros((1, C, B), np.float32)
        assert L.irb_engine_set_ir(hh, 2, _p(h), None, len(h)) == eng.IRB_ERR_ARG            # ir_id out of range
        assert L.irb_engine_set_ir(hh, 0, _p(h), None, P * B + 1) == eng.IRB_ERR_ARG          # more taps than max_partitions holds
        assert L.irb_engine_set_ir(hh, 0, None, None, 10) == eng.IRB_ERR_ARG
        assert L.irb_engine_stage_ir(hh, 1, _p(h), None, len(h), P + 1) == eng.IRB_ERR_ARG    # more partitions than the ring
        assert L.irb_engine_stage_ir(hh, 1, _p(h), None, 0, 0) == eng.IRB_ERR_ARG
        assert L.irb_engine_bind(hh, 0, C + 1, 0) == eng.IRB_ERR_ARG
        assert L.irb_engine_bind(hh, 2, 1, 0) == eng.IRB_ERR_ARG
        assert L.irb_engine_bind(hh, 0, C, 5) == eng.IRB_ERR_ARG
        assert L.irb_engine_process(hh, None, _p(buf), 1) == eng.IRB_ERR_ARG
        assert L.irb_engine_process(hh, _p(buf), _p(buf), -1) == eng.IRB_ERR_ARG
        assert L.irb_engine_process_callback(hh, _p(buf), _p(buf), P + 1) == eng.IRB_ERR_ARG   # more blocks than FDL slots
        assert L.irb_engine_set_mac_split(hh, 2, 0) == eng.IRB_ERR_ARG
        assert L.irb_engine_read_ir_spectrum(hh, 0, P, _p(buf)) == eng.IRB_ERR_ARG
        assert L.irb_engine_process(hh, _p(buf), _p(buf), 0) == 0                             # zero blocks: nothing to do
        assert b"" != L.irb_last_error()
        assert np.array_equal(e.process_stream(x), want)                                      # state untouched by the rejected calls
    for fn in ("irb_engine_reset", "irb_engine_synchronize"):
        assert getattr(L, fn)(None) == eng.IRB_ERR_ARG
    assert L.irb_engine_destroy(None) == 0


def test_offline_functions_reject_bad_arguments(eng):
    L = eng.lib()
    x = np.ones(100, np.float32)
    out = np.zeros(4096, np.float32)
    assert L.irb_convolve_periodic(_p(x), 1, 0, _p(x), 1, 10, 16, _p(out)) == eng.IRB_ERR_ARG       # empty audio
    assert L.irb_convolve_periodic(_p(x), 1, 100, _p(x), 1, 10, 0, _p(out)) == eng.IRB_ERR_ARG
    assert L.irb_convolve_periodic(None, 1, 100, _p(x), 1, 10, 16, _p(out)) == eng.IRB_ERR_ARG
    assert L.irb_convolve_nonperiodic(_p(x), 1, 100, _p(x), 1, 0, _p(out)) == eng.IRB_ERR_ARG
    assert L.irb_deconvolve(_p(x), 5, _p(x), 5, ctypes.c_double(48000.0), 0, 1, 1, _p(out)) == eng.IRB_ERR_ARG     # N = 8 < 16
    assert L.irb_deconvolve_batch(_p(x), 0, 100, _p(x), 100, ctypes.c_double(48000.0), 0, 1, 1, _p(out)) == eng.IRB_ERR_ARG
    assert L.irb_fft_inv_transform(_p(x), 1, 96, _p(out)) == eng.IRB_ERR_ARG                        # N = 48 is not a power of two
    assert L.irb_ir_to_real_fft_raw(_p(x), 100, 48, _p(out)) == eng.IRB_ERR_ARG
    assert L.irb_ess_generate(ctypes.c_double(1.0), ctypes.c_double(48000.0), ctypes.c_double(100.0), ctypes.c_double(50.0),
                              ctypes.c_double(0.0), 0, None, 0) == eng.IRB_ERR_ARG                  # f2 < f1
    assert L.irb_averaging_filter(_p(x), 1, 64, ctypes.c_double(-1.0), ctypes.c_double(48000.0), 1, 1, 1) == eng.IRB_ERR_ARG
    # the reference's silent cases stay silent: unsupported layout -> cleared output; non-power-of-two spectrum -> untouched
    x3 = np.ones((3, 40), np.float32)
    y = np.ones((3, 49), np.float32)
    assert L.irb_convolve_periodic(_p(x3), 3, 40, _p(x), 1, 10, 16, _p(y)) == eng.IRB_ERR_LAYOUT and not y.any()
    odd = np.ones(100, np.float32)
    assert L.irb_averaging_filter(_p(odd), 1, 100, ctypes.c_double(0.1), ctypes.c_double(48000.0), 1, 1, 1) == 0 and (odd == 1).all()


def test_many_engines_and_repeated_create_destroy_do_not_leak(eng):
    import torch
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(20):
        with eng.Engine(256, 64, 512, 4) as e:
            e.set_ir(0, synth.decaying_ir(2000, 5000))
            e.stage_ir(1, synth.decaying_ir(2001, 4000))
            e.process(np.zeros((1, 512, 256), np.float32))
    es = [eng.Engine(128, 8, 16, 1) for _ in range(8)]
    for i, e in enumerate(es):
        e.set_ir(0, synth.decaying_ir(2000 + i, 700, i))
    x = synth.white_noise(1002, 0, 128 * 16).reshape(16, 1, 128).repeat(16, axis=1)
    outs = [e.process(np.ascontiguousarray(x)) for e in es]
    assert not np.array_equal(outs[0], outs[1])                       # engines are independent
    for e in es:
        e.close()
    torch.cuda.synchronize()
    assert free0 - torch.cuda.mem_get_info()[0] < 64 << 20            # twiddle tables stay cached, nothing else


def test_offline_scratch_pool_recycles_and_releases(eng):
    import torch
    x = synth.white_noise(1001, 0, 1 << 16)
    h = synth.decaying_ir(2000, 1 << 14)
    eng.release_workspace()
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    a = eng.convolve_nonperiodic(x, h)
    b = eng.convolve_nonperiodic(x, h)                       # second call runs out of recycled blocks
    c = eng.deconvolve(a[0, :1 << 16], x, 48000.0, True)
    assert np.array_equal(a, b) and c.shape == (1, 1 << 16)
    held = free0 - torch.cuda.mem_get_info()[0]
    assert held > 1 << 20                                    # the pool keeps the scratch of the calls above
    released = eng.release_workspace()
    assert released > 1 << 20 and eng.release_workspace() == 0
    assert free0 - torch.cuda.mem_get_info()[0] < 16 << 20   # back to where it started (tables stay cached)
