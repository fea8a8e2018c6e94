"""CPU: the C-ABI library loads without a GPU and exports every symbol include/irb_b200.h declares;
argument validation and the no-CPU-fallback rule are observable without a device."""
import ctypes
import os
import re

import numpy as np
import pytest


def _declared_symbols(header, prefix="irb_"):
    txt = re.sub(r"/\*.*?\*/", "", open(header).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_all_exported(eng):
    names = _declared_symbols(eng.HEADER_PATH)
    assert len(names) >= 20
    L = ctypes.CDLL(eng.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    # and the python binding covers the same surface
    assert sorted(eng._SIGS) == names


def test_measurement_aids_live_outside_the_product_abi(eng):
    """include/irb_b200.h carries no bandwidth probe, no tuning switch and no bare-MAC hook; those are irbx_* entry points of
    include/irb_b200_bench.h (a translation unit of their own), and no library source reads the environment."""
    prod = open(eng.HEADER_PATH).read()
    assert "probe" not in prod.lower() and "irbx_" not in prod and "getenv" not in prod and "mac_only" not in prod
    names = _declared_symbols(eng.BENCH_HEADER_PATH, "irbx_")
    L = ctypes.CDLL(eng.LIB_PATH)
    assert names and not [n for n in names if not hasattr(L, n)]
    assert sorted(eng._BENCH_SIGS) == names
    csrc = os.path.join(os.path.dirname(eng.LIB_PATH), "csrc")
    for f in os.listdir(csrc):
        assert "getenv" not in open(os.path.join(csrc, f)).read(), f
    # the knobs are reachable by name only; unknown names are rejected
    assert eng.get_tuning("mac_persistent") in (0, 1)
    with pytest.raises(eng.IrbError):
        eng.set_tuning("no_such_knob", 1)


def test_library_is_built_for_sm_100a(eng):
    import subprocess
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", eng.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_hot_kernel_uses_bulk_tma_and_128bit_loads(eng):
    """SASS evidence that the MAC kernel stages IR spectra with cp.async.bulk (UBLKCP) and reads the FDL
    with 128-bit loads, and that nothing in the library touches tensor cores or cuFFT."""
    import subprocess
    sass = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", eng.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "SYNCS.ARRIVE.TRANS64" in sass          # cp.async.bulk + mbarrier expect_tx
    assert re.search(r"LDG\.E\.[A-Z0-9.]*256", sass)                    # 32-byte FDL loads (sm_100 LDG.256, L2 evict-first)
    assert re.search(r"LDG\.E\.[A-Z.]*128", sass)                       # the 16-byte form is still there for A/B
    assert "HMMA" not in sass and "UTCHMMA" not in sass
    ldd = subprocess.run(["ldd", eng.LIB_PATH], capture_output=True, text=True).stdout
    assert "cufft" not in ldd.lower()


def test_headline_kernel_streams_fdl_and_ir_through_tma(eng):
    """SASS of k_mac_tma<512> (the block step behind the headline number): the FDL slots and the IR partitions arrive by bulk
    TMA copies (the FDL ones with an L2 cache-policy descriptor), there is no global load wider than the 8-byte twiddle /
    4-byte audio loads of the FFT prologue, and the stage release is predicated on one lane (the arrive whose address
    carries the data dependency on the values read, see mbar_arrive_after)."""
    import subprocess
    sass = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", "-fun", "_ZN3irb9k_mac_tmaILi512EEEvNS_7MacArgsE", eng.LIB_PATH],
                          capture_output=True, text=True).stdout
    assert sass.count("UBLKCP") >= 3 and re.search(r"UBLKCP[^;]*desc\[", sass)
    assert not re.search(r"LDG\.E\.[A-Z0-9.]*(128|256)", sass)
    assert "FENCE.VIEW.ASYNC" in sass                                   # generic -> async proxy hand-over of stage 0's memory
    assert re.search(r"@!?P\d SYNCS\.ARRIVE", sass) and "LDS.128" in sass


def test_persistent_block_step_kernel_sass(eng):
    """SASS of k_mac_p<512> (shared IR) and k_mac_p<1024, PERROW> (per-stream IRs), the kernels behind the headline and the
    configs[3] numbers: FDL slots and IR partitions arrive by bulk TMA copies only (no 16- or 32-byte global load), units come
    from a global atomic counter, the ring stage is released behind a generic->async proxy fence by one predicated arrive,
    the inner loop reads shared memory with LDS.128, and there is no tensor-core instruction."""
    import subprocess
    for fun in ("_ZN3irb7k_mac_pILi512ELb0EEEvNS_7MacArgsE", "_ZN3irb7k_mac_pILi1024ELb1EEEvNS_7MacArgsE"):
        sass = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", "-fun", fun, eng.LIB_PATH], capture_output=True, text=True).stdout
        assert sass.count("UBLKCP") >= 3 and re.search(r"UBLKCP[^;]*desc\[", sass), fun
        assert not re.search(r"LDG\.E\.[A-Z0-9.]*(128|256)", sass), fun
        assert "ATOMG" in sass and "FENCE.VIEW.ASYNC" in sass and "LDS.128" in sass, fun
        assert re.search(r"@!?P\d SYNCS\.ARRIVE", sass), fun
        assert "HMMA" not in sass and "UTCHMMA" not in sass, fun


def test_argument_validation_without_a_device(eng):
    L = eng.lib()
    h = ctypes.c_void_p()
    assert L.irb_engine_create(ctypes.byref(h), 0, 0, 4, 1, 1) == eng.IRB_ERR_ARG
    assert L.irb_engine_create(ctypes.byref(h), 0, 4096, 4, 1, 1) == eng.IRB_ERR_ARG
    assert L.irb_engine_create(ctypes.byref(h), 0, 256, 0, 1, 1) == eng.IRB_ERR_ARG
    assert b"block_size" in L.irb_last_error() or b"must be" in L.irb_last_error()
    assert L.irb_max_block_size() == 2048 and L.irb_version() >= 100
    x = np.zeros((3, 10), np.float32)
    out = eng.convolve_periodic(x, np.ones(4, np.float32), 16)     # rejected layout: cleared buffer, like the reference
    assert out.shape == (3, 13) and not out.any()


def test_no_cpu_fallback(eng):
    """Without a GPU every compute entry point must fail loudly rather than compute on the host."""
    try:
        n = eng.device_count()
    except eng.IrbError:
        n = 0
    if n > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(eng.IrbError):
        eng.convolve_periodic(np.ones(100, np.float32), np.ones(10, np.float32), 16)
    with pytest.raises(eng.IrbError):
        eng.Engine(256, 8, 2)


def test_product_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "irbaboon_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "irb_oracle" not in txt and "libirb_ref" not in txt, f
