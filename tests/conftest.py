import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Everything compiled (CUDA library cross-compiles without a GPU)."""
    import __graft_entry__ as ge

    ge.build()
    return True


@pytest.fixture(scope="session")
def orc(built):
    import oracle

    return oracle.Oracle()


@pytest.fixture(scope="session")
def ref(built):
    import oracle

    if not oracle.have_reference():
        pytest.skip("oracle/_ref/libirb_ref.so not built (no /root/reference here)")
    return oracle.Reference()


@pytest.fixture(scope="session")
def eng(built):
    from irbaboon_b200 import engine

    engine.lib()
    return engine


def parity(got, want):
    """(max-abs error / full scale, relative L2) with full scale := max(1, max|want|)  (SURVEY 8d)."""
    import numpy as np

    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    fs = max(1.0, float(np.abs(want).max())) if want.size else 1.0
    den = float(np.sqrt((want * want).sum()))
    l2 = float(np.sqrt(((got - want) ** 2).sum())) / den if den > 0 else float(np.sqrt(((got - want) ** 2).sum()))
    return (float(np.abs(got - want).max()) / fs if want.size else 0.0), l2


TOL = 1e-5          # north_star: max-abs <= 1e-5 of full scale and relative L2 <= 1e-5, FP32


def conditioned_bound(fn, inputs, want, floor_abs=TOL, floor_l2=TOL, rel=6e-8, seed=0):
    """Tolerance for an ill-conditioned float32 computation: (max-abs/FS, relative L2) by which the REFERENCE's own result
    `want = fn(*inputs)` moves when every input sample changes by half a float32 ulp (seeded signs), doubled, and never below
    the north-star tolerance.  Spectral division by a sweep whose spectrum falls to 1e-4 of its peak amplifies rounding that
    much: at N = 2^20 the reference itself sits 2.2e-5 (relative L2) from the float64 result."""
    import numpy as np

    rng = np.random.default_rng(seed)
    pert = [(np.asarray(v) * (1 + rel * rng.choice([-1.0, 1.0], np.asarray(v).shape))).astype(np.float32) for v in inputs]
    e, l2 = parity(fn(*pert), want)
    return max(floor_abs, 2 * e), max(floor_l2, 2 * l2)
