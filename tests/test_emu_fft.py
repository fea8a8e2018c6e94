"""CPU: the device FFT building blocks (irbaboon_b200/csrc/irb_fft.cuh), compiled for the host by tests/emu,
against numpy's float64 FFT and the oracle's transform.  Guards the Stockham index math, the packed
real split/merge and the per-bin MAC without a GPU."""
import ctypes
import os

import numpy as np
import pytest

from irbaboon_b200 import synth

_f32p = ctypes.POINTER(ctypes.c_float)


@pytest.fixture(scope="module")
def emu(built):
    return ctypes.CDLL(os.path.join(os.path.dirname(__file__), "emu", "libemu_fft.so"))


def _p(a):
    return a.ctypes.data_as(_f32p)


def _packed_ref(x, M):
    X = np.fft.rfft(np.concatenate([x.astype(np.float64), np.zeros(2 * M - len(x))]))
    ref = np.zeros(M, np.complex128)
    ref[1:] = X[1:M]
    ref[0] = X[0].real + 1j * X[M].real
    return ref


@pytest.mark.parametrize("M", [16, 32, 64, 128, 256, 512, 1024, 2048])
@pytest.mark.parametrize("frac", [2.0, 1.0, 0.37])
def test_forward_and_inverse(emu, M, frac):
    ln = max(1, int(M * frac))
    x = synth.white_noise(11, M, ln)
    out = np.zeros(2 * M, np.float32)
    assert emu.emu_real_forward(M, _p(x), ln, _p(out)) == 0
    ref = _packed_ref(x, M)
    got = out[0::2] + 1j * out[1::2]
    assert np.abs(got - ref).max() / np.abs(ref).max() <= 5e-7
    y = np.zeros(2 * M, np.float32)
    assert emu.emu_real_inverse(M, _p(out), _p(y)) == 0
    assert np.abs(y[:ln] - x).max() <= 2e-6
    assert np.abs(y[ln:]).max() <= 2e-6 if ln < 2 * M else True


def test_forward_matches_oracle_transform(emu, orc):
    M = 512
    x = synth.white_noise(12, 0, M)
    out = np.zeros(2 * M, np.float32)
    emu.emu_real_forward(M, _p(x), M, _p(out))
    buf = np.zeros(4 * M, np.float32)
    buf[:M] = x
    want = orc.real_forward(buf, 2 * M)
    assert np.abs(out[2:] - want[2:2 * M]).max() <= 2e-5 * np.abs(want).max()
    assert abs(out[0] - want[0]) <= 1e-4 and abs(out[1] - want[2 * M]) <= 1e-4


def test_block_convolution_through_the_emulated_path(emu):
    """zero-padded FFT -> packed MAC -> inverse == linear convolution of one block with one partition."""
    M = 256
    x, h = synth.white_noise(13, 0, M), synth.white_noise(13, 1, M)
    X, H, acc = (np.zeros(2 * M, np.float32) for _ in range(3))
    emu.emu_real_forward(M, _p(x), M, _p(X))
    emu.emu_real_forward(M, _p(h), M, _p(H))
    emu.emu_mac(M, _p(acc), _p(X), _p(H))
    y = np.zeros(2 * M, np.float32)
    emu.emu_real_inverse(M, _p(acc), _p(y))
    want = np.convolve(x.astype(np.float64), h.astype(np.float64))
    assert np.abs(y[:2 * M - 1] - want).max() <= 1e-5 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("M", [128, 256, 512, 1024, 2048])
def test_exchange_is_bank_conflict_free(emu, M):
    """The XOR swizzle of the Stockham exchange buffer (fft_sw): for rows of >= 16 threads every scatter and gather of
    every pass touches 16 distinct bank pairs per half-warp.  Counted from the addresses the device code itself
    produces (tests/emu), so the kernel and this check cannot drift apart."""
    out = np.full(8, -1.0)
    assert emu.emu_exchange_conflicts(M, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double))) == 0
    used = out[out >= 0]
    assert len(used) >= 4 and np.all(used == 1.0), out
