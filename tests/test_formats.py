"""Wire / on-disk formats and the capture -> filter chain (N3) of the C++ facade.

CPU: the WAV writer is checked against Python's own `wave` reader and the byte layout of RIFF/WAVE, the ".sweepandir"
container (2-channel 24-bit WAV: recorded sweep, IR) round-trips, the TSV spectrum dump parses back to the spectrum.
GPU: consolidate -> deconvolve -> createIRFilt -> IRchop -> normalize against the same chain built from the oracle's
(reference's) functions."""
import ctypes
import os
import subprocess
import wave

import numpy as np
import pytest

from conftest import TOL, parity
from irbaboon_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_f32p = ctypes.POINTER(ctypes.c_float)


def _fp(a):
    return a.ctypes.data_as(_f32p)


@pytest.fixture(scope="module")
def fac(built):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")])
    return ctypes.CDLL(os.path.join(ROOT, "tests", "cpp", "libfacade_capi.so"))


def _read(fac, path, cap=1 << 20):
    out = np.zeros(cap, np.float32)
    ch, n, sr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = fac.fac_read_wav(path.encode(), _fp(out), cap, ctypes.byref(ch), ctypes.byref(n), ctypes.byref(sr))
    return rc, out[: ch.value * n.value].reshape(ch.value, n.value).copy() if rc == 0 else None, sr.value


@pytest.mark.parametrize("bits", [16, 24])
@pytest.mark.parametrize("ch,n", [(1, 1), (2, 1000), (2, 65536), (3, 7)])
def test_wav_writer_is_standard_pcm(fac, tmp_path, bits, ch, n):
    rng = np.random.default_rng(ch * 1000 + n)
    x = (rng.random((ch, n), dtype=np.float32) * 2.4 - 1.2).astype(np.float32)       # some samples beyond full scale: clipped
    x[0, 0] = 1.0
    path = str(tmp_path / "t.wav")
    assert fac.fac_write_wav(path.encode(), _fp(x), ch, n, 48000, bits) == 0
    with wave.open(path, "rb") as w:                                                   # an independent reader accepts the file
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (ch, bits // 8, 48000, n)
        raw = w.readframes(n)
    assert os.path.getsize(path) == 44 + len(raw) + (len(raw) & 1)
    b = np.frombuffer(raw, np.uint8).reshape(n, ch, bits // 8).astype(np.int64)
    q = sum(b[..., i] << (8 * i) for i in range(bits // 8))
    q = np.where(q >= 1 << (bits - 1), q - (1 << bits), q).T                           # [ch][n] signed integers
    want = np.rint(np.clip(x.astype(np.float64), -1.0, 1.0) * 2147483647.0).astype(np.int64) >> (32 - bits)
    assert np.array_equal(q, want)
    rc, back, sr = _read(fac, path)
    assert rc == 0 and sr == 48000 and back.shape == (ch, n)
    assert np.array_equal(back, (want / float(1 << (bits - 1))).astype(np.float32))
    assert np.abs(back - np.clip(x, -1, 1)).max() <= 2.0 ** -(bits - 1)


def _juce_quantise(x32, bits):
    """The JUCE 6 integer-WAV sample rule restated in plain Python (no numpy arithmetic, no C++): double product, round half to
    even, saturation at the 32-bit limits, then an arithmetic right shift to the file's word length
    (AudioFormatWriter::convertFloatsToInts, AudioData::Int24::setAsInt32LE / Int16::setAsInt32LE)."""
    samp = float(x32)                       # the float32 value, exactly, as a Python (IEEE double) float
    if samp <= -1.0:
        q = -(1 << 31)
    elif samp >= 1.0:
        q = (1 << 31) - 1
    else:
        q = round(2147483647.0 * samp)      # Python's round(): half to even, on the same IEEE double product
    return q >> (32 - bits)                 # Python's >> on a negative int is arithmetic (floor)


@pytest.mark.parametrize("bits", [16, 24])
def test_wav_quantisation_matches_the_juce_rule_on_edge_vectors(fac, tmp_path, bits):
    """Fixed vectors, bit for bit: full scale, one ulp inside and outside it, the smallest normal and denormal floats, signed
    zeros, values on both sides of a quantisation step, and the step's exact half-way points."""
    f = np.float32
    one = f(1.0)
    step = 2.0 ** -(bits - 1)
    vec = [f(0.0), f(-0.0), one, -one, np.nextafter(one, f(0)), np.nextafter(-one, f(0)), np.nextafter(one, f(2)), np.nextafter(-one, f(-2)), f(1.5), f(-1.5),
           f(3.4e38), f(-3.4e38), np.finfo(f).tiny, -np.finfo(f).tiny, f(1e-45), f(-1e-45), f(2.0 ** -31), f(-2.0 ** -31), f(2.0 ** -32), f(-2.0 ** -33),
           f(step), f(-step), f(step / 2), f(-step / 2), np.nextafter(f(step / 2), f(1)), np.nextafter(f(-step / 2), f(-1)), f(1.5 * step), f(-1.5 * step),
           f(2.5 * step), f(-2.5 * step), f(0.5), f(-0.5), f(0.25) + f(step / 2), f(1.0 / 3.0), f(-2.0 / 3.0), f(0.999), f(-0.999), f(12345 * step), f(-12345 * step - step / 4)]
    x = np.array(vec, dtype=np.float32).reshape(1, -1)
    n = x.shape[1]
    path = str(tmp_path / "edge.wav")
    assert fac.fac_write_wav(path.encode(), _fp(x), 1, n, 48000, bits) == 0
    raw = open(path, "rb").read()[44:44 + n * (bits // 8)]
    got = [int.from_bytes(raw[i * (bits // 8):(i + 1) * (bits // 8)], "little", signed=True) for i in range(n)]
    want = [_juce_quantise(v, bits) for v in x[0]]
    assert got == want, [(float(v), g, w) for v, g, w in zip(x[0], got, want) if g != w]
    top = (1 << (bits - 1)) - 1
    assert got[2] == top and got[3] == -top - 1 and got[6] == top and got[7] == -top - 1          # +-1.0 and beyond: the extreme codes
    assert got[0] == 0 and got[1] == 0 and got[12] == 0 and got[14] == 0                           # zeros, FLT_MIN, a positive denormal
    

def test_wav_reader_takes_float_and_32_bit_files_and_rejects_garbage(fac, tmp_path):
    import struct
    x = synth.white_noise(5, 0, 300).reshape(1, 300)
    for fmt, bits, payload in ((3, 32, x.astype("<f4").tobytes()), (1, 32, np.rint(x.astype(np.float64) * 2147483647.0).astype("<i4").tobytes())):
        p = str(tmp_path / ("f%d.wav" % fmt))
        hdr = b"RIFF" + struct.pack("<I", 36 + 12 + len(payload)) + b"WAVE" + b"LIST" + struct.pack("<I", 4) + b"abcd"      # an extra chunk before fmt
        hdr += b"fmt " + struct.pack("<IHHIIHH", 16, fmt, 1, 44100, 44100 * bits // 8, bits // 8, bits) + b"data" + struct.pack("<I", len(payload))
        open(p, "wb").write(hdr + payload)
        rc, back, sr = _read(fac, p)
        assert rc == 0 and sr == 44100 and np.abs(back - x).max() <= 1e-7
    bad = str(tmp_path / "bad.wav")
    open(bad, "wb").write(b"not a wav file at all")
    assert _read(fac, bad)[0] == -1
    assert _read(fac, str(tmp_path / "missing.wav"))[0] == -1


def test_sweepandir_container_round_trip(fac, tmp_path):
    n = 65536                                                                          # totalSweepBreakSamples
    sweep = (synth.exp_sine_sweep(49152 / 48000.0, 48000.0, 20.0, 24000.0).astype(np.float32) * np.float32(0.5))[:49152]
    ir = synth.decaying_ir(3000, 70000)                                               # longer than the container: truncated
    ir = (ir / np.abs(ir).max() * 0.9).astype(np.float32)
    path = str(tmp_path / "2026-10-18 120000 IR Target.sweepandir")
    assert fac.fac_write_sweep_and_ir(path.encode(), _fp(sweep), len(sweep), _fp(ir), len(ir), 48000, n) == 0
    with wave.open(path, "rb") as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getnframes()) == (2, 3, n)
    s2, i2 = np.ones(n, np.float32), np.ones(n, np.float32)
    assert fac.fac_read_sweep_and_ir(path.encode(), _fp(s2), _fp(i2), n) == 0
    assert np.abs(s2[:49152] - sweep).max() <= 2.0 ** -23 and not s2[49152:].any()     # zero-padded to the container length
    assert np.abs(i2 - ir[:n]).max() <= 2.0 ** -23
    assert fac.fac_read_sweep_and_ir(str(tmp_path / "nope.sweepandir").encode(), _fp(s2), _fp(i2), n) == -1


def test_spectrum_tsv_and_raw_text_dumps(fac, orc, tmp_path):
    x = synth.white_noise(6, 0, 512)
    spec = orc.fft_transform(x)[0]                                                      # [2N] interleaved, N = 512
    path = str(tmp_path / "fft_IR_target.tsv")
    assert fac.fac_write_spectrum_tsv(path.encode(), b"fft IR target", _fp(spec), len(spec), 48000) == 0
    lines = open(path).read().split("\n")
    assert lines[0] == "freq\tfft IR target[lin]\tfft IR target[dB]\tbin\tfft IR target phase[rad]"
    rows = [l.split("\t") for l in lines[1:] if l]
    assert len(rows) == 257 and [int(r[3]) for r in rows] == list(range(257))
    c = spec[0:514:2] + 1j * spec[1:514:2]
    freq = np.array([float(r[0]) for r in rows])
    assert np.allclose(freq, np.arange(257) * 24000.0 / 256, rtol=2e-6)
    assert np.allclose([float(r[1]) for r in rows], np.abs(c), rtol=1e-6)
    assert np.allclose([float(r[2]) for r in rows], 20 * np.log10(np.abs(c)), atol=1e-4)
    assert np.allclose([float(r[4]) for r in rows], np.angle(c), atol=1e-6)
    assert fac.fac_write_spectrum_tsv(path.encode(), b"x", _fp(spec), 1000, 48000) == -1           # not a power of two: refused
    packed = np.arange(64, dtype=np.float32)
    raw = str(tmp_path / "IR.txt")
    assert fac.fac_write_raw_text(raw.encode(), _fp(packed), 64, 16) == 0
    txt = open(raw).read()
    assert txt.startswith("\n\n0, 1, 2, 3, 4, 5, 6, 7, \n8, ") and txt.count("\n\n") == 2 and txt.rstrip().endswith("63,")


@pytest.mark.gpu
def test_capture_to_filter_chain_matches_the_reference_functions(fac, orc, ref):
    """The plug-in's capture flow at its own sizes: 65 536-sample capture in host blocks of 512, deconvolved against the
    sweep with the default smoothing, target / base -> IR filter, chopped and normalised."""
    n, H, sr = 65536, 512, 48000.0
    sweep = np.zeros(n, np.float32)
    sweep[:49152] = synth.exp_sine_sweep(49153 / 48000.0, 48000.0, 20.0, 24000.0).astype(np.float32)[:49152]
    fac.fac_capture_to_ir.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, _f32p, ctypes.c_int, ctypes.c_double, _f32p, _f32p]
    fac.fac_create_ir_filt.argtypes = [_f32p, ctypes.c_int, _f32p, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int, _f32p]
    fac.fac_chop_and_normalize.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int, _f32p]
    irs, wants = [], []
    for j in range(2):                                                                # target and base rooms
        h = synth.decaying_ir(3000 + j, 6000 + 2000 * j, j)
        cap = np.ascontiguousarray(orc.convolve_nonperiodic(sweep, h)[0, :n] + synth.white_noise(4000 + j, 0, n) * np.float32(1e-3))
        rec, ir = np.zeros(n, np.float32), np.zeros(n, np.float32)
        assert fac.fac_capture_to_ir(_fp(cap), n // H, H, _fp(sweep), n, sr, _fp(rec), _fp(ir)) == n
        assert np.array_equal(rec, cap)                                                # consolidate() restores the capture
        want = orc.deconvolve(cap, sweep, sr, True)
        e, l2 = parity(ir[None, :], want)
        assert e <= TOL and l2 <= TOL, (e, l2)
        irs.append(ir); wants.append(want[0])
    filt = np.zeros(n, np.float32)
    assert fac.fac_create_ir_filt(_fp(irs[0]), n, _fp(irs[1]), n, sr, 1, 1, _fp(filt)) == n
    want_f = orc.deconvolve(irs[0], irs[1], sr, True, True, True)                      # same inputs: isolates this step
    e, l2 = parity(filt[None, :], want_f)
    # Dividing one measured IR's spectrum by another's and smoothing the quotient is ill-conditioned (both are noise where the
    # sweep has no energy): the REFERENCE's own output moves by 1e-4 .. 1e-3 relative L2 when its inputs change by one float32
    # ulp.  The bound is therefore the reference's response to exactly that perturbation (seeded), not a fixed number.
    rng = np.random.default_rng(0)
    pert = [(v * (1 + 1.2e-7 * rng.choice([-1.0, 1.0], n))).astype(np.float32) for v in irs]
    ce, cl2 = parity(orc.deconvolve(pert[0], pert[1], sr, True, True, True), want_f)
    assert e <= max(5e-5, 2 * ce) and l2 <= max(5e-4, 2 * cl2), (e, l2, ce, cl2)
    got = np.zeros(2048, np.float32)
    assert fac.fac_chop_and_normalize(_fp(filt), n, 2048, ctypes.c_float(-60.0), 50, _fp(got)) == 2048
    chopped = ref.ir_chop(filt, 2048, -60.0, 50)[0]
    want_c = chopped * (np.float32(1.0) / np.abs(chopped).max())
    assert np.abs(got - want_c).max() <= 1e-6 and abs(np.abs(got).max() - 1.0) <= 1e-6
