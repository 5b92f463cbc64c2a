"""BC6H -> 8-bit tone path of the decode kernels (detex_b200/csrc/tone.cuh) compiled for the CPU
(tests/emu/host_emu.h restates the rounding-mode intrinsics with fesetround) and compared with the
checker's conversion chain FLOAT_RGBX16[_HDR] -> RGBX16 -> RGBX8 for EVERY half value the unsigned
BC6H decoder can produce (0 .. 0x7BFF), with and without HDR range parameters.  Runs without a GPU;
the kernels themselves are covered by tests/test_gpu_tone.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import harness as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
PF = H.PIXEL_FORMATS

# (range_min, range_max): identity, the validation program's (validate.c:177), negative minimum,
# narrow range, full half range, inverted range (negative factor), awkward fractions.
# range_min == range_max is outside the domain: the reference divides by zero (hdr.c:133) and what
# its -ffinite-math-only build does with the resulting NaN is unspecified.
RANGES = [(0.0, 1.0), (0.0, 2.0), (-1.0, 1.0), (0.25, 0.75), (0.0, 65504.0), (1.0, 0.0),
          (0.1, 3.3), (-0.003, 0.7001), (1e-5, 1e-3)]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "libemu_tone.so")
    subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-I", EMU, "-I",
                    os.path.join(ROOT, "detex_b200", "csrc"), "-o", so, os.path.join(EMU, "emu_bptc.cpp")],
                   check=True)
    lib = C.CDLL(so)
    lib.emu_tone8_tiles.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p]
    lib.emu_tone8_tiles.restype = None
    return lib


@pytest.fixture(scope="module")
def cpu():
    return H.best_cpu_decoder()[0]


def all_half_pixels():
    """FLOAT_RGBX16 pixels covering every non-negative finite half in each colour channel
    (R ascending, G descending, B scrambled), X = 0 as the BC6H decoder writes it; padded to whole tiles"""
    n = 0x7C00
    h = np.arange(n, dtype=np.uint16)
    px = np.zeros((n, 4), dtype=np.uint16)
    px[:, 0] = h
    px[:, 1] = h[::-1]
    px[:, 2] = (h.astype(np.uint32) * 7919 % n).astype(np.uint16)
    return px.reshape(-1).view(np.uint8).copy(), n


def emu_tone(emu, src, n, hdr, lo, hi):
    assert n % 16 == 0
    out = np.empty(n * 4, dtype=np.uint8)
    src = np.ascontiguousarray(src)
    emu.emu_tone8_tiles(src.ctypes.data, n // 16, int(hdr), lo, hi, out.ctypes.data)
    return out


def test_plain_chain_every_half(emu, cpu):
    src, n = all_half_pixels()
    ok, want = H.convert_pixels(cpu, src, n, PF["FLOAT_RGBX16"], PF["RGBX8"])
    assert ok
    got = emu_tone(emu, src, n, 0, 0.0, 1.0)
    assert np.array_equal(got, want)
    assert (got[3::4] == 0xFF).all()


@pytest.mark.parametrize("lo,hi", RANGES)
def test_hdr_chain_every_half(emu, cpu, lo, hi):
    src, n = all_half_pixels()
    cpu.detexSetHDRParameters(1.0, lo, hi)
    try:
        ok, want = H.convert_pixels(cpu, src, n, PF["FLOAT_RGBX16_HDR"], PF["RGBX8"])
    finally:
        cpu.detexSetHDRParameters(1.0, 0.0, 1.0)
    assert ok
    got = emu_tone(emu, src, n, 1, lo, hi)
    bad = np.flatnonzero(got != want)
    assert bad.size == 0, (lo, hi, bad[:8], got[bad[:8]], want[bad[:8]])
