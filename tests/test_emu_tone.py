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
    lib.emu_tone8_pixels.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p]
    lib.emu_tone8_pixels.restype = None
    lib.emu_tone8_pixels_gamma.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p]
    lib.emu_tone8_pixels_gamma.restype = None
    lib.emu_bc6h_decode_tone8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    lib.emu_bc6h_decode_tone8.restype = None
    lib.emu_bc6h_decode_tone16.argtypes = lib.emu_bc6h_decode_tone8.argtypes
    lib.emu_bc6h_decode_tone16.restype = None
    return lib


@pytest.fixture(scope="module")
def cpu():
    return H.best_cpu_decoder()[0]


def all_half_pixels():
    """FLOAT_RGBX16 pixels covering every non-negative finite half in each colour channel
    (R ascending, G descending, B scrambled), X = 0 as the BC6H decoder writes it"""
    n = 0x7C00
    h = np.arange(n, dtype=np.uint16)
    px = np.zeros((n, 4), dtype=np.uint16)
    px[:, 0] = h
    px[:, 1] = h[::-1]
    px[:, 2] = (h.astype(np.uint32) * 7919 % n).astype(np.uint16)
    return px.reshape(-1).view(np.uint8).copy(), n


def emu_tone(emu, src, n, hdr, lo, hi):
    out = np.empty(n * 4, dtype=np.uint8)
    src = np.ascontiguousarray(src)
    emu.emu_tone8_pixels(src.ctypes.data, n, int(hdr), lo, hi, out.ctypes.data)
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


@pytest.mark.parametrize("hdr,lo,hi", [(0, 0.0, 1.0), (1, 0.0, 2.0), (1, -1.0, 1.0), (1, 0.0, 65504.0)])
def test_decoder_with_8bit_sink_matches_decode_then_convert(emu, cpu, hdr, lo, hi):
    """the BC6H decoder writing through the 8-bit pixel sink == decode to FLOAT_RGBX16, then the
    conversion chain; every mode (forced) and uniform bits (12.5 % invalid blocks)"""
    for forced in (True, False):
        b = H.synth_blocks("BPTC_FLOAT", 14 * 1024, seed=41)
        if forced:
            b = H.force_modes("BPTC_FLOAT", b)
        n = b.size // 16
        px, ok, _ = H.cpu_decode(cpu, "BPTC_FLOAT", b, 4 * n, 4, linear=False)
        cpu.detexSetHDRParameters(1.0, lo, hi)
        try:
            ok_c, want = H.convert_pixels(cpu, px, n * 16, PF["FLOAT_RGBX16_HDR" if hdr else "FLOAT_RGBX16"], PF["RGBX8"])
        finally:
            cpu.detexSetHDRParameters(1.0, 0.0, 1.0)
        assert ok_c
        got = np.empty(n * 64, dtype=np.uint8)
        valid = np.empty(n, dtype=np.uint8)
        b = np.ascontiguousarray(b)
        emu.emu_bc6h_decode_tone8(b.ctypes.data, n, hdr, lo, hi, got.ctypes.data, valid.ctypes.data)
        assert np.array_equal(got, want), (hdr, lo, hi, forced)
        assert bool(valid.all()) == ok
        assert forced or not ok
        # the 16-bit stage on its own (target RGBX16)
        cpu.detexSetHDRParameters(1.0, lo, hi)
        try:
            ok_c, want16 = H.convert_pixels(cpu, px, n * 16, PF["FLOAT_RGBX16_HDR" if hdr else "FLOAT_RGBX16"], PF["RGBX16"])
        finally:
            cpu.detexSetHDRParameters(1.0, 0.0, 1.0)
        got16 = np.empty(n * 128, dtype=np.uint8)
        emu.emu_bc6h_decode_tone16(b.ctypes.data, n, hdr, lo, hi, got16.ctypes.data, valid.ctypes.data)
        assert ok_c and np.array_equal(got16, want16), (hdr, lo, hi, forced)


@pytest.mark.parametrize("gamma,lo,hi", [(2.2, 0.0, 1.0), (2.2, 0.0, 4.0), (0.45, 0.01, 2.0), (1.8, -0.5, 8.0)])
def test_gamma_variant_every_half_within_one_step(emu, cpu, gamma, lo, hi):
    """the gamma variant (hdr.c:143-164) replaces the reference's powf table by 2^(e log2 f): not bit-exact,
    but never more than one 8-bit step away, for every half value"""
    src, n = all_half_pixels()
    cpu.detexSetHDRParameters(gamma, lo, hi)
    try:
        # (primed: the special-gamma path inherits its thread's rounding mode, DESIGN.md section 2, deviation 2)
        ok, want = H.convert_pixels(cpu, src, n, PF["FLOAT_RGBX16_HDR"], PF["RGBX8"], primed=True)
    finally:
        cpu.detexSetHDRParameters(1.0, 0.0, 1.0)
    assert ok
    got = np.empty(n * 4, dtype=np.uint8)
    emu.emu_tone8_pixels_gamma(np.ascontiguousarray(src).ctypes.data, n, gamma, lo, hi, got.ctypes.data)
    d = np.abs(got.astype(int) - want.astype(int))
    assert d.max() <= 1, (gamma, lo, hi, int(d.max()))
    assert (d != 0).mean() < 0.01
