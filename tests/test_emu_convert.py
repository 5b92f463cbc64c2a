"""The conversion kernel's own arithmetic on the CPU: detex_b200/csrc/convert_steps.cuh (the reference's
conversion table and chain search restated as data, and the per-pixel step functions with their explicit
rounding modes) compiled with g++ through tests/emu/host_emu.h and compared with the golden vectors the
compiled reference produced for EVERY connected format pair (tests/golden/convert_golden.npz), the HDR
parameter sets included, and with the compiled reference itself on larger buffers.  Runs without a GPU; the
kernel around these functions (staging, tails, alignment) is covered by tests/test_gpu_convert.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import harness as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
GOLDEN = os.path.join(ROOT, "tests", "golden", "convert_golden.npz")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "libemu_convert.so")
    subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-I", EMU, "-I",
                    os.path.join(ROOT, "detex_b200", "csrc"), "-o", so, os.path.join(EMU, "emu_convert.cpp")],
                   check=True)
    lib = C.CDLL(so)
    lib.emu_convert.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint32, C.c_float, C.c_float,
                                C.c_float, C.POINTER(C.c_uint32)]
    lib.emu_convert.restype = C.c_int
    return lib


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def _emu(emu, src, n, sf, tf, hdr=(1.0, 0.0, 1.0)):
    src = np.ascontiguousarray(src, dtype=np.uint8)
    out = np.zeros(n * H.pixel_size(tf) + 16, dtype=np.uint8)
    chain = (C.c_uint32 * 5)()
    steps = emu.emu_convert(src.ctypes.data, n, sf, out.ctypes.data, tf, hdr[0], hdr[1], hdr[2], chain)
    return steps, out[: n * H.pixel_size(tf)], [int(chain[i]) for i in range(max(steps, 0) + 1)]


def _close(tf, got, want, gamma):
    if gamma == 1.0:
        return np.array_equal(got, want)
    if tf & 0x2000:
        return np.allclose(got.view(np.float32), want.view(np.float32), rtol=4e-6, atol=1e-7)
    dt = np.uint16 if (tf & 3) == 1 else np.uint8
    return np.abs(got.view(dt).astype(int) - want.view(dt).astype(int)).max() <= 1


def test_every_pair_of_the_reference_golden(emu, golden):
    """connected pairs: same bytes as the reference; the others: no chain"""
    oracle = H.load_oracle()
    done = 0
    for sf, tf, ok in golden["pairs"]:
        sf, tf = int(sf), int(tf)
        key = "%04x_%04x" % (sf, tf)
        if not ok:
            steps, _, _ = _emu(emu, np.zeros(16 * H.pixel_size(sf), np.uint8), 16, sf, tf)
            assert steps == -1, key
            continue
        src, want = golden[key + "/src"], golden[key + "/out"]
        steps, got, chain = _emu(emu, src, len(src) // H.pixel_size(sf), sf, tf)
        assert steps >= 0 and np.array_equal(got, want), key
        assert chain == H.conversion_chain(oracle, sf, tf), key        # the same chain as the restated search
        done += 1
    assert done > 400


def test_hdr_parameter_sets_of_the_reference_golden(emu, golden):
    for k, (g, lo, hi) in enumerate(golden["hdr_sets"]):
        keys = [f[: -len("/src")] for f in golden.files if f.startswith("hdr%d/" % k) and f.endswith("/src")]
        assert keys
        for key in keys:
            sf, tf = (int(x, 16) for x in key.split("/")[1].split("_"))
            src, want = golden[key + "/src"], golden[key + "/out"]
            steps, got, _ = _emu(emu, src, len(src) // H.pixel_size(sf), sf, tf, (float(g), float(lo), float(hi)))
            assert steps >= 1 and _close(tf, got, want, float(g)), key


@pytest.mark.skipif(not H.have_reference(), reason="compiled reference not present")
def test_float_and_half_steps_against_the_compiled_reference_at_size(emu):
    """the rounding-sensitive steps on 50 000 pixels of mixed magnitudes each"""
    ref = H.load_reference()
    PF = H.PIXEL_FORMATS
    for a, b in (("FLOAT_RGBX32", "FLOAT_RGBX16"), ("FLOAT_RGBX16", "FLOAT_RGBX32"), ("FLOAT_RGBX16", "RGBX16"),
                 ("FLOAT_RGBX32", "RGBX16"), ("RGBX16", "FLOAT_RGBX16"), ("FLOAT_R32", "R8"), ("FLOAT_RG16", "RG8"),
                 ("R16", "FLOAT_R16"), ("FLOAT_RGBX16", "BGRA8")):
        for seed in (3, 4):
            src = H.synth_pixels(PF[a], 50000, seed=seed)
            # (fresh rounding mode: the integer -> half step is the one place where the reference's result
            #  depends on what its thread converted before, DESIGN.md section 2, deviation 2)
            ok, want = H.convert_pixels(ref, src, 50000, PF[a], PF[b])
            steps, got, _ = _emu(emu, src, 50000, PF[a], PF[b])
            assert ok and steps >= 1 and np.array_equal(got, want), (a, b, seed)
