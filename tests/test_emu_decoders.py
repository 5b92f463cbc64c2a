"""Register arithmetic of ALL 19 CUDA block decoders (detex_b200/csrc/decode_*.cuh) compiled for the CPU
(tests/emu/host_emu.h restates the PTX intrinsics they use) and compared bit-for-bit with the checker:
uniform random blocks (which contain every invalid-block case the formats have), every BC7 / BC6H mode,
and the mode_mask / flags semantics of the per-block API.  Runs without a GPU; the texture kernels
around these decoders are covered by the -m gpu tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import harness as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "libemu_all.so")
    subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-I", EMU, "-I",
                    os.path.join(ROOT, "detex_b200", "csrc"), "-o", so, os.path.join(EMU, "emu_all.cpp")],
                   check=True)
    lib = C.CDLL(so)
    lib.emu_decode.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    lib.emu_decode.restype = C.c_int
    return lib


@pytest.fixture(scope="module")
def cpu():
    return H.best_cpu_decoder()[0]


def _emu(emu, fmt, blocks, mode_mask=0xFFFFFFFF, flags=0):
    tf, bs, ps = H.FORMATS[fmt]
    n = blocks.size // bs
    px = np.empty(n * 16 * ps, dtype=np.uint8)
    valid = np.empty(n, dtype=np.uint8)
    blocks = np.ascontiguousarray(blocks)
    tile = emu.emu_decode(tf >> 24, blocks.ctypes.data, n, mode_mask, flags, px.ctypes.data, valid.ctypes.data)
    assert tile == 16 * ps
    return px, valid


@pytest.mark.parametrize("fmt", H.FORMAT_NAMES)
def test_every_decoder_matches_the_checker_on_uniform_blocks(emu, cpu, fmt):
    n = 1 << 15
    for seed in (1, 29):
        b = H.synth_blocks(fmt, n, seed=seed)
        want, ok, _ = H.cpu_decode(cpu, fmt, b, 4 * n, 4, linear=False)
        got, valid = _emu(emu, fmt, b)
        assert np.array_equal(got, want), (fmt, seed)
        assert bool(valid.all()) == ok


@pytest.mark.parametrize("fmt", ["BPTC", "BPTC_FLOAT", "BPTC_SIGNED_FLOAT"])
def test_every_mode(emu, cpu, fmt):
    n_modes = 8 if fmt == "BPTC" else 14
    for mode in list(range(n_modes)) + [None]:
        base = H.synth_blocks(fmt, 4096, seed=7 + (mode or 0))
        b = H.force_modes(fmt, base, None if mode is None else [mode])
        want, ok, _ = H.cpu_decode(cpu, fmt, b, 4 * 4096, 4, linear=False)
        got, valid = _emu(emu, fmt, b)
        assert ok and valid.all()
        assert np.array_equal(got, want), (fmt, mode)


@pytest.mark.parametrize("fmt", ["BC1A", "BC2", "BC3", "BPTC", "BPTC_FLOAT", "BPTC_SIGNED_FLOAT", "ETC1", "ETC2",
                                 "ETC2_PUNCHTHROUGH", "ETC2_EAC", "EAC_R11", "EAC_SIGNED_RG11"])
def test_per_block_flags_and_masks(emu, cpu, fmt):
    """mode_mask / flags of the per-block API (detex.h:383-411): same verdict as the checker's
    detexDecompressBlock<FMT>, same pixels wherever the block is accepted"""
    tf, bs, ps = H.FORMATS[fmt]
    n = 384
    b = H.synth_blocks(fmt, n, seed=21)
    if fmt.startswith("BPTC"):
        b = H.force_modes(fmt, b)
    fn = getattr(cpu, "detexDecompressBlock" + fmt)
    o = np.zeros(256, dtype=np.uint8)
    for flags in (0, 1, 2, 4):
        for mask in (0xFFFFFFFF, 0x1, 0x2, 0x4, 0x8, 0x10, 0x15, 0xAA, 0x1F00):
            got, valid = _emu(emu, fmt, b, mask, flags)
            tiles = got.reshape(n, 16 * ps)
            for i in range(n):
                r = fn(b.ctypes.data + i * bs, mask, flags, o.ctypes.data)
                assert bool(valid[i]) == bool(r), (fmt, flags, mask, i)
                if r:
                    assert np.array_equal(tiles[i], o[: 16 * ps]), (fmt, flags, mask, i)


@pytest.mark.parametrize("variant", ["-DDETEXB200_BC7_TABLE2", "-DDETEXB200_BC7_PERPIXEL"])
def test_bc7_both_pixel_paths_are_bit_exact(tmp_path, cpu, variant):
    """decode_bptc.cuh has two ways to turn endpoints + indices into pixels: PRMT byte-table look-ups for modes
    2, 3, 4, 5, 7 (the default build) and the per-pixel interpolation for every mode (-DDETEXB200_BC7_PERPIXEL,
    kept for A/B runs): the same bytes as the checker for every mode, 3 x 32 768 blocks each (all 64 partitions,
    all anchors, all four rotations occur)."""
    so = str(tmp_path / "libemu_variant.so")
    subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", variant, "-I", EMU, "-I",
                    os.path.join(ROOT, "detex_b200", "csrc"), "-o", so, os.path.join(EMU, "emu_all.cpp")], check=True)
    lib = C.CDLL(so)
    lib.emu_decode.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    lib.emu_decode.restype = C.c_int
    for mode in (2, 3, 5, 7, 0, 1, 4, 6, None):
        for seed in (1, 2, 3):
            base = H.synth_blocks("BPTC", 1 << 15, seed=seed)
            b = base if mode is None else H.force_modes("BPTC", base, [mode])
            want, ok, _ = H.cpu_decode(cpu, "BPTC", b, 4 * (1 << 15), 4, linear=False)
            got, valid = _emu(lib, "BPTC", b)
            assert np.array_equal(got, want), (mode, seed)
            assert bool(valid.all()) == ok
