"""BC7 register arithmetic of the CUDA decoders, compiled for the CPU (tests/emu/host_emu.h restates
the few PTX intrinsics) and compared bit-for-bit with the oracle: both the per-pixel decoder
(decode_bptc.cuh) and the palette + selector decoder (decode_bptc_palette.cuh) the BC7 texture
kernel uses.  Runs without a GPU; the kernels themselves are covered by the -m gpu tests."""
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
    so = str(tmp_path_factory.mktemp("emu") / "libemu_bptc.so")
    subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-I", EMU, "-I",
                    os.path.join(ROOT, "detex_b200", "csrc"), "-o", so, os.path.join(EMU, "emu_bptc.cpp")],
                   check=True)
    lib = C.CDLL(so)
    lib.emu_bc7_decode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.emu_bc7_decode.restype = C.c_int
    lib.emu_bc6h_decode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.emu_bc6h_decode.restype = C.c_int
    return lib


def _emu_decode(lib, blocks, swap_rb=False, path=0):
    n = blocks.size // 16
    px = np.empty(n * 64, dtype=np.uint8)
    valid = np.empty(n, dtype=np.uint8)
    entries = np.empty(n, dtype=np.int32)
    blocks = np.ascontiguousarray(blocks)
    rc = lib.emu_bc7_decode(blocks.ctypes.data, n, int(swap_rb), path, px.ctypes.data, valid.ctypes.data,
                            entries.ctypes.data)
    assert rc == 0, "selector out of range in block %d" % (-rc - 1)
    return px, valid, entries


def _oracle_tiles(cpu, blocks, pixel_format=None):
    n = blocks.size // 16
    out, ok, _ = H.cpu_decode(cpu, "BPTC", blocks, 4 * n, 4, linear=False, pixel_format=pixel_format)
    return out, ok


@pytest.fixture(scope="module")
def cpu():
    return H.best_cpu_decoder()[0]


@pytest.mark.parametrize("path", [0, 1])
@pytest.mark.parametrize("mode", list(range(8)) + [None])
def test_bc7_emulated_decoder_matches_oracle(emu, cpu, mode, path):
    base = H.synth_blocks("BPTC", 32768, seed=11 + (mode or 0))
    b = base if mode is None else H.force_modes("BPTC", base, [mode])
    got, valid, entries = _emu_decode(emu, b, path=path)
    want, ok = _oracle_tiles(cpu, b)
    assert np.array_equal(got, want), (mode, path)
    assert bool(valid.all()) == ok
    if mode is not None and path == 0:
        want_entries = {0: 16, 1: 16, 2: 12, 3: 8, 4: 16, 5: 16, 6: 16, 7: 8}[mode]
        assert (entries == want_entries).all()


def test_bc7_emulated_every_partition(emu, cpu):
    """all 64 partition sets x all modes: every anchor position and subset map"""
    base = H.force_modes("BPTC", H.synth_blocks("BPTC", 64 * 8 * 16, seed=23)).reshape(-1, 16)
    for i in range(base.shape[0]):
        m, p = i % 8, (i // 8) % 64
        nbits = [4, 6, 6, 6, 0, 0, 0, 6][m]
        v = int.from_bytes(base[i].tobytes(), "little")
        mask = ((1 << nbits) - 1) << (m + 1)
        v = (v & ~mask) | ((p << (m + 1)) & mask)
        base[i] = np.frombuffer(v.to_bytes(16, "little"), dtype=np.uint8)
    b = base.reshape(-1)
    want, _ = _oracle_tiles(cpu, b)
    for path in (0, 1):
        got, _, _ = _emu_decode(emu, b, path=path)
        assert np.array_equal(got, want), path


def test_bc7_emulated_swap_rb(emu, cpu):
    b = H.force_modes("BPTC", H.synth_blocks("BPTC", 8192, seed=5))
    want, _ = _oracle_tiles(cpu, b, pixel_format=H.PF["BGRA8"])
    for path in (0, 1):
        got, _, _ = _emu_decode(emu, b, swap_rb=True, path=path)
        assert np.array_equal(got, want), path


def test_bc7_emulated_extreme_endpoints(emu, cpu):
    """all-zero / all-one payloads per mode: lane borrow limits of the packed interpolation"""
    rows = []
    for fill in (0x00, 0xFF, 0xAA, 0x55, 0x0F, 0xF0):
        rows += [bytes([fill] * 16)] * 8		# force_modes cycles block i through mode i % 8
    b = H.force_modes("BPTC", np.frombuffer(b"".join(rows), dtype=np.uint8).copy(), None)
    want, _ = _oracle_tiles(cpu, b)
    got, _, _ = _emu_decode(emu, b)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("split", [0, 1])
@pytest.mark.parametrize("fmt", ["BPTC_FLOAT", "BPTC_SIGNED_FLOAT"])
def test_bc6h_emulated_decoder_matches_oracle(emu, cpu, fmt, split):
    """BC6H (decode_bptc_float.cuh) on the CPU: uniform blocks (12.5 % invalid), every forced mode,
    every partition set, extreme payloads.  split = 1: the decoder in the two halves the SPLIT kernels run
    (mode-dependent half -> 48-byte record -> mode-independent half)"""
    def check(b):
        n = b.size // 16
        got = np.empty(n * 128, dtype=np.uint8)
        valid = np.empty(n, dtype=np.uint8)
        b = np.ascontiguousarray(b)
        assert emu.emu_bc6h_decode(b.ctypes.data, n, int(fmt == "BPTC_SIGNED_FLOAT"), split, got.ctypes.data, valid.ctypes.data) == 0
        want, ok, _ = H.cpu_decode(cpu, fmt, b, 4 * n, 4, linear=False)
        assert np.array_equal(got, want)
        assert bool(valid.all()) == ok
    base = H.synth_blocks(fmt, 32768, seed=31)
    check(base)
    for m in range(14):
        check(H.force_modes(fmt, base[:16 * 8192], [m]))
    blocks = H.force_modes(fmt, H.synth_blocks(fmt, 14 * 32 * 8, seed=37)).reshape(-1, 16)
    for i in range(blocks.shape[0]):            # partition field at bits 77-81
        v = int.from_bytes(blocks[i].tobytes(), "little")
        v = (v & ~(0x1F << 77)) | (((i // 14) % 32) << 77)
        blocks[i] = np.frombuffer(v.to_bytes(16, "little"), dtype=np.uint8)
    check(blocks.reshape(-1))
    rows = []
    for fill in (0x00, 0xFF, 0xAA, 0x55, 0x0F, 0xF0, 0x7F, 0x80):
        rows += [bytes([fill] * 16)] * 14
    check(H.force_modes(fmt, np.frombuffer(b"".join(rows), dtype=np.uint8).copy(), None))
