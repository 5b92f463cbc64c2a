"""GPU parity tests: the CUDA path (through the C ABI of libdetex_b200.so) against the CPU
checker -- the compiled reference oracle/_ref when it travelled to this box, else the oracle
restatement -- on the same seeded inputs, bit-exact for every format (the BC6H formats emit
half-float bit patterns through integer arithmetic, so their tolerance is 0 ULP as well).
Also checks the committed golden vectors and the edge cases of the reference API
(clipping, invalid blocks, flags / mode masks, error messages)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import detex_b200 as D
from oracle import harness as H

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def cpu():
    lib, kind = H.best_cpu_decoder()
    return lib


@pytest.fixture(scope="module", autouse=True)
def _need_cuda_and_native_lib():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    assert D.load().detexb200DeviceCount() > 0, D.detexGetErrorMessage()


def _cmp(fmt, blocks, w, h, cpu, linear, threads=4):
    want, ok_w, _ = H.cpu_decode(cpu, fmt, blocks, w, h, linear=linear, threads=threads)
    fn = D.detexDecompressTextureLinear if linear else D.detexDecompressTextureTiled
    got, ok_g = fn(fmt, blocks, w, h)
    assert ok_g == ok_w, (fmt, w, h, linear, D.detexGetErrorMessage())
    if not np.array_equal(got, want):
        ps = H.FORMATS[fmt][2]
        bad = np.nonzero(got != want)[0]
        raise AssertionError("%s %dx%d linear=%s: %d bytes differ, first at byte %d (pixel %d)" %
                             (fmt, w, h, linear, bad.size, bad[0], bad[0] // ps))
    return ok_g


# ---- host-pointer API (reference signature), uniform random blocks ---------------------------
@pytest.mark.parametrize("fmt", H.FORMAT_NAMES)
def test_uniform_blocks_linear_and_tiled(fmt, cpu):
    for seed in (1, 2):
        b = H.synth_blocks(fmt, 65536, seed=seed)
        _cmp(fmt, b, 1024, 1024, cpu, True)
        _cmp(fmt, b, 1024, 1024, cpu, False)


@pytest.mark.parametrize("fmt", H.FORMAT_NAMES)
def test_golden_vectors(fmt, golden):
    g = golden["uniform"][fmt]
    b = H.synth_blocks(fmt, g["blocks"], seed=g["seed"])
    lin, ok = D.detexDecompressTextureLinear(fmt, b, 256, 256)
    til, ok2 = D.detexDecompressTextureTiled(fmt, b, 256, 256)
    assert ok == g["ok"] and ok2 == g["ok"]
    assert "%016x" % H.fnv1a64(lin) == g["linear"]
    assert "%016x" % H.fnv1a64(til) == g["tiled"]
    tiles = til.reshape(g["blocks"], g["tile_bytes"])
    assert int((~tiles.any(axis=1)).sum()) >= g["invalid"]


def test_bundled_fixtures(fixtures):
    names = sorted({k.split("/")[0] for k in fixtures.files})
    assert "BC1" in names and len(names) == 17
    for name in names:
        w, h, ok_ref = (int(v) for v in fixtures[name + "/meta"])
        out, ok = D.detexDecompressTextureLinear(name, fixtures[name + "/blocks"], w, h)
        assert ok == bool(ok_ref), name
        assert np.array_equal(out, fixtures[name + "/pixels"]), name


@pytest.mark.parametrize("fmt", ["BPTC", "BPTC_FLOAT", "BPTC_SIGNED_FLOAT"])
def test_forced_modes(fmt, cpu, golden):
    n_modes = 8 if fmt == "BPTC" else 14
    base = H.synth_blocks(fmt, 16384, seed=3)
    for m in list(range(n_modes)) + [None]:
        b = H.force_modes(fmt, base, None if m is None else [m])
        assert _cmp(fmt, b, 512, 512, cpu, True), (fmt, m)	# no invalid blocks in distribution M
        _cmp(fmt, b, 512, 512, cpu, False)
    for mode, g in golden["forced_modes"][fmt].items():
        b = H.force_modes(fmt, H.synth_blocks(fmt, 4096, seed=g["seed"]), None if mode == "cycle" else [int(mode)])
        out, ok = D.detexDecompressTextureLinear(fmt, b, 256, 256)
        assert ok == g["ok"] and "%016x" % H.fnv1a64(out) == g["linear"], (fmt, mode)


@pytest.mark.parametrize("fmt", ["BPTC", "BPTC_FLOAT", "BPTC_SIGNED_FLOAT"])
def test_every_partition_and_anchor(fmt, cpu):
    """All 64 partition sets x all modes: exercises every anchor position."""
    n_modes = 8 if fmt == "BPTC" else 14
    base = H.synth_blocks(fmt, 64 * n_modes * 8, seed=17).reshape(-1, 16)
    blocks = H.force_modes(fmt, base.reshape(-1)).reshape(-1, 16)
    # sweep the partition field: BC7 partition sits right after the mode bits; BC6H at bits 77-81
    for i in range(blocks.shape[0]):
        p = (i // n_modes) % 64
        if fmt == "BPTC":
            m = i % 8
            nbits = [4, 6, 6, 6, 0, 0, 0, 6][m]
            v = int.from_bytes(blocks[i].tobytes(), "little")
            mask = ((1 << nbits) - 1) << (m + 1)
            v = (v & ~mask) | ((p << (m + 1)) & mask)
            blocks[i] = np.frombuffer(v.to_bytes(16, "little"), dtype=np.uint8)
        else:
            v = int.from_bytes(blocks[i].tobytes(), "little")
            mask = 0x1F << 77
            v = (v & ~mask) | ((p & 0x1F) << 77)
            blocks[i] = np.frombuffer(v.to_bytes(16, "little"), dtype=np.uint8)
    n = blocks.shape[0]
    w = 64 * 4
    h = (n // 64) * 4
    _cmp(fmt, blocks.reshape(-1), w, h, cpu, True)


@pytest.mark.parametrize("fmt", H.FORMAT_NAMES)
def test_clipped_and_odd_sizes(fmt, cpu, golden):
    for (w, h) in [(1021, 37), (5, 3), (1, 1), (130, 67), (64, 61), (4, 4), (12, 8), (36, 4), (124, 4)]:
        b = H.synth_blocks(fmt, ((w + 3) // 4) * ((h + 3) // 4), seed=5)
        _cmp(fmt, b, w, h, cpu, True, threads=1)
        _cmp(fmt, b, w, h, cpu, False, threads=1)
    for size, g in golden["clipped"][fmt].items():
        w, h = (int(v) for v in size.split("x"))
        b = H.synth_blocks(fmt, ((w + 3) // 4) * ((h + 3) // 4), seed=g["seed"])
        out, ok = D.detexDecompressTextureLinear(fmt, b, w, h)
        assert ok == g["ok"] and "%016x" % H.fnv1a64(out) == g["linear"], (fmt, size)


def test_empty_texture():
    out, ok = D.detexDecompressTextureLinear("BC1", np.zeros(0, dtype=np.uint8), 0, 0)
    assert ok and out.size == 0
    out, ok = D.detexDecompressTextureTiled("BPTC", np.zeros(0, dtype=np.uint8), 0, 0)
    assert ok and out.size == 0


def test_invalid_blocks_zero_tile_and_message(cpu):
    # BC7 block with byte 0 == 0 is invalid (Q7); texture-level decode zero-fills it, carries
    # on, returns false and leaves the reference's message (texture.c:62-66, 90-93).
    b = H.synth_blocks("BPTC", 64, seed=4).reshape(64, 16).copy()
    b[5, 0] = 0
    b[63, 0] = 0
    out, ok = D.detexDecompressTextureTiled("BPTC", b.reshape(-1), 32, 32)
    assert not ok
    assert "Decompress function for format 0x0B800334 returned error" in D.detexGetErrorMessage()
    tiles = out.reshape(64, 64)
    assert not tiles[5].any() and not tiles[63].any() and tiles[6].any()
    want, ok_w, _ = H.cpu_decode(cpu, "BPTC", b.reshape(-1), 32, 32, linear=False)
    assert not ok_w and np.array_equal(out, want)


def test_uncompressed_texture_is_converted_not_decoded():
    # texture.c:108-111: uncompressed formats go straight to detexConvertPixels
    px = np.arange(4 * 4 * 4, dtype=np.uint8)
    out, ok = D.detexDecompressTextureLinear(D.PIXEL_FORMATS["RGBA8"], px, 4, 4, pixel_format=D.PIXEL_FORMATS["RGBA8"])
    assert ok and np.array_equal(out, px)
    out, ok = D.detexDecompressTextureLinear(D.PIXEL_FORMATS["RGBA8"], px, 4, 4, pixel_format=D.PIXEL_FORMATS["BGRA8"])
    assert ok and np.array_equal(out.reshape(-1, 4)[:, [2, 1, 0, 3]].reshape(-1), px)
    out, ok = D.detexDecompressTextureTiled(D.PIXEL_FORMATS["RGBA8"], px, 4, 4, pixel_format=D.PIXEL_FORMATS["RGBA8"])
    assert not ok and "Cannot handle uncompressed" in D.detexGetErrorMessage()


# ---- per-block API with flags and mode masks ---------------------------------------------------
@pytest.mark.parametrize("fmt", H.FORMAT_NAMES)
def test_per_block_api_flags_and_masks(fmt, cpu):
    bs = H.FORMATS[fmt][1]
    b = H.synth_blocks(fmt, 24, seed=21)
    if fmt in ("BPTC", "BPTC_FLOAT", "BPTC_SIGNED_FLOAT"):
        b = H.force_modes(fmt, b)
    f_ref = getattr(cpu, "detexDecompressBlock" + fmt)
    o = np.zeros(256, dtype=np.uint8)
    combos = [(0xFFFFFFFF, 0), (0xFFFFFFFF, 1), (0xFFFFFFFF, 2), (0xFFFFFFFF, 4), (0x1, 0), (0x2, 0),
              (0x15, 0), (0xAA, 0), (0x1F00, 0)]
    for mask, flags in combos:
        for i in range(24):
            blk = b[i * bs:(i + 1) * bs]
            r = f_ref(blk.ctypes.data, mask, flags, o.ctypes.data)
            got, ok = D.detexDecompressBlockFormat(fmt, blk, mask, flags)
            assert ok == r, (fmt, mask, flags, i)
            if r:
                assert np.array_equal(got, o[:got.size]), (fmt, mask, flags, i)


def test_generic_block_entry_point(cpu):
    b = H.synth_blocks("ETC2_EAC", 4, seed=8)
    tf = H.FORMATS["ETC2_EAC"][0]
    o = np.zeros(64, dtype=np.uint8)
    for i in range(4):
        blk = b[i * 16:(i + 1) * 16]
        r = cpu.detexDecompressBlock(blk.ctypes.data, tf, 0xFFFFFFFF, 0, o.ctypes.data, tf & 0xFFFF)
        got, ok = D.detexDecompressBlock(blk, tf, 0xFFFFFFFF, 0, tf & 0xFFFF)
        assert ok == r and np.array_equal(got, o)
    got, ok = D.detexDecompressBlock(b[:16], 0x14800334, 0xFFFFFFFF, 0, 0x334)	# ASTC: no decoder
    assert not ok and "no decompress function" in D.detexGetErrorMessage()


# ---- device-pointer API -------------------------------------------------------------------------
@pytest.mark.parametrize("fmt", H.FORMAT_NAMES)
def test_device_api_and_generator(fmt, cpu):
    tf, bs, ps = H.FORMATS[fmt]
    n = 256 * 96
    w, h = 1024, 384
    d_blocks = torch.empty(n * bs, dtype=torch.uint8, device="cuda")
    D.generate_blocks_device(d_blocks.data_ptr(), n, bs, seed=7)
    host_blocks = H.synth_blocks(fmt, n, seed=7)
    assert np.array_equal(d_blocks.cpu().numpy(), host_blocks)		# device generator == CPU generator
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    for layout, linear in ((D.LAYOUT_LINEAR, True), (D.LAYOUT_TILED, False)):
        d_out = torch.full((w * h * ps,), 0xCD, dtype=torch.uint8, device="cuda")
        flag.zero_()
        D.decode_device(fmt, d_blocks.data_ptr(), w, h, d_out.data_ptr(), layout=layout, d_invalid=flag.data_ptr())
        torch.cuda.synchronize()
        want, ok_w, _ = H.cpu_decode(cpu, fmt, host_blocks, w, h, linear=linear, threads=4)
        assert np.array_equal(d_out.cpu().numpy(), want), (fmt, linear)
        assert bool(flag.item() == 0) == ok_w


def test_device_api_unaligned_output_and_odd_width(cpu):
    # output pointer off by 4 bytes and width % 4 != 0 -> clipped general path
    fmt = "BC3"
    tf, bs, ps = H.FORMATS[fmt]
    w, h = 203, 45
    wib, hib = 51, 12
    b = H.synth_blocks(fmt, wib * hib, seed=31)
    d_blocks = torch.from_numpy(b.copy()).cuda()
    d_out = torch.zeros(w * h * ps + 64, dtype=torch.uint8, device="cuda")
    D.decode_device(fmt, d_blocks.data_ptr(), w, h, d_out.data_ptr() + 4, layout=D.LAYOUT_LINEAR)
    torch.cuda.synchronize()
    want, _, _ = H.cpu_decode(cpu, fmt, b, w, h, linear=True)
    got = d_out.cpu().numpy()
    assert np.array_equal(got[4:4 + want.size], want) and not got[:4].any() and not got[4 + want.size:].any()


def test_forced_mode_generator_matches_host(cpu):
    for fmt, nm in (("BPTC", 8), ("BPTC_FLOAT", 14)):
        n = 4096
        d = torch.empty(n * 16, dtype=torch.uint8, device="cuda")
        D.generate_blocks_device(d.data_ptr(), n, 16, seed=3, force_modes=nm)
        want = H.force_modes(fmt, H.synth_blocks(fmt, n, seed=3))
        assert np.array_equal(d.cpu().numpy(), want)


def test_device_hash_is_the_tree_hash():
    for n in (0, 1, 4095, 4096, 4097, 1 << 20, (1 << 20) + 13):
        a = np.random.default_rng(n).integers(0, 256, size=n, dtype=np.uint8)
        d = torch.from_numpy(a).cuda() if n else torch.empty(0, dtype=torch.uint8, device="cuda")
        assert D.hash_device(d.data_ptr() if n else 0, n) == H.tree_hash(a), n


# ---- conversions fused behind the texture API ------------------------------------
@pytest.mark.parametrize("fmt,target", [("BC1", "BGRX8"), ("BC1", "RGB8"), ("BC1", "RGBA8"), ("BC3", "BGRA8"),
                                        ("BC1", "BGRA8"), ("BC3", "RGB8"), ("BC3", "BGRX8"),
                                        ("BPTC", "BGRA8"), ("BPTC", "RGBX8"), ("ETC2", "RGB8"),
                                        ("BPTC_FLOAT", "FLOAT_RGB16"), ("BPTC_FLOAT", "FLOAT_BGRX16")])
def test_texture_decode_with_conversion(fmt, target, cpu):
    # FLOAT_RGBX16 -> FLOAT_RGB16 writes through an uninitialised pointer in the reference
    # (convert.c:707, undefined behaviour); the intended result (drop X) is what the oracle
    # restates, so that pair is checked against the oracle.
    if target == "FLOAT_RGB16":
        cpu = H.load_oracle()
    pf = D.PIXEL_FORMATS[target]
    for (w, h) in ((256, 128), (61, 35)):
        b = H.synth_blocks(fmt, ((w + 3) // 4) * ((h + 3) // 4), seed=13)
        if fmt.startswith("BPTC"):
            b = H.force_modes(fmt, b)
        want, ok_w, _ = H.cpu_decode(cpu, fmt, b, w, h, linear=True, pixel_format=pf)
        got, ok = D.detexDecompressTextureLinear(fmt, b, w, h, pixel_format=pf)
        assert ok == ok_w and np.array_equal(got, want), (fmt, target, w, h)
        want, ok_w, _ = H.cpu_decode(cpu, fmt, b, w, h, linear=False, pixel_format=pf)
        got, ok = D.detexDecompressTextureTiled(fmt, b, w, h, pixel_format=pf)
        assert ok == ok_w and np.array_equal(got, want), (fmt, target, w, h, "tiled")


def test_unsupported_conversion_fails_loudly():
    b = H.synth_blocks("BPTC_SIGNED_FLOAT", 16, seed=1)
    out, ok = D.detexDecompressTextureLinear("BPTC_SIGNED_FLOAT", b, 16, 16, pixel_format=D.PIXEL_FORMATS["RGBA8"])
    assert not ok and "conversion path" in D.detexGetErrorMessage()


# ---- BASELINE.json configurations at full size ------------------------------------------------------
def _full_size(fmt, w, h, cpu, forced=0, threads=None):
    import os
    threads = threads or min(64, os.cpu_count() or 1)
    tf, bs, ps = H.FORMATS[fmt]
    n = (w // 4) * (h // 4)
    d_blocks = torch.empty(n * bs, dtype=torch.uint8, device="cuda")
    D.generate_blocks_device(d_blocks.data_ptr(), n, bs, seed=1, force_modes=forced)
    d_out = torch.empty(w * h * ps, dtype=torch.uint8, device="cuda")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    D.decode_device(fmt, d_blocks.data_ptr(), w, h, d_out.data_ptr(), layout=D.LAYOUT_LINEAR, d_invalid=flag.data_ptr())
    torch.cuda.synchronize()
    host_blocks = d_blocks.cpu().numpy()
    want, ok_w, _ = H.cpu_decode(cpu, fmt, host_blocks, w, h, linear=True, threads=threads)
    got = d_out.cpu().numpy()
    assert bool(flag.item() == 0) == ok_w
    assert np.array_equal(got, want), fmt
    # the host-pointer API on the same texture (pipelined slabs) must agree too
    got2, ok2 = D.detexDecompressTextureLinear(fmt, host_blocks, w, h)
    assert ok2 == ok_w and np.array_equal(got2, want)


def test_config_bc1_8192(cpu):
    _full_size("BC1", 8192, 8192, cpu)


def test_config_bc7_16384(cpu):
    _full_size("BPTC", 16384, 16384, cpu, forced=8)


def test_config_bc6h_8192(cpu):
    _full_size("BPTC_FLOAT", 8192, 8192, cpu, forced=14)


def test_host_api_tiled_multi_slab(cpu):
    """Flat block arrays through detexDecompressTextureTiled are cut into block-range slabs
    (not rows): 1.5M EAC_RG11 blocks given as one 'row' of blocks -> several pipelined slabs."""
    fmt = "EAC_RG11"
    n = 1_500_000 + 37
    b = H.synth_blocks(fmt, n, seed=4)
    tf, bs, ps = H.FORMATS[fmt]
    out = np.zeros(n * 16 * ps, dtype=np.uint8)
    tex = D.DetexTexture(tf, b.ctypes.data, 4 * n, 4, n, 1)
    ok = D.load().detexDecompressTextureTiled(C.byref(tex), out.ctypes.data, tf & 0xFFFF)
    want, ok_w, _ = H.cpu_decode(cpu, fmt, b[: (n - 37) * bs], 4 * 1000, 4 * ((n - 37) // 1000), linear=False, threads=16)
    assert ok == ok_w
    assert np.array_equal(out[: want.size], want)
    tail, _, _ = H.cpu_decode(cpu, fmt, b[(n - 37) * bs:], 4 * 37, 4, linear=False)
    assert np.array_equal(out[want.size:], tail)


def test_config_mixed_tiled_reduced(cpu):
    """configs[4] at 1/50 scale per format (2e6 blocks each), TILED layout, per-shard tree hashes
    computed on the device against the same construction on the host."""
    for fmt, seed in (("ETC2", 1), ("EAC_RG11", 2), ("RGTC2", 3)):
        tf, bs, ps = H.FORMATS[fmt]
        n = 2_000_000
        d_blocks = torch.empty(n * bs, dtype=torch.uint8, device="cuda")
        D.generate_blocks_device(d_blocks.data_ptr(), n, bs, seed=seed)
        d_out = torch.empty(n * 16 * ps, dtype=torch.uint8, device="cuda")
        D.decode_device(fmt, d_blocks.data_ptr(), 4 * 2000, 4 * 1000, d_out.data_ptr(), layout=D.LAYOUT_TILED, wib=2000, hib=1000)
        torch.cuda.synchronize()
        want, _, _ = H.cpu_decode(cpu, fmt, d_blocks.cpu().numpy(), 8000, 4000, linear=False, threads=16)
        assert D.hash_device(d_out.data_ptr(), d_out.numel()) == H.tree_hash(want), fmt
        assert np.array_equal(d_out.cpu().numpy(), want), fmt


def test_bc7_flow_kernel():
    """The alternative BC7 kernel (palette + selectors, barrier-free dataflow; kernels.cu
    decode_bc7_flow_kernel) is selected per process with DETEXB200_BC7_KERNEL=flow: check it in a
    child process (tests/bc7_flow_check.py)."""
    if b"+experiments" not in D.load().detexb200Version():
        pytest.skip("library built without -DDETEXB200_EXPERIMENTS (the default): the dataflow kernel is not in it")
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, DETEXB200_BC7_KERNEL="flow")
    r = subprocess.run([sys.executable, os.path.join(here, "bc7_flow_check.py")], env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().startswith("OK"), r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("split", ["0", "1"])
def test_bc6h_kernel_structures(split):
    """BC6H runs one of two structures of the mode-sorted kernel depending on the target format (kernels.cu,
    bc6h_split_selected); DETEXB200_BC6H_SPLIT forces one for every BC6H launch of a process.  Both are checked
    against the CPU checker in a child process (tests/bc6h_structure_check.py)."""
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, DETEXB200_BC6H_SPLIT=split)
    r = subprocess.run([sys.executable, os.path.join(here, "bc6h_structure_check.py")], env=env, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.strip().startswith("OK"), r.stdout[-2000:] + r.stderr[-4000:]


# ---- round 2: narrow images, pageable / pinned callers, threads, device sets, configs[4] at full size ----------
def test_linear_image_narrower_than_its_block_grid(cpu):
    """width may be anything in ((wib - 1) * 4, wib * 4] for the reference (texture.c:133-136 clips the last
    block column); a multiple of four below wib * 4 used to take the unclipped fast path and write past the rows
    (ADVICE r1).  Device API with explicit wib, guard bytes behind the image, every format family."""
    for fmt in ("BC1", "BPTC", "BPTC_FLOAT", "ETC2", "EAC_R11", "RGTC1"):
        tf, bs, ps = H.FORMATS[fmt]
        for wib, hib, width, height in ((8, 8, 29, 32), (8, 8, 30, 30), (300, 20, 1197, 77), (2, 1, 5, 4), (1100, 4, 4397, 16)):
            n = wib * hib
            blocks = H.synth_blocks(fmt, n, seed=wib + hib)
            if fmt.startswith("BPTC"):
                blocks = H.force_modes(fmt, blocks)
            # the checker decodes the full block grid; the image is its top-left width x height corner
            full, _, _ = H.cpu_decode(cpu, fmt, blocks, wib * 4, hib * 4, linear=True)
            want = full.reshape(hib * 4, wib * 4 * ps)[:height, : width * ps].reshape(-1)
            d_blocks = torch.from_numpy(blocks.copy()).cuda()
            d_out = torch.full((width * height * ps + 4096,), 0xA5, dtype=torch.uint8, device="cuda")
            D.decode_device(fmt, d_blocks.data_ptr(), width, height, d_out.data_ptr(), layout=D.LAYOUT_LINEAR, wib=wib, hib=hib)
            torch.cuda.synchronize()
            got = d_out.cpu().numpy()
            assert np.array_equal(got[: want.size], want), (fmt, wib, hib, width, height)
            assert (got[want.size:] == 0xA5).all(), ("write past the image", fmt, wib, hib, width, height)
    # a width that is a multiple of four but SMALLER than the grid (undefined in the reference, legal for
    # detexb200DecodeDevice's validation): the columns that exist are decoded, nothing else is touched
    for fmt in ("BC1", "BPTC"):
        tf, bs, ps = H.FORMATS[fmt]
        wib, hib, width, height = 1024, 4, 4, 16
        blocks = H.force_modes(fmt, H.synth_blocks(fmt, wib * hib, seed=2)) if fmt == "BPTC" else H.synth_blocks(fmt, wib * hib, seed=2)
        full, _, _ = H.cpu_decode(cpu, fmt, blocks, wib * 4, hib * 4, linear=True)
        want = full.reshape(hib * 4, wib * 4 * ps)[:height, : width * ps].reshape(-1)
        d_out = torch.full((width * height * ps + 65536,), 0xA5, dtype=torch.uint8, device="cuda")
        D.decode_device(fmt, torch.from_numpy(blocks.copy()).cuda().data_ptr(), width, height, d_out.data_ptr(),
                        layout=D.LAYOUT_LINEAR, wib=wib, hib=hib)
        torch.cuda.synchronize()
        got = d_out.cpu().numpy()
        assert np.array_equal(got[: want.size], want) and (got[want.size:] == 0xA5).all(), fmt


def _host_call(fmt, blocks, w, h, linear, pf, pinned):
    """detexDecompressTexture{Linear,Tiled} on pageable (numpy) or pinned (torch) caller buffers."""
    tf, bs, ps_native = H.FORMATS[fmt]
    ps = H.pixel_size(pf)
    wib, hib = (w + 3) // 4, (h + 3) // 4
    n_out = w * h * ps if linear else wib * hib * 16 * ps
    if pinned:
        hb = torch.from_numpy(blocks.copy()).pin_memory()
        out = torch.zeros(n_out, dtype=torch.uint8).pin_memory()
        in_ptr, out_ptr = hb.data_ptr(), out.data_ptr()
    else:
        hb = blocks.copy()
        out = np.zeros(n_out, dtype=np.uint8)
        in_ptr, out_ptr = hb.ctypes.data, out.ctypes.data
    tex = D.DetexTexture(tf, in_ptr, w, h, wib, hib)
    fn = D.load().detexDecompressTextureLinear if linear else D.load().detexDecompressTextureTiled
    ok = fn(C.byref(tex), out_ptr, pf)
    return (out.numpy() if pinned else out), bool(ok)


@pytest.mark.parametrize("pinned", [False, True])
def test_host_api_pageable_and_pinned_callers(cpu, pinned):
    """The host-pointer API takes two routes (api.cu): pinned caller buffers are DMA'd directly from the calling
    thread's lanes; pageable ones (what malloc'ing callers such as detex-convert.c:303-310 pass) go through the copy
    pool's pinned bounce buffers.  Same bytes either way: several slabs, ragged sizes, both layouts, invalid blocks,
    fused and two-pass conversions."""
    PF = H.PIXEL_FORMATS
    cases = [("BC1", 2048, 2052, None, 0), ("BPTC", 2048, 1030, None, 0), ("BPTC", 1024, 2048, None, 8),
             ("BPTC_FLOAT", 1024, 1024, None, 14), ("ETC2_EAC", 1532, 1000, None, 0), ("EAC_SIGNED_R11", 4096, 1024, None, 0),
             ("BC3", 2048, 1024, PF["BGRA8"], 0), ("BC1", 2048, 1024, PF["RGBA16"], 0), ("ETC1", 1024, 2048, PF["RGB8"], 0)]
    for fmt, w, h, pf, force in cases:
        tf, bs, _ = H.FORMATS[fmt]
        pf = (tf & 0xFFFF) if pf is None else pf
        n = ((w + 3) // 4) * ((h + 3) // 4)
        blocks = H.synth_blocks(fmt, n, seed=w + h)
        if force:
            blocks = H.force_modes(fmt, blocks)
        for linear in (True, False):
            want, ok_w, _ = H.cpu_decode(cpu, fmt, blocks, w, h, linear=linear, pixel_format=pf, threads=8)
            got, ok = _host_call(fmt, blocks, w, h, linear, pf, pinned)
            assert ok == ok_w, (fmt, w, h, linear, D.detexGetErrorMessage())
            assert np.array_equal(got, want), (fmt, w, h, linear, pinned)


def test_host_api_from_many_threads(cpu):
    """The reference is re-entrant and lock-free on this path (TLS only: convert.c:870-873, misc.c:75); callers may
    decode from many threads at once.  Six threads, mixed formats, layouts, sizes on both sides of
    the copy-pool threshold, pinned and pageable buffers, each checked against the checker; one thread decodes an
    invalid texture and must see ITS error, the others none."""
    import threading
    jobs = [("BC1", 2048, 1024, True, False), ("BPTC", 1024, 1024, True, False), ("BPTC_FLOAT", 512, 1024, False, False),
            ("ETC2", 256, 256, True, False), ("EAC_RG11", 1024, 512, False, True), ("BC3", 1028, 516, True, True)]
    prepared = []
    for fmt, w, h, linear, pinned in jobs:
        n = ((w + 3) // 4) * ((h + 3) // 4)
        blocks = H.synth_blocks(fmt, n, seed=len(prepared) + 1)
        if fmt.startswith("BPTC") and fmt != "BPTC":
            blocks = H.force_modes(fmt, blocks)
        want, ok_w, _ = H.cpu_decode(cpu, fmt, blocks, w, h, linear=linear, threads=4)
        prepared.append((fmt, w, h, linear, pinned, blocks, want, ok_w))
    errors = []

    def work(k):
        fmt, w, h, linear, pinned, blocks, want, ok_w = prepared[k]
        try:
            for _ in range(4):
                got, ok = _host_call(fmt, blocks, w, h, linear, H.FORMATS[fmt][0] & 0xFFFF, pinned)
                msg = D.detexGetErrorMessage()
                if ok != ok_w or not np.array_equal(got, want):
                    errors.append((fmt, "pixels or verdict differ"))
                if not ok and ("0x%08X" % H.FORMATS[fmt][0]) not in msg:
                    errors.append((fmt, "error text of another thread: " + msg))
            D.release_thread_resources()
        except Exception as exc:       # noqa: BLE001
            errors.append((fmt, repr(exc)))

    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(prepared))]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    assert not prepared[1][7], "the BC7 job is meant to contain invalid blocks (distribution U)"


def test_device_set_spreads_one_call_over_gpus(cpu):
    """detexb200SetDevices (SURVEY.md 8e behind the C ABI): one detexDecompressTextureLinear call, slabs dealt
    round-robin to the listed devices.  With one GPU the set {0} still takes the pool route; with more, all of them."""
    n_dev = D.load().detexb200DeviceCount()
    sets = [[0]] + ([list(range(n_dev))] if n_dev > 1 else []) + ([[1, 0]] if n_dev > 1 else [])
    try:
        for devs in sets:
            D.set_devices(devs)
            for fmt, w, h, linear, force in (("BC1", 4096, 2048, True, 0), ("BPTC", 2048, 2048, True, 8), ("EAC_RG11", 2048, 1024, False, 0)):
                n = (w // 4) * (h // 4)
                blocks = H.synth_blocks(fmt, n, seed=9)
                if force:
                    blocks = H.force_modes(fmt, blocks)
                want, ok_w, _ = H.cpu_decode(cpu, fmt, blocks, w, h, linear=linear, threads=8)
                for pinned in (True, False):
                    got, ok = _host_call(fmt, blocks, w, h, linear, H.FORMATS[fmt][0] & 0xFFFF, pinned)
                    assert ok == ok_w and np.array_equal(got, want), (devs, fmt, pinned)
        with pytest.raises(D.DetexError):
            D.set_devices([n_dev])
        with pytest.raises(D.DetexError):
            D.set_devices([0, 0])
    finally:
        D.set_devices([])


def test_config_mixed_tiled_full_size(cpu):
    """configs[4] at FULL size on one GPU: 33 333 334 ETC2 + 33 333 333 EAC_RG11 + 33 333 333 RGTC2 blocks (seeds 1 / 2 / 3),
    TILED.  Compared through the hash of hashes (device: detexb200HashDevice; host: the same construction over the
    checker's output) and byte for byte over the first and last 32 MiB."""
    threads = os.cpu_count() or 1
    for fmt, seed, n in (("ETC2", 1, 33_333_334), ("EAC_RG11", 2, 33_333_333), ("RGTC2", 3, 33_333_333)):
        tf, bs, ps = H.FORMATS[fmt]
        n_gen = n + (n & 1)                        # the generator works in 16-byte units
        d_blocks = torch.empty(n_gen * bs, dtype=torch.uint8, device="cuda")
        D.generate_blocks_device(d_blocks.data_ptr(), n_gen, bs, seed=seed)
        d_out = torch.empty(n * 16 * ps, dtype=torch.uint8, device="cuda")
        flag = torch.zeros(1, dtype=torch.int32, device="cuda")
        D.decode_device(fmt, d_blocks.data_ptr(), 4 * n, 4, d_out.data_ptr(), layout=D.LAYOUT_TILED, wib=n, hib=1,
                        d_invalid=flag.data_ptr())
        torch.cuda.synchronize()
        host_blocks = d_blocks[: n * bs].cpu().numpy()
        assert np.array_equal(host_blocks[: 1 << 20], H.synth_blocks(fmt, (1 << 20) // bs, seed=seed))
        # the checker decodes the flat array as rows of 1000 blocks + a ragged tail
        rows = n // 1000
        want = np.empty(n * 16 * ps, dtype=np.uint8)
        body, ok1, _ = H.cpu_decode(cpu, fmt, host_blocks[: rows * 1000 * bs], 4000, 4 * rows, linear=False, threads=threads)
        want[: body.size] = body
        ok2 = True
        if n > rows * 1000:
            tail, ok2, _ = H.cpu_decode(cpu, fmt, host_blocks[rows * 1000 * bs:], 4 * (n - rows * 1000), 4, linear=False)
            want[body.size:] = tail
        del body
        assert bool(flag.item() == 0) == (ok1 and ok2), fmt
        assert D.hash_device(d_out.data_ptr(), d_out.numel()) == H.tree_hash(want, threads=threads), fmt
        m = 32 << 20
        assert np.array_equal(d_out[:m].cpu().numpy(), want[:m]) and np.array_equal(d_out[-m:].cpu().numpy(), want[-m:]), fmt
        del d_blocks, d_out, want
        torch.cuda.empty_cache()
