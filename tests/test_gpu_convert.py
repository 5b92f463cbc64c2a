"""Device pixel-format conversion against the oracle and the
reference-generated golden vectors: every pair of formats the reference can connect, chains of
up to four steps in one kernel, odd sizes and unaligned buffers, in-place calls, HDR parameters,
and the decode + conversion path of the texture API for targets that change the pixel size."""
import ctypes as C
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import detex_b200 as D  # noqa: E402
from oracle import harness as H  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "convert_golden.npz")


@pytest.fixture(scope="module")
def oracle():
    return H.load_oracle()


def _device_convert(src, n, sf, tf, offset_in=0, offset_out=0):
    d_in = torch.zeros(len(src) + 64, dtype=torch.uint8, device="cuda")
    d_in[offset_in:offset_in + len(src)] = torch.from_numpy(np.ascontiguousarray(src)).cuda()
    out_bytes = n * H.pixel_size(tf)
    d_out = torch.full((out_bytes + 128,), 0xA5, dtype=torch.uint8, device="cuda")
    D.convert_pixels_device(d_in.data_ptr() + offset_in, n, sf, d_out.data_ptr() + offset_out, tf)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    # nothing outside the target range may be touched
    assert (got[:offset_out] == 0xA5).all() and (got[offset_out + out_bytes:] == 0xA5).all()
    return got[offset_out:offset_out + out_bytes]


def _close(tf, got, want, gamma):
    if gamma == 1.0:
        return np.array_equal(got, want)
    if tf & 0x2000:
        return np.allclose(got.view(np.float32), want.view(np.float32), rtol=4e-6, atol=1e-7)
    dt = np.uint16 if (tf & 3) == 1 else np.uint8
    return np.abs(got.view(dt).astype(int) - want.view(dt).astype(int)).max() <= 1


def test_every_connected_pair_matches_reference_golden():
    g = np.load(GOLDEN)
    done = 0
    for sf, tf, ok in g["pairs"]:
        sf, tf = int(sf), int(tf)
        if not ok or sf == tf:
            continue
        key = "%04x_%04x" % (sf, tf)
        src, want = g[key + "/src"], g[key + "/out"]
        got = _device_convert(src, len(src) // H.pixel_size(sf), sf, tf)
        assert np.array_equal(got, want), key
        done += 1
    assert done > 400


def test_every_connected_pair_matches_oracle_at_size(oracle):
    """70 001 pixels (68 full tiles + a tail), bit-exact for every pair incl. the float ones."""
    n = 70001
    done = 0
    for a in H.PIXEL_FORMATS.values():
        for b in H.PIXEL_FORMATS.values():
            if a == b or H.conversion_chain(oracle, a, b) is None:
                continue
            src = H.synth_pixels(a, n, seed=a * 7 + b)
            ok, want = H.convert_pixels(oracle, src, n, a, b)
            assert ok
            got = _device_convert(src, n, a, b)
            assert np.array_equal(got, want), (hex(a), hex(b))
            done += 1
    assert done > 400


@pytest.mark.parametrize("n", [0, 1, 3, 1023, 1024, 1025, 4097])
def test_sizes_and_unaligned_buffers(oracle, n):
    for (a, b) in ((0x334, 0x220), (0x220, 0x328), (0x2721, 0x2521), (0x101, 0x0), (0x2B22, 0x2521), (0x0, 0x2101)):
        src = H.synth_pixels(a, max(n, 1), seed=n + 1)[: n * H.pixel_size(a)]
        ok, want = H.convert_pixels(oracle, src, n, a, b) if n else (True, np.zeros(0, np.uint8))
        for off_in, off_out in ((0, 0), (1, 0), (0, 3), (5, 7), (16, 32)):
            if n == 0:
                d = torch.full((64,), 0xA5, dtype=torch.uint8, device="cuda")
                D.convert_pixels_device(d.data_ptr(), 0, a, d.data_ptr() + 32, b)
                torch.cuda.synchronize()
                assert (d.cpu().numpy() == 0xA5).all()
                continue
            got = _device_convert(src, n, a, b, off_in, off_out)
            assert ok and np.array_equal(got, want), (hex(a), hex(b), n, off_in, off_out)


def test_in_place_on_device(oracle):
    n = 5000
    for (a, b) in ((0x334, 0x328), (0x2721, 0x2729), (0x1101, 0x101), (0x2721, 0x721), (0x110, 0x1110)):
        src = H.synth_pixels(a, n, seed=5)
        ok, want = H.convert_pixels(oracle, src, n, a, b)
        d = torch.from_numpy(src.copy()).cuda()
        D.convert_pixels_device(d.data_ptr(), n, a, d.data_ptr(), b)
        torch.cuda.synchronize()
        assert ok and np.array_equal(d.cpu().numpy(), want), (hex(a), hex(b))


def test_host_api_convert_and_in_place_rules(oracle):
    lib = D.lib.load()
    n = 3001
    for (a, b) in ((0x334, 0x220), (0x735, 0x334), (0x2F22, 0x721), (0x0, 0x2302)):
        src = H.synth_pixels(a, n, seed=9)
        ok, want = H.convert_pixels(oracle, src, n, a, b)
        got, ok2 = D.detexConvertPixels(src, n, a, b)
        assert ok == ok2 and np.array_equal(got, want), (hex(a), hex(b))
    buf = H.synth_pixels(0x334, 64)
    keep = buf.copy()
    lib.detexConvertPixelsInPlace.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
    lib.detexConvertPixelsInPlace.restype = C.c_bool
    assert lib.detexConvertPixelsInPlace(buf.ctypes.data, 64, 0x334, 0x328)
    assert np.array_equal(buf.reshape(-1, 4)[:, [2, 1, 0, 3]], keep.reshape(-1, 4))
    assert not lib.detexConvertPixelsInPlace(buf.ctypes.data, 64, 0x334, 0x220)
    assert "in-place" in D.detexGetErrorMessage()
    _, ok = D.detexConvertPixels(buf, 16, 0x334, 0x4)            # RGBA8 -> A8: no table entry
    assert not ok and "Unable to find conversion path" in D.detexGetErrorMessage()


def test_host_texture_api_with_conversions_and_invalid_blocks(oracle):
    for fmt, target in (("EAC_SIGNED_R11", "R8"), ("ETC1", "RGBX16"), ("EAC_SIGNED_R11", "FLOAT_R16"), ("ETC1", "RGB8")):
        pf = H.PIXEL_FORMATS[target]
        for (w, h) in ((512, 256), (61, 35)):
            b = H.synth_blocks(fmt, ((w + 3) // 4) * ((h + 3) // 4), seed=17)
            for linear in (True, False):
                want, ok_w, _ = H.cpu_decode(oracle, fmt, b, w, h, linear=linear, pixel_format=pf)
                fn = D.detexDecompressTextureLinear if linear else D.detexDecompressTextureTiled
                got, ok = fn(fmt, b, w, h, pixel_format=pf)
                assert ok == ok_w and np.array_equal(got, want), (fmt, target, w, h, linear)
                assert not ok_w or w < 512        # the large size does contain invalid blocks


def test_hdr_parameters(oracle):
    lib = D.lib.load()
    lib.detexSetHDRParameters.argtypes = [C.c_float, C.c_float, C.c_float]
    lib.detexSetHDRParameters.restype = None
    g = np.load(GOLDEN)
    try:
        for k, (gamma, lo, hi) in enumerate(g["hdr_sets"]):
            lib.detexSetHDRParameters(float(gamma), float(lo), float(hi))
            oracle.detexSetHDRParameters(float(gamma), float(lo), float(hi))
            keys = [f[: -len("/src")] for f in g.files if f.startswith("hdr%d/" % k) and f.endswith("/src")]
            for key in keys:
                sf, tf = (int(x, 16) for x in key.split("/")[1].split("_"))
                src, want = g[key + "/src"], g[key + "/out"]
                got = _device_convert(src, len(src) // H.pixel_size(sf), sf, tf)
                assert _close(tf, got, want, float(gamma)), key
            # larger buffer against the oracle
            for (a, b) in ((0x6101, 0x101), (0x6721, 0x721), (0x6B22, 0x2B22)):
                src = H.synth_pixels(a, 20000, seed=k)
                ok, want = H.convert_pixels(oracle, src, 20000, a, b)
                got = _device_convert(src, 20000, a, b)
                assert ok and _close(b, got, want, float(gamma)), (k, hex(a), hex(b))
    finally:
        lib.detexSetHDRParameters(1.0, 0.0, 1.0)
        oracle.detexSetHDRParameters(1.0, 0.0, 1.0)


# fused into the decode kernels' store epilogue (post op / post class) ...
FUSED = [("BC1", "RGB8"), ("BC1A", "RGB8"), ("BC2", "RGB8"), ("BC3", "RGB8"), ("ETC1", "RGB8"), ("ETC2", "RGB8"),
         ("ETC2_PUNCHTHROUGH", "RGB8"), ("ETC2_EAC", "RGB8"), ("BPTC", "RGB8"), ("BPTC_FLOAT", "FLOAT_RGB16"),
         ("EAC_R11", "R8"), ("EAC_RG11", "RG8"), ("EAC_SIGNED_R11", "R8"), ("EAC_SIGNED_RG11", "RG8"),
         ("SIGNED_RGTC1", "R8"), ("SIGNED_RGTC2", "RG8"), ("RGTC1", "SIGNED_R8"), ("RGTC2", "SIGNED_RG8"),
         ("EAC_SIGNED_R11", "R16"), ("EAC_R11", "SIGNED_R16"), ("SIGNED_RGTC2", "RG16"), ("EAC_RG11", "SIGNED_RG16")]
# ... and through the decode + conversion pass over L2-resident slabs
# (ETC1 -> RGBX16 and EAC_SIGNED_R11 -> FLOAT_R16: uniform random blocks contain invalid ones, and the
# conversion of a zero pixel is not zero there -- the target must still be zero, texture.c:90-93)
TWO_PASS = [("BC3", "R8"), ("ETC2_EAC", "RG8"), ("RGTC2", "RGBX8"), ("RGTC1", "R16"), ("BPTC", "RGBA16"),
            ("BPTC_FLOAT", "RGBX16"), ("BPTC_FLOAT", "FLOAT_RGBX32"), ("EAC_R11", "FLOAT_R16"),
            ("ETC1", "RGBX16"), ("EAC_SIGNED_R11", "FLOAT_R16"), ("EAC_SIGNED_RG11", "RGBX8")]


@pytest.mark.parametrize("fmt,target", FUSED + TWO_PASS)
def test_device_decode_with_conversion(fmt, target, oracle):
    """detexb200DecodeDevice with a target format other than the native one; 2048 x 1536 gives
    several slabs on the two-pass path and several tiles per CTA on the fused one; the last case
    has a target that is only 16-byte aligned (fused classes fall back to two passes)."""
    pf = H.PIXEL_FORMATS[target]
    tf, bs, ps = H.FORMATS[fmt]
    for (w, h, layout, shift) in ((2048, 1536, D.LAYOUT_LINEAR, 0), (2048, 1536, D.LAYOUT_TILED, 0), (3000, 20, D.LAYOUT_LINEAR, 0), (256, 264, D.LAYOUT_LINEAR, 0),
                                  (250, 61, D.LAYOUT_LINEAR, 0), (256, 64, D.LAYOUT_TILED, 16), (256, 64, D.LAYOUT_LINEAR, 16)):
        wib, hib = (w + 3) // 4, (h + 3) // 4
        b = H.synth_blocks(fmt, wib * hib, seed=21)
        if fmt.startswith("BPTC"):
            b = H.force_modes(fmt, b)
        d_blocks = torch.from_numpy(b.copy()).cuda()
        linear = layout == D.LAYOUT_LINEAR
        n_px = w * h if linear else wib * hib * 16
        d_out = torch.full((n_px * H.pixel_size(pf) + 64,), 0x5A, dtype=torch.uint8, device="cuda")
        flag = torch.zeros(1, dtype=torch.int32, device="cuda")
        D.decode_device(fmt, d_blocks.data_ptr(), w, h, d_out.data_ptr() + shift, pixel_format=pf, layout=layout,
                        d_invalid=flag.data_ptr())
        torch.cuda.synchronize()
        want, ok_w, _ = H.cpu_decode(oracle, fmt, b, w, h, linear=linear, pixel_format=pf, threads=32)
        got = d_out.cpu().numpy()
        assert np.array_equal(got[shift:shift + want.size], want), (fmt, target, w, h, layout, shift)
        assert (got[:shift] == 0x5A).all() and (got[shift + want.size:] == 0x5A).all()
        assert bool(flag.item() == 0) == ok_w
