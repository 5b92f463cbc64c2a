"""BC6H straight to 8-bit pixels: the tone classes of the decode kernels
(detex_b200/csrc/tone.cuh) through detexb200DecodeDevice, the host texture API and
detexb200DecodeHDRDevice, against the checker's decode followed by its conversion chain."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import detex_b200 as D  # noqa: E402
from oracle import harness as H  # noqa: E402

PF = H.PIXEL_FORMATS
TARGETS_8BIT = ["RGBX8", "RGBA8", "BGRX8", "BGRA8", "RGB8", "RGBX16"]      # (and the 16-bit stage on its own)
# (w, h, layout, byte offset of the target): sorted kernel with several tiles per CTA, plain kernels,
# clipped kernel, and a target that is only 16-byte aligned (falls back to decode + conversion pass)
SHAPES = [(2048, 1536, D.LAYOUT_LINEAR, 0), (2048, 1536, D.LAYOUT_TILED, 0), (8192, 24, D.LAYOUT_LINEAR, 0),
          (3000, 20, D.LAYOUT_LINEAR, 0),       # 750 blocks per row: staged tiles straddle block rows at odd offsets
          (256, 264, D.LAYOUT_LINEAR, 0),       # 64 blocks per row: a staged tile covers many block rows
          (128, 64, D.LAYOUT_LINEAR, 0), (128, 64, D.LAYOUT_TILED, 0), (250, 61, D.LAYOUT_LINEAR, 0),
          (256, 64, D.LAYOUT_TILED, 16), (256, 64, D.LAYOUT_LINEAR, 16)]


@pytest.fixture(scope="module")
def oracle():
    return H.load_oracle()


def _blocks(w, h, seed, forced):
    b = H.synth_blocks("BPTC_FLOAT", ((w + 3) // 4) * ((h + 3) // 4), seed=seed)
    return H.force_modes("BPTC_FLOAT", b) if forced else b


def _run(call, w, h, layout, shift, ps):
    linear = layout == D.LAYOUT_LINEAR
    n_px = w * h if linear else ((w + 3) // 4) * ((h + 3) // 4) * 16
    d_out = torch.full((n_px * ps + 64,), 0x5A, dtype=torch.uint8, device="cuda")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    before = D.launch_count()
    call(d_out.data_ptr() + shift, flag.data_ptr())
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    assert (got[:shift] == 0x5A).all() and (got[shift + n_px * ps:] == 0x5A).all()
    return got[shift:shift + n_px * ps], bool(flag.item() == 0), D.launch_count() - before


@pytest.mark.parametrize("target", TARGETS_8BIT)
@pytest.mark.parametrize("forced", [False, True])
def test_bc6h_decode_to_8bit_is_one_fused_launch(target, forced, oracle):
    """detexb200DecodeDevice(BPTC_FLOAT, 8-bit target) == the checker's detexDecompressTexture*
    with that target (chain FLOAT_RGBX16 -> RGBX16 -> RGBX8 -> ...), invalid blocks zero in the
    target format; one kernel launch when the target is 32-byte aligned."""
    pf = PF[target]
    for (w, h, layout, shift) in SHAPES:
        b = _blocks(w, h, 31, forced)
        d_blocks = torch.from_numpy(b.copy()).cuda()
        got, ok, launches = _run(lambda out, flag: D.decode_device("BPTC_FLOAT", d_blocks.data_ptr(), w, h, out,
                                                                   pixel_format=pf, layout=layout, d_invalid=flag),
                                 w, h, layout, shift, H.pixel_size(pf))
        want, ok_w, _ = H.cpu_decode(oracle, "BPTC_FLOAT", b, w, h, linear=layout == D.LAYOUT_LINEAR,
                                     pixel_format=pf, threads=32)
        assert np.array_equal(got, want), (target, forced, w, h, layout, shift)
        assert ok == ok_w
        assert launches == 1 or shift != 0, (target, w, h, layout, launches)


def test_host_texture_api_bc6h_to_rgbx8(oracle):
    for (w, h) in ((1024, 512), (96, 40)):
        b = _blocks(w, h, 5, False)
        for linear in (True, False):
            want, ok_w, _ = H.cpu_decode(oracle, "BPTC_FLOAT", b, w, h, linear=linear, pixel_format=PF["RGBX8"])
            fn = D.detexDecompressTextureLinear if linear else D.detexDecompressTextureTiled
            got, ok = fn("BPTC_FLOAT", b, w, h, pixel_format=PF["RGBX8"])
            assert ok == ok_w and np.array_equal(got, want), (w, h, linear)


def _reference_sequence(oracle, b, w, h, linear, pf, gamma, lo, hi):
    """decode to FLOAT_RGBX16, then detexConvertPixels from FLOAT_RGBX16_HDR under the parameters"""
    px, ok, _ = H.cpu_decode(oracle, "BPTC_FLOAT", b, w, h, linear=linear, threads=32)
    n = px.size // 8
    oracle.detexSetHDRParameters(gamma, lo, hi)
    try:
        ok_c, out = H.convert_pixels(oracle, px, n, PF["FLOAT_RGBX16_HDR"], pf)
    finally:
        oracle.detexSetHDRParameters(1.0, 0.0, 1.0)
    assert ok_c
    return out, ok


HDR_SETS = [(1.0, 0.0, 1.0), (1.0, 0.0, 2.0), (1.0, -1.0, 1.0), (1.0, 0.25, 0.75), (1.0, 0.0, 65504.0), (2.2, 0.0, 4.0),
            (2.2, 0.0, 1.0), (0.45, 0.01, 2.0), (1.8, -0.5, 8.0)]


@pytest.mark.parametrize("gamma,lo,hi", HDR_SETS)
@pytest.mark.parametrize("target", ["RGBX8", "BGRX8", "RGB8", "RGBX16"])
def test_decode_hdr_device(gamma, lo, hi, target, oracle):
    """detexb200DecodeHDRDevice == decode + detexSetHDRParameters + detexConvertPixels(FLOAT_RGBX16_HDR -> target).
    One fused launch for every parameter set.  gamma 1 is bit-exact; with gamma 2.2 the kernel evaluates the
    power as 2^(e log2 f) on the special-function unit where the reference tabulates powf, and may differ
    by one integer step of the target."""
    pf = PF[target]
    lib = D.lib.load()
    lib.detexSetHDRParameters.argtypes = [C.c_float, C.c_float, C.c_float]
    lib.detexSetHDRParameters.restype = None
    lib.detexSetHDRParameters(1.0, 0.125, 0.5)          # the caller's own parameters must survive the call
    try:
        for (w, h, layout, shift) in SHAPES[:2] + SHAPES[3:8]:
            b = _blocks(w, h, 77, False)                # uniform bits: 12.5 % invalid blocks
            d_blocks = torch.from_numpy(b.copy()).cuda()
            got, ok, launches = _run(lambda out, flag: D.decode_hdr_device("BPTC_FLOAT", d_blocks.data_ptr(), w, h, out, pf,
                                                                           gamma=gamma, range_min=lo, range_max=hi,
                                                                           layout=layout, d_invalid=flag),
                                     w, h, layout, shift, H.pixel_size(pf))
            want, ok_w = _reference_sequence(oracle, b, w, h, layout == D.LAYOUT_LINEAR, pf, gamma, lo, hi)
            assert ok == ok_w
            assert launches == 1
            if gamma == 1.0:
                assert np.array_equal(got, want), (gamma, lo, hi, target, w, h, layout)
            else:
                dt = np.uint16 if target == "RGBX16" else np.uint8
                assert np.abs(got.view(dt).astype(int) - want.view(dt).astype(int)).max() <= 1
        # the thread's parameters are untouched: a plain HDR conversion still uses (1, 0.125, 0.5)
        src = H.synth_pixels(PF["FLOAT_R16_HDR"], 4096, seed=3)
        oracle.detexSetHDRParameters(1.0, 0.125, 0.5)
        ok_c, want = H.convert_pixels(oracle, src, 4096, PF["FLOAT_R16_HDR"], PF["R16"])
        d_in = torch.from_numpy(src.copy()).cuda()
        d_o = torch.empty(4096 * 2, dtype=torch.uint8, device="cuda")
        D.convert_pixels_device(d_in.data_ptr(), 4096, PF["FLOAT_R16_HDR"], d_o.data_ptr(), PF["R16"])
        torch.cuda.synchronize()
        assert ok_c and np.array_equal(d_o.cpu().numpy(), want)
    finally:
        lib.detexSetHDRParameters(1.0, 0.0, 1.0)
        oracle.detexSetHDRParameters(1.0, 0.0, 1.0)


def test_decode_hdr_device_rejects_other_formats():
    d = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    for fmt in ("BPTC", "BC1", "BPTC_SIGNED_FLOAT"):
        with pytest.raises(D.DetexError):
            D.decode_hdr_device(fmt, d.data_ptr(), 16, 16, d.data_ptr() + 2048, PF["RGBX8"])
