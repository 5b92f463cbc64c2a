"""KTX / DDS file -> decoded mip chain in device memory against the oracle."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import container_files as CF  # noqa: E402
import detex_b200 as D  # noqa: E402
from oracle import harness as H  # noqa: E402

GL = {"BC1": 0x83F0, "BC3": 0x83F3, "BPTC": 0x8E8C, "BPTC_FLOAT": 0x8E8F, "ETC2": 0x9274, "ETC2_EAC": 0x9278,
      "EAC_RG11": 0x9272, "RGTC1": 0x8DBB, "EAC_SIGNED_R11": 0x9271}


def _device_bytes(ptr, n):
    """n bytes at a raw device pointer that torch does not own (identity "conversion" = device copy)."""
    out = torch.empty(n, dtype=torch.uint8, device="cuda")
    D.convert_pixels_device(ptr, n, 0x0, out.data_ptr(), 0x0)
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("fmt", list(GL))
@pytest.mark.parametrize("container", ["ktx", "dds"])
def test_file_to_device_mip_chain(fmt, container, tmp_path):
    tf, bs, ps = H.FORMATS[fmt]
    oracle = H.load_oracle()
    dxgi = {"BC1": 71, "BC3": 77, "BPTC": 98, "BPTC_FLOAT": 95, "RGTC1": 80}.get(fmt)
    if container == "dds" and dxgi is None:
        pytest.skip("format has no DDS identifier in the reference's table")
    w, h, levels = 200, 120, 6
    dims = CF.level_dims(w, h, 4, levels)
    pay = [H.synth_blocks(fmt, wib * hib, seed=40 + i).tobytes() for i, (_, _, wib, hib) in enumerate(dims)]
    path = str(tmp_path / ("chain." + container))
    if container == "ktx":
        CF.write_ktx(path, GL[fmt], 0, 0, w, h, bs, 4, pay)
    else:
        CF.write_dds(path, w, h, pay, four_cc=b"DX10", dx10_format=dxgi)
    native = tf & 0xFFFF
    targets = [0]
    if native in (0x320, 0x334):
        targets += [H.PIXEL_FORMATS["BGRA8"], H.PIXEL_FORMATS["RGB8"], H.PIXEL_FORMATS["RGBA16"]]
    if native == 0x2721:
        targets += [H.PIXEL_FORMATS["FLOAT_RGB16"]]
    for pf in targets:
        for layout in (D.LAYOUT_LINEAR, D.LAYOUT_TILED):
            imgs, got_tf, all_valid, free = D.load_texture_file_to_device(path, 32, pf, layout)
            try:
                assert got_tf == tf and len(imgs) == levels
                ok_all = True
                for i, (lw, lh, wib, hib) in enumerate(dims):
                    im = imgs[i]
                    assert (im.width, im.height, im.width_in_blocks, im.height_in_blocks) == (lw, lh, wib, hib)
                    if wib * hib == 0:
                        continue
                    blocks = np.frombuffer(pay[i], dtype=np.uint8)
                    want, ok, _ = H.cpu_decode(oracle, fmt, blocks, lw, lh, linear=(layout == D.LAYOUT_LINEAR),
                                               pixel_format=(pf or native))
                    ok_all &= ok
                    assert im.pixel_bytes == want.size
                    assert np.array_equal(_device_bytes(im.d_pixels, want.size), want), (fmt, container, hex(pf), layout, i)
                assert all_valid == ok_all
            finally:
                free()


def test_uncompressed_file_is_converted_on_device(tmp_path):
    oracle = H.load_oracle()
    w, h, levels = 33, 20, 3
    dims = CF.level_dims(w, h, 1, levels)
    rng = np.random.default_rng(2)
    pay = [rng.integers(0, 256, a * b * 4, dtype=np.uint8).tobytes() for (a, b, _, _) in dims]
    path = str(tmp_path / "rgba.ktx")
    CF.write_ktx(path, 0x1908, 0x1908, 0x1401, w, h, 4, 1, pay)
    for pf in (0, H.PIXEL_FORMATS["BGRA8"], H.PIXEL_FORMATS["RGB8"], H.PIXEL_FORMATS["RGBA16"], H.PIXEL_FORMATS["FLOAT_RGBX16"]):
        imgs, tf, all_valid, free = D.load_texture_file_to_device(path, 8, pf)
        try:
            assert tf == 0x334 and len(imgs) == levels and all_valid
            for i, (a, b, _, _) in enumerate(dims):
                src = np.frombuffer(pay[i], dtype=np.uint8)
                ok, want = H.convert_pixels(oracle, src, a * b, 0x334, pf or 0x334)
                assert ok and np.array_equal(_device_bytes(imgs[i].d_pixels, want.size), want), (hex(pf), i)
        finally:
            free()


def test_file_errors_surface(tmp_path):
    with pytest.raises(D.DetexError, match="Do not recognize filename extension"):
        D.load_texture_file_to_device(str(tmp_path / "x.png"))
    with pytest.raises(D.DetexError, match="Could not open file"):
        D.load_texture_file_to_device(str(tmp_path / "missing.ktx"))
    p = str(tmp_path / "bc1.ktx")
    CF.write_ktx(p, 0x83F0, 0, 0, 16, 16, 8, 4, [bytes(128)])
    with pytest.raises(D.DetexError, match="Unable to find conversion path"):
        D.load_texture_file_to_device(p, 1, H.PIXEL_FORMATS["A8"])
