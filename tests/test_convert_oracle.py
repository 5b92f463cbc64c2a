"""Pixel-format conversion, CPU side: the oracle's restatement of
convert.c / half-float.c / hdr.c against golden vectors produced by the compiled reference
(tests/golden/make_convert_golden.py), against the compiled reference itself when it is present,
and the product library's host-side chain search against both."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import harness as H

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "convert_golden.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.fixture(scope="module")
def oracle():
    return H.load_oracle()


def _close(tf, got, want, gamma):
    """Bit-exact, except behind a gamma != 1 powf: one integer step / a few float ulps."""
    if gamma == 1.0:
        return np.array_equal(got, want)
    if tf & 0x2000:
        a, b = got.view(np.float32), want.view(np.float32)
        return np.allclose(a, b, rtol=4e-6, atol=1e-7)
    dt = np.uint16 if (tf & 3) == 1 else np.uint8
    return np.abs(got.view(dt).astype(int) - want.view(dt).astype(int)).max() <= 1


def test_oracle_matches_reference_golden_for_every_pair(golden, oracle):
    pairs = golden["pairs"]
    assert len(pairs) > 1300 and pairs[:, 2].sum() > 400
    for sf, tf, ok in pairs:
        sf, tf = int(sf), int(tf)
        key = "%04x_%04x" % (sf, tf)
        if not ok:
            got_ok, _ = H.convert_pixels(oracle, np.zeros(16 * H.pixel_size(sf), np.uint8), 16, sf, tf)
            assert not got_ok, key
            assert b"Unable to find conversion path" in oracle.detexGetErrorMessage()
            continue
        src, want = golden[key + "/src"], golden[key + "/out"]
        n = len(src) // H.pixel_size(sf)
        got_ok, got = H.convert_pixels(oracle, src, n, sf, tf)
        assert got_ok and np.array_equal(got, want), key


def test_oracle_matches_reference_golden_hdr(golden, oracle):
    try:
        for k, (g, lo, hi) in enumerate(golden["hdr_sets"]):
            oracle.detexSetHDRParameters(float(g), float(lo), float(hi))
            keys = [f[: -len("/src")] for f in golden.files if f.startswith("hdr%d/" % k) and f.endswith("/src")]
            assert keys
            for key in keys:
                sf, tf = (int(x, 16) for x in key.split("/")[1].split("_"))
                src, want = golden[key + "/src"], golden[key + "/out"]
                ok, got = H.convert_pixels(oracle, src, len(src) // H.pixel_size(sf), sf, tf)
                assert ok and _close(tf, got, want, float(g)), key
    finally:
        oracle.detexSetHDRParameters(1.0, 0.0, 1.0)


def test_table_order_is_the_references(golden, oracle):
    """Every reference table entry is a one-step chain of the oracle and of the product."""
    lib = H_product()
    out = (C.c_uint32 * 5)()
    for s, t in golden["table"]:
        assert H.conversion_chain(oracle, int(s), int(t)) == [int(s), int(t)]
        assert lib.detexb200ConversionPlan(int(s), int(t), out) == 1 and out[1] == int(t)
    assert len(golden["table"]) == 73


def H_product():
    import detex_b200 as D
    lib = D.lib.load()
    lib.detexb200ConversionPlan.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
    lib.detexb200ConversionPlan.restype = C.c_int
    return lib


def test_product_chain_search_equals_oracle_for_every_pair(oracle):
    """Host-only part of the product: which formats a conversion goes through."""
    lib = H_product()
    out = (C.c_uint32 * 5)()
    n_paths = 0
    for a in H.PIXEL_FORMATS.values():
        for b in H.PIXEL_FORMATS.values():
            n = lib.detexb200ConversionPlan(a, b, out)
            chain = H.conversion_chain(oracle, a, b)
            if chain is None:
                assert n == -1, (hex(a), hex(b))
            else:
                assert [int(out[i]) for i in range(n + 1)] == chain, (hex(a), hex(b))
                n_paths += 1
    assert n_paths > 450


@pytest.mark.skipif(not H.have_reference(), reason="compiled reference not present")
def test_oracle_equals_compiled_reference_on_larger_buffers(oracle):
    ref = H.load_reference()
    n = 4099
    checked = 0
    for a in H.PIXEL_FORMATS.values():
        for b in H.PIXEL_FORMATS.values():
            chain = H.conversion_chain(oracle, a, b)
            if chain is None or len(chain) < 2:
                continue
            if any((chain[i], chain[i + 1]) in H.REFERENCE_UB_STEPS for i in range(len(chain) - 1)):
                continue
            src = H.synth_pixels(a, n, seed=a * 31 + b)
            okr, want = H.convert_pixels(ref, src, n, a, b)
            oko, got = H.convert_pixels(oracle, src, n, a, b)
            assert okr and oko and np.array_equal(got, want), (hex(a), hex(b))
            checked += 1
    assert checked > 400


def test_in_place_rules(oracle):
    buf = H.synth_pixels(0x334, 64)
    keep = buf.copy()
    # RGBA8 -> BGRX8: two same-size steps, allowed in place (convert.c:1100-1113)
    assert oracle.detexConvertPixelsInPlace(buf.ctypes.data, 64, 0x334, 0x328)
    assert np.array_equal(buf.reshape(-1, 4)[:, [2, 1, 0, 3]], keep.reshape(-1, 4))
    # RGBA8 -> RGB8 changes the pixel size
    assert not oracle.detexConvertPixelsInPlace(buf.ctypes.data, 64, 0x334, 0x220)
    assert b"in-place" in oracle.detexGetErrorMessage()
