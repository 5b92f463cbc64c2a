"""CPU-side checks of the product library: it builds, loads, exports every symbol the public
headers declare, fails loudly without a GPU, and its packed tables match the oracle's."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import detex_b200 as D
from oracle import harness as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for hdr in ("detex.h", "detex_b200.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        for m in re.finditer(r"DETEX(?:B200)?_API\s+[\w\s\*]+?\b(detex\w+)\s*\(", text):
            if "##" not in m.group(0):
                names.add(m.group(1))
        for m in re.finditer(r"DETEX_DECLARE_BLOCK_DECODER\((\w+)\)", text):
            if m.group(1) != "NAME":		# the macro definition itself
                names.add("detexDecompressBlock" + m.group(1))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    lib = D.load()
    names = _declared_symbols()
    assert len(names) >= 19 + 6 + 7
    for n in names:
        assert hasattr(lib, n), "libdetex_b200.so does not export " + n


def test_version_string():
    assert b"sm_100a" in D.load().detexb200Version()
    from detex_b200 import build as B
    # the build id in the version string is the hash of the sources the library was compiled from
    assert D.build_id() == B.build_id(), "libdetex_b200.so is older than its sources: rebuild (python -m detex_b200.build)"


def test_header_lut_data_symbols_match_the_reference():
    """The four arrays behind the inline helpers of the reference header (detex.h:933-982: detexClamp0To255,
    detexDivide0To767By3, ...By7, ...By5) are exported with the reference's contents, so that a host compiled
    against the reference's own header links against this library alone."""
    lib = D.load()
    ref = H.load_reference()
    for name, n in (("detex_clamp0to255_table", 767), ("detex_division_by_3_table", 768),
                    ("detex_division_by_7_table", 1792), ("detex_division_by_5_table", 1280)):
        mine = bytes((C.c_uint8 * n).in_dll(lib, name))
        if name.startswith("detex_clamp"):
            assert mine == bytes(min(255, max(0, i - 255)) for i in range(n))
        else:
            d = int(name.split("_")[3])
            assert mine == bytes(i // d for i in range(n))
        if ref is not None:
            assert mine == bytes((C.c_uint8 * n).in_dll(ref, name)), name


REF_HEADER = "/root/reference/detex.h"


@pytest.mark.skipif(not os.path.exists(REF_HEADER), reason="reference header only exists in the build container")
def test_c_host_built_against_the_reference_header_links_this_library_alone(tmp_path):
    """A C99 host written against the REFERENCE's detex.h -- using the inline helpers that pull in the LUT data
    symbols, the texture-level API and the error message -- compiles, links against libdetex_b200.so and nothing
    else of detex, and runs: without a GPU the decode call fails loudly, the helpers work either way."""
    src = tmp_path / "host.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "detex.h"
int main(void) {
	if (detexClamp0To255(-7) != 0 || detexClamp0To255(300) != 255 || detexClamp0To255(77) != 77) return 2;
	if (detexDivide0To767By3(767) != 255 || detexDivide0To1791By7(1791) != 255 || detexDivide0To1279By5(1279) != 255) return 3;
	if (detexDivideMinus895To895By7(-895) != -127 || detexDivideMinus639To639By5(-639) != -127) return 4;
	uint8_t blocks[2 * 8] = { 0 }, pixels[8 * 4 * 4];
	detexTexture t = { DETEX_TEXTURE_FORMAT_BC1, blocks, 8, 4, 2, 1 };
	bool ok = detexDecompressTextureLinear(&t, pixels, DETEX_PIXEL_FORMAT_RGBX8);
	printf("%d %s\n", (int)ok, ok ? "" : detexGetErrorMessage());
	return 0;
}
''')
    exe = tmp_path / "host"
    libdir = os.path.join(ROOT, "detex_b200")
    subprocess.run(["gcc", "-std=c99", "-D_POSIX_C_SOURCE=200809L", "-I", os.path.dirname(REF_HEADER), str(src), "-o", str(exe),
                    "-L", libdir, "-l:libdetex_b200.so", "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    ok, _, msg = r.stdout.strip().partition(" ")
    if D.load().detexb200DeviceCount() > 0:
        assert ok == "1"
    else:
        assert ok == "0" and "no CPU fallback" in msg


def test_header_constants_match_reference_values():
    """include/detex.h must carry the reference's numeric values (dumped from the reference
    header into tests/golden/reference_enum_values.txt by the build container)."""
    ref = {}
    for line in open(os.path.join(ROOT, "tests", "golden", "reference_enum_values.txt")):
        k, v = line.split()
        ref[k] = int(v, 16)
    text = open(os.path.join(ROOT, "include", "detex.h")).read()
    found = 0
    for m in re.finditer(r"\b(DETEX_[A-Z0-9_]+)\s*=?\s*(0[xX][0-9A-Fa-f]+|\d+)\b", text):
        name, val = m.group(1), int(m.group(2), 0)
        if name in ref:
            assert ref[name] == val, name
            found += 1
    assert found > 80
    for name, (tf, bs, ps) in H.FORMATS.items():
        assert D.FORMATS[name] == tf == ref["DETEX_TEXTURE_FORMAT_" + name]
        assert D.block_size(tf) == bs and D.pixel_size(tf & 0xFFFF) == ps


def test_packed_bptc_tables_match_oracle_tables():
    """bptc_tables.cuh (generated from the compiled reference's symbols) must equal the same
    packing derived from the oracle's independent copy of the partition sets."""
    src = open(os.path.join(ROOT, "oracle", "detex_oracle.c")).read()

    def strings(name):
        body = src.split(name + "[64] = {")[1].split("};")[0]
        return re.findall(r'"([012]{16})"', body)

    def ints(name):
        body = src.split(name + "[64] = {")[1].split("};")[0]
        return [int(v) for v in re.findall(r"\d+", body)]

    p2, p3 = strings("partitions2"), strings("partitions3")
    a2, a23, a33 = ints("anchor2_of_2"), ints("anchor2_of_3"), ints("anchor3_of_3")
    assert len(p2) == len(p3) == 64 and len(a2) == len(a23) == len(a33) == 64
    cuh = open(os.path.join(ROOT, "detex_b200", "csrc", "bptc_tables.cuh")).read()
    two = [int(v, 16) for v in re.findall(r"0x([0-9A-F]{8})u,", cuh.split("bptc_two_global")[1].split("};")[0])]
    three = re.findall(r"\{ 0x([0-9A-F]{8})u, 0x([0-9A-F]{4})u \}", cuh.split("bptc_three_global")[1])
    assert len(two) == 64 and len(three) == 64
    for s in range(64):
        mask = sum(int(p2[s][i]) << i for i in range(16))
        assert two[s] == (mask | (a2[s] << 16))
        bits = sum(int(p3[s][i]) << (2 * i) for i in range(16))
        assert int(three[s][0], 16) == bits and int(three[s][1], 16) == (a23[s] | (a33[s] << 8))
        # anchors belong to their subsets, so "pixel i is an anchor" == "i in {0, a2, a3}"
        assert p2[s][a2[s]] == "1" and p3[s][a23[s]] == "1" and p3[s][a33[s]] == "2"
        assert p2[s][0] == "0" and p3[s][0] == "0"


def test_no_cpu_fallback_without_device():
    """Without a CUDA device every decode entry point must fail loudly, not fall back."""
    lib = D.load()
    if lib.detexb200DeviceCount() > 0:
        pytest.skip("a CUDA device is present")
    out, ok = D.detexDecompressTextureLinear("BC1", np.zeros(8, dtype=np.uint8), 4, 4)
    assert not ok and "no CPU fallback" in D.detexGetErrorMessage()
    out, ok = D.detexDecompressBlockFormat("BPTC", np.zeros(16, dtype=np.uint8))
    assert not ok
    assert not out.any()


def test_identity_conversion_is_host_memcpy():
    src = np.arange(64, dtype=np.uint8)
    out, ok = D.detexConvertPixels(src, 16, D.PIXEL_FORMATS["RGBA8"], D.PIXEL_FORMATS["RGBA8"])
    assert ok and np.array_equal(out, src)


def test_product_sources_do_not_use_the_oracle():
    """The product path must never route through oracle/ (only tests, smoke and the bench CPU
    legs may): no product source imports, includes, links or dlopens anything from it."""
    pat = re.compile(r"(import\s+oracle|from\s+oracle|#include\s+\"[^\"]*oracle|liboracle|libdetex_ref|cpu_harness)")
    for root, _, files in os.walk(os.path.join(ROOT, "detex_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(root, f)).read()), (root, f)


def test_hdr_decode_entry_point_validates_on_the_host():
    """detexb200DecodeHDRDevice rejects every texture format without an _HDR pixel format before it
    touches the device (so this runs without a GPU), and the chains it fuses are the reference's."""
    for fmt in ("BC1", "BPTC", "BPTC_SIGNED_FLOAT", "ETC2"):
        with pytest.raises(D.DetexError) as e:
            D.decode_hdr_device(fmt, 256, 16, 16, 512, D.PIXEL_FORMATS["RGBX8"])
        assert "does not decode to an HDR pixel format" in str(e.value)
    lib = D.load()
    out = (C.c_uint32 * 5)()
    hdr, plain = 0x6721, 0x2721                 # FLOAT_RGBX16_HDR / FLOAT_RGBX16
    for src in (hdr, plain):
        assert lib.detexb200ConversionPlan(src, D.PIXEL_FORMATS["RGBX8"], out) == 2
        assert [out[i] for i in range(3)] == [src, D.PIXEL_FORMATS["RGBX16"], D.PIXEL_FORMATS["RGBX8"]]
        assert lib.detexb200ConversionPlan(src, D.PIXEL_FORMATS["RGB8"], out) == 3
        assert [out[i] for i in range(4)] == [src, D.PIXEL_FORMATS["RGBX16"], D.PIXEL_FORMATS["RGBX8"], D.PIXEL_FORMATS["RGB8"]]
        assert lib.detexb200ConversionPlan(src, D.PIXEL_FORMATS["RGBX16"], out) == 1
