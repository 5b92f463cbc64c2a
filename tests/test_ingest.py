"""KTX / DDS loaders, host side: the product's detexLoad*File* functions against
containers synthesised by tests/container_files.py, and against the compiled reference's own
loaders (ktx.c, dds.c, file-info.c) on the same files when oracle/_ref is present."""
import ctypes as C
import os

import numpy as np
import pytest

import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import container_files as CF  # noqa: E402
from oracle import harness as H  # noqa: E402

import detex_b200 as D

# (name, texture/pixel format, unit bytes, block, KTX ids, DDS FourCC, DXGI format)
FILE_FORMATS = [
    ("BC1", 0x01000320, 8, 4, (0x83F0, 0, 0), b"DXT1", 71),
    ("BC1A", 0x02000334, 8, 4, (0x83F1, 0, 0), None, None),
    ("BC2", 0x03800334, 16, 4, (0x83F2, 0, 0), b"DXT3", 74),
    ("BC3", 0x04800334, 16, 4, (0x83F3, 0, 0), b"DXT5", 77),
    ("RGTC1", 0x05000000, 8, 4, (0x8DBB, 0, 0), b"ATI1", 80),
    ("SIGNED_RGTC1", 0x06001101, 8, 4, (0x8DBC, 0, 0), b"BC4S", 81),
    ("RGTC2", 0x07800110, 16, 4, (0x8DBD, 0, 0), b"ATI2", 83),
    ("SIGNED_RGTC2", 0x08801311, 16, 4, (0x8DBE, 0, 0), b"BC5S", 84),
    ("BPTC_FLOAT", 0x09802721, 16, 4, (0x8E8F, 0, 0), None, 95),
    ("BPTC_SIGNED_FLOAT", 0x0A803721, 16, 4, (0x8E8E, 0, 0), None, 96),
    ("BPTC", 0x0B800334, 16, 4, (0x8E8C, 0, 0), None, 98),
    ("ETC1", 0x0C000320, 8, 4, (0x8D64, 0, 0), None, None),
    ("ETC2", 0x0D000320, 8, 4, (0x9274, 0, 0), None, None),
    ("ETC2_PUNCHTHROUGH", 0x0E000334, 8, 4, (0x9275, 0, 0), None, None),
    ("ETC2_EAC", 0x0F800334, 16, 4, (0x9278, 0, 0), None, None),
    ("EAC_R11", 0x10000101, 8, 4, (0x9270, 0, 0), None, None),
    ("EAC_SIGNED_R11", 0x11001101, 8, 4, (0x9271, 0, 0), None, None),
    ("EAC_RG11", 0x12800311, 16, 4, (0x9272, 0, 0), None, None),
    ("EAC_SIGNED_RG11", 0x13801311, 16, 4, (0x9273, 0, 0), None, None),
    ("RGBA8", 0x334, 4, 1, (0x1908, 0x1908, 0x1401), None, 28),
    ("RGB8", 0x220, 3, 1, (0x1907, 0x1907, 0x1401), None, None),
    ("R16", 0x101, 2, 1, (0x822A, 0x1903, 0x1403), None, 56),
    ("FLOAT_RGBA16", 0x6735, 8, 1, (0x881A, 0x1908, 0x140B), None, 10),      # GL_RGBA16F synonym
    ("FLOAT_RGB32", 0x2B22, 12, 1, (0x8815, 0x1907, 0x1406), None, 6),
]


def _lib():
    lib = D.lib.load()
    tpp = C.POINTER(C.POINTER(C.POINTER(H.DetexTexture)))
    for name in ("detexLoadKTXFileWithMipmaps", "detexLoadDDSFileWithMipmaps", "detexLoadTextureFileWithMipmaps"):
        getattr(lib, name).argtypes = [C.c_char_p, C.c_int, tpp, C.POINTER(C.c_int)]
        getattr(lib, name).restype = C.c_bool
    for name in ("detexLoadKTXFile", "detexLoadDDSFile", "detexLoadTextureFile"):
        getattr(lib, name).argtypes = [C.c_char_p, C.POINTER(C.POINTER(H.DetexTexture))]
        getattr(lib, name).restype = C.c_bool
    return lib


def _ref():
    if not H.have_reference():
        return None
    ref = H.load_reference()
    tpp = C.POINTER(C.POINTER(C.POINTER(H.DetexTexture)))
    for name in ("detexLoadKTXFileWithMipmaps", "detexLoadDDSFileWithMipmaps", "detexLoadTextureFileWithMipmaps"):
        getattr(ref, name).argtypes = [C.c_char_p, C.c_int, tpp, C.POINTER(C.c_int)]
        getattr(ref, name).restype = C.c_bool
    return ref


def _load(lib, fn, path, max_levels):
    """-> list of (format, width, height, wib, hib, payload bytes) or the error text."""
    texs = C.POINTER(C.POINTER(H.DetexTexture))()
    n = C.c_int(0)
    if not getattr(lib, fn)(path.encode(), max_levels, C.byref(texs), C.byref(n)):
        return lib.detexGetErrorMessage().decode()
    out = []
    for i in range(n.value):
        t = texs[i].contents
        unit = (8 + ((t.format & 0x00800000) >> 20)) if (t.format >> 24) else 1 + ((t.format & 0xF00) >> 8)
        size = t.width_in_blocks * t.height_in_blocks * unit
        data = bytes(C.cast(t.data, C.POINTER(C.c_uint8 * max(size, 1))).contents)[:size]
        out.append((t.format, t.width, t.height, t.width_in_blocks, t.height_in_blocks, data))
    return out


def _payloads(rng, width, height, unit, block, levels):
    dims = CF.level_dims(width, height, block, levels)
    return dims, [rng.integers(0, 256, wib * hib * unit, dtype=np.uint8).tobytes() for (_, _, wib, hib) in dims]


@pytest.mark.parametrize("entry", FILE_FORMATS, ids=[e[0] for e in FILE_FORMATS])
def test_ktx_round_trip_all_formats(entry, tmp_path):
    name, fmt, unit, block, (gl_i, gl_f, gl_t), _, _ = entry
    lib, ref = _lib(), _ref()
    rng = np.random.default_rng(len(name))
    for (w, h, levels, meta, be) in ((64, 64, 1, b"", False), (100, 36, 5, b"KTXorientation\0S=r,T=d\0\0" * 1, False),
                                    (17, 9, 6, b"", True)):
        dims, pay = _payloads(rng, w, h, unit, block, levels)
        path = str(tmp_path / ("t_%s_%d.ktx" % (name, w)))
        meta_block = (len(meta).to_bytes(4, "big" if be else "little") + meta) if meta else b""
        CF.write_ktx(path, gl_i, gl_f, gl_t, w, h, unit, block, pay, metadata=meta_block, big_endian=be)
        want = [(fmt, d[0], d[1], d[2], d[3], p) for d, p in zip(dims, pay)]
        for max_levels in (1, 3, 32):
            got = _load(lib, "detexLoadKTXFileWithMipmaps", path, max_levels)
            assert got == want[:max_levels], (name, w, max_levels, got if isinstance(got, str) else "payload/dims differ")
            assert _load(lib, "detexLoadTextureFileWithMipmaps", path, max_levels) == got
            if ref is not None:
                assert _load(ref, "detexLoadKTXFileWithMipmaps", path, max_levels) == got


@pytest.mark.parametrize("entry", [e for e in FILE_FORMATS if e[5] or e[6]], ids=[e[0] for e in FILE_FORMATS if e[5] or e[6]])
def test_dds_round_trip(entry, tmp_path):
    name, fmt, unit, block, _, four_cc, dxgi = entry
    lib, ref = _lib(), _ref()
    rng = np.random.default_rng(len(name) + 100)
    variants = []
    if four_cc:
        variants.append(dict(four_cc=four_cc))
    if dxgi is not None:
        variants.append(dict(four_cc=b"DX10", dx10_format=dxgi))
    for k, v in enumerate(variants):
        for (w, h, levels) in ((64, 32, 1), (40, 24, 4)):
            dims, pay = _payloads(rng, w, h, unit, block, levels)
            path = str(tmp_path / ("t_%s_%d_%d.dds" % (name, k, w)))
            CF.write_dds(path, w, h, pay, **v)
            want = [(fmt, d[0], d[1], d[2], d[3], p) for d, p in zip(dims, pay)]
            for max_levels in (1, 32):
                got = _load(lib, "detexLoadDDSFileWithMipmaps", path, max_levels)
                assert got == want[:max_levels], (name, v, w, got if isinstance(got, str) else "differs")
                if ref is not None:
                    assert _load(ref, "detexLoadDDSFileWithMipmaps", path, max_levels) == got


def test_dds_uncompressed_masks_and_synonyms(tmp_path):
    lib, ref = _lib(), _ref()
    rng = np.random.default_rng(5)
    cases = [
        (0x334, 4, dict(pf_flags=0x41, bitcount=32, masks=(0xFF, 0xFF00, 0xFF0000, 0xFF000000))),     # RGBA8 by masks
        (0x328, 4, dict(pf_flags=0x40, bitcount=32, masks=(0xFF0000, 0xFF00, 0xFF, 0))),               # BGRX8 by masks
        (0x220, 3, dict(pf_flags=0x40, bitcount=24, masks=(0xFF, 0xFF00, 0xFF0000, 0))),               # RGB8
        (0x4, 1, dict(pf_flags=0x2, bitcount=8, masks=(0, 0, 0, 0xFF))),                               # old alpha format
        (0x334, 4, dict(four_cc=b"DX10", dx10_format=27)),                                             # synonym
        (0x01000320, 8, dict(four_cc=b"DX10", dx10_format=70)),                                        # BC1 typeless
        (0x05000000, 8, dict(four_cc=b"BC4U")),
        (0x09802721, 16, dict(four_cc=b"DX10", dx10_format=94)),
    ]
    for k, (fmt, unit, kw) in enumerate(cases):
        block = 4 if fmt >> 24 else 1
        dims, pay = _payloads(rng, 20, 12, unit, block, 2)
        path = str(tmp_path / ("m%d.dds" % k))
        CF.write_dds(path, 20, 12, pay, **kw)
        got = _load(lib, "detexLoadDDSFileWithMipmaps", path, 8)
        assert got == [(fmt, d[0], d[1], d[2], d[3], p) for d, p in zip(dims, pay)], (k, got if isinstance(got, str) else "")
        if ref is not None:
            assert _load(ref, "detexLoadDDSFileWithMipmaps", path, 8) == got


def test_error_cases_match_reference_messages(tmp_path):
    lib, ref = _lib(), _ref()
    rng = np.random.default_rng(9)
    dims, pay = _payloads(rng, 32, 32, 8, 4, 3)
    files = {}
    p = str(tmp_path / "good.ktx")
    CF.write_ktx(p, 0x83F0, 0, 0, 32, 32, 8, 4, pay)
    raw = open(p, "rb").read()
    files["nosig.ktx"] = b"XXXX" + raw[4:]
    files["short_header.ktx"] = raw[:40]
    files["truncated.ktx"] = raw[:-20]
    files["unsupported.ktx"] = raw[:28] + (0x1234).to_bytes(4, "little") + raw[32:]
    p2 = str(tmp_path / "badsize.ktx")
    CF.write_ktx(p2, 0x83F0, 0, 0, 32, 32, 8, 4, pay, bad_size_level=1)
    files["badsize.ktx"] = open(p2, "rb").read()
    pd = str(tmp_path / "good.dds")
    CF.write_dds(pd, 32, 32, pay, four_cc=b"DXT1")
    rawd = open(pd, "rb").read()
    files["nosig.dds"] = b"DDX " + rawd[4:]
    files["truncated.dds"] = rawd[:-8]
    files["unsupported.dds"] = rawd[:84] + b"ZZZZ" + rawd[88:]
    p3 = str(tmp_path / "cube.dds")
    CF.write_dds(p3, 32, 32, pay, four_cc=b"DX10", dx10_format=71, resource_dimension=4)
    files["cube.dds"] = open(p3, "rb").read()
    for name, content in files.items():
        path = str(tmp_path / ("e_" + name))
        open(path, "wb").write(content)
        got = _load(lib, "detexLoadTextureFileWithMipmaps", path, 8)
        assert isinstance(got, str) and got, name
        if ref is not None:
            assert _load(ref, "detexLoadTextureFileWithMipmaps", path, 8) == got, name
    for lb in (lib, ref):
        if lb is None:
            continue
        assert _load(lb, "detexLoadTextureFileWithMipmaps", str(tmp_path / "missing.ktx"), 1).endswith("missing.ktx")
        assert _load(lb, "detexLoadTextureFileWithMipmaps", str(tmp_path / "file.png"), 1) == \
            "detexLoadTextureFileWithMipmaps: Do not recognize filename extension"


def test_single_level_entry_points(tmp_path):
    lib = _lib()
    rng = np.random.default_rng(3)
    dims, pay = _payloads(rng, 24, 24, 16, 4, 3)
    pk, pd = str(tmp_path / "a.KTX"), str(tmp_path / "a.Dds")      # extension match is case-insensitive
    CF.write_ktx(pk, 0x8E8C, 0, 0, 24, 24, 16, 4, pay)
    CF.write_dds(pd, 24, 24, pay, four_cc=b"DX10", dx10_format=98)
    for fn, path in (("detexLoadKTXFile", pk), ("detexLoadDDSFile", pd), ("detexLoadTextureFile", pk), ("detexLoadTextureFile", pd)):
        t = C.POINTER(H.DetexTexture)()
        assert getattr(lib, fn)(path.encode(), C.byref(t)), fn
        tx = t.contents
        assert (tx.format, tx.width, tx.height, tx.width_in_blocks, tx.height_in_blocks) == (0x0B800334, 24, 24, 6, 6)
        assert bytes(C.cast(tx.data, C.POINTER(C.c_uint8 * len(pay[0]))).contents) == pay[0]


def test_crafted_headers_are_rejected_not_trusted(tmp_path):
    """Sizes in a KTX / DDS header are untrusted input: dimensions whose block count overflows 32-bit
    arithmetic, negative dimensions and negative level counts must end in an error return -- never in a huge or
    wrapped allocation, a backwards seek or a crash.  (Only the product is called: the reference computes these
    sizes in `int` and is not safe to run on such files.)"""
    lib = _lib()
    rng = np.random.default_rng(5)
    dims, pay = _payloads(rng, 32, 32, 8, 4, 1)
    pk, pd = str(tmp_path / "good.ktx"), str(tmp_path / "good.dds")
    CF.write_ktx(pk, 0x83F0, 0, 0, 32, 32, 8, 4, pay)
    CF.write_dds(pd, 32, 32, pay, four_cc=b"DXT1")
    ktx, dds = open(pk, "rb").read(), open(pd, "rb").read()

    def u32(v):
        return (v & 0xFFFFFFFF).to_bytes(4, "little")

    def ktx_with(width=None, height=None, levels=None, kv=None, image_size=None):
        b = bytearray(ktx)
        for off, v in ((36, width), (40, height), (56, levels), (60, kv)):     # header words 9, 10, 14, 15
            if v is not None:
                b[off:off + 4] = u32(v)
        if image_size is not None:
            b[64:68] = u32(image_size)
        return bytes(b)

    def dds_with(width=None, height=None, levels=None):
        b = bytearray(dds)
        for off, v in ((16, width), (12, height)):      # dwWidth at 4 + 12, dwHeight at 4 + 8
            if v is not None:
                b[off:off + 4] = u32(v)
        if levels is not None:
            b[8:12] = u32(int.from_bytes(b[8:12], "little") | 0x20000)
            b[28:32] = u32(levels)
        return bytes(b)

    cases = {
        # wib * hib = 65536 * 65535 wraps negative in int; times 8 bytes it wraps again and can match image_size
        "wrap.ktx": ktx_with(width=262144, height=262140, image_size=(65536 * 65535 * 8) & 0xFFFFFFFF),
        "huge.ktx": ktx_with(width=0x7FFFFFFF, height=0x7FFFFFFF, image_size=8),
        "negw.ktx": ktx_with(width=-32, image_size=len(pay[0])),
        "negh.ktx": ktx_with(height=-4),
        "neglevels.ktx": ktx_with(levels=-1),
        "negkv.ktx": ktx_with(kv=-64),
        "hugekv.ktx": ktx_with(kv=0x7FFFFFF0),
        "wrap.dds": dds_with(width=262144, height=262140),
        "huge.dds": dds_with(width=0x7FFFFFFF, height=0x7FFFFFFF),
        "negw.dds": dds_with(width=-32),
        "negh.dds": dds_with(height=-4),
        "manylevels.dds": dds_with(levels=0x7FFFFFFF),
    }
    for name, content in cases.items():
        path = str(tmp_path / name)
        open(path, "wb").write(content)
        got = _load(lib, "detexLoadTextureFileWithMipmaps", path, 16)
        if name == "neglevels.ktx":
            assert got == [] or isinstance(got, str), name      # zero levels: nothing loaded, nothing trusted
        else:
            assert isinstance(got, str) and got, (name, got if isinstance(got, str) else len(got))
    # the untouched files still load
    assert len(_load(lib, "detexLoadTextureFileWithMipmaps", pk, 16)) == 1
    assert len(_load(lib, "detexLoadTextureFileWithMipmaps", pd, 16)) == 1
