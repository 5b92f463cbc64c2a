"""TEST / BENCH INFRASTRUCTURE ONLY -- the checker, never the product.

ctypes loaders for the CPU decoders that pin parity:

* ``load_reference()``  -> oracle/_ref/libdetex_ref.so, the UNMODIFIED reference
  compiled by oracle/Makefile from the sources under /root/reference (prebuilt
  file travels to the GPU box; it is never rebuilt there).
* ``load_oracle()``     -> oracle/liboracle.so, the C restatement
  (oracle/detex_oracle.c); buildable anywhere gcc is.
* ``load_harness()``    -> oracle/libcpu_harness.so: pthread row-sharding driver,
  FNV-1a-64 and the splitmix64 block generator of SURVEY.md 8c.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libdetex_ref.so")
ORACLE_SO = os.path.join(HERE, "liboracle.so")
HARNESS_SO = os.path.join(HERE, "libcpu_harness.so")

# ---------------------------------------------------------------------------
# Format table.  Values are those of the reference's enums (detex.h:83-379 for
# pixel formats, detex.h:575-727 for texture formats): texture format =
# (index << 24) | 128-bit-block flag 0x00800000 | native pixel format.
# name -> (texture_format, block_bytes, pixel_bytes)
# ---------------------------------------------------------------------------
PF = {
    "RGBA8": 0x334, "BGRA8": 0x33C, "RGBX8": 0x320, "BGRX8": 0x328,
    "RGB8": 0x220, "BGR8": 0x228, "R8": 0x000, "SIGNED_R8": 0x1000,
    "RG8": 0x110, "SIGNED_RG8": 0x1110, "R16": 0x101, "SIGNED_R16": 0x1101,
    "RG16": 0x311, "SIGNED_RG16": 0x1311, "RGB16": 0x521, "RGBX16": 0x721,
    "RGBA16": 0x735, "FLOAT_RGBX16": 0x2721, "SIGNED_FLOAT_RGBX16": 0x3721,
    "FLOAT_RGB16": 0x2521,
}

_FMT_ROWS = [
    # name, index, 128-bit, native pixel format
    ("BC1", 1, 0, "RGBX8"), ("BC1A", 2, 0, "RGBA8"), ("BC2", 3, 1, "RGBA8"),
    ("BC3", 4, 1, "RGBA8"), ("RGTC1", 5, 0, "R8"), ("SIGNED_RGTC1", 6, 0, "SIGNED_R16"),
    ("RGTC2", 7, 1, "RG8"), ("SIGNED_RGTC2", 8, 1, "SIGNED_RG16"),
    ("BPTC_FLOAT", 9, 1, "FLOAT_RGBX16"), ("BPTC_SIGNED_FLOAT", 10, 1, "SIGNED_FLOAT_RGBX16"),
    ("BPTC", 11, 1, "RGBA8"), ("ETC1", 12, 0, "RGBX8"), ("ETC2", 13, 0, "RGBX8"),
    ("ETC2_PUNCHTHROUGH", 14, 0, "RGBA8"), ("ETC2_EAC", 15, 1, "RGBA8"),
    ("EAC_R11", 16, 0, "R16"), ("EAC_SIGNED_R11", 17, 0, "SIGNED_R16"),
    ("EAC_RG11", 18, 1, "RG16"), ("EAC_SIGNED_RG11", 19, 1, "SIGNED_RG16"),
]


def pixel_size(pf):
    return 1 + ((pf & 0xF00) >> 8)


FORMATS = {}
for _n, _i, _w, _p in _FMT_ROWS:
    _tf = (_i << 24) | (0x00800000 if _w else 0) | PF[_p]
    FORMATS[_n] = (_tf, 16 if _w else 8, pixel_size(PF[_p]))
FORMAT_NAMES = [r[0] for r in _FMT_ROWS]


class DetexTexture(C.Structure):
    """Same layout as the reference's detexTexture (detex.h:729-736)."""
    _fields_ = [("format", C.c_uint32), ("data", C.c_void_p), ("width", C.c_int),
                ("height", C.c_int), ("width_in_blocks", C.c_int),
                ("height_in_blocks", C.c_int)]


def _build(target=None):
    cmd = ["make", "-s", "-C", HERE] + ([target] if target else ["all"])
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)


def _bind_detex_api(lib):
    blk = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    for n in FORMAT_NAMES:
        f = getattr(lib, "detexDecompressBlock" + n)
        f.argtypes, f.restype = blk, C.c_bool
    lib.detexDecompressBlock.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                         C.c_void_p, C.c_uint32]
    lib.detexDecompressBlock.restype = C.c_bool
    for n in ("detexDecompressTextureLinear", "detexDecompressTextureTiled"):
        f = getattr(lib, n)
        f.argtypes, f.restype = [C.POINTER(DetexTexture), C.c_void_p, C.c_uint32], C.c_bool
    lib.detexConvertPixels.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32]
    lib.detexConvertPixels.restype = C.c_bool
    lib.detexConvertPixelsInPlace.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
    lib.detexConvertPixelsInPlace.restype = C.c_bool
    lib.detexGetErrorMessage.argtypes, lib.detexGetErrorMessage.restype = [], C.c_char_p
    lib.detexSetHDRParameters.argtypes = [C.c_float, C.c_float, C.c_float]
    lib.detexSetHDRParameters.restype = None
    if hasattr(lib, "oracle_conversion_chain"):
        lib.oracle_conversion_chain.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
        lib.oracle_conversion_chain.restype = C.c_int
    return lib


# ---- pixel-format conversion helpers (tests only) -------------------------------------------
PIXEL_FORMATS = {
    "RGBA8": 0x334, "BGRA8": 0x33C, "RGBX8": 0x320, "BGRX8": 0x328, "RGB8": 0x220,
    "R8": 0x0, "SIGNED_R8": 0x1000, "RG8": 0x110, "SIGNED_RG8": 0x1110,
    "R16": 0x101, "SIGNED_R16": 0x1101, "RG16": 0x311, "SIGNED_RG16": 0x1311,
    "RGB16": 0x521, "RGBX16": 0x721, "RGBA16": 0x735,
    "FLOAT_R16": 0x2101, "FLOAT_R16_HDR": 0x6101, "FLOAT_RG16": 0x2311, "FLOAT_RG16_HDR": 0x6311,
    "FLOAT_RGBX16": 0x2721, "FLOAT_RGBX16_HDR": 0x6721, "FLOAT_RGBA16": 0x6735,
    "FLOAT_RGB16": 0x2521, "FLOAT_RGB16_HDR": 0x6521, "FLOAT_BGRX16": 0x2729,
    "FLOAT_R32": 0x2302, "FLOAT_R32_HDR": 0x6302, "FLOAT_RG32": 0x2712, "FLOAT_RG32_HDR": 0x6712,
    "FLOAT_RGB32": 0x2B22, "FLOAT_RGB32_HDR": 0x6B22, "FLOAT_RGBX32": 0x2F22, "FLOAT_RGBX32_HDR": 0x6F22,
    # in no table entry: conversions from / to these must fail
    "BGR8": 0x228, "A8": 0x4, "FLOAT_RGBA32": 0x2F36,
}
# (source, target) table entries whose reference implementation dereferences an uninitialised
# pointer (convert.c:704: `uint16_t *target_pixel16_buffer = (uint16_t *)target_pixel16_buffer;`):
# never call the compiled reference on a chain that contains one of them.
REFERENCE_UB_STEPS = {(0x2721, 0x2521), (0x6721, 0x6521)}


def pixel_size(pf):
    return 1 + ((pf & 0xF00) >> 8)


def synth_pixels(pf, n, seed=1):
    """n pixels of format pf: raw random bits for integer formats; for float formats a mix of
    in-range, out-of-range, tiny, huge, zero and infinite values (no NaNs: the reference is built
    with -ffinite-math-only and its NaN behaviour is unspecified)."""
    rng = np.random.default_rng(seed)
    nbytes = n * pixel_size(pf)
    if not (pf & 0x2000):
        return rng.integers(0, 256, nbytes, dtype=np.uint8)
    if (pf & 3) == 2:
        m = nbytes // 4
        v = rng.uniform(-0.5, 1.5, m).astype(np.float32)
        k = rng.integers(0, 16, m)
        v = np.where(k == 0, rng.uniform(-70000, 70000, m), v).astype(np.float32)
        v = np.where(k == 1, rng.uniform(-1e-4, 1e-4, m), v).astype(np.float32)
        v = np.where(k == 2, rng.uniform(-1e-7, 1e-7, m), v).astype(np.float32)
        v = np.where(k == 3, np.float32(0.0), v)
        v = np.where(k == 4, np.float32(-0.0), v)
        v = np.where(k == 5, np.float32(np.inf), v)
        v = np.where(k == 6, np.float32(-np.inf), v)
        v = np.where(k == 7, np.float32(1e-41), v).astype(np.float32)
        v = np.where(k == 8, rng.integers(0, 65536, m).astype(np.float32) / np.float32(65535.0), v).astype(np.float32)
        return v.view(np.uint8).copy()
    h = rng.integers(0, 65536, nbytes // 2).astype(np.uint16)
    nan = ((h & 0x7C00) == 0x7C00) & ((h & 0x3FF) != 0)
    h = np.where(nan, h & 0xFC00, h).astype(np.uint16)        # NaN -> infinity of the same sign
    return h.view(np.uint8).copy()


_libm = None


def reset_rounding_mode():
    """The reference leaves the calling thread in FE_DOWNWARD after some conversions
    (half-float.c:306,317; hdr.c:122,171); put it back to nearest before the next call."""
    global _libm
    if _libm is None:
        _libm = C.CDLL("libm.so.6")
    _libm.fesetround(0)


def convert_pixels(lib, src, n, spf, tpf, primed=False):
    """detexConvertPixels through lib; returns (ok, uint8 array of n target pixels).
    primed=True runs the call with the thread in FE_DOWNWARD, the state the reference leaves
    behind after any float -> uint16 conversion (half-float.c:306)."""
    reset_rounding_mode()
    if primed:
        _libm.fesetround(0x400)
    src = np.ascontiguousarray(src, dtype=np.uint8).copy()
    out = np.zeros(n * pixel_size(tpf) + 64, dtype=np.uint8)
    ok = lib.detexConvertPixels(src.ctypes.data, n, spf, out.ctypes.data, tpf)
    reset_rounding_mode()
    return bool(ok), out[: n * pixel_size(tpf)]


def conversion_chain(oracle, spf, tpf):
    f = (C.c_uint32 * 5)()
    n = oracle.oracle_conversion_chain(spf, tpf, f)
    return None if n < 0 else [int(f[i]) for i in range(n + 1)]


_cache = {}


def have_reference():
    return os.path.exists(REF_SO)


def load_reference():
    """The compiled, unmodified reference (None when oracle/_ref is absent)."""
    if "ref" not in _cache:
        if not os.path.exists(REF_SO) and os.path.exists("/root/reference/detex.h"):
            _build("ref")
        _cache["ref"] = _bind_detex_api(C.CDLL(REF_SO)) if os.path.exists(REF_SO) else None
    return _cache["ref"]


def load_oracle():
    if "oracle" not in _cache:
        src = os.path.join(HERE, "detex_oracle.c")
        if (not os.path.exists(ORACLE_SO)
                or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src)):
            _build(ORACLE_SO)
        _cache["oracle"] = _bind_detex_api(C.CDLL(ORACLE_SO))
    return _cache["oracle"]


def load_harness():
    if "harness" not in _cache:
        src = os.path.join(HERE, "cpu_harness.c")
        if (not os.path.exists(HARNESS_SO)
                or os.path.getmtime(HARNESS_SO) < os.path.getmtime(src)):
            _build(HARNESS_SO)
        h = C.CDLL(HARNESS_SO)
        h.cpu_harness_decode.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_int,
                                         C.c_int, C.POINTER(C.c_double)]
        h.cpu_harness_decode.restype = C.c_int
        h.cpu_harness_fnv1a64.argtypes = [C.c_void_p, C.c_size_t]
        h.cpu_harness_fnv1a64.restype = C.c_uint64
        h.cpu_harness_splitmix64.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
        h.cpu_harness_splitmix64.restype = None
        h.cpu_harness_tree_hash.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
        h.cpu_harness_tree_hash.restype = C.c_uint64
        _cache["harness"] = h
    return _cache["harness"]


def best_cpu_decoder():
    """(lib, kind): the compiled reference when present, else the restatement."""
    ref = load_reference()
    if ref is not None:
        return ref, "reference"
    return load_oracle(), "port"


# ---------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------
def fnv1a64(buf):
    a = np.ascontiguousarray(buf).view(np.uint8).reshape(-1)
    return int(load_harness().cpu_harness_fnv1a64(a.ctypes.data, a.size))


def tree_hash(buf, seg=4096, threads=None):
    """CPU statement of detexb200HashDevice (detex_b200/csrc/aux.cu): FNV-1a-64 every
    4096-byte segment, treat the little-endian array of segment hashes as the next buffer,
    repeat until one hash is left.  threads=None: the numpy statement below; threads=N: the same
    construction in C on N host threads (cpu_harness.c, for multi-gigabyte buffers)."""
    a = np.ascontiguousarray(buf).view(np.uint8).reshape(-1)
    if threads is not None and seg == 4096:
        return int(load_harness().cpu_harness_tree_hash(a.ctypes.data, a.size, int(threads)))
    while True:
        n = a.size
        n_seg = max(1, (n + seg - 1) // seg)
        full = n // seg
        h = np.full(n_seg, 0xcbf29ce484222325, dtype=np.uint64)
        prime = np.uint64(0x100000001b3)
        with np.errstate(over="ignore"):
            if full:
                m = a[:full * seg].reshape(full, seg)
                for j in range(seg):
                    h[:full] = (h[:full] ^ m[:, j].astype(np.uint64)) * prime
            for j in range(full * seg, n):          # ragged last segment
                h[-1] = (h[-1] ^ np.uint64(a[j])) * prime
        if n_seg == 1:
            return int(h[0])
        a = h.view(np.uint8)


def synth_blocks(fmt, n_blocks, seed=1, first_block=0):
    """Distribution U of SURVEY.md 8d: raw splitmix64 words, uint8[n_blocks*bs]."""
    bs = FORMATS[fmt][1]
    wpb = bs // 8
    out = np.empty(n_blocks * wpb, dtype=np.uint64)
    load_harness().cpu_harness_splitmix64(seed, first_block * wpb, out.size, out.ctypes.data)
    return out.view(np.uint8)


BC6H_MODE_CODES = [0, 1, 2, 6, 10, 14, 18, 22, 26, 30, 3, 7, 11, 15]


def force_modes(fmt, blocks, modes=None):
    """Distribution M of SURVEY.md 8d: overwrite the mode field so block i has
    mode modes[i % len(modes)] (BC7: 8 modes, unary prefix; BC6H: 14 modes, the
    2/5-bit codes of decompress-bptc-float.c:660-662).  Returns a new array."""
    b = blocks.copy().reshape(-1, FORMATS[fmt][1])
    n = b.shape[0]
    if fmt == "BPTC":
        modes = list(range(8)) if modes is None else modes
        m = np.array(modes, dtype=np.uint32)[np.arange(n) % len(modes)]
        keep = (0xFF << (m + 1)) & 0xFF
        b[:, 0] = ((b[:, 0] & keep) | (1 << m)).astype(np.uint8)
    elif fmt in ("BPTC_FLOAT", "BPTC_SIGNED_FLOAT"):
        modes = list(range(14)) if modes is None else modes
        m = np.array(modes, dtype=np.uint32)[np.arange(n) % len(modes)]
        code = np.array(BC6H_MODE_CODES, dtype=np.uint32)[m]
        two_bit = m < 2
        b0 = b[:, 0].astype(np.uint32)
        b0 = np.where(two_bit, (b0 & 0xFC) | code, (b0 & 0xE0) | code)
        b[:, 0] = b0.astype(np.uint8)
    else:
        raise ValueError("no mode forcing for " + fmt)
    return b.reshape(-1)


def cpu_decode(lib, fmt, blocks, width, height, linear=True, pixel_format=None, threads=1, out=None):
    """Decode with a CPU library through detexDecompressTexture{Linear,Tiled}.
    Returns (pixels uint8[], ok, seconds).  out: a caller-owned (warm) output buffer to reuse."""
    tf, bs, ps_native = FORMATS[fmt]
    pf = (tf & 0xFFFF) if pixel_format is None else pixel_format
    ps = pixel_size(pf)
    wib, hib = (width + 3) // 4, (height + 3) // 4
    blocks = np.ascontiguousarray(blocks).view(np.uint8).reshape(-1)
    assert blocks.size == wib * hib * bs
    n_out = width * height * ps if linear else wib * hib * 16 * ps
    if out is None:
        out = np.zeros(n_out, dtype=np.uint8)
    assert out.dtype == np.uint8 and out.size == n_out and out.flags["C_CONTIGUOUS"]
    fn = lib.detexDecompressTextureLinear if linear else lib.detexDecompressTextureTiled
    secs = C.c_double(0)
    ok = load_harness().cpu_harness_decode(C.cast(fn, C.c_void_p), tf, blocks.ctypes.data,
                                           width, height, wib, hib, out.ctypes.data, pf,
                                           1 if linear else 0, threads, C.byref(secs))
    return out, bool(ok), secs.value
