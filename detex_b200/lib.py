"""ctypes host-side mirror of the detex C API served by libdetex_b200.so.

Function names and argument meaning follow reference detex.h (:435-531 per-block
decoders, :747-765 texture-level API, :777-784 conversion, :806 error message).
numpy arrays stand in for the caller-owned host buffers of the C API.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# DETEXB200_LIB selects another build of the same library (kernel A/B measurements)
LIB_PATH = os.environ.get("DETEXB200_LIB") or os.path.join(HERE, "libdetex_b200.so")

LAYOUT_TILED, LAYOUT_LINEAR = 0, 1

# reference detex.h:83-379
PIXEL_FORMATS = {
    "RGBA8": 0x334, "BGRA8": 0x33C, "RGBX8": 0x320, "BGRX8": 0x328, "RGB8": 0x220, "BGR8": 0x228,
    "R8": 0x000, "SIGNED_R8": 0x1000, "RG8": 0x110, "SIGNED_RG8": 0x1110, "R16": 0x101,
    "SIGNED_R16": 0x1101, "RG16": 0x311, "SIGNED_RG16": 0x1311, "RGB16": 0x521, "RGBX16": 0x721,
    "RGBA16": 0x735, "FLOAT_RGBX16": 0x2721, "SIGNED_FLOAT_RGBX16": 0x3721, "FLOAT_RGB16": 0x2521,
    "FLOAT_BGRX16": 0x2729, "SIGNED_FLOAT_BGRX16": 0x3729,
}
# reference detex.h:575-727: name -> texture format value
FORMATS = {
    "BC1": 0x01000320, "BC1A": 0x02000334, "BC2": 0x03800334, "BC3": 0x04800334,
    "RGTC1": 0x05000000, "SIGNED_RGTC1": 0x06001101, "RGTC2": 0x07800110,
    "SIGNED_RGTC2": 0x08801311, "BPTC_FLOAT": 0x09802721, "BPTC_SIGNED_FLOAT": 0x0A803721,
    "BPTC": 0x0B800334, "ETC1": 0x0C000320, "ETC2": 0x0D000320, "ETC2_PUNCHTHROUGH": 0x0E000334,
    "ETC2_EAC": 0x0F800334, "EAC_R11": 0x10000101, "EAC_SIGNED_R11": 0x11001101,
    "EAC_RG11": 0x12800311, "EAC_SIGNED_RG11": 0x13801311,
}
FORMAT_NAMES = list(FORMATS)


def pixel_size(pixel_format):            # detex.h:879-881
    return 1 + ((pixel_format & 0xF00) >> 8)


def block_size(texture_format):          # detex.h:918-920
    return 8 + ((texture_format & 0x00800000) >> 20)


def native_pixel_format(texture_format):  # detex.h:928-930
    return texture_format & 0xFFFF


def _tf(fmt):
    return FORMATS[fmt] if isinstance(fmt, str) else int(fmt)


class DetexError(RuntimeError):
    pass


class DetexTexture(C.Structure):
    """detexTexture, reference detex.h:729-736."""
    _fields_ = [("format", C.c_uint32), ("data", C.c_void_p), ("width", C.c_int),
                ("height", C.c_int), ("width_in_blocks", C.c_int), ("height_in_blocks", C.c_int)]


_lib = None


def load():
    """Load libdetex_b200.so (building it first when a toolchain is present).
    Fails loudly: there is no Python or CPU fallback for the decode path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    blk = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    for n in FORMAT_NAMES:
        f = getattr(lib, "detexDecompressBlock" + n)
        f.argtypes, f.restype = blk, C.c_bool
    lib.detexDecompressBlock.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32]
    lib.detexDecompressBlock.restype = C.c_bool
    for n in ("detexDecompressTextureLinear", "detexDecompressTextureTiled"):
        f = getattr(lib, n)
        f.argtypes, f.restype = [C.POINTER(DetexTexture), C.c_void_p, C.c_uint32], C.c_bool
    lib.detexConvertPixels.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32]
    lib.detexConvertPixels.restype = C.c_bool
    lib.detexConvertPixelsInPlace.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
    lib.detexConvertPixelsInPlace.restype = C.c_bool
    lib.detexGetErrorMessage.argtypes, lib.detexGetErrorMessage.restype = [], C.c_char_p
    lib.detexb200DecodeDevice.argtypes = [C.c_uint32, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32,
                                          C.c_void_p, C.c_void_p]
    lib.detexb200DecodeDevice.restype = C.c_int
    lib.detexb200DecodeHDRDevice.argtypes = [C.c_uint32, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_void_p, C.c_uint32, C.c_int, C.c_float, C.c_float, C.c_float,
                                             C.c_void_p, C.c_void_p]
    lib.detexb200DecodeHDRDevice.restype = C.c_int
    lib.detexb200ConvertPixelsDevice.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
    lib.detexb200ConvertPixelsDevice.restype = C.c_int
    lib.detexb200GenerateBlocksDevice.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
    lib.detexb200GenerateBlocksDevice.restype = C.c_int
    lib.detexb200HashDevice.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p]
    lib.detexb200HashDevice.restype = C.c_int
    lib.detexb200DeviceCount.argtypes, lib.detexb200DeviceCount.restype = [], C.c_int
    lib.detexb200Version.argtypes, lib.detexb200Version.restype = [], C.c_char_p
    lib.detexb200LaunchCount.argtypes, lib.detexb200LaunchCount.restype = [], C.c_uint64
    lib.detexb200SetDevices.argtypes, lib.detexb200SetDevices.restype = [C.c_int, C.POINTER(C.c_int)], C.c_int
    lib.detexb200SetHostThreads.argtypes, lib.detexb200SetHostThreads.restype = [C.c_int], C.c_int
    lib.detexb200ReleaseThreadResources.argtypes, lib.detexb200ReleaseThreadResources.restype = [], None
    _lib = lib
    return lib


def detexGetErrorMessage():
    m = load().detexGetErrorMessage()
    return m.decode() if m else ""


def _host_u8(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint8).reshape(-1)


# ---- reference host-pointer API ---------------------------------------------------------
def detexDecompressBlockFormat(fmt, bitstring, mode_mask=0xFFFFFFFF, flags=0):
    """detexDecompressBlock<FMT>(bitstring, mode_mask, flags, pixel_buffer) -> (pixels, ok)."""
    tf = _tf(fmt)
    name = fmt if isinstance(fmt, str) else [k for k, v in FORMATS.items() if v == tf][0]
    b = _host_u8(bitstring)
    out = np.zeros(16 * pixel_size(tf & 0xFFFF), dtype=np.uint8)
    ok = getattr(load(), "detexDecompressBlock" + name)(b.ctypes.data, mode_mask, flags, out.ctypes.data)
    return out, bool(ok)


def detexDecompressBlock(bitstring, texture_format, mode_mask, flags, pixel_format):
    """detexDecompressBlock (texture.c:55-70) -> (pixels, ok)."""
    tf = _tf(texture_format)
    b = _host_u8(bitstring)
    out = np.zeros(16 * pixel_size(pixel_format), dtype=np.uint8)
    ok = load().detexDecompressBlock(b.ctypes.data, tf, mode_mask, flags, out.ctypes.data, pixel_format)
    return out, bool(ok)


def _texture_call(fn_name, fmt, blocks, width, height, pixel_format, linear, out=None):
    tf = _tf(fmt)
    pf = native_pixel_format(tf) if pixel_format is None else pixel_format
    wib, hib = (width + 3) // 4, (height + 3) // 4
    b = _host_u8(blocks)
    if (tf >> 24) != 0 and b.size != wib * hib * block_size(tf):
        raise ValueError("block array has %d bytes, expected %d" % (b.size, wib * hib * block_size(tf)))
    n_out = width * height * pixel_size(pf) if linear else wib * hib * 16 * pixel_size(pf)
    if out is None:
        out = np.zeros(n_out, dtype=np.uint8)
    tex = DetexTexture(tf, b.ctypes.data, width, height, wib, hib)
    ok = getattr(load(), fn_name)(C.byref(tex), out.ctypes.data, pf)
    return out, bool(ok)


def detexDecompressTextureLinear(fmt, blocks, width, height, pixel_format=None, out=None):
    """detexDecompressTextureLinear (texture.c:105-145) -> (row-major pixels, ok)."""
    return _texture_call("detexDecompressTextureLinear", fmt, blocks, width, height, pixel_format, True, out)


def detexDecompressTextureTiled(fmt, blocks, width, height, pixel_format=None, out=None):
    """detexDecompressTextureTiled (texture.c:77-98) -> (tile-major pixels, ok)."""
    return _texture_call("detexDecompressTextureTiled", fmt, blocks, width, height, pixel_format, False, out)


def detexConvertPixels(source, nu_pixels, source_pixel_format, target_pixel_format):
    """detexConvertPixels (convert.c:1082-1166) -> (pixels, ok)."""
    s = _host_u8(source)
    out = np.zeros(nu_pixels * pixel_size(target_pixel_format), dtype=np.uint8)
    ok = load().detexConvertPixels(s.ctypes.data, nu_pixels, source_pixel_format, out.ctypes.data,
                                   target_pixel_format)
    return out, bool(ok)


# ---- device-pointer API (include/detex_b200.h) ---------------------------------------------
def _check(rc):
    if rc != 0:
        raise DetexError("detex_b200 error %d: %s" % (rc, detexGetErrorMessage()))


def decode_device(fmt, d_blocks, width, height, d_pixels, pixel_format=None, layout=LAYOUT_LINEAR,
                  mode_mask=0xFFFFFFFF, flags=0, d_invalid=0, stream=0, wib=None, hib=None):
    """detexb200DecodeDevice on raw device pointers (ints, e.g. torch.Tensor.data_ptr())."""
    tf = _tf(fmt)
    pf = native_pixel_format(tf) if pixel_format is None else pixel_format
    wib = (width + 3) // 4 if wib is None else wib
    hib = (height + 3) // 4 if hib is None else hib
    _check(load().detexb200DecodeDevice(tf, d_blocks, width, height, wib, hib, d_pixels, pf, layout,
                                        mode_mask, flags, d_invalid, stream))


def decode_hdr_device(fmt, d_blocks, width, height, d_pixels, pixel_format, gamma=1.0, range_min=0.0, range_max=1.0,
                      layout=LAYOUT_LINEAR, d_invalid=0, stream=0, wib=None, hib=None):
    """detexb200DecodeHDRDevice: BC6H blocks -> HDR-converted pixels (decode, then the conversion from the
    _HDR twin of the native format under the given gamma / range) on raw device pointers."""
    tf = _tf(fmt)
    wib = (width + 3) // 4 if wib is None else wib
    hib = (height + 3) // 4 if hib is None else hib
    _check(load().detexb200DecodeHDRDevice(tf, d_blocks, width, height, wib, hib, d_pixels, pixel_format, layout,
                                           gamma, range_min, range_max, d_invalid, stream))


def convert_pixels_device(d_src, nu_pixels, source_pixel_format, d_dst, target_pixel_format, stream=0):
    _check(load().detexb200ConvertPixelsDevice(d_src, nu_pixels, source_pixel_format, d_dst,
                                               target_pixel_format, stream))


def generate_blocks_device(d_blocks, n_blocks, block_bytes, seed=1, first_block=0, force_modes=0, stream=0):
    _check(load().detexb200GenerateBlocksDevice(d_blocks, n_blocks, block_bytes, seed, first_block,
                                                force_modes, stream))


def hash_device(d_buf, n_bytes, stream=0):
    h = C.c_uint64(0)
    _check(load().detexb200HashDevice(d_buf, n_bytes, C.byref(h), stream))
    return h.value


def launch_count():
    return int(load().detexb200LaunchCount())


def set_devices(device_ids):
    """detexb200SetDevices: the GPUs the host-pointer texture calls spread their slabs over
    (empty list = the calling thread's current device)."""
    ids = list(device_ids)
    arr = (C.c_int * max(1, len(ids)))(*ids)
    _check(load().detexb200SetDevices(len(ids), arr))


def set_host_threads(n):
    _check(load().detexb200SetHostThreads(n))


def release_thread_resources():
    load().detexb200ReleaseThreadResources()


def build_id():
    """The source hash the loaded library was built from (detexb200Version(): "... build <id>...")."""
    v = load().detexb200Version().decode()
    return v.split("build ")[1].split(";")[0].split(")")[0].strip() if "build " in v else "unknown"


# ---- KTX / DDS ingest straight to device memory (include/detex_b200.h) ----------------------------
class DeviceImage(C.Structure):
    _fields_ = [("d_pixels", C.c_void_p), ("pixel_bytes", C.c_size_t), ("width", C.c_int), ("height", C.c_int),
                ("width_in_blocks", C.c_int), ("height_in_blocks", C.c_int)]


def load_texture_file_to_device(filename, max_mipmaps=32, pixel_format=0, layout=LAYOUT_LINEAR, stream=0):
    """detexb200LoadTextureFileToDevice -> (levels, texture_format, all_valid, free).  levels is a list
    of DeviceImage (device pointers into one allocation); call free() when done."""
    lib = load()
    lib.detexb200LoadTextureFileToDevice.argtypes = [C.c_char_p, C.c_int, C.c_uint32, C.c_int, C.POINTER(DeviceImage),
                                                    C.POINTER(C.c_int), C.POINTER(C.c_uint32), C.POINTER(C.c_int), C.c_void_p]
    lib.detexb200LoadTextureFileToDevice.restype = C.c_int
    lib.detexb200FreeDeviceImages.argtypes = [C.POINTER(DeviceImage), C.c_int]
    lib.detexb200FreeDeviceImages.restype = None
    levels = (DeviceImage * max_mipmaps)()
    n, tf, valid = C.c_int(0), C.c_uint32(0), C.c_int(0)
    _check(lib.detexb200LoadTextureFileToDevice(filename.encode(), max_mipmaps, pixel_format, layout, levels,
                                                C.byref(n), C.byref(tf), C.byref(valid), stream))
    return [levels[i] for i in range(n.value)], tf.value, bool(valid.value), lambda: lib.detexb200FreeDeviceImages(levels, n.value)
