"""detex_b200 -- B200-native 4x4 texture-block decompressor behind the detex C API.

The product is the C-ABI shared library ``libdetex_b200.so`` (include/detex.h,
include/detex_b200.h).  This package is only the thin Python host-side mirror used
by the tests and the benchmark: ctypes bindings with the reference's function names
and argument meaning.  There is no CPU decode path; importing works without a GPU
(so the CPU test-suite can check that the library loads and exports its symbols) but
every decode call needs a CUDA device and raises / returns False loudly otherwise.
"""
from .lib import (  # noqa: F401
    DetexTexture, FORMATS, FORMAT_NAMES, PIXEL_FORMATS, LAYOUT_LINEAR, LAYOUT_TILED,
    DetexError, load, pixel_size, block_size, native_pixel_format,
    detexDecompressBlock, detexDecompressBlockFormat, detexDecompressTextureLinear,
    detexDecompressTextureTiled, detexConvertPixels, detexGetErrorMessage,
    decode_device, decode_hdr_device, generate_blocks_device, hash_device, convert_pixels_device, launch_count,
    load_texture_file_to_device, DeviceImage, set_devices, set_host_threads, release_thread_resources, build_id,
)
