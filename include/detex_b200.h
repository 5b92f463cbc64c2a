/*
 * detex_b200.h -- GPU-side C ABI of libdetex_b200.so (additions outside the
 * reference API; SURVEY.md section 8b "GPU-specific additions").
 *
 * Plain C linkage, plain pointers and sizes, no CUDA or torch types in the
 * signatures (streams travel as void* = cudaStream_t, 0 = default stream).
 * The classic host-pointer detex API (include/detex.h) is implemented on top
 * of these entry points with pinned staging and asynchronous copies.
 *
 * Every function returns 0 on success or a negative DETEXB200_E* code; the text
 * of the last failure on the calling thread is available from
 * detexGetErrorMessage() (same thread-local convention as reference misc.c:75-93).
 */
#ifndef DETEX_B200_H
#define DETEX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DETEXB200_API __attribute__((visibility("default")))
#else
#define DETEXB200_API
#endif

enum {
	DETEXB200_OK = 0,
	DETEXB200_EINVAL = -1,		/* bad argument (format, layout, null pointer, size) */
	DETEXB200_ECUDA = -2,		/* CUDA runtime error (no device, launch failure, ...) */
	DETEXB200_EUNSUPPORTED = -3,	/* pixel-format conversion not implemented on device */
};

enum {
	DETEXB200_LAYOUT_TILED = 0,	/* block i -> 16 pixels at out + i*16*ps   (texture.c:77-98)   */
	DETEXB200_LAYOUT_LINEAR = 1,	/* row-major image, edge blocks clipped    (texture.c:105-145) */
};

/*
 * Decode a block array that is already resident in device memory.
 *
 *   texture_format  DETEX_TEXTURE_FORMAT_* (index 1..19)
 *   d_blocks        device pointer, width_in_blocks*height_in_blocks blocks of 8/16
 *                   bytes, row-major over blocks; must be 16-byte aligned
 *   width, height   image size in pixels (only used by the LINEAR layout for edge
 *                   clipping; TILED ignores them)
 *   d_pixels        device pointer to the output: LINEAR width*height*ps bytes,
 *                   TILED wib*hib*16*ps bytes, ps = pixel size of pixel_format
 *   pixel_format    DETEX_PIXEL_FORMAT_*: the texture format's native pixel
 *                   format, or one of the fused conversions
 *   mode_mask,flags as for the per-block decoders (detex.h:383-411); the texture
 *                   level reference API always passes DETEX_MODE_MASK_ALL, 0
 *   d_invalid       device pointer to one uint32 the kernels set to 1 when any
 *                   block is invalid (such blocks decode to zero pixels, Q16 /
 *                   texture.c:90-93); never cleared by this call; may be NULL
 *   stream          cudaStream_t as void*; the call only enqueues work
 *
 * Replaces the loops of reference texture.c:77-98 and texture.c:105-145.
 */
DETEXB200_API int detexb200DecodeDevice(uint32_t texture_format, const void *d_blocks,
	int width, int height, int width_in_blocks, int height_in_blocks,
	void *d_pixels, uint32_t pixel_format, int layout,
	uint32_t mode_mask, uint32_t flags, uint32_t *d_invalid, void *stream);

/*
 * BC6H straight to displayable pixels: the bytes the reference produces with
 *     detexDecompressTextureLinear / ...Tiled(texture, tmp, DETEX_PIXEL_FORMAT_FLOAT_RGBX16);
 *     detexSetHDRParameters(gamma, range_min, range_max);                        (hdr.c:34-38)
 *     detexConvertPixels(tmp, n, DETEX_PIXEL_FORMAT_FLOAT_RGBX16_HDR, out, pixel_format);
 * i.e. the decoded half floats go through the HDR half -> 16-bit integer step (hdr.c:119-171)
 * and whatever the conversion chain to pixel_format holds after it.  An invalid block leaves the
 * conversion of zero pixels (that is what the sequence above does), and *d_invalid is raised.
 * texture_format must be DETEX_TEXTURE_FORMAT_BPTC_FLOAT (the only compressed format whose pixel
 * format has an _HDR twin).  For pixel_format RGBX16, RGBX8, RGBA8, BGRX8, BGRA8 and RGB8 the
 * whole chain runs inside the decode kernel (one pass over HBM: 16 bytes in, 128, 64 or 48 out
 * per block), bit-exact for gamma == 1; for other gamma values the kernel evaluates the power on
 * the special-function unit and agrees with the reference's powf table within one integer step
 * of the target.  Other targets go through the L2-resident conversion pass of
 * detexb200DecodeDevice.  The HDR parameters of the calling thread (detexSetHDRParameters) are
 * left unchanged.
 */
DETEXB200_API int detexb200DecodeHDRDevice(uint32_t texture_format, const void *d_blocks,
	int width, int height, int width_in_blocks, int height_in_blocks,
	void *d_pixels, uint32_t pixel_format, int layout,
	float gamma, float range_min, float range_max, uint32_t *d_invalid, void *stream);

/*
 * Convert nu_pixels pixels between two pixel formats on the device
 * (reference convert.c:1082-1166 for the supported pairs).
 * d_src and d_dst may be equal when both formats have the same pixel size.
 */
DETEXB200_API int detexb200ConvertPixelsDevice(const void *d_src, uint64_t nu_pixels,
	uint32_t source_pixel_format, void *d_dst, uint32_t target_pixel_format, void *stream);

/*
 * The chain of formats detexConvertPixels goes through for a (source, target) pair: the
 * reference's search over its conversion table (convert.c:887-1045) restated.  Returns the
 * number of steps (0 for identical formats, at most 4) and writes the steps + 1 formats of
 * the chain, source first, to formats_out[5]; returns -1 when there is no path.  Host only.
 */
DETEXB200_API int detexb200ConversionPlan(uint32_t source_pixel_format, uint32_t target_pixel_format,
	uint32_t formats_out[5]);

/*
 * Fill a device buffer with the counter-based splitmix64 block generator of
 * SURVEY.md 8c/8d (word j = mix(seed + (first_word + j + 1) * 0x9E3779B97F4A7C15)),
 * so that every shard / GPU can synthesise its own index range in HBM.
 * force_modes: 0 = raw bits (distribution U); 8 = BC7 block i gets mode i % 8;
 * 14 = BC6H block i gets mode i % 14 (distribution M).  block_bytes is 8 or 16.
 */
DETEXB200_API int detexb200GenerateBlocksDevice(void *d_blocks, uint64_t n_blocks, int block_bytes,
	uint64_t seed, uint64_t first_block, int force_modes, void *stream);

/* FNV-1a-64 of a device buffer computed on the device (per-shard parity hashes,
 * SURVEY.md 8d): the byte-serial hash is restated as a parallel reduction over
 * fixed 4 KiB segments; the host-side reference of the same
 * construction lives in oracle/harness.py. */
DETEXB200_API int detexb200HashDevice(const void *d_buf, uint64_t n_bytes, uint64_t *h_hash_out,
	void *stream);

/*
 * KTX / DDS file -> decoded mip chain in device memory.  The header is parsed
 * on the host like detexLoadTextureFileWithMipmaps does (ktx.c:36-176, dds.c:32-143), the
 * payload of every level goes through one pinned buffer and one asynchronous copy, and every
 * level is decoded (compressed formats) or converted (uncompressed formats) on `stream` into ONE
 * device allocation.  pixel_format 0 = the file's native pixel format.  levels_out must have
 * room for max_mipmaps entries; levels_out[0].d_pixels is the base of the allocation, release
 * it with detexb200FreeDeviceImages.  Synchronous: the pixels are complete on return.
 * *all_valid_out is 0 when some block was invalid (its pixels are zero, as in texture.c:90-93).
 */
typedef struct {
	void *d_pixels;
	size_t pixel_bytes;
	int width, height, width_in_blocks, height_in_blocks;
} detexb200DeviceImage;
DETEXB200_API int detexb200LoadTextureFileToDevice(const char *filename, int max_mipmaps,
	uint32_t pixel_format, int layout, detexb200DeviceImage *levels_out, int *nu_levels_out,
	uint32_t *texture_format_out, int *all_valid_out, void *stream);
DETEXB200_API void detexb200FreeDeviceImages(detexb200DeviceImage *levels, int nu_levels);

/*
 * Which GPUs the host-pointer texture calls of include/detex.h (detexDecompressTextureLinear / ...Tiled,
 * reference texture.c:77-145) use.  n == 0 (the default): the calling thread's current CUDA device.  n >= 1: every
 * call cuts the texture into slabs of block rows (LINEAR) or block runs (TILED) and deals them round-robin to the
 * listed devices, each moving its slabs over its own PCIe link (index-range sharding, no exchange
 * between devices; results land in the caller's one pixel buffer).  Process-wide; also settable without touching
 * the host program through the environment: DETEXB200_DEVICES=all | <count> | <ordinal>,<ordinal>,...
 */
DETEXB200_API int detexb200SetDevices(int n, const int *device_ids);
/*
 * Threads of the copy pool that serves pageable (malloc'd) caller buffers and multi-device calls: each thread
 * stages one slab at a time through its own pinned bounce buffers.  0 = default (one per hardware thread,
 * between 4 and 32, at least 3 per device).  Environment: DETEXB200_HOST_THREADS.
 */
DETEXB200_API int detexb200SetHostThreads(int n);
/*
 * Give the calling thread's streams, device buffers and pinned buffers back to the driver now (they are also
 * released when the thread exits).  The next call on this thread sets them up again, on whatever CUDA device is
 * current then.
 */
DETEXB200_API void detexb200ReleaseThreadResources(void);

/* Library / device information. */
DETEXB200_API int detexb200DeviceCount(void);
DETEXB200_API const char *detexb200Version(void);
/* Number of kernel launches this library has issued on this process (all
 * threads) since load; bench.py reports the delta over the timed region. */
DETEXB200_API uint64_t detexb200LaunchCount(void);

#ifdef __cplusplus
}
#endif
#endif
