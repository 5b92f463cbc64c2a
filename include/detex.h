/*
 * detex.h -- drop-in C API header for libdetex_b200.so.
 *
 * This header re-declares, for the block-decode hot path only, the public C API
 * of hglm/detex v0.1.2 (reference detex.h).  Every constant has the same
 * numeric value and every function the same signature and error behaviour as
 * in the reference, so existing C hosts (e.g. the reference's detex-convert.c:310
 * or detex-view.c:182) recompile and relink against libdetex_b200.so unchanged.
 * The decode itself runs as sm_100a CUDA kernels; there is no CPU fallback.
 *
 * Scope: the 19 per-block decoders, detexDecompressBlock,
 * detexDecompressTextureTiled / Linear, detexConvertPixels(InPlace) and
 * detexGetErrorMessage.  File I/O, metadata, HDR and Get/SetMode entry points
 * are NOT provided here; a host that needs them keeps linking the reference's
 * own objects for those.
 *
 * Constant values were taken from the compiled reference header
 * (tests/golden/reference_enum_values.txt); the citations give the reference line.
 */
#ifndef DETEX_B200_COMPAT_DETEX_H
#define DETEX_B200_COMPAT_DETEX_H

#include <stdbool.h>
#include <stdint.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DETEX_API __attribute__((visibility("default")))
#else
#define DETEX_API
#endif

#define DETEX_MAX_BLOCK_SIZE 256			/* reference detex.h:79 */

/* ---- pixel formats: bit-field layout, reference detex.h:83-379 ------------- */
/* bits 0-1 component size (1 << n bytes), bit 2 alpha, bit 3 BGR order,
 * bits 4-5 components-1, bits 8-11 pixel bytes-1, 0x1000 signed, 0x2000 float,
 * 0x4000 HDR. */
#define DETEX_PIXEL_FORMAT_16BIT_COMPONENT_BIT		0x1
#define DETEX_PIXEL_FORMAT_32BIT_COMPONENT_BIT		0x2
#define DETEX_PIXEL_FORMAT_ALPHA_COMPONENT_BIT		0x4
#define DETEX_PIXEL_FORMAT_RGB_COMPONENT_ORDER_BIT	0x0
#define DETEX_PIXEL_FORMAT_BGR_COMPONENT_ORDER_BIT	0x8
#define DETEX_PIXEL_FORMAT_SIGNED_BIT			0x1000
#define DETEX_PIXEL_FORMAT_FLOAT_BIT			0x2000
#define DETEX_PIXEL_FORMAT_HDR_BIT			0x4000

enum {
	DETEX_PIXEL_FORMAT_RGBA8 = 0x334, DETEX_PIXEL_FORMAT_BGRA8 = 0x33C,
	DETEX_PIXEL_FORMAT_RGBX8 = 0x320, DETEX_PIXEL_FORMAT_BGRX8 = 0x328,
	DETEX_PIXEL_FORMAT_RGB8 = 0x220, DETEX_PIXEL_FORMAT_BGR8 = 0x228,
	DETEX_PIXEL_FORMAT_R8 = 0x0, DETEX_PIXEL_FORMAT_SIGNED_R8 = 0x1000,
	DETEX_PIXEL_FORMAT_RG8 = 0x110, DETEX_PIXEL_FORMAT_SIGNED_RG8 = 0x1110,
	DETEX_PIXEL_FORMAT_R16 = 0x101, DETEX_PIXEL_FORMAT_SIGNED_R16 = 0x1101,
	DETEX_PIXEL_FORMAT_RG16 = 0x311, DETEX_PIXEL_FORMAT_SIGNED_RG16 = 0x1311,
	DETEX_PIXEL_FORMAT_RGB16 = 0x521, DETEX_PIXEL_FORMAT_RGBX16 = 0x721,
	DETEX_PIXEL_FORMAT_RGBA16 = 0x735,
	DETEX_PIXEL_FORMAT_FLOAT_R16 = 0x2101, DETEX_PIXEL_FORMAT_FLOAT_R16_HDR = 0x6101,
	DETEX_PIXEL_FORMAT_FLOAT_RG16 = 0x2311, DETEX_PIXEL_FORMAT_FLOAT_RG16_HDR = 0x6311,
	DETEX_PIXEL_FORMAT_FLOAT_RGBX16 = 0x2721, DETEX_PIXEL_FORMAT_FLOAT_RGBX16_HDR = 0x6721,
	/* sic: the reference swaps the HDR bit on these two (detex.h:262-277). */
	DETEX_PIXEL_FORMAT_FLOAT_RGBA16 = 0x6735, DETEX_PIXEL_FORMAT_FLOAT_RGBA16_HDR = 0x2735,
	DETEX_PIXEL_FORMAT_FLOAT_RGB16 = 0x2521, DETEX_PIXEL_FORMAT_FLOAT_RGB16_HDR = 0x6521,
	DETEX_PIXEL_FORMAT_FLOAT_BGRX16 = 0x2729, DETEX_PIXEL_FORMAT_FLOAT_BGRX16_HDR = 0x6729,
	DETEX_PIXEL_FORMAT_SIGNED_FLOAT_RGBX16 = 0x3721, DETEX_PIXEL_FORMAT_SIGNED_FLOAT_BGRX16 = 0x3729,
	DETEX_PIXEL_FORMAT_FLOAT_R32 = 0x2302, DETEX_PIXEL_FORMAT_FLOAT_R32_HDR = 0x6302,
	DETEX_PIXEL_FORMAT_FLOAT_RG32 = 0x2712, DETEX_PIXEL_FORMAT_FLOAT_RG32_HDR = 0x6712,
	DETEX_PIXEL_FORMAT_FLOAT_RGB32 = 0x2B22, DETEX_PIXEL_FORMAT_FLOAT_RGB32_HDR = 0x6B22,
	DETEX_PIXEL_FORMAT_FLOAT_RGBX32 = 0x2F22, DETEX_PIXEL_FORMAT_FLOAT_RGBX32_HDR = 0x6F22,
	DETEX_PIXEL_FORMAT_FLOAT_RGBA32 = 0x2F36, DETEX_PIXEL_FORMAT_FLOAT_RGBA32_HDR = 0x6F36,
	DETEX_PIXEL_FORMAT_A8 = 0x4,
};

/* ---- mode masks and decompression flags, reference detex.h:383-411 --------- */
enum {
	DETEX_MODE_MASK_ETC_INDIVIDUAL = 0x1, DETEX_MODE_MASK_ETC_DIFFERENTIAL = 0x2,
	DETEX_MODE_MASK_ETC_T = 0x4, DETEX_MODE_MASK_ETC_H = 0x8, DETEX_MODE_MASK_ETC_PLANAR = 0x10,
	DETEX_MODE_MASK_ALL_MODES_ETC1 = 0x3, DETEX_MODE_MASK_ALL_MODES_ETC2 = 0x1F,
	DETEX_MODE_MASK_ALL_MODES_ETC2_PUNCHTHROUGH = 0x1E,
	DETEX_MODE_MASK_ALL_MODES_BPTC = 0xFF, DETEX_MODE_MASK_ALL_MODES_BPTC_FLOAT = 0x3FFF,
	DETEX_MODE_MASK_ALL = 0xFFFFFFFF,
};
enum {
	DETEX_DECOMPRESS_FLAG_ENCODE = 0x1,		/* reject blocks an encoder may not emit */
	DETEX_DECOMPRESS_FLAG_OPAQUE_ONLY = 0x2,	/* reject non-opaque modes */
	DETEX_DECOMPRESS_FLAG_NON_OPAQUE_ONLY = 0x4,	/* reject opaque modes */
};

/* ---- texture formats, reference detex.h:575-727 ----------------------------- */
/* (index << 24) | 128-bit-block flag | native pixel format. */
#define DETEX_TEXTURE_FORMAT_COMPRESSED_FORMAT_BITS(n)	((uint32_t)(n) << 24)
#define DETEX_TEXTURE_FORMAT_PIXEL_FORMAT_MASK		0x0000FFFF
#define DETEX_TEXTURE_FORMAT_128BIT_BLOCK_BIT		0x00800000

enum {
	DETEX_COMPRESSED_TEXTURE_FORMAT_INDEX_UNCOMPRESSED = 0,
	DETEX_COMPRESSED_TEXTURE_FORMAT_INDEX_BC1 = 1, DETEX_COMPRESSED_TEXTURE_FORMAT_INDEX_BC1A = 2,
	DETEX_COMPRESSED_TEXTURE_FORMAT_INDEX_BC2 = 3, DETEX_COMPRESSED_TEXTURE_FORMAT_INDEX_BC3 = 4,
	DETEX_COMPRESSED_TEXTURE_FORMAT_INDEX_RGTC1 = 5, DETEX_COMPRESSED_TEXTURE_FORMAT_INDEX_SIGNED_RGTC1 = 6,
	DETEX_COMPRESSED_TEXTURE_FORMAT_INDEX_RGTC2 = 7, DETEX_COMPRESSED_TEXTURE_FORMAT_INDEX_SIGNED_RGTC2 = 8,
	DETEX_COMPRESSED_TEXTURE_FORMAT_INDEX_BPTC_FLOAT = 9,
	DETEX_COMPRESSED_TEXTURE_FORMAT_INDEX_BPTC_SIGNED_FLOAT = 10,
	DETEX_COMPRESSED_TEXTURE_FORMAT_INDEX_BPTC = 11, DETEX_COMPRESSED_TEXTURE_FORMAT_INDEX_ETC1 = 12,
	DETEX_COMPRESSED_TEXTURE_FORMAT_INDEX_ETC2 = 13,
	DETEX_COMPRESSED_TEXTURE_FORMAT_INDEX_ETC2_PUNCHTHROUGH = 14,
	DETEX_COMPRESSED_TEXTURE_FORMAT_INDEX_ETC2_EAC = 15, DETEX_COMPRESSED_TEXTURE_FORMAT_INDEX_EAC_R11 = 16,
	DETEX_COMPRESSED_TEXTURE_FORMAT_INDEX_EAC_SIGNED_R11 = 17,
	DETEX_COMPRESSED_TEXTURE_FORMAT_INDEX_EAC_RG11 = 18,
	DETEX_COMPRESSED_TEXTURE_FORMAT_INDEX_EAC_SIGNED_RG11 = 19,
	DETEX_COMPRESSED_TEXTURE_FORMAT_INDEX_ASTC_4X4 = 20,	/* declared, no decoder (detex.h:610) */
};
enum {
	DETEX_TEXTURE_FORMAT_BC1 = 0x01000320, DETEX_TEXTURE_FORMAT_BC1A = 0x02000334,
	DETEX_TEXTURE_FORMAT_BC2 = 0x03800334, DETEX_TEXTURE_FORMAT_BC3 = 0x04800334,
	DETEX_TEXTURE_FORMAT_RGTC1 = 0x05000000, DETEX_TEXTURE_FORMAT_SIGNED_RGTC1 = 0x06001101,
	DETEX_TEXTURE_FORMAT_RGTC2 = 0x07800110, DETEX_TEXTURE_FORMAT_SIGNED_RGTC2 = 0x08801311,
	DETEX_TEXTURE_FORMAT_BPTC_FLOAT = 0x09802721, DETEX_TEXTURE_FORMAT_BPTC_SIGNED_FLOAT = 0x0A803721,
	DETEX_TEXTURE_FORMAT_BPTC = 0x0B800334, DETEX_TEXTURE_FORMAT_ETC1 = 0x0C000320,
	DETEX_TEXTURE_FORMAT_ETC2 = 0x0D000320, DETEX_TEXTURE_FORMAT_ETC2_PUNCHTHROUGH = 0x0E000334,
	DETEX_TEXTURE_FORMAT_ETC2_EAC = 0x0F800334, DETEX_TEXTURE_FORMAT_EAC_R11 = 0x10000101,
	DETEX_TEXTURE_FORMAT_EAC_SIGNED_R11 = 0x11001101, DETEX_TEXTURE_FORMAT_EAC_RG11 = 0x12800311,
	DETEX_TEXTURE_FORMAT_EAC_SIGNED_RG11 = 0x13801311, DETEX_TEXTURE_FORMAT_ASTC_4X4 = 0x14800334,
	/* aliases used by the reference's own callers */
	DETEX_TEXTURE_FORMAT_BC4_UNORM = DETEX_TEXTURE_FORMAT_RGTC1,
	DETEX_TEXTURE_FORMAT_BC5_UNORM = DETEX_TEXTURE_FORMAT_RGTC2,
	DETEX_TEXTURE_FORMAT_BC6H_UF16 = DETEX_TEXTURE_FORMAT_BPTC_FLOAT,
	DETEX_TEXTURE_FORMAT_BC6H_SF16 = DETEX_TEXTURE_FORMAT_BPTC_SIGNED_FLOAT,
	DETEX_TEXTURE_FORMAT_BC7 = DETEX_TEXTURE_FORMAT_BPTC,
};

/* Caller-owned texture descriptor; layout of reference detex.h:729-736. */
typedef struct {
	uint32_t format;
	uint8_t *data;
	int width;
	int height;
	int width_in_blocks;
	int height_in_blocks;
} detexTexture;

/* ---- per-block decoders, reference detex.h:435-531 --------------------------
 * bitstring: one 8- or 16-byte block (host memory); pixel_buffer: 16 pixels,
 * row-major, in the format's native pixel format.  Returns false for an
 * invalid / masked-out block (output then undefined, as in the reference). */
#define DETEX_DECLARE_BLOCK_DECODER(NAME) \
	DETEX_API bool detexDecompressBlock##NAME(const uint8_t *bitstring, uint32_t mode_mask, \
		uint32_t flags, uint8_t *pixel_buffer);
DETEX_DECLARE_BLOCK_DECODER(BC1)			/* decompress-bc.c:23   -> RGBX8  */
DETEX_DECLARE_BLOCK_DECODER(BC1A)			/* decompress-bc.c:87   -> RGBA8  */
DETEX_DECLARE_BLOCK_DECODER(BC2)			/* decompress-bc.c:136  -> RGBA8  */
DETEX_DECLARE_BLOCK_DECODER(BC3)			/* decompress-bc.c:175  -> RGBA8  */
DETEX_DECLARE_BLOCK_DECODER(RGTC1)			/* decompress-rgtc.c:64 -> R8     */
DETEX_DECLARE_BLOCK_DECODER(SIGNED_RGTC1)		/* decompress-rgtc.c:134 -> SIGNED_R16 */
DETEX_DECLARE_BLOCK_DECODER(RGTC2)			/* decompress-rgtc.c:72 -> RG8    */
DETEX_DECLARE_BLOCK_DECODER(SIGNED_RGTC2)		/* decompress-rgtc.c:141 -> SIGNED_RG16 */
DETEX_DECLARE_BLOCK_DECODER(BPTC_FLOAT)			/* decompress-bptc-float.c:631 -> FLOAT_RGBX16 */
DETEX_DECLARE_BLOCK_DECODER(BPTC_SIGNED_FLOAT)		/* decompress-bptc-float.c:640 -> SIGNED_FLOAT_RGBX16 */
DETEX_DECLARE_BLOCK_DECODER(BPTC)			/* decompress-bptc.c:354 -> RGBA8 */
DETEX_DECLARE_BLOCK_DECODER(ETC1)			/* decompress-etc.c:89  -> RGBX8  */
DETEX_DECLARE_BLOCK_DECODER(ETC2)			/* decompress-etc.c:321 -> RGBX8  */
DETEX_DECLARE_BLOCK_DECODER(ETC2_PUNCHTHROUGH)		/* decompress-etc.c:653 -> RGBA8  */
DETEX_DECLARE_BLOCK_DECODER(ETC2_EAC)			/* decompress-eac.c:54  -> RGBA8  */
DETEX_DECLARE_BLOCK_DECODER(EAC_R11)			/* decompress-eac.c:132 -> R16    */
DETEX_DECLARE_BLOCK_DECODER(EAC_SIGNED_R11)		/* decompress-eac.c:206 -> SIGNED_R16 */
DETEX_DECLARE_BLOCK_DECODER(EAC_RG11)			/* decompress-eac.c:143 -> RG16   */
DETEX_DECLARE_BLOCK_DECODER(EAC_SIGNED_RG11)		/* decompress-eac.c:217 -> SIGNED_RG16 */
#undef DETEX_DECLARE_BLOCK_DECODER

/* ---- texture-level API, reference detex.h:747-765 / texture.c:55-145 -------- */
DETEX_API bool detexDecompressBlock(const uint8_t *bitstring, uint32_t texture_format,
	uint32_t mode_mask, uint32_t flags, uint8_t *pixel_buffer, uint32_t pixel_format);
/* Flat block array -> flat array of 16-pixel tiles. */
DETEX_API bool detexDecompressTextureTiled(const detexTexture *texture, uint8_t *pixel_buffer,
	uint32_t pixel_format);
/* Block array -> row-major image (stride width * pixel size, edge blocks clipped). */
DETEX_API bool detexDecompressTextureLinear(const detexTexture *texture, uint8_t *pixel_buffer,
	uint32_t pixel_format);

/* ---- pixel conversion, reference detex.h:777-784 / convert.c:1082-1171 ------ */
DETEX_API bool detexConvertPixels(uint8_t *source_pixel_buffer, uint32_t nu_pixels,
	uint32_t source_pixel_format, uint8_t *target_pixel_buffer, uint32_t target_pixel_format);
DETEX_API bool detexConvertPixelsInPlace(uint8_t *source_pixel_buffer, uint32_t nu_pixels,
	uint32_t source_pixel_format, uint32_t target_pixel_format);

/* ---- KTX / DDS loaders, reference detex.h:826-867 (ktx.c:36-176, dds.c:32-143, misc.c:98-121) --
 * textures_out receives a malloc'd array of malloc'd textures (each with malloc'd data);
 * release everything with free(), as with the reference. */
DETEX_API bool detexLoadKTXFileWithMipmaps(const char *filename, int max_mipmaps,
	detexTexture ***textures_out, int *nu_levels_out);
DETEX_API bool detexLoadKTXFile(const char *filename, detexTexture **texture_out);
DETEX_API bool detexLoadDDSFileWithMipmaps(const char *filename, int max_mipmaps,
	detexTexture ***textures_out, int *nu_levels_out);
DETEX_API bool detexLoadDDSFile(const char *filename, detexTexture **texture_out);
DETEX_API bool detexLoadTextureFileWithMipmaps(const char *filename, int max_mipmaps,
	detexTexture ***textures_out, int *nu_levels_out);
DETEX_API bool detexLoadTextureFile(const char *filename, detexTexture **texture_out);

/* Range / gamma used by the *_HDR source formats, reference detex.h:792-800 / hdr.c:26-38
 * (thread-local; defaults gamma 1, range 0..1). */
DETEX_API void detexSetHDRParameters(float gamma, float range_min, float range_max);

/* Thread-local message of the last failure, reference detex.h:806 / misc.c:75-93. */
DETEX_API const char *detexGetErrorMessage(void);

/* ---- inline format helpers, reference detex.h:879-930 ----------------------- */
static inline int detexGetPixelSize(uint32_t pixel_format) { return 1 + ((pixel_format & 0xF00) >> 8); }
static inline int detexGetNumberOfComponents(uint32_t pixel_format) { return 1 + ((pixel_format & 0x30) >> 4); }
static inline int detexGetComponentSize(uint32_t pixel_format) { return 1 << (pixel_format & 0x3); }
static inline uint32_t detexTextureSize(uint32_t width_in_blocks, uint32_t height_in_blocks,
uint32_t pixel_format) {
	return width_in_blocks * height_in_blocks * detexGetPixelSize(pixel_format) * 16;
}
static inline uint32_t detexFormatHasAlpha(uint32_t pixel_format) { return (pixel_format & 0x4) != 0; }
static inline uint32_t detexGetCompressedFormat(uint32_t texture_format) { return texture_format >> 24; }
static inline uint32_t detexGetCompressedBlockSize(uint32_t texture_format) {
	return 8 + ((texture_format & DETEX_TEXTURE_FORMAT_128BIT_BLOCK_BIT) >> 20);
}
static inline uint32_t detexFormatIsCompressed(uint32_t texture_format) { return (texture_format >> 24) != 0; }
static inline uint32_t detexGetPixelFormat(uint32_t texture_format) { return texture_format & 0xFFFF; }

#ifdef __cplusplus
}
#endif
#endif
