/*
 * TEST INFRASTRUCTURE ONLY.  CPU restatement ("oracle") of the hglm/detex block
 * decoders and texture drivers -- see detex_oracle.c.  It exports the same C
 * symbol names as the reference library so one ctypes binding drives either.
 * Never linked into, or called by, the product library (libdetex_b200.so).
 */
#ifndef DETEX_ORACLE_H
#define DETEX_ORACLE_H
#include <stdbool.h>
#include <stdint.h>

typedef struct {			/* reference detex.h:729-736 */
	uint32_t format;
	uint8_t *data;
	int width, height, width_in_blocks, height_in_blocks;
} oracleTexture;

#define ORACLE_BLOCK_FN(NAME) \
	bool detexDecompressBlock##NAME(const uint8_t *bitstring, uint32_t mode_mask, \
		uint32_t flags, uint8_t *pixel_buffer)

ORACLE_BLOCK_FN(BC1);  ORACLE_BLOCK_FN(BC1A);  ORACLE_BLOCK_FN(BC2);  ORACLE_BLOCK_FN(BC3);
ORACLE_BLOCK_FN(RGTC1);  ORACLE_BLOCK_FN(SIGNED_RGTC1);
ORACLE_BLOCK_FN(RGTC2);  ORACLE_BLOCK_FN(SIGNED_RGTC2);
ORACLE_BLOCK_FN(BPTC_FLOAT);  ORACLE_BLOCK_FN(BPTC_SIGNED_FLOAT);  ORACLE_BLOCK_FN(BPTC);
ORACLE_BLOCK_FN(ETC1);  ORACLE_BLOCK_FN(ETC2);  ORACLE_BLOCK_FN(ETC2_PUNCHTHROUGH);
ORACLE_BLOCK_FN(ETC2_EAC);
ORACLE_BLOCK_FN(EAC_R11);  ORACLE_BLOCK_FN(EAC_SIGNED_R11);
ORACLE_BLOCK_FN(EAC_RG11);  ORACLE_BLOCK_FN(EAC_SIGNED_RG11);

bool detexDecompressBlock(const uint8_t *bitstring, uint32_t texture_format, uint32_t mode_mask,
	uint32_t flags, uint8_t *pixel_buffer, uint32_t pixel_format);
bool detexDecompressTextureTiled(const oracleTexture *texture, uint8_t *pixel_buffer,
	uint32_t pixel_format);
bool detexDecompressTextureLinear(const oracleTexture *texture, uint8_t *pixel_buffer,
	uint32_t pixel_format);
bool detexConvertPixels(uint8_t *source_pixel_buffer, uint32_t nu_pixels,
	uint32_t source_pixel_format, uint8_t *target_pixel_buffer, uint32_t target_pixel_format);
const char *detexGetErrorMessage(void);

#endif
