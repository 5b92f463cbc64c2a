// BC6H pixels -> displayable 8-bit pixels inside the decode kernels.
//
// What a caller of the reference does to show or export a BC6H texture is a chain of three
// whole-buffer passes: decode to FLOAT_RGBX16, half -> 16-bit integer (plain:
// detexConvertNormalizedHalfFloatToUInt16, half-float.c:304-312; with range parameters:
// detexConvertHDRHalfFloatToUInt16Gamma1, hdr.c:119-141), 16 -> 8 bit (convert.c:300-316).
// Every step is a per-component function, so the chain is applied here to the decoder's 16 pixels
// while they are still in registers; HBM sees 16 bytes in and 64 (48) bytes out per block instead
// of 16 + 128 + 128 + 128 + 64.
//
// The arithmetic is the reference's, operation by operation, with its rounding mode
// (FE_DOWNWARD) spelled out; the only liberties are ones that cannot change a result:
//   * half -> float as a multiply of the shifted bit pattern by 2^112 (exact for every
//     non-negative finite half, denormals included; the unsigned BC6H decoder produces nothing
//     else: (x * 31) >> 6 <= 0x7BFF);
//   * lrintf(t) under FE_DOWNWARD for 0 <= t < 2^23 read from the mantissa of t + 2^23 rounded
//     down;
//   * (v + 127) * 255 / 65535 == ((v + 127) * 65281) >> 24 (kernels.cu, narrow16), folded with the
//     removal of the 2^23 offset into one multiply-add.
// Gamma other than 1 (powf per component, hdr.c:143-164) is not fused; api.cu sends it through
// the separate conversion pass.
#pragma once
#include "common.cuh"

namespace detexb200 {

// Kernel argument.  hdr = 0: plain normalised half -> integer; 1: rescale by the range first.
// zero_target = 1: an invalid block leaves zeros in the TARGET format (texture-level decode,
// texture.c:90-93); 0: it leaves the conversion of zero FLOAT_RGBX16 pixels (decode, then
// detexConvertPixels from the _HDR format -- what the reference's two calls produce).
struct ToneArgs {
	float range_min, range_max;
	uint32_t hdr, zero_target;
};

struct Tone {
	float lo, factor;
	bool rescale;
};

__device__ __forceinline__ Tone tone_setup(const ToneArgs &a) {
	Tone t;
	t.rescale = a.hdr != 0u && !(a.range_min == 0.0f && a.range_max == 1.0f);	// hdr.c:124-131
	t.lo = a.range_min;
	// computed after fesetround(FE_DOWNWARD), hdr.c:123,133
	t.factor = t.rescale ? __fdiv_rd(1.0f, __fsub_rd(a.range_max, a.range_min)) : 1.0f;
	return t;
}

// hs = (half bit pattern) << 13, half in [0, 0x7BFF].  Returns a word whose TOP BYTE is the 8-bit
// component; the lower bytes are garbage.
__device__ __forceinline__ uint32_t tone8_top(uint32_t hs, const Tone &t) {
	float f = __fmul_rn(__uint_as_float(hs), 0x1p112f);		// exact
	if (t.rescale)
		f = __fmul_rd(__fsub_rd(f, t.lo), t.factor);		// hdr.c:134-136
	// detexClamp0To1 (detex.h:941-948); NaN (0 * inf when range_min == range_max) and everything
	// not above zero -> 0 as in the shipped build.  The sign of a zero cannot reach the result.
	f = fminf(fmaxf(f, 0.0f), 1.0f);
	const float u = __fadd_rd(__fmul_rd(f, 65535.0f), 0.5f);	// in [0.5, 65535.5]
	const uint32_t bits = __float_as_uint(__fadd_rd(u, 8388608.0f));	// 0x4B000000 + floor(u)
	return bits * 65281u + (127u - 0x4B000000u) * 65281u;		// ((floor(u) + 127) * 65281), top byte = / 2^24
}

// 16 FLOAT_RGBX16 pixels (R | G << 16, B | X << 16 per pixel) -> 16 RGBX8 pixels with X = 0xFF
// (convert.c:305-309 packs alpha 0xFF).
__device__ __forceinline__ void tone8_tile(const uint32_t (&px)[32], uint32_t (&q)[16], const Tone &t) {
#pragma unroll
	for (int i = 0; i < 16; i++) {
		const uint32_t lo = px[2 * i], hi = px[2 * i + 1];
		const uint32_t r = tone8_top((lo << 13) & 0x0FFFE000u, t);
		const uint32_t g = tone8_top((lo >> 3) & 0x0FFFE000u, t);
		const uint32_t b = tone8_top((hi << 13) & 0x0FFFE000u, t);
		q[i] = prmt(prmt(r, g, 0x7773), b | 0xFFu, 0x4710);	// R, G, B, 0xFF
	}
}

}  // namespace detexb200
