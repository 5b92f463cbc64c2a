// BC6H pixels -> displayable 8-bit pixels inside the decode kernels.
//
// What a caller of the reference does to show or export a BC6H texture is a chain of three
// whole-buffer passes: decode to FLOAT_RGBX16, half -> 16-bit integer (plain:
// detexConvertNormalizedHalfFloatToUInt16, half-float.c:304-312; with range parameters:
// detexConvertHDRHalfFloatToUInt16Gamma1, hdr.c:119-141), 16 -> 8 bit (convert.c:300-316).
// Every step is a per-component function, so the chain is applied here to the decoder's pixels
// while they are still in registers (a pixel sink of decode_bptc_float_to); HBM sees 16 bytes in
// and 64 (48) bytes out per block instead of 16 + 128 + 128 + 128 + 64.  Stopping after the second
// step gives the RGBX16 target (the HDR step of hdr.c on its own).
//
// The arithmetic is the reference's, with its rounding mode (FE_DOWNWARD) spelled out, restated
// only where the restatement provably cannot change a result (tests/test_emu_tone.py checks the
// composition against the compiled reference for every half value the decoder can produce, on the
// CPU, through tests/emu/host_emu.h):
//   * half -> float is a multiply of the shifted bit pattern by 2^112: exact for every
//     non-negative finite half, denormals included, and the unsigned BC6H decoder produces nothing
//     else ((x * 31) >> 6 <= 0x7BFF); detexClamp0To1 (detex.h:941-948) rides on the multiply that
//     produces its argument as a .sat modifier (clamps to [0, 1], NaN -> 0: the value the shipped
//     -Ofast build gives everything that is not above zero);
//   * u = lrintf(c * 65535.0f + 0.5f) under FE_DOWNWARD is floor(rd(rd(p) + 0.5)) with p = c * 65535.
//     Rounding down to a grid that contains every integer and every integer + 0.5 (any float grid
//     below 2^23) moves a value across neither, so u = floor(p + 0.5) with p exact;
//     fma.rm(c, 131070, 1) = rd(2p + 1), and adding 2^24 (ulp 2) rounds that down to an even integer:
//     the mantissa field of the sum is floor((2p + 1) / 2) = u.  Two instructions instead of
//     multiply, add, float -> int;
//   * (u + 127) * 255 / 65535 == ((u + 127) * 65281) >> 24 (kernels.cu, narrow16), folded with the
//     removal of the exponent bits into one multiply-add.
// Gamma other than 1 (hdr.c:143-164) replaces the reference's powf table by 2^(e * log2 f) on the
// special-function unit: that variant is not bit-exact, it agrees with the reference within one
// integer step of the target (as the separate conversion pass with CUDA's powf does).
#pragma once
#include "common.cuh"
#include "decode_bptc_float.cuh"

namespace detexb200 {

// Three variants, compiled separately (kernels.cu "post class"; the host picks one per call):
//   TONE_PLAIN  normalised half -> integer (half-float.c:304-312, and hdr.c:124-131 for range [0, 1])
//   TONE_RANGE  gamma 1, value rescaled by the range first (hdr.c:132-140)
//   TONE_GAMMA  gamma-corrected value rescaled by the gamma-corrected range (hdr.c:143-164)
constexpr int TONE_PLAIN = 0, TONE_RANGE = 1, TONE_GAMMA = 2;

// Kernel argument.  zero_target = 1: an invalid block leaves zeros in the TARGET format
// (texture-level decode, texture.c:90-93); 0: it leaves the conversion of zero FLOAT_RGBX16 pixels
// (decode, then detexConvertPixels from the _HDR format -- what the reference's two calls produce).
// TONE_RANGE reads range_min / range_max; TONE_GAMMA reads inv_gamma and the corrected range the host
// computed the way the reference does (powf on the host, hdr.c:149-158): gamma_lo = signed
// pow(range_min, 1 / gamma), gamma_factor = 1 / (pow(range_max) - pow(range_min)).
struct ToneArgs {
	float range_min, range_max;
	float inv_gamma, gamma_lo, gamma_factor;
	uint32_t zero_target;
};

struct Tone {
	float lo, factor, e;
};

#ifdef DETEXB200_HOST_EMU
__device__ __forceinline__ float mul_sat(float a, float b) { return emu_sat(__fmul_rn(a, b)); }
__device__ __forceinline__ float mul_rd_sat(float a, float b) { return emu_sat(__fmul_rd(a, b)); }
__device__ __forceinline__ float fast_pow(float f, float e) { return exp2f(e * log2f(f)); }
#else
__device__ __forceinline__ float mul_sat(float a, float b) {
	float r;
	asm("mul.rn.sat.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
	return r;
}
__device__ __forceinline__ float mul_rd_sat(float a, float b) {
	float r;
	asm("mul.rm.sat.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
	return r;
}
// f >= 0: 2^(e * log2 f) on the special-function unit (two MUFU per component); pow(0) = 0 for e > 0
__device__ __forceinline__ float fast_pow(float f, float e) {
	float l, r;
	// .ftz: the argument is a half's value (never a float denormal), and a result below 2^-126 is zero
	// to every target anyway; without it each MUFU is wrapped in a denormal-scaling sequence
	asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(f));
	asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fmul_rn(e, l)));
	return r;
}
#endif

template <int MODE>
__device__ __forceinline__ Tone tone_setup(const ToneArgs &a) {
	Tone t;
	t.lo = 0.0f;
	t.factor = 1.0f;
	t.e = 1.0f;
	if constexpr (MODE == TONE_RANGE) {
		t.lo = a.range_min;
		// computed after fesetround(FE_DOWNWARD), hdr.c:123,133
		t.factor = __fdiv_rd(1.0f, __fsub_rd(a.range_max, a.range_min));
	}
	else if constexpr (MODE == TONE_GAMMA) {
		t.lo = a.gamma_lo;
		t.factor = a.gamma_factor;
		t.e = a.inv_gamma;
	}
	return t;
}

// hs = (half bit pattern) << 13, half in [0, 0x7BFF].  Returns 0x4B800000 + u, u the 16-bit integer
// component.
template <int MODE>
__device__ __forceinline__ uint32_t tone16_bits(uint32_t hs, const Tone &t) {
	float c;
	if constexpr (MODE == TONE_RANGE)
		c = mul_rd_sat(__fsub_rd(__fmul_rn(__uint_as_float(hs), 0x1p112f), t.lo), t.factor);	// hdr.c:134-136
	else if constexpr (MODE == TONE_GAMMA)	// hdr.c:160-163 (the table look-up of :161 is powf(f, 1 / gamma), :40-50)
		c = mul_sat(__fsub_rn(fast_pow(__fmul_rn(__uint_as_float(hs), 0x1p112f), t.e), t.lo), t.factor);
	else
		c = mul_sat(__uint_as_float(hs), 0x1p112f);
	// u = lrintf(c * 65535 + 0.5) rounding down (see the head of the file)
	return __float_as_uint(__fadd_rd(__fmaf_rd(c, 131070.0f, 1.0f), 16777216.0f));
}

// The same down to 8 bits: a word whose TOP BYTE is the component; the lower bytes are garbage.
template <int MODE>
__device__ __forceinline__ uint32_t tone8_top(uint32_t hs, const Tone &t) {
	// (u + 127) * 65281: top byte = (u + 127) / 257
	return tone16_bits<MODE>(hs, t) * 65281u + (127u - 0x4B800000u) * 65281u;
}

// one RGBX8 pixel from three words holding a component in their top byte; X = 0xFF (convert.c:305-309)
__device__ __forceinline__ uint32_t tone8_pack(uint32_t r, uint32_t g, uint32_t b) {
	return prmt(prmt(r, g, 0x7773), b | 0xFFu, 0x4710);
}

// The RGBX8 pixel a zero FLOAT_RGBX16 pixel converts to (what an invalid block leaves behind when
// the caller asked for decode + conversion semantics).
template <int MODE>
__device__ __forceinline__ uint32_t tone8_of_zero(const Tone &t) {
	const uint32_t z = tone8_top<MODE>(0u, t);
	return tone8_pack(z, z, z);
}

// Pixel sink of the BC6H decoder: interpolated values -> RGBX8.  The half pattern of a value is
// (v * 31) >> 6 (decompress-bptc-float.c:616-621); shifted left by 13 that is v * 31 * 128 with the
// low 13 bits cleared.
template <int N, int MODE>
struct Bc6hTone8Sink {
	static constexpr bool FLOAT_PIPE = false, TAKES_SCALED = false;
	__device__ __forceinline__ void emit(int, float, float, float) const {}
	__device__ __forceinline__ void emit_scaled(int, uint32_t, uint32_t, uint32_t) const {}
	uint32_t (&q)[N];		// the first 16 words receive the pixels
	const Tone &tone;
	__device__ __forceinline__ void operator()(int i, int v0, int v1, int v2) const {
		const uint32_t r = tone8_top<MODE>(((uint32_t)v0 * 3968u) & 0xFFFFE000u, tone);
		const uint32_t g = tone8_top<MODE>(((uint32_t)v1 * 3968u) & 0xFFFFE000u, tone);
		const uint32_t b = tone8_top<MODE>(((uint32_t)v2 * 3968u) & 0xFFFFE000u, tone);
		q[i] = tone8_pack(r, g, b);
	}
};

// The 16-bit stage alone (target RGBX16): pixels R | G << 16, B | X << 16 with X the conversion of the
// decoder's zero (FLOAT_RGBX16 -> RGBX16 converts all four components, convert.c:604-608).
template <int MODE>
__device__ __forceinline__ uint32_t tone16_of_zero(const Tone &t) {
	return tone16_bits<MODE>(0u, t) & 0xFFFFu;
}
template <int MODE>
struct Bc6hTone16Sink {
	static constexpr bool FLOAT_PIPE = false, TAKES_SCALED = false;
	__device__ __forceinline__ void emit(int, float, float, float) const {}
	__device__ __forceinline__ void emit_scaled(int, uint32_t, uint32_t, uint32_t) const {}
	uint32_t (&q)[32];
	const Tone &tone;
	uint32_t x16;
	__device__ __forceinline__ void operator()(int i, int v0, int v1, int v2) const {
		const uint32_t r = tone16_bits<MODE>(((uint32_t)v0 * 3968u) & 0xFFFFE000u, tone);
		const uint32_t g = tone16_bits<MODE>(((uint32_t)v1 * 3968u) & 0xFFFFE000u, tone);
		const uint32_t b = tone16_bits<MODE>(((uint32_t)v2 * 3968u) & 0xFFFFE000u, tone);
		q[2 * i] = prmt(r, g, 0x5410);
		q[2 * i + 1] = prmt(b, x16, 0x5410);
	}
};
template <int MODE>
__device__ __forceinline__ bool decode_bptc_float_tone16(const uint4 blk, uint32_t mode_mask, const BptcTables &tab,
const Tone &tone, uint32_t (&q)[32]) {
	return decode_bptc_float_to<false>(blk, mode_mask, tab, Bc6hTone16Sink<MODE>{ q, tone, tone16_of_zero<MODE>(tone) });
}

// BPTC_FLOAT block -> 16 RGBX8 pixels in q[0..15].  Returns false for an invalid block (q is then undefined).
template <int MODE, int N>
__device__ __forceinline__ bool decode_bptc_float_tone8(const uint4 blk, uint32_t mode_mask, const BptcTables &tab,
const Tone &tone, uint32_t (&q)[N]) {
	static_assert(N >= 16, "room for 16 pixels");
	return decode_bptc_float_to<false>(blk, mode_mask, tab, Bc6hTone8Sink<N, MODE>{ q, tone });
}

}  // namespace detexb200
