// Pixel-format conversion: the reference's table and chain search restated as data, and the per-pixel
// step functions of the conversion kernel (convert.cu) -- kept in a header so that tests/emu can compile
// the very same arithmetic for the CPU and compare it with the reference (tests/test_emu_convert.py).
#pragma once
#include "common.cuh"

namespace detexb200 {

// Pixel format constants (reference detex.h:83-379; include/detex.h).
enum : uint32_t {
	PF_RGBA8 = 0x334, PF_BGRA8 = 0x33C, PF_RGBX8 = 0x320, PF_BGRX8 = 0x328, PF_RGB8 = 0x220,
	PF_R8 = 0x0, PF_SIGNED_R8 = 0x1000, PF_RG8 = 0x110, PF_SIGNED_RG8 = 0x1110,
	PF_R16 = 0x101, PF_SIGNED_R16 = 0x1101, PF_RG16 = 0x311, PF_SIGNED_RG16 = 0x1311,
	PF_RGB16 = 0x521, PF_RGBX16 = 0x721, PF_RGBA16 = 0x735,
	PF_FLOAT_R16 = 0x2101, PF_FLOAT_R16_HDR = 0x6101, PF_FLOAT_RG16 = 0x2311, PF_FLOAT_RG16_HDR = 0x6311,
	PF_FLOAT_RGBX16 = 0x2721, PF_FLOAT_RGBX16_HDR = 0x6721, PF_FLOAT_RGBA16 = 0x6735,
	PF_FLOAT_RGB16 = 0x2521, PF_FLOAT_RGB16_HDR = 0x6521, PF_FLOAT_BGRX16 = 0x2729,
	PF_FLOAT_R32 = 0x2302, PF_FLOAT_R32_HDR = 0x6302, PF_FLOAT_RG32 = 0x2712, PF_FLOAT_RG32_HDR = 0x6712,
	PF_FLOAT_RGB32 = 0x2B22, PF_FLOAT_RGB32_HDR = 0x6B22, PF_FLOAT_RGBX32 = 0x2F22, PF_FLOAT_RGBX32_HDR = 0x6F22,
};

// What one step does to the component registers c[0..3] of a pixel.
enum StepOp : uint8_t {
	ST_KEEP,		// same bytes (X <-> A relabelling, dropping or keeping leading components)
	ST_SWAP_RB,		// c0 <-> c2
	ST_RGB_TO_BGRX,		// (r, g, b) -> (b, g, r, 0xFF)
	ST_FLIP_SIGN,		// unsigned <-> signed integer: toggle the top bit of every component
	ST_PAD,			// more components: missing colour components 0, fourth component "one"
	ST_NARROW,		// 16-bit -> 8-bit integer: (v + 127) * 255 / 65535
	ST_WIDEN,		// 8-bit -> 16-bit integer: v * 65535 / 255
	ST_INT_TO_HALF,		// 16-bit integer -> float v / 65535 -> half
	ST_HALF_TO_INT,		// normalised half -> 16-bit integer
	ST_HALF_HDR_TO_INT,	// half with range / gamma parameters -> 16-bit integer
	ST_FLOAT_TO_HALF, ST_FLOAT_TO_INT, ST_HALF_TO_FLOAT,
	ST_FLOAT_HDR,		// float with range / gamma parameters -> normalised float
};

struct TableEntry { uint32_t source, target; StepOp op; };

// Same order as detex_conversion_table (convert.c:765-864): plan() reproduces the reference's
// choice among several possible chains only if it scans the entries in the same order.
static const TableEntry conversion_table[] = {
	{ PF_RGBX8, PF_RGBA8, ST_KEEP }, { PF_RGBA8, PF_RGBX8, ST_KEEP },			// :767-770
	{ PF_BGRX8, PF_BGRA8, ST_KEEP }, { PF_BGRA8, PF_BGRX8, ST_KEEP },
	{ PF_RGBX8, PF_BGRX8, ST_SWAP_RB }, { PF_BGRX8, PF_RGBX8, ST_SWAP_RB },			// :772-777
	{ PF_RGBA8, PF_BGRA8, ST_SWAP_RB }, { PF_BGRA8, PF_RGBA8, ST_SWAP_RB },
	{ PF_FLOAT_RGBX16, PF_FLOAT_BGRX16, ST_SWAP_RB }, { PF_FLOAT_BGRX16, PF_FLOAT_RGBX16, ST_SWAP_RB },
	{ PF_RGB8, PF_BGRX8, ST_RGB_TO_BGRX },							// :779
	{ PF_R8, PF_SIGNED_R8, ST_FLIP_SIGN }, { PF_RG8, PF_SIGNED_RG8, ST_FLIP_SIGN },		// :782-789
	{ PF_SIGNED_R8, PF_R8, ST_FLIP_SIGN }, { PF_SIGNED_RG8, PF_RG8, ST_FLIP_SIGN },
	{ PF_R16, PF_SIGNED_R16, ST_FLIP_SIGN }, { PF_RG16, PF_SIGNED_RG16, ST_FLIP_SIGN },
	{ PF_SIGNED_R16, PF_R16, ST_FLIP_SIGN }, { PF_SIGNED_RG16, PF_RG16, ST_FLIP_SIGN },
	{ PF_RGBA8, PF_R8, ST_KEEP }, { PF_RGBA8, PF_RG8, ST_KEEP },				// :791-794
	{ PF_RGB8, PF_R8, ST_KEEP }, { PF_RGB8, PF_RG8, ST_KEEP },
	{ PF_R8, PF_RGBX8, ST_PAD }, { PF_RG8, PF_RGBX8, ST_PAD },				// :797-798
	{ PF_R16, PF_R8, ST_NARROW }, { PF_RG16, PF_RG8, ST_NARROW }, { PF_RGB16, PF_RGB8, ST_NARROW },	// :800-809
	{ PF_RGBX16, PF_RGBX8, ST_NARROW }, { PF_RGBA16, PF_RGBA8, ST_NARROW },
	{ PF_R8, PF_R16, ST_WIDEN }, { PF_RG8, PF_RG16, ST_WIDEN }, { PF_RGB8, PF_RGB16, ST_WIDEN },
	{ PF_RGBX8, PF_RGBX16, ST_WIDEN }, { PF_RGBA8, PF_RGBA16, ST_WIDEN },
	{ PF_R16, PF_FLOAT_R16, ST_INT_TO_HALF }, { PF_RG16, PF_FLOAT_RG16, ST_INT_TO_HALF },	// :812-815
	{ PF_RGB16, PF_FLOAT_RGB16, ST_INT_TO_HALF }, { PF_RGBX16, PF_FLOAT_RGBX16, ST_INT_TO_HALF },
	{ PF_FLOAT_R16, PF_R16, ST_HALF_TO_INT }, { PF_FLOAT_RG16, PF_RG16, ST_HALF_TO_INT },	// :817-821
	{ PF_FLOAT_RGB16, PF_RGB16, ST_HALF_TO_INT }, { PF_FLOAT_RGBX16, PF_RGBX16, ST_HALF_TO_INT },
	{ PF_FLOAT_RGBA16, PF_RGBA16, ST_HALF_TO_INT },
	{ PF_FLOAT_R16_HDR, PF_R16, ST_HALF_HDR_TO_INT }, { PF_FLOAT_RG16_HDR, PF_RG16, ST_HALF_HDR_TO_INT },	// :823-825
	{ PF_FLOAT_RGBX16_HDR, PF_RGBX16, ST_HALF_HDR_TO_INT },
	{ PF_FLOAT_R32, PF_FLOAT_R16, ST_FLOAT_TO_HALF }, { PF_FLOAT_RG32, PF_FLOAT_RG16, ST_FLOAT_TO_HALF },	// :828-831
	{ PF_FLOAT_RGB32, PF_FLOAT_RGB16, ST_FLOAT_TO_HALF }, { PF_FLOAT_RGBX32, PF_FLOAT_RGBX16, ST_FLOAT_TO_HALF },
	{ PF_FLOAT_R32, PF_R16, ST_FLOAT_TO_INT }, { PF_FLOAT_RG32, PF_RG16, ST_FLOAT_TO_INT },	// :833-836
	{ PF_FLOAT_RGB32, PF_RGB16, ST_FLOAT_TO_INT }, { PF_FLOAT_RGBX32, PF_RGBX16, ST_FLOAT_TO_INT },
	{ PF_FLOAT_R16, PF_FLOAT_R32, ST_HALF_TO_FLOAT }, { PF_FLOAT_RG16, PF_FLOAT_RG32, ST_HALF_TO_FLOAT },	// :839-842
	{ PF_FLOAT_RGB16, PF_FLOAT_RGB32, ST_HALF_TO_FLOAT }, { PF_FLOAT_RGBX16, PF_FLOAT_RGBX32, ST_HALF_TO_FLOAT },
	{ PF_FLOAT_R32_HDR, PF_FLOAT_R32, ST_FLOAT_HDR }, { PF_FLOAT_RG32_HDR, PF_FLOAT_RG32, ST_FLOAT_HDR },	// :844-848
	{ PF_FLOAT_RGB32_HDR, PF_FLOAT_RGB32, ST_FLOAT_HDR }, { PF_FLOAT_RGBX32_HDR, PF_FLOAT_RGBX32, ST_FLOAT_HDR },
	{ PF_RGB8, PF_RGBX8, ST_PAD }, { PF_RGBX8, PF_RGB8, ST_KEEP },				// :851-852
	{ PF_FLOAT_RGB16, PF_FLOAT_RGBX16, ST_PAD }, { PF_FLOAT_RGBX16, PF_FLOAT_RGB16, ST_KEEP },	// :854-857
	{ PF_FLOAT_RGB16_HDR, PF_FLOAT_RGBX16_HDR, ST_PAD }, { PF_FLOAT_RGBX16_HDR, PF_FLOAT_RGB16_HDR, ST_KEEP },
	{ PF_FLOAT_RGB32, PF_FLOAT_RGBX32, ST_PAD }, { PF_FLOAT_RGBX32, PF_FLOAT_RGB32, ST_KEEP },	// :859-862
	{ PF_FLOAT_RGB32_HDR, PF_FLOAT_RGBX32_HDR, ST_PAD }, { PF_FLOAT_RGBX32_HDR, PF_FLOAT_RGB32_HDR, ST_KEEP },
};
constexpr int TABLE_SIZE = (int)(sizeof(conversion_table) / sizeof(conversion_table[0]));

// detex.h:884-898
static inline int nominal_components(uint32_t pf) { return 1 + (int)((pf & 0x30u) >> 4); }
static inline int component_bytes(uint32_t pf) { return 1 << (pf & 0x3u); }
static inline int component_precision(uint32_t pf) {
	const int cs = component_bytes(pf);
	return cs * 8 - ((pf & 0x2000u) ? 5 * (1 + (cs == 4)) : 0);
}
static inline int pixel_bytes(uint32_t pf) { return 1 + (int)((pf & 0xF00u) >> 8); }

// detexMatchConversion, convert.c:887-1045.  Returns the number of steps (0 for identical
// formats) or -1; steps[] receives table indices.
static inline int convert_plan_search(uint32_t s, uint32_t d, int steps[4]) {
	if (s == d)
		return 0;
	const TableEntry *T = conversion_table;
	for (int i = 0; i < TABLE_SIZE; i++)				// direct, :906-912
		if (T[i].target == d && T[i].source == s) {
			steps[0] = i;
			return 1;
		}
	int min_comp = nominal_components(s) < nominal_components(d) ? nominal_components(s) : nominal_components(d);
	int min_prec = component_precision(s) < component_precision(d) ? component_precision(s) : component_precision(d);
	auto keeps = [&](uint32_t pf) { return nominal_components(pf) >= min_comp && component_precision(pf) >= min_prec; };
	for (int i = 0; i < TABLE_SIZE; i++)				// two steps, :922-943
		if (T[i].target == d && keeps(T[i].source))
			for (int j = 0; j < TABLE_SIZE; j++)
				if (T[j].target == T[i].source && T[j].source == s) {
					steps[0] = j;
					steps[1] = i;
					return 2;
				}
	for (int i = 0; i < TABLE_SIZE; i++)				// three steps, :945-987
		if (T[i].source == s && keeps(T[i].target))
			for (int j = 0; j < TABLE_SIZE; j++)
				if (T[j].target == d && keeps(T[j].source))
					for (int k = 0; k < TABLE_SIZE; k++)
						if (T[k].target == T[j].source && T[k].source == T[i].target) {
							steps[0] = i;
							steps[1] = k;
							steps[2] = j;
							return 3;
						}
	for (int i = 0; i < TABLE_SIZE; i++)				// four steps, :989-1043
		if (T[i].source == s && keeps(T[i].target))
			for (int j = 0; j < TABLE_SIZE; j++)
				if (T[j].target == d && keeps(T[j].source))
					for (int k = 0; k < TABLE_SIZE; k++)
						if (T[k].source == T[i].target && keeps(T[k].target))
							for (int l = 0; l < TABLE_SIZE; l++)
								if (T[l].target == T[j].source && T[l].source == T[k].target) {
									steps[0] = i;
									steps[1] = k;
									steps[2] = l;
									steps[3] = j;
									return 4;
								}
	return -1;
}


// ---- device side ---------------------------------------------------------------------------------
struct ConvertProgram {
	int n_steps;
	uint16_t format[5];		// pixel format before step i / after the last step (low 16 bits)
	uint8_t op[4];
	float gamma, range_min, range_max;
};

// float -> half exactly as the reference's integer routine does it (half-float.c:93-161): zero and
// denormal inputs give signed zero, NaN gives 0xFE00, rounding is "add the first dropped bit".
__device__ __forceinline__ uint32_t float_bits_to_half(uint32_t x) {
	if ((x & 0x7FFFFFFFu) == 0u)
		return x >> 16;
	const uint32_t sign = (x & 0x80000000u) >> 16, xe = x & 0x7F800000u;
	uint32_t xm = x & 0x007FFFFFu;
	if (xe == 0u)
		return sign;
	if (xe == 0x7F800000u)
		return xm == 0u ? (sign | 0x7C00u) : 0xFE00u;
	const int he = (int)(xe >> 23) - 127 + 15;
	if (he >= 0x1F)
		return sign | 0x7C00u;
	if (he <= 0) {
		if (14 - he > 24)
			return sign;
		xm |= 0x00800000u;
		uint32_t hm = xm >> (14 - he);
		hm += (xm >> (13 - he)) & 1u;
		return (sign | hm) & 0xFFFFu;
	}
	const uint32_t h = sign | ((uint32_t)he << 10) | (xm >> 13);
	return (h + ((xm >> 12) & 1u)) & 0xFFFFu;
}

// half -> float (half-float.c:179-256; the reference reads it from a 65536-entry table built by
// this routine, half-float.c:263-271).
__device__ __forceinline__ uint32_t half_to_float_bits(uint32_t h) {
	if ((h & 0x7FFFu) == 0u)
		return h << 16;
	const uint32_t sign = (h & 0x8000u) << 16, he = h & 0x7C00u, hm = h & 0x03FFu;
	if (he == 0u) {
		const int top = 31 - __clz((int)hm);		// position of the leading mantissa bit, 0..9
		return sign | ((uint32_t)(103 + top) << 23) | (((hm << (10 - top)) & 0x03FFu) << 13);
	}
	if (he == 0x7C00u)
		return hm == 0u ? (sign | 0x7F800000u) : 0xFFC00000u;
	return sign | (((he >> 10) + 112u) << 23) | (hm << 13);
}

// detexClamp0To1 (detex.h:941-948) as the shipped -Ofast build evaluates it with max / min
// instructions: everything that is not greater than zero, -0.0 included, becomes +0.0.
// Decided on the bit pattern so that no min / max instruction selection can bring the sign of a
// negative zero or a negative denormal back.
__device__ __forceinline__ float clamp01(float f) {
	const uint32_t b = __float_as_uint(f);
	if ((b >> 31) != 0u || (b << 1) == 0u || (b << 1) > 0xFF000000u)	// negative, zero or NaN
		return 0.0f;
	return f > 1.0f ? 1.0f : f;
}

// lrintf(clamp(f) * 65535 + 0.5) with the rounding mode the reference selects, FE_DOWNWARD
// (half-float.c:306-310, :317-321): product, sum and float -> int all round towards -infinity.
__device__ __forceinline__ uint32_t normalized_float_to_u16(float f) {
	const float t = __fadd_rd(__fmul_rd(clamp01(f), 65535.0f), 0.5f);
	return (uint32_t)__float2int_rd(t) & 0xFFFFu;
}

// sign-preserving power, hdr.c:46-49
__device__ __forceinline__ float signed_pow(float f, float e) { return f >= 0.0f ? powf(f, e) : -powf(-f, e); }

__device__ __forceinline__ int stored_components(uint32_t pf) {
	return (1 + (int)((pf & 0xF00u) >> 8)) >> (pf & 0x3u);
}

// One step of the chain on the component registers of K pixels.  sf / tf: formats before and after.  The operation is
// the same for every pixel of a launch: the switch is taken once per K pixels, the bodies are unrolled over them.
template <int K>
__device__ __forceinline__ void apply_step_k(uint32_t op, uint32_t sf, uint32_t tf, uint32_t (&px)[K][4], const ConvertProgram &p) {
	const int n_in = stored_components(sf), n_out = stored_components(tf);
	const uint32_t cs_out = 1u << (tf & 3u);
	const uint32_t one = cs_out == 1u ? 0xFFu : (cs_out == 2u ? ((tf & 0x2000u) ? 0x3C00u : 0xFFFFu) : 0x3F800000u);
#define DETEX_EACH_PIXEL _Pragma("unroll") for (int k_ = 0; k_ < K; k_++)
	switch (op) {
	case ST_KEEP:
		break;
	case ST_SWAP_RB:
		DETEX_EACH_PIXEL {
			uint32_t (&c)[4] = px[k_];
			const uint32_t t = c[0];
			c[0] = c[2];
			c[2] = t;
		}
		break;
	case ST_RGB_TO_BGRX:					// convert.c:73-86
		DETEX_EACH_PIXEL {
			uint32_t (&c)[4] = px[k_];
			const uint32_t t = c[0];
			c[0] = c[2];
			c[2] = t;
			c[3] = 0xFFu;
		}
		break;
	case ST_FLIP_SIGN: {					// convert.c:89-186: -128 / +128, -32768 / +32768
		const uint32_t top = (sf & 1u) ? 0x8000u : 0x80u;
		DETEX_EACH_PIXEL {
#pragma unroll
			for (int j = 0; j < 4; j++)
				px[k_][j] ^= top;
		}
		break;
	}
	case ST_PAD:						// convert.c:237-260, :660-669, :692-698, :719-735
		DETEX_EACH_PIXEL {
#pragma unroll
			for (int j = 0; j < 4; j++)
				if (j >= n_in)
					px[k_][j] = j == 3 ? one : 0u;
		}
		break;
	case ST_NARROW:						// convert.c:263-331
		DETEX_EACH_PIXEL {
			uint32_t (&c)[4] = px[k_];
#pragma unroll
			for (int j = 0; j < 4; j++)
				c[j] = ((c[j] + 127u) * 255u) / 65535u;		// 16-bit inputs only; kernels.cu has the cheap form
			if (n_out == 4 && !(tf & 0x4u))
				c[3] = 0xFFu;				// RGBX16 -> RGBX8 writes alpha 0xFF (:312-316)
		}
		break;
	case ST_WIDEN:						// convert.c:333-404
		DETEX_EACH_PIXEL {
			uint32_t (&c)[4] = px[k_];
#pragma unroll
			for (int j = 0; j < 4; j++)
				c[j] = c[j] * 257u;			// v * 65535 / 255
			if (n_out == 4 && !(tf & 0x4u))
				c[3] = 0xFFFFu;				// RGBX8 -> RGBX16 writes 0xFFFF (:384)
		}
		break;
	case ST_INT_TO_HALF: {					// convert.c:473-586
		// RGBX16 reads its components through int16_t (:565-567) and writes 1.0 into X (:574)
		const bool quirk_signed = n_in == 4;
		DETEX_EACH_PIXEL {
			uint32_t (&c)[4] = px[k_];
#pragma unroll
			for (int j = 0; j < 4; j++) {
				const float v = quirk_signed ? (float)(int)(short)(unsigned short)c[j] : (float)c[j];
				c[j] = float_bits_to_half(__float_as_uint(__fmul_rn(v, 1.0f / 65535.0f)));
			}
			if (quirk_signed)
				c[3] = 0x3C00u;
		}
		break;
	}
	case ST_HALF_TO_INT:					// convert.c:589-608 -> half-float.c:304-312
		DETEX_EACH_PIXEL {
#pragma unroll
			for (int j = 0; j < 4; j++)
				px[k_][j] = normalized_float_to_u16(__uint_as_float(half_to_float_bits(px[k_][j])));
		}
		break;
	case ST_HALF_HDR_TO_INT: case ST_FLOAT_HDR: {		// hdr.c:119-214
		const bool half_in = op == ST_HALF_HDR_TO_INT;
		float lo = p.range_min, hi = p.range_max, e = 1.0f;
		const bool identity = p.gamma == 1.0f && lo == 0.0f && hi == 1.0f;
		float factor;
		if (p.gamma == 1.0f)
			factor = __fdiv_rd(1.0f, __fsub_rd(hi, lo));	// computed after fesetround(FE_DOWNWARD), :125-137
		else {
			e = 1.0f / p.gamma;
			lo = signed_pow(lo, e);
			hi = signed_pow(hi, e);
			factor = 1.0f / (hi - lo);
		}
		DETEX_EACH_PIXEL {
			uint32_t (&c)[4] = px[k_];
#pragma unroll
			for (int j = 0; j < 4; j++) {
				float f = half_in ? __uint_as_float(half_to_float_bits(c[j])) : __uint_as_float(c[j]);
				if (p.gamma == 1.0f) {
					if (!identity)
						f = __fmul_rd(__fsub_rd(f, lo), factor);
				}
				else {
					// the half path corrects every value (table, hdr.c:40-50); the float path only rescales
					if (half_in)
						f = signed_pow(f, e);
					f = (f - lo) * factor;
				}
				c[j] = half_in ? normalized_float_to_u16(f) : __float_as_uint(clamp01(f));
			}
		}
		break;
	}
	case ST_FLOAT_TO_HALF:					// convert.c:407-425
		DETEX_EACH_PIXEL {
#pragma unroll
			for (int j = 0; j < 4; j++)
				px[k_][j] = float_bits_to_half(px[k_][j]);
		}
		break;
	case ST_FLOAT_TO_INT:					// convert.c:428-446 -> half-float.c:315-322
		DETEX_EACH_PIXEL {
#pragma unroll
			for (int j = 0; j < 4; j++)
				px[k_][j] = normalized_float_to_u16(__uint_as_float(px[k_][j]));
		}
		break;
	case ST_HALF_TO_FLOAT:					// convert.c:449-467
		DETEX_EACH_PIXEL {
#pragma unroll
			for (int j = 0; j < 4; j++)
				px[k_][j] = half_to_float_bits(px[k_][j]);
		}
		break;
	}
#undef DETEX_EACH_PIXEL
}

// one pixel
__device__ __forceinline__ void apply_step(uint32_t op, uint32_t sf, uint32_t tf, uint32_t (&c)[4], const ConvertProgram &p) {
	uint32_t px[1][4] = { { c[0], c[1], c[2], c[3] } };
	apply_step_k<1>(op, sf, tf, px, p);
#pragma unroll
	for (int j = 0; j < 4; j++)
		c[j] = px[0][j];
}


// Components of one staged pixel <-> registers.  Pixels of 4, 8, 12 or 16 bytes are whole words
// (the tile is 16-byte aligned and the pixel size a multiple of 4): one to four 32-bit shared
// accesses instead of one per component; the other sizes go component by component.
__device__ __forceinline__ void load_pixel(const uint8_t *in, uint32_t ps, uint32_t cs, int n, uint32_t (&c)[4]) {
	if ((ps & 3u) == 0u) {
		const uint32_t *w = reinterpret_cast<const uint32_t *>(in);
		if (cs == 4u) {
#pragma unroll
			for (int j = 0; j < 4; j++)
				if (j < n)
					c[j] = w[j];
		}
		else if (cs == 2u) {
			const uint32_t a = w[0];
			c[0] = a & 0xFFFFu;
			c[1] = a >> 16;
			if (ps == 8u) {
				const uint32_t b = w[1];
				c[2] = b & 0xFFFFu;
				c[3] = b >> 16;
			}
		}
		else {
			const uint32_t a = w[0];
			c[0] = a & 0xFFu;
			c[1] = (a >> 8) & 0xFFu;
			c[2] = (a >> 16) & 0xFFu;
			c[3] = a >> 24;
		}
		return;
	}
#pragma unroll
	for (int j = 0; j < 4; j++)
		if (j < n)
			c[j] = cs == 1u ? in[j] : reinterpret_cast<const uint16_t *>(in)[j];
}

__device__ __forceinline__ void store_pixel(uint8_t *out, uint32_t ps, uint32_t cs, int n, const uint32_t (&c)[4]) {
	if ((ps & 3u) == 0u) {
		uint32_t *w = reinterpret_cast<uint32_t *>(out);
		if (cs == 4u) {
#pragma unroll
			for (int j = 0; j < 4; j++)
				if (j < n)
					w[j] = c[j];
		}
		else if (cs == 2u) {
			w[0] = (c[0] & 0xFFFFu) | (c[1] << 16);
			if (ps == 8u)
				w[1] = (c[2] & 0xFFFFu) | (c[3] << 16);
		}
		else
			w[0] = (c[0] & 0xFFu) | ((c[1] & 0xFFu) << 8) | ((c[2] & 0xFFu) << 16) | (c[3] << 24);
		return;
	}
#pragma unroll
	for (int j = 0; j < 4; j++)
		if (j < n) {
			if (cs == 1u)
				out[j] = (uint8_t)c[j];
			else
				reinterpret_cast<uint16_t *>(out)[j] = (uint16_t)c[j];
		}
}


}  // namespace detexb200
