// BC1 / BC1A / BC2 / BC3 and RGTC1/2 (BC4/BC5, unsigned + signed) block decoders.
// Behaviour follows reference decompress-bc.c:23-240 and decompress-rgtc.c:26-147;
// the division / clamp LUTs of the reference (division-tables.c) are replaced by
// exact multiply-shift arithmetic (their entries equal the closed forms, Q17).
#pragma once
#include "common.cuh"

namespace detexb200 {

// ---- S3TC colour block -----------------------------------------------------------
// Four palette entries as packed R | G<<8 | B<<16 (alpha byte clear).
// 565 -> 888 is a plain shift, no bit replication (Q1, decompress-bc.c:34-39).
// R and B are processed together in two 16-bit lanes of one register.
struct S3tcPalette {
	uint32_t p0, p1, p2, p3;
};

template <bool ALLOW_THREE_COLOUR>
__device__ __forceinline__ S3tcPalette s3tc_palette(uint32_t colors, bool &four_colour) {
	const uint32_t c0 = colors & 0xFFFFu, c1 = colors >> 16;
	four_colour = ALLOW_THREE_COLOUR ? (c0 > c1) : true;		// decompress-bc.c:40
	const uint32_t rb0 = (c0 >> 11) | ((c0 & 0x1Fu) << 16);		// R5 | B5 << 16
	const uint32_t rb1 = (c1 >> 11) | ((c1 & 0x1Fu) << 16);
	const uint32_t g0 = (c0 >> 5) & 0x3Fu, g1 = (c1 >> 5) & 0x3Fu;	// G6
	S3tcPalette pal;
	pal.p0 = (rb0 << 3) | (g0 << 10);
	pal.p1 = (rb1 << 3) | (g1 << 10);
	// Four-colour mode: floor((2a+b)/3) with a = 8*R5 is floor(8k/3), k = 2*R5a + R5b <= 93,
	// which equals (k*683) >> 8 (683/256 - 8/3 = 1.3e-3, k*1.3e-3 < 1/3); for green
	// floor(4k/3), k <= 189, equals (k*683) >> 9.  Products stay inside their 16-bit lane.
	const uint32_t rb2 = (((2u * rb0 + rb1) * 683u) >> 8) & 0x00FF00FFu;
	const uint32_t rb3 = (((rb0 + 2u * rb1) * 683u) >> 8) & 0x00FF00FFu;
	const uint32_t g2 = ((2u * g0 + g1) * 683u) >> 9;
	const uint32_t g3 = ((g0 + 2u * g1) * 683u) >> 9;
	pal.p2 = rb2 | (g2 << 8);
	pal.p3 = rb3 | (g3 << 8);
	if (ALLOW_THREE_COLOUR && !four_colour) {
		// (a+b)/2 with a = 8*R5: 4*(R5a+R5b); green 2*(G6a+G6b)   (decompress-bc.c:49-53)
		pal.p2 = ((rb0 + rb1) << 2) | ((g0 + g1) << 9);
		pal.p3 = 0u;
	}
	return pal;
}

__device__ __forceinline__ uint32_t s3tc_pick(const S3tcPalette &pal, uint32_t indices, int i) {
	const bool b0 = (indices >> (2 * i)) & 1u, b1 = (indices >> (2 * i + 1)) & 1u;
	const uint32_t lo = b0 ? pal.p1 : pal.p0, hi = b0 ? pal.p3 : pal.p2;
	return b1 ? hi : lo;
}

// BC1: alpha is 0xFF for every pixel, including the black entry of three-colour
// mode (Q2).  decompress-bc.c:23-61.
__device__ __forceinline__ bool decode_bc1(const uint2 blk, uint32_t, uint32_t, uint32_t (&px)[16]) {
	bool four;
	S3tcPalette pal = s3tc_palette<true>(blk.x, four);
	pal.p0 |= 0xFF000000u;	pal.p1 |= 0xFF000000u;	pal.p2 |= 0xFF000000u;	pal.p3 |= 0xFF000000u;
#pragma unroll
	for (int i = 0; i < 16; i++)
		px[i] = s3tc_pick(pal, blk.y, i);
	return true;
}

// BC1A: the black entry of three-colour mode is transparent (Q2); honours the
// OPAQUE_ONLY / NON_OPAQUE_ONLY flags.  decompress-bc.c:87-132.
__device__ __forceinline__ bool decode_bc1a(const uint2 blk, uint32_t, uint32_t flags, uint32_t (&px)[16]) {
	bool opaque;
	S3tcPalette pal = s3tc_palette<true>(blk.x, opaque);
	pal.p0 |= 0xFF000000u;	pal.p1 |= 0xFF000000u;	pal.p2 |= 0xFF000000u;
	if (opaque)
		pal.p3 |= 0xFF000000u;
#pragma unroll
	for (int i = 0; i < 16; i++)
		px[i] = s3tc_pick(pal, blk.y, i);
	return !((opaque && (flags & FLAG_NON_OPAQUE_ONLY)) || (!opaque && (flags & FLAG_OPAQUE_ONLY)));
}

// BC2: explicit 4-bit alpha scaled by 17 (= a*255/15), colour always four-colour
// (Q3).  decompress-bc.c:136-171.
__device__ __forceinline__ bool decode_bc2(const uint4 blk, uint32_t, uint32_t flags, uint32_t (&px)[16]) {
	bool four;
	const S3tcPalette pal = s3tc_palette<false>(blk.z, four);
#pragma unroll
	for (int i = 0; i < 16; i++) {
		const uint32_t a4 = ((i < 8 ? blk.x : blk.y) >> (4 * (i & 7))) & 0xFu;
		px[i] = a4 * 0x11000000u + s3tc_pick(pal, blk.w, i);
	}
	// ENCODE flag rejects c0 <= c1 (decompress-bc.c:146-149).
	return !(((blk.z & 0xFFFFu) <= (blk.z >> 16)) && (flags & FLAG_ENCODE));
}

// ---- 8-entry interpolated ramp shared by BC3 alpha and RGTC ----------------------
// Unsigned: a0 > a1: floor(((8-k)a0 + (k-1)a1)/7), k = 2..7; else
// floor(((6-k)a0 + (k-1)a1)/5), k = 2..5, then 0 and 255 (Q4; decompress-bc.c:212-234,
// decompress-rgtc.c:36-57).  x/7 == (x*9363)>>16 for x <= 1785 and x/5 == (x*13108)>>16
// for x <= 1275 (error terms < 1/7 resp. 1/5).  Result: bytes ramp[0..3] in lo,
// ramp[4..7] in hi, ready for a byte-permute lookup.
__device__ __forceinline__ void ramp8_unsigned(uint32_t a0, uint32_t a1, uint32_t &lo, uint32_t &hi) {
	const bool seven = a0 > a1;
	const uint32_t m = seven ? 9363u : 13108u;
	const uint32_t step = a1 - a0;			// wraps; only sums are used
	uint32_t t = (seven ? 6u : 4u) * a0 + a1;	// k = 2
	const uint32_t r2 = (t * m) >> 16;	t += step;
	const uint32_t r3 = (t * m) >> 16;	t += step;
	const uint32_t r4 = (t * m) >> 16;	t += step;
	const uint32_t r5 = (t * m) >> 16;	t += step;
	uint32_t r6 = (t * m) >> 16;		t += step;
	uint32_t r7 = (t * m) >> 16;
	if (!seven) {
		r6 = 0u;
		r7 = 255u;
	}
	lo = a0 | (a1 << 8) | (r2 << 16) | (r3 << 24);
	hi = r4 | (r5 << 8) | (r6 << 16) | (r7 << 24);
}

// 3-bit code of pixel i from the 48-bit code stream {s_lo (32 bits), s_hi (16 bits)}.
__device__ __forceinline__ uint32_t code3(uint32_t s_lo, uint32_t s_hi, int i) {
	const int sh = 3 * i;
	const uint32_t v = sh < 32 ? __funnelshift_r(s_lo, s_hi, sh) : (s_hi >> (sh - 32));
	return v & 7u;
}

// Four consecutive 3-bit codes (one pixel row) spread into the four selector nibbles
// of a byte permute.
__device__ __forceinline__ uint32_t row_selector(uint32_t s_lo, uint32_t s_hi, int row) {
	const int sh = 12 * row;
	const uint32_t x = sh < 32 ? __funnelshift_r(s_lo, s_hi, sh) : (s_hi >> (sh - 32));
	return (x & 0x7u) | ((x & 0x38u) << 1) | ((x & 0x1C0u) << 2) | ((x & 0xE00u) << 3);
}

// BC3: interpolated alpha + four-colour block.  decompress-bc.c:175-240.
__device__ __forceinline__ bool decode_bc3(const uint4 blk, uint32_t, uint32_t flags, uint32_t (&px)[16]) {
	const uint32_t a0 = blk.x & 0xFFu, a1 = (blk.x >> 8) & 0xFFu;
	uint32_t rlo, rhi;
	ramp8_unsigned(a0, a1, rlo, rhi);
	const uint32_t s_lo = __funnelshift_r(blk.x, blk.y, 16), s_hi = blk.y >> 16;
	bool four;
	const S3tcPalette pal = s3tc_palette<false>(blk.z, four);
#pragma unroll
	for (int i = 0; i < 16; i++) {
		const uint32_t a = prmt(rlo, rhi, code3(s_lo, s_hi, i));	// byte 0 = alpha
		px[i] = prmt(s3tc_pick(pal, blk.w, i), a, 0x4210);
	}
	bool ok = !(a0 > a1 && (flags & FLAG_OPAQUE_ONLY));			// :179-180
	ok = ok && !(((blk.z & 0xFFFFu) <= (blk.z >> 16)) && (flags & FLAG_ENCODE));	// :188-191
	return ok;
}

// ---- RGTC ------------------------------------------------------------------------
// One unsigned channel: the four pixel rows as four words of bytes.
__device__ __forceinline__ void rgtc_channel_u8(uint32_t w0, uint32_t w1, uint32_t (&rows)[4]) {
	uint32_t rlo, rhi;
	ramp8_unsigned(w0 & 0xFFu, (w0 >> 8) & 0xFFu, rlo, rhi);
	const uint32_t s_lo = __funnelshift_r(w0, w1, 16), s_hi = w1 >> 16;
#pragma unroll
	for (int r = 0; r < 4; r++)
		rows[r] = prmt(rlo, rhi, row_selector(s_lo, s_hi, r));
}

// RGTC1 -> R8 (4 words).  decompress-rgtc.c:64-68.
__device__ __forceinline__ bool decode_rgtc1(const uint2 blk, uint32_t, uint32_t, uint32_t (&px)[4]) {
	rgtc_channel_u8(blk.x, blk.y, px);
	return true;
}

// RGTC2 -> RG8 (8 words, R in the low byte of each pixel).  decompress-rgtc.c:72-77.
__device__ __forceinline__ bool decode_rgtc2(const uint4 blk, uint32_t, uint32_t, uint32_t (&px)[8]) {
	uint32_t r[4], g[4];
	rgtc_channel_u8(blk.x, blk.y, r);
	rgtc_channel_u8(blk.z, blk.w, g);
#pragma unroll
	for (int row = 0; row < 4; row++) {
		px[2 * row] = prmt(r[row], g[row], 0x5140);		// R0 G0 R1 G1
		px[2 * row + 1] = prmt(r[row], g[row], 0x7362);	// R2 G2 R3 G3
	}
	return true;
}

// One signed channel.  Endpoints are int8 with -128 clamped to -127; (-127,-128)
// is invalid; interpolation divides truncating toward zero (C '/' == the
// reference's LUT * sign, detex.h:966-982); each value v in [-127,127] maps to
// (v+127)*65535/254 - 32768 (Q5, decompress-rgtc.c:84-130).  The eight 16-bit results are
// computed once, split into a low-byte and a high-byte table and looked up four pixels at a
// time.  rows[y][0] = pixels x=0,1 of row y (16 bits each), rows[y][1] = pixels x=2,3.
__device__ __forceinline__ bool rgtc_channel_s16(uint32_t w0, uint32_t w1, uint32_t (&rows)[4][2]) {
	int l0 = (int)(int8_t)(w0 & 0xFFu), l1 = (int)(int8_t)((w0 >> 8) & 0xFFu);
	const bool valid = !(l0 == -127 && l1 == -128);
	l0 = max(l0, -127);
	l1 = max(l1, -127);
	const bool seven = l0 > l1;
	int v[8];
	v[0] = l0;
	v[1] = l1;
	const int step = l1 - l0;
	int t = (seven ? 6 : 4) * l0 + l1;
	// t / 7 resp. t / 5 truncating toward zero: |t| <= 889, so |t|/7 == (|t|*9363)>>16 and
	// |t|/5 == (|t|*13108)>>16 (checked exhaustively over all endpoint pairs); sign restored.
	const uint32_t m = seven ? 9363u : 13108u;
#pragma unroll
	for (int k = 2; k < 8; k++) {
		const int sgn = t >> 31;
		const int q = (int)(((uint32_t)((t ^ sgn) - sgn) * m) >> 16);
		v[k] = (q ^ sgn) - sgn;
		t += step;
	}
	if (!seven) {
		v[6] = -127;
		v[7] = 127;
	}
	uint32_t e[8];
#pragma unroll
	for (int k = 0; k < 8; k++)
		e[k] = (((uint32_t)(v[k] + 127) * 65535u) / 254u - 32768u) & 0xFFFFu;
	const uint32_t e01 = e[0] | (e[1] << 16), e23 = e[2] | (e[3] << 16), e45 = e[4] | (e[5] << 16), e67 = e[6] | (e[7] << 16);
	const uint32_t l_lo = prmt(e01, e23, 0x6420), l_hi = prmt(e45, e67, 0x6420);	// low bytes, byte k = entry k
	const uint32_t h_lo = prmt(e01, e23, 0x7531), h_hi = prmt(e45, e67, 0x7531);	// high bytes
	const uint32_t s_lo = __funnelshift_r(w0, w1, 16), s_hi = w1 >> 16;
#pragma unroll
	for (int r = 0; r < 4; r++) {
		const uint32_t sel = row_selector(s_lo, s_hi, r);
		const uint32_t l4 = prmt(l_lo, l_hi, sel), h4 = prmt(h_lo, h_hi, sel);	// byte x = pixel x
		rows[r][0] = prmt(l4, h4, 0x5140);		// pixel 0 | pixel 1 << 16
		rows[r][1] = prmt(l4, h4, 0x7362);		// pixel 2 | pixel 3 << 16
	}
	return valid;
}

// SIGNED_RGTC1 -> SIGNED_R16 (8 words).  decompress-rgtc.c:134-137.
__device__ __forceinline__ bool decode_signed_rgtc1(const uint2 blk, uint32_t, uint32_t, uint32_t (&px)[8]) {
	uint32_t rows[4][2];
	const bool ok = rgtc_channel_s16(blk.x, blk.y, rows);
#pragma unroll
	for (int r = 0; r < 4; r++) {
		px[2 * r] = rows[r][0];
		px[2 * r + 1] = rows[r][1];
	}
	return ok;
}

// SIGNED_RGTC2 -> SIGNED_RG16 (16 words).  decompress-rgtc.c:141-147.
__device__ __forceinline__ bool decode_signed_rgtc2(const uint4 blk, uint32_t, uint32_t, uint32_t (&px)[16]) {
	uint32_t r[4][2], g[4][2];
	const bool ok_r = rgtc_channel_s16(blk.x, blk.y, r);
	const bool ok_g = rgtc_channel_s16(blk.z, blk.w, g);
#pragma unroll
	for (int y = 0; y < 4; y++) {
		px[4 * y + 0] = prmt(r[y][0], g[y][0], 0x5410);
		px[4 * y + 1] = prmt(r[y][0], g[y][0], 0x7632);
		px[4 * y + 2] = prmt(r[y][1], g[y][1], 0x5410);
		px[4 * y + 3] = prmt(r[y][1], g[y][1], 0x7632);
	}
	return ok_r && ok_g;
}

}  // namespace detexb200
