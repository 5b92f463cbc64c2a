// ETC1 / ETC2 / ETC2 punch-through / ETC2+EAC alpha / EAC R11, RG11 (unsigned and
// signed) block decoders.  Behaviour follows reference decompress-etc.c:21-717 and
// decompress-eac.c:21-231.  ETC/EAC blocks are big-endian byte strings; blocks are
// held as little-endian words so byte k of the string is byte k&3 of word k>>2.
// Pixels are numbered column-major in the stream (x = i>>2, y = i&3) and stored
// row-major (Q12).  The clamp LUT of the reference is replaced by min/max (Q17).
#pragma once
#include "common.cuh"

namespace detexb200 {

enum : uint32_t { ETC_INDIVIDUAL = 1, ETC_DIFFERENTIAL = 2, ETC_T = 4, ETC_H = 8, ETC_PLANAR = 16 };

// Intensity modifier tables (decompress-etc.c:25-34: rows {a, b, -a, -b}) as byte LUTs
// indexed by the 3-bit table codeword: small = a, large = b.
__device__ __forceinline__ uint32_t etc_lut8(uint32_t lo, uint32_t hi, uint32_t cw) {
	// byte cw of the 64-bit constant hi:lo
	return prmt(lo, hi, cw) & 0xFFu;
}
__device__ __forceinline__ uint32_t etc_small(uint32_t cw) { return etc_lut8(0x0D090502u, 0x2F211812u, cw); }
__device__ __forceinline__ uint32_t etc_large(uint32_t cw) { return etc_lut8(0x2A1D1108u, 0xB76A503Cu, cw); }
// etc2_distance_table {3,6,11,16,23,32,41,64} (decompress-etc.c:200)
__device__ __forceinline__ uint32_t etc2_distance(uint32_t k) { return etc_lut8(0x100B0603u, 0x40292017u, k); }

// 5-bit base + 3-bit signed delta of one colour byte, as an 8-bit-scaled sum; the
// differential mode overflows when the sum leaves [0,248] (decompress-etc.c:104-110).
__device__ __forceinline__ int etc_diff_sum(uint32_t byte) {
	return (int)(byte & 0xF8u) + bfe_s(byte, 0, 3) * 8;
}
__device__ __forceinline__ bool etc_overflow(int sum) { return (uint32_t)sum > 248u; }

// Mode selected by a block whose differential bit is set (decompress-etc.c:328-360).
__device__ __forceinline__ uint32_t etc2_mode(uint32_t w0) {
	if (etc_overflow(etc_diff_sum(w0 & 0xFFu)))
		return ETC_T;
	if (etc_overflow(etc_diff_sum((w0 >> 8) & 0xFFu)))
		return ETC_H;
	if (etc_overflow(etc_diff_sum((w0 >> 16) & 0xFFu)))
		return ETC_PLANAR;
	return ETC_DIFFERENTIAL;
}

// 2-bit selector of stream pixel i: LSB = bit i, MSB = bit 16+i of the big-endian
// index word (decompress-etc.c:75-76; the pixel-15 MSB is honoured like the shipped
// build does).
__device__ __forceinline__ uint32_t etc_selector(uint32_t idx, int i) {
	return ((idx >> i) & 1u) | (((idx >> (16 + i)) & 1u) << 1);
}
__host__ __device__ constexpr int etc_dest(int i) { return (i & 3) * 4 + (i >> 2); }

__device__ __forceinline__ uint32_t pack_rgb(int r, int g, int b) {
	return (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16);
}

// Individual / differential sub-block modes (decompress-etc.c:89-180); PUNCH = the
// non-opaque punch-through variant (modifier 0 for selectors 0 and 2, selector 2 is
// a fully transparent black pixel, Q13, decompress-etc.c:472-499).
template <bool PUNCH>
__device__ __forceinline__ void etc_subblocks(uint32_t w0, uint32_t idx, bool differential, uint32_t (&px)[16]) {
	int base[2][3];
#pragma unroll
	for (int c = 0; c < 3; c++) {
		const uint32_t byte = (w0 >> (8 * c)) & 0xFFu;
		if (differential) {
			const int v1 = (int)(byte & 0xF8u), v2 = etc_diff_sum(byte);
			base[0][c] = v1 | (v1 >> 5);
			base[1][c] = v2 | ((v2 & 0xE0) >> 5);
		}
		else {
			base[0][c] = (int)((byte & 0xF0u) | (byte >> 4));
			base[1][c] = (int)((byte & 0x0Fu) | ((byte & 0x0Fu) << 4));
		}
	}
	const uint32_t b3 = w0 >> 24;
	const bool flip = b3 & 1u;
	int small_[2], large_[2];
	small_[0] = PUNCH ? 0 : (int)etc_small(b3 >> 5);
	large_[0] = (int)etc_large(b3 >> 5);
	small_[1] = PUNCH ? 0 : (int)etc_small((b3 >> 2) & 7u);
	large_[1] = (int)etc_large((b3 >> 2) & 7u);
#pragma unroll
	for (int i = 0; i < 16; i++) {
		// flip 0: columns 0-1 | 2-3 (i >= 8); flip 1: rows 0-1 | 2-3 ((i & 3) >= 2)
		const bool second = flip ? ((i & 3) >= 2) : (i >= 8);
		const uint32_t sel = etc_selector(idx, i);
		int m = (sel & 1u) ? (second ? large_[1] : large_[0]) : (second ? small_[1] : small_[0]);
		if (sel & 2u)
			m = -m;
		const int r = clamp255((second ? base[1][0] : base[0][0]) + m);
		const int g = clamp255((second ? base[1][1] : base[0][1]) + m);
		const int b = clamp255((second ? base[1][2] : base[0][2]) + m);
		uint32_t p = pack_rgb(r, g, b) | 0xFF000000u;
		if (PUNCH && sel == 2u)
			p = 0u;
		px[etc_dest(i)] = p;
	}
}

__device__ __forceinline__ int nibble8(uint32_t v) { return (int)((v << 4) | v); }

// T and H modes (decompress-etc.c:202-285, punch-through variant :565-649).
template <bool PUNCH>
__device__ __forceinline__ void etc2_t_or_h(uint32_t w0, uint32_t idx, bool h_mode, uint32_t (&px)[16]) {
	const uint32_t b0 = w0 & 0xFFu, b1 = (w0 >> 8) & 0xFFu, b2 = (w0 >> 16) & 0xFFu, b3 = w0 >> 24;
	int c1[3], c2[3];
	uint32_t paint[4];
	if (!h_mode) {
		c1[0] = nibble8(((b0 & 0x18u) >> 1) | (b0 & 0x3u));
		c1[1] = nibble8(b1 >> 4);
		c1[2] = nibble8(b1 & 0xFu);
		c2[0] = nibble8(b2 >> 4);
		c2[1] = nibble8(b2 & 0xFu);
		c2[2] = nibble8(b3 >> 4);
		const int d = (int)etc2_distance(((b3 & 0x0Cu) >> 1) | (b3 & 1u));
		paint[0] = pack_rgb(c1[0], c1[1], c1[2]);
		paint[1] = pack_rgb(clamp255(c2[0] + d), clamp255(c2[1] + d), clamp255(c2[2] + d));
		paint[2] = pack_rgb(c2[0], c2[1], c2[2]);
		paint[3] = pack_rgb(clamp255(c2[0] - d), clamp255(c2[1] - d), clamp255(c2[2] - d));
	}
	else {
		c1[0] = nibble8((b0 & 0x78u) >> 3);
		c1[1] = nibble8(((b0 & 0x07u) << 1) | ((b1 & 0x10u) >> 4));
		c1[2] = nibble8((b1 & 0x08u) | ((b1 & 0x03u) << 1) | (b2 >> 7));
		c2[0] = nibble8((b2 & 0x78u) >> 3);
		c2[1] = nibble8(((b2 & 0x07u) << 1) | (b3 >> 7));
		c2[2] = nibble8((b3 & 0x78u) >> 3);
		const int v1 = (c1[0] << 16) + (c1[1] << 8) + c1[2], v2 = (c2[0] << 16) + (c2[1] << 8) + c2[2];
		const int d = (int)etc2_distance((b3 & 0x04u) | ((b3 & 0x01u) << 1) | (v1 >= v2 ? 1u : 0u));
		paint[0] = pack_rgb(clamp255(c1[0] + d), clamp255(c1[1] + d), clamp255(c1[2] + d));
		paint[1] = pack_rgb(clamp255(c1[0] - d), clamp255(c1[1] - d), clamp255(c1[2] - d));
		paint[2] = pack_rgb(clamp255(c2[0] + d), clamp255(c2[1] + d), clamp255(c2[2] + d));
		paint[3] = pack_rgb(clamp255(c2[0] - d), clamp255(c2[1] - d), clamp255(c2[2] - d));
	}
	paint[0] |= 0xFF000000u;
	paint[1] |= 0xFF000000u;
	paint[2] = PUNCH ? 0u : (paint[2] | 0xFF000000u);	// Q13
	paint[3] |= 0xFF000000u;
#pragma unroll
	for (int i = 0; i < 16; i++) {
		const bool s0 = (idx >> i) & 1u, s1 = (idx >> (16 + i)) & 1u;
		const uint32_t lo = s0 ? paint[1] : paint[0], hi = s0 ? paint[3] : paint[2];
		px[etc_dest(i)] = s1 ? hi : lo;
	}
}

// Planar mode (decompress-etc.c:287-317): 6-7-6 O/H/V colours, bilinear ramp.
__device__ __forceinline__ void etc2_planar(uint32_t w0, uint32_t w1, uint32_t (&px)[16]) {
	const uint32_t hi = bswap32(w0), lo = bswap32(w1);	// big-endian halves: hi = stream bits 63..32
	int o[3], h[3], v[3];
	const uint32_t ro = (hi >> 25) & 0x3Fu;
	const uint32_t go = (((hi >> 24) & 1u) << 6) | ((hi >> 17) & 0x3Fu);
	const uint32_t bo = (((hi >> 16) & 1u) << 5) | (((hi >> 11) & 3u) << 3) | ((hi >> 7) & 7u);
	const uint32_t rh = (((hi >> 2) & 0x1Fu) << 1) | (hi & 1u);
	const uint32_t gh = (lo >> 25) & 0x7Fu;
	const uint32_t bh = (lo >> 19) & 0x3Fu;
	const uint32_t rv = (lo >> 13) & 0x3Fu;
	const uint32_t gv = (lo >> 6) & 0x7Fu;
	const uint32_t bv = lo & 0x3Fu;
	o[0] = (int)((ro << 2) | (ro >> 4));	o[1] = (int)((go << 1) | (go >> 6));	o[2] = (int)((bo << 2) | (bo >> 4));
	h[0] = (int)((rh << 2) | (rh >> 4));	h[1] = (int)((gh << 1) | (gh >> 6));	h[2] = (int)((bh << 2) | (bh >> 4));
	v[0] = (int)((rv << 2) | (rv >> 4));	v[1] = (int)((gv << 1) | (gv >> 6));	v[2] = (int)((bv << 2) | (bv >> 4));
	int dx[3], dy[3], base[3];
#pragma unroll
	for (int c = 0; c < 3; c++) {
		dx[c] = h[c] - o[c];
		dy[c] = v[c] - o[c];
		base[c] = 4 * o[c] + 2;
	}
#pragma unroll
	for (int y = 0; y < 4; y++)
#pragma unroll
		for (int x = 0; x < 4; x++) {
			const int r = clamp255((x * dx[0] + y * dy[0] + base[0]) >> 2);
			const int g = clamp255((x * dx[1] + y * dy[1] + base[1]) >> 2);
			const int b = clamp255((x * dx[2] + y * dy[2] + base[2]) >> 2);
			px[y * 4 + x] = pack_rgb(r, g, b) | 0xFF000000u;
		}
}

// ETC1 -> RGBX8.  Differential overflow is an error in this format only (Q11).
__device__ __forceinline__ bool decode_etc1(const uint2 blk, uint32_t mode_mask, uint32_t, uint32_t (&px)[16]) {
	const bool differential = (blk.x >> 24) & 2u;
	bool ok = mode_mask & (differential ? ETC_DIFFERENTIAL : ETC_INDIVIDUAL);
	if (differential)
		ok = ok && etc2_mode(blk.x) == ETC_DIFFERENTIAL;
	etc_subblocks<false>(blk.x, bswap32(blk.y), differential, px);
	return ok;
}

// ETC2 -> RGBX8.  decompress-etc.c:321-367.
__device__ __forceinline__ bool decode_etc2(const uint2 blk, uint32_t mode_mask, uint32_t, uint32_t (&px)[16]) {
	const uint32_t idx = bswap32(blk.y);
	const bool differential_bit = (blk.x >> 24) & 2u;
	const uint32_t mode = differential_bit ? etc2_mode(blk.x) : (uint32_t)ETC_INDIVIDUAL;
	if (mode == ETC_T || mode == ETC_H)
		etc2_t_or_h<false>(blk.x, idx, mode == ETC_H, px);
	else if (mode == ETC_PLANAR)
		etc2_planar(blk.x, blk.y, px);
	else
		etc_subblocks<false>(blk.x, idx, differential_bit, px);
	return (mode_mask & mode) != 0u;
}

// ETC2 punch-through -> RGBA8.  Bit 1 of byte 3 is the opaque flag
// (decompress-etc.c:653-717).
__device__ __forceinline__ bool decode_etc2_punchthrough(const uint2 blk, uint32_t mode_mask, uint32_t flags,
uint32_t (&px)[16]) {
	const uint32_t idx = bswap32(blk.y);
	const bool opaque = (blk.x >> 24) & 2u;
	const uint32_t mode = etc2_mode(blk.x);
	bool ok = (mode_mask & mode) != 0u;
	ok = ok && !(opaque && (flags & FLAG_NON_OPAQUE_ONLY)) && !(!opaque && (flags & FLAG_OPAQUE_ONLY));
	if (mode == ETC_T || mode == ETC_H) {
		if (opaque)
			etc2_t_or_h<false>(blk.x, idx, mode == ETC_H, px);
		else
			etc2_t_or_h<true>(blk.x, idx, mode == ETC_H, px);
	}
	else if (mode == ETC_PLANAR) {
		ok = ok && !(flags & FLAG_NON_OPAQUE_ONLY);		// :695-697
		etc2_planar(blk.x, blk.y, px);
	}
	else {
		if (opaque)
			etc_subblocks<false>(blk.x, idx, true, px);
		else
			etc_subblocks<true>(blk.x, idx, true, px);
	}
	return ok;
}

// ---- EAC ---------------------------------------------------------------------------
// eac_modifier_table (decompress-eac.c:21-38), one row of eight int8 per 64-bit entry.
__device__ __constant__ unsigned long long eac_modifier_rows[16] = {
	0x0E080502F1F7FAFDull, 0x0C090602F3F6F9FDull, 0x0C070401F3F8FBFEull, 0x0C050301F3FAFCFEull,
	0x0B070502F4F8FAFDull, 0x0A080602F5F7F9FDull, 0x0A070603F5F8F9FCull, 0x0A070402F5F8FBFDull,
	0x09070501F6F8FAFEull, 0x09070401F6F8FBFEull, 0x09070301F6F8FCFEull, 0x09060401F6F9FBFEull,
	0x09060302F6F9FCFDull, 0x09020100F6FDFEFFull, 0x08070503F7F8FAFCull, 0x08060402F7F9FBFDull,
};

// Modifier (sign-extended) of 3-bit selector sel from table row {lo, hi}.
__device__ __forceinline__ int eac_modifier(uint32_t lo, uint32_t hi, uint32_t sel) {
	return (int)(int8_t)(prmt(lo, hi, sel) & 0xFFu);
}

// 3-bit selector of stream pixel i: bits 47-3i .. 45-3i of the big-endian qword whose
// halves are hi (bits 63..32) and lo (bits 31..0)  (decompress-eac.c:46, :121).
__device__ __forceinline__ uint32_t eac_selector(uint32_t hi, uint32_t lo, int i) {
	const int sh = 45 - 3 * i;
	const uint32_t v = sh >= 32 ? (hi >> (sh - 32)) : __funnelshift_r(lo, hi, sh);
	return v & 7u;
}

// ETC2_EAC -> RGBA8: ETC2 colour in bytes 8..15, EAC alpha in bytes 0..7
// (decompress-eac.c:54-86).  Multiplier 0 is legal on decode (Q15).
__device__ __forceinline__ bool decode_etc2_eac(const uint4 blk, uint32_t mode_mask, uint32_t flags,
uint32_t (&px)[16]) {
	const bool ok_colour = decode_etc2(make_uint2(blk.z, blk.w), mode_mask, flags, px);
	const uint32_t hi = bswap32(blk.x), lo = bswap32(blk.y);
	const int base = (int)(hi >> 24), mult = (int)((hi >> 20) & 0xFu);
	const unsigned long long row = eac_modifier_rows[(hi >> 16) & 0xFu];
	const uint32_t rlo = (uint32_t)row, rhi = (uint32_t)(row >> 32);
#pragma unroll
	for (int i = 0; i < 16; i++) {
		const int a = clamp255(base + eac_modifier(rlo, rhi, eac_selector(hi, lo, i)) * mult);
		px[etc_dest(i)] = (px[etc_dest(i)] & 0x00FFFFFFu) | ((uint32_t)a << 24);
	}
	return ok_colour && !(mult == 0 && (flags & FLAG_ENCODE));
}

// One 11-bit EAC channel -> sixteen 16-bit values (row-major order).
// Unsigned: clamp(base*8+4 + mod*mult8, 0, 2047), replicated (v<<5)|(v>>6);
// signed: base int8 (-128 invalid), clamp(base*8 + mod*mult8, +-1023), magnitude
// replicated (m<<5)|(m>>5) then re-signed; multiplier 0 means x1 (Q14;
// decompress-eac.c:111-128, :167-202).
template <bool SIGNED>
__device__ __forceinline__ bool eac11_channel(uint32_t w0, uint32_t w1, uint32_t (&out16)[16]) {
	const uint32_t hi = bswap32(w0), lo = bswap32(w1);
	const unsigned long long row = eac_modifier_rows[(hi >> 16) & 0xFu];
	const uint32_t rlo = (uint32_t)row, rhi = (uint32_t)(row >> 32);
	int mult8 = (int)((hi >> 20) & 0xFu) * 8;
	if (mult8 == 0)
		mult8 = 1;
	int base;
	bool ok = true;
	if (SIGNED) {
		base = (int)(int8_t)(hi >> 24);
		ok = base != -128;
		base *= 8;
	}
	else
		base = (int)(hi >> 24) * 8 + 4;
#pragma unroll
	for (int i = 0; i < 16; i++) {
		int v = base + eac_modifier(rlo, rhi, eac_selector(hi, lo, i)) * mult8;
		uint32_t o;
		if (SIGNED) {
			v = min(max(v, -1023), 1023);
			const int mag = abs(v);
			const int rep = (mag << 5) | (mag >> 5);
			o = (uint32_t)(v < 0 ? -rep : rep) & 0xFFFFu;
		}
		else {
			v = min(max(v, 0), 2047);
			o = (uint32_t)((v << 5) | (v >> 6));
		}
		out16[etc_dest(i)] = o;
	}
	return ok;
}

// EAC_R11 / EAC_SIGNED_R11 -> R16 / SIGNED_R16 (8 words).
template <bool SIGNED>
__device__ __forceinline__ bool decode_eac_r11(const uint2 blk, uint32_t, uint32_t, uint32_t (&px)[8]) {
	uint32_t h[16];
	const bool ok = eac11_channel<SIGNED>(blk.x, blk.y, h);
#pragma unroll
	for (int i = 0; i < 8; i++)
		px[i] = h[2 * i] | (h[2 * i + 1] << 16);
	return ok;
}

// EAC_RG11 / EAC_SIGNED_RG11 -> RG16 / SIGNED_RG16 (16 words, R in the low half).
template <bool SIGNED>
__device__ __forceinline__ bool decode_eac_rg11(const uint4 blk, uint32_t, uint32_t, uint32_t (&px)[16]) {
	uint32_t r[16], g[16];
	const bool ok_r = eac11_channel<SIGNED>(blk.x, blk.y, r);
	const bool ok_g = eac11_channel<SIGNED>(blk.z, blk.w, g);
#pragma unroll
	for (int i = 0; i < 16; i++)
		px[i] = r[i] | (g[i] << 16);
	return ok_r && ok_g;
}

}  // namespace detexb200
