// ETC1 / ETC2 / ETC2 punch-through / ETC2+EAC alpha / EAC R11, RG11 (unsigned and
// signed) block decoders.  Behaviour follows reference decompress-etc.c:21-717 and
// decompress-eac.c:21-231.  ETC/EAC blocks are big-endian byte strings; blocks are
// held as little-endian words so byte k of the string is byte k&3 of word k>>2.
// Pixels are numbered column-major in the stream (x = i>>2, y = i&3) and stored
// row-major (Q12).  The clamp LUT of the reference is replaced by SIMD min/max (Q17).
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
	// select chain, no control flow (see etc_palette_pixels)
	const bool ov_r = etc_overflow(etc_diff_sum(w0 & 0xFFu));
	const bool ov_g = etc_overflow(etc_diff_sum((w0 >> 8) & 0xFFu));
	const bool ov_b = etc_overflow(etc_diff_sum((w0 >> 16) & 0xFFu));
	return ov_r ? (uint32_t)ETC_T : (ov_g ? (uint32_t)ETC_H : (ov_b ? (uint32_t)ETC_PLANAR : (uint32_t)ETC_DIFFERENTIAL));
}

// ---- palette engine shared by the individual, differential, T and H modes ----------------------
// In all four modes a pixel is one of at most eight colours, chosen by (sub-block g, 2-bit
// selector s): clamp(B[g][s] + M[g][s]) with
//   individual / differential: B = base colour of sub-block g, M = {a, b, -a, -b} of its table
//                              (decompress-etc.c:25-34, :72-85)
//   T: B = {c1, c2, c2, c2}, M = {0, d, 0, -d}                       (decompress-etc.c:209-237)
//   H: B = {c1, c1, c2, c2}, M = {d, -d, d, -d}                      (decompress-etc.c:238-273)
// The eight colours are computed once per block with 16-bit SIMD add-clamp instructions
// (VIADDMNMX.S16x2.RELU), laid out as per-channel byte tables (two registers per channel), and
// every pixel row is then four table look-ups by byte permute plus a 4x4 byte transpose --
// identical instruction stream for all four modes, so mixed warps do not diverge here.
// Colours are kept as two registers with 16-bit lanes: rb = R | B<<16, ga = G | A<<16.

// packed bytes (x, y, z, *) -> lanes x | z<<16 and y | 255<<16
__device__ __forceinline__ void etc_lanes(uint32_t rgb, uint32_t &rb, uint32_t &ga) {
	rb = prmt(rgb, 0u, 0x4240);
	ga = prmt(rgb, 0xFFu, 0x5451);
}
// clamp(colour + m) on R, G, B (alpha lane untouched)
__device__ __forceinline__ void etc_add_clamp(uint32_t rb, uint32_t ga, int m, uint32_t &orb, uint32_t &oga) {
	const uint32_t m2 = prmt((uint32_t)m, 0u, 0x1010), m1 = (uint32_t)m & 0xFFFFu;
	orb = __viaddmin_s16x2_relu(rb, m2, 0x00FF00FFu);
	oga = __viaddmin_s16x2_relu(ga, m1, 0x00FF00FFu);
}

// mode: ETC_INDIVIDUAL / ETC_DIFFERENTIAL / ETC_T / ETC_H.  punch: non-opaque punch-through
// block (modifier 0 for selectors 0 and 2, selector 2 -> transparent black, Q13).
// KIND narrows the modes the caller can pass, at compile time: ETC_KIND_ANY all four, ETC_KIND_SIMPLE individual /
// differential only (the T / H halves of every select fold away: the ETC1 decoder, and the main pass of the ETC2-family
// texture kernels, which queue T, H and planar blocks for a deferred pass), ETC_KIND_TH T / H only (the deferred pass).
enum EtcKind : int { ETC_KIND_ANY = 0, ETC_KIND_SIMPLE = 1, ETC_KIND_TH = 2 };
template <bool ALPHA, int KIND = ETC_KIND_ANY>
__device__ __forceinline__ void etc_palette_pixels(uint32_t w0, uint32_t idx, uint32_t mode, bool punch, uint32_t (&px)[16]) {
	// Opaque to the optimiser: a caller that knows the mode on one of its paths (e.g. "differential
	// bit clear -> individual") would otherwise get a specialised clone of this whole function on
	// that path, and warps holding both kinds of block would execute both clones.
	asm volatile("" : "+r"(mode));
	const uint32_t b0 = w0 & 0xFFu, b1 = (w0 >> 8) & 0xFFu, b2 = (w0 >> 16) & 0xFFu, b3 = w0 >> 24;
	const bool is_t = KIND != ETC_KIND_SIMPLE && mode == ETC_T;
	const bool is_h = KIND != ETC_KIND_SIMPLE && (KIND == ETC_KIND_TH ? !is_t : mode == ETC_H);
	const bool is_th = KIND == ETC_KIND_TH || (KIND != ETC_KIND_SIMPLE && (is_t || is_h));
	// Everything below is straight-line code: both variants of every mode-dependent quantity are
	// computed and one is picked with a select, so a warp with mixed modes runs it exactly once.
	// (With if/else on the mode the compiler clones the whole pixel loop per branch and a mixed
	// warp then executes it once per mode present -- measured 2.7x the instructions.)
	// -- sub-block base colours as packed bytes R, G, B (three channels at once)
	const uint32_t w = w0 & 0x00FFFFFFu;
	// differential: 5-bit base + 3-bit signed delta per byte.  The byte-parallel sum is evaluated
	// modulo 2^32; whenever this mode is really selected no channel overflows (every byte of the
	// true result is in [0,248]), so carries between bytes cancel and the word is exact.
	const uint32_t v1 = w & 0xF8F8F8u, d3 = w & 0x070707u;
	const uint32_t v2 = v1 + (d3 << 3) - ((d3 & 0x040404u) << 4);
	const bool differential = KIND != ETC_KIND_TH && mode == ETC_DIFFERENTIAL;
	const uint32_t base0 = differential ? (v1 | ((v1 >> 5) & 0x070707u)) : ((w & 0xF0F0F0u) | ((w >> 4) & 0x0F0F0Fu));
	const uint32_t base1 = differential ? (v2 | ((v2 >> 5) & 0x070707u)) : ((w & 0x0F0F0Fu) * 17u);
	// -- T / H base colours (4-bit channels replicated to 8 bits) and distance
	const uint32_t h_r1 = (b0 & 0x78u) >> 3, h_g1 = ((b0 & 0x07u) << 1) | ((b1 & 0x10u) >> 4);
	const uint32_t h_b1 = (b1 & 0x08u) | ((b1 & 0x03u) << 1) | (b2 >> 7);
	const uint32_t h_r2 = (b2 & 0x78u) >> 3, h_g2 = ((b2 & 0x07u) << 1) | (b3 >> 7), h_b2 = (b3 & 0x78u) >> 3;
	// ordering of the replicated 24-bit colours == ordering of the 12-bit nibble triples
	const uint32_t h_ge = ((h_r1 << 8) | (h_g1 << 4) | h_b1) >= ((h_r2 << 8) | (h_g2 << 4) | h_b2) ? 1u : 0u;
	const uint32_t t_r1 = ((b0 & 0x18u) >> 1) | (b0 & 0x3u);
	const uint32_t dist_index = is_h ? ((b3 & 0x04u) | ((b3 & 0x01u) << 1) | h_ge) : (((b3 & 0x0Cu) >> 1) | (b3 & 1u));
	const uint32_t dist = etc2_distance(dist_index);
	const uint32_t n1 = is_h ? (h_r1 | (h_g1 << 8) | (h_b1 << 16)) : (t_r1 | ((b1 >> 4) << 8) | ((b1 & 0xFu) << 16));
	const uint32_t n2 = is_h ? (h_r2 | (h_g2 << 8) | (h_b2 << 16)) : ((b2 >> 4) | ((b2 & 0xFu) << 8) | ((b3 >> 4) << 16));
	const uint32_t c1 = n1 * 17u, c2 = n2 * 17u;
	// -- the eight palette colours -> per-channel byte tables (entry n = g*4 + s)
	uint32_t tab_rb[2][2], tab_ga[2][2];	// [g][pair]: bytes (R_s, R_s+1, B_s, B_s+1) / (G.., A..)
#pragma unroll
	for (int g = 0; g < 2; g++) {
		const uint32_t cw = g == 0 ? (b3 >> 5) : ((b3 >> 2) & 7u);
		const int sm = punch ? 0 : (int)etc_small(cw), lg = (int)etc_large(cw), d = (int)dist;
		const uint32_t x = is_th ? c1 : (g == 0 ? base0 : base1);
		const uint32_t y = is_th ? c2 : (g == 0 ? base0 : base1);
		const int m0 = is_th ? (is_h ? d : 0) : sm;
		const int m1 = is_th ? (is_h ? -d : d) : lg;
		const int m2 = is_th ? (is_h ? d : 0) : -sm;
		const int m3 = is_th ? -d : -lg;
		uint32_t xrb, xga, yrb, yga;
		etc_lanes(x, xrb, xga);
		etc_lanes(y, yrb, yga);
		uint32_t erb[4], ega[4];
		etc_add_clamp(xrb, xga, m0, erb[0], ega[0]);
		etc_add_clamp(is_t ? yrb : xrb, is_t ? yga : xga, m1, erb[1], ega[1]);
		etc_add_clamp(yrb, yga, m2, erb[2], ega[2]);
		etc_add_clamp(yrb, yga, m3, erb[3], ega[3]);
		if (ALPHA && punch) {			// selector 2 is a fully transparent black pixel
			erb[2] = 0u;
			ega[2] = 0u;
		}
		tab_rb[g][0] = prmt(erb[0], erb[1], 0x6240);	// R0 R1 B0 B1
		tab_rb[g][1] = prmt(erb[2], erb[3], 0x6240);	// R2 R3 B2 B3
		tab_ga[g][0] = prmt(ega[0], ega[1], 0x6240);	// G0 G1 A0 A1
		tab_ga[g][1] = prmt(ega[2], ega[3], 0x6240);
	}
	uint32_t tr[2], tg[2], tb[2], ta[2];		// [g]: bytes = entries s = 0..3
#pragma unroll
	for (int g = 0; g < 2; g++) {
		tr[g] = prmt(tab_rb[g][0], tab_rb[g][1], 0x5410);
		tb[g] = prmt(tab_rb[g][0], tab_rb[g][1], 0x7632);
		tg[g] = prmt(tab_ga[g][0], tab_ga[g][1], 0x5410);
		ta[g] = prmt(tab_ga[g][0], tab_ga[g][1], 0x7632);
	}
	// -- pixels: output row y holds stream pixels i = 4x + y (Q12); selector nibble = s | g<<2
	const bool flip = b3 & 1u;
#pragma unroll
	for (int y = 0; y < 4; y++) {
		// LSBs of the four selectors sit at bits y, y+4, y+8, y+12 of idx, MSBs 16 higher:
		// exactly nibble spacing.  Sub-block: flip 0 -> columns 2,3 (nibbles 2,3), flip 1 -> rows 2,3.
		const uint32_t gsel = flip ? (y >= 2 ? 0x4444u : 0u) : 0x4400u;
		const uint32_t sel = ((idx >> y) & 0x1111u) | ((idx >> (15 + y)) & 0x2222u) | gsel;
		const uint32_t rq = prmt(tr[0], tr[1], sel), gq = prmt(tg[0], tg[1], sel), bq = prmt(tb[0], tb[1], sel);
		const uint32_t aq = ALPHA ? prmt(ta[0], ta[1], sel) : 0xFFFFFFFFu;
		const uint32_t rg01 = prmt(rq, gq, 0x5140), rg23 = prmt(rq, gq, 0x7362);
		const uint32_t ba01 = prmt(bq, aq, 0x5140), ba23 = prmt(bq, aq, 0x7362);
		px[4 * y + 0] = prmt(rg01, ba01, 0x5410);
		px[4 * y + 1] = prmt(rg01, ba01, 0x7632);
		px[4 * y + 2] = prmt(rg23, ba23, 0x5410);
		px[4 * y + 3] = prmt(rg23, ba23, 0x7632);
	}
}

// Planar mode (decompress-etc.c:287-317): 6-7-6 O/H/V colours, bilinear ramp
// clamp((x(H-O) + y(V-O) + 4O + 2) >> 2), evaluated with 16-bit SIMD lanes (R | B<<16 and G):
// clamp(t >> 2, 0, 255) == clamp(t, 0, 1023) >> 2, and the >> 2 + byte extraction is done by
// scaling with 64 so that the wanted byte is byte 1 of each lane.
__device__ __forceinline__ void etc2_planar(uint32_t w0, uint32_t w1, uint32_t (&px)[16]) {
	const uint32_t hi = bswap32(w0), lo = bswap32(w1);	// big-endian halves: hi = stream bits 63..32
	const uint32_t ro = (hi >> 25) & 0x3Fu;
	const uint32_t go = (((hi >> 24) & 1u) << 6) | ((hi >> 17) & 0x3Fu);
	const uint32_t bo = (((hi >> 16) & 1u) << 5) | (((hi >> 11) & 3u) << 3) | ((hi >> 7) & 7u);
	const uint32_t rh = (((hi >> 2) & 0x1Fu) << 1) | (hi & 1u);
	const uint32_t gh = (lo >> 25) & 0x7Fu;
	const uint32_t bh = (lo >> 19) & 0x3Fu;
	const uint32_t rv = (lo >> 13) & 0x3Fu;
	const uint32_t gv = (lo >> 6) & 0x7Fu;
	const uint32_t bv = lo & 0x3Fu;
	const uint32_t o_r = (ro << 2) | (ro >> 4), o_g = (go << 1) | (go >> 6), o_b = (bo << 2) | (bo >> 4);
	const uint32_t h_r = (rh << 2) | (rh >> 4), h_g = (gh << 1) | (gh >> 6), h_b = (bh << 2) | (bh >> 4);
	const uint32_t v_r = (rv << 2) | (rv >> 4), v_g = (gv << 1) | (gv >> 6), v_b = (bv << 2) | (bv >> 4);
	// per-lane signed 16-bit quantities (R low lane, B high lane; G alone)
	const uint32_t o_rb = o_r | (o_b << 16), h_rb = h_r | (h_b << 16), v_rb = v_r | (v_b << 16);
	const uint32_t dx_rb = __vsub2(h_rb, o_rb), dy_rb = __vsub2(v_rb, o_rb);
	const uint32_t dx_g = (h_g - o_g) & 0xFFFFu, dy_g = (v_g - o_g) & 0xFFFFu;
	uint32_t row_rb = o_rb * 4u + 0x00020002u, row_g = o_g * 4u + 2u;		// 4*O + 2 (<= 1022 per lane)
#pragma unroll
	for (int y = 0; y < 4; y++) {
		uint32_t t_rb = row_rb, t_g = row_g;
#pragma unroll
		for (int x = 0; x < 4; x++) {
			const uint32_t c_rb = __vimin_s16x2_relu(t_rb, 0x03FF03FFu) * 64u;	// byte 1 / 3 = (t >> 2)
			const uint32_t c_g = __vimin_s16x2_relu(t_g, 0x03FF03FFu) * 64u;
			px[4 * y + x] = prmt(c_rb, c_g, 0x7351) | 0xFF000000u;
			t_rb = __vadd2(t_rb, dx_rb);
			t_g = __vadd2(t_g, dx_g);
		}
		row_rb = __vadd2(row_rb, dy_rb);
		row_g = __vadd2(row_g, dy_g);
	}
}

// ETC1 -> RGBX8.  Differential overflow is an error in this format only (Q11).
__device__ __forceinline__ bool decode_etc1(const uint2 blk, uint32_t mode_mask, uint32_t, uint32_t (&px)[16]) {
	const bool differential = (blk.x >> 24) & 2u;
	bool ok = mode_mask & (differential ? ETC_DIFFERENTIAL : ETC_INDIVIDUAL);
	if (differential)
		ok = ok && etc2_mode(blk.x) == ETC_DIFFERENTIAL;
	etc_palette_pixels<false, ETC_KIND_SIMPLE>(blk.x, bswap32(blk.y), differential ? (uint32_t)ETC_DIFFERENTIAL : (uint32_t)ETC_INDIVIDUAL, false, px);
	return ok;
}

// Which code paths of an ETC2-family decoder get compiled in.  The texture kernels decode the individual / differential
// blocks (the ETC1 subset: 91 % of uniformly random blocks) in their main pass and queue the T, H and planar blocks for
// a deferred pass, so each pass only carries the code it can reach: ETC_PATHS_SIMPLE / ETC_PATHS_RARE.  (Round 1 deferred
// planar only -- ETC_PATHS_NO_PLANAR / ETC_PATHS_PLANAR_ONLY -- and its main pass computed both the sub-block and the
// T / H variant of every palette quantity for every block: ETC2 ran at 52 G blocks/s where ETC1 ran at 76.)
enum EtcPaths : int { ETC_PATHS_ALL = 0, ETC_PATHS_NO_PLANAR = 1, ETC_PATHS_PLANAR_ONLY = 2, ETC_PATHS_SIMPLE = 3, ETC_PATHS_RARE = 4 };

// ETC2 -> RGBX8.  decompress-etc.c:321-367.
template <int PATHS = ETC_PATHS_ALL>
__device__ __forceinline__ bool decode_etc2(const uint2 blk, uint32_t mode_mask, uint32_t, uint32_t (&px)[16]) {
	const uint32_t idx = bswap32(blk.y);
	const bool differential_bit = (blk.x >> 24) & 2u;
	const uint32_t mode_if_diff = etc2_mode(blk.x);
	const uint32_t mode = differential_bit ? mode_if_diff : (uint32_t)ETC_INDIVIDUAL;
	if constexpr (PATHS == ETC_PATHS_SIMPLE)
		etc_palette_pixels<false, ETC_KIND_SIMPLE>(blk.x, idx, mode, false, px);
	else if constexpr (PATHS == ETC_PATHS_RARE) {
		if (mode == ETC_PLANAR)
			etc2_planar(blk.x, blk.y, px);
		else
			etc_palette_pixels<false, ETC_KIND_TH>(blk.x, idx, mode, false, px);
	}
	else if (PATHS != ETC_PATHS_NO_PLANAR && (PATHS == ETC_PATHS_PLANAR_ONLY || mode == ETC_PLANAR))
		etc2_planar(blk.x, blk.y, px);
	else if (PATHS != ETC_PATHS_PLANAR_ONLY)
		etc_palette_pixels<false>(blk.x, idx, mode, false, px);
	return (mode_mask & mode) != 0u;
}

// ETC2 punch-through -> RGBA8.  Bit 1 of byte 3 is the opaque flag
// (decompress-etc.c:653-717).
template <int PATHS = ETC_PATHS_ALL>
__device__ __forceinline__ bool decode_etc2_punchthrough(const uint2 blk, uint32_t mode_mask, uint32_t flags,
uint32_t (&px)[16]) {
	const uint32_t idx = bswap32(blk.y);
	const bool opaque = (blk.x >> 24) & 2u;
	const uint32_t mode = etc2_mode(blk.x);
	bool ok = (mode_mask & mode) != 0u;
	ok = ok && !(opaque && (flags & FLAG_NON_OPAQUE_ONLY)) && !(!opaque && (flags & FLAG_OPAQUE_ONLY));
	if constexpr (PATHS == ETC_PATHS_SIMPLE)
		etc_palette_pixels<true, ETC_KIND_SIMPLE>(blk.x, idx, mode, !opaque, px);	// (always differential in this format)
	else if constexpr (PATHS == ETC_PATHS_RARE) {
		if (mode == ETC_PLANAR) {
			ok = ok && !(flags & FLAG_NON_OPAQUE_ONLY);		// :695-697
			etc2_planar(blk.x, blk.y, px);
		}
		else
			etc_palette_pixels<true, ETC_KIND_TH>(blk.x, idx, mode, !opaque, px);
	}
	else if (PATHS != ETC_PATHS_NO_PLANAR && (PATHS == ETC_PATHS_PLANAR_ONLY || mode == ETC_PLANAR)) {
		ok = ok && !(flags & FLAG_NON_OPAQUE_ONLY);		// :695-697
		etc2_planar(blk.x, blk.y, px);
	}
	else if (PATHS != ETC_PATHS_PLANAR_ONLY)
		etc_palette_pixels<true>(blk.x, idx, mode, !opaque, px);
	return ok;
}

// ---- EAC ---------------------------------------------------------------------------
// eac_modifier_table (decompress-eac.c:21-38).  Every row is {-a,-b,-c,-d, a-1,b-1,c-1,d-1};
// the magnitudes a..d of the 16 rows are kept as packed bytes.
// (global memory + read-only cache: the 64-byte table is one line, so a warp's 32 differently
// indexed reads are a single L1 wavefront; __constant__ would serialise them per distinct row)
__device__ __align__(64) const uint32_t eac_magnitudes[16] = {
	0x0F090603u, 0x0D0A0703u, 0x0D080502u, 0x0D060402u, 0x0C080603u, 0x0B090703u, 0x0B080704u, 0x0B080503u,
	0x0A080602u, 0x0A080502u, 0x0A080402u, 0x0A070502u, 0x0A070403u, 0x0A030201u, 0x09080604u, 0x09070503u,
};

// A pixel of an EAC channel takes one of eight values (3-bit selector), so -- as for the
// ETC palette -- the eight results are computed once per block, stored as byte tables and
// looked up four pixels at a time with byte permutes.
//
// Selector nibbles of output row y, REVERSED in x (nibble 0 = pixel x=3 ... nibble 3 = x=0):
// stream pixel i = 4x + y has its selector at bits 47-3i..45-3i of the big-endian qword
// {hi, lo}; after shifting right by 9-3y the four selectors sit 12 bits apart.
__device__ __forceinline__ uint32_t eac_row_selector_rev(uint32_t hi, uint32_t lo, int y) {
	const int sh = 9 - 3 * y;			// 9, 6, 3, 0
	const uint32_t q_lo = sh ? __funnelshift_r(lo, hi, sh) : lo, q_hi = hi >> sh;
	return (q_lo & 0x7u) | ((q_lo >> 8) & 0x70u) | ((q_lo >> 16) & 0x700u) | (__funnelshift_r(q_lo, q_hi, 24) & 0x7000u);
}

// ETC2_EAC -> RGBA8: ETC2 colour in bytes 8..15, EAC alpha in bytes 0..7
// (decompress-eac.c:54-86): alpha = clamp(base + modifier * multiplier, 0, 255).
// Multiplier 0 is legal on decode (Q15).
template <int PATHS = ETC_PATHS_ALL>
__device__ __forceinline__ bool decode_etc2_eac(const uint4 blk, uint32_t mode_mask, uint32_t flags,
uint32_t (&px)[16]) {
	const bool ok_colour = decode_etc2<PATHS>(make_uint2(blk.z, blk.w), mode_mask, flags, px);
	const uint32_t hi = bswap32(blk.x), lo = bswap32(blk.y);
	const uint32_t base = hi >> 24, mult = (hi >> 20) & 0xFu;
	const uint32_t mag = __ldg(&eac_magnitudes[(hi >> 16) & 0xFu]);
	// eight alpha values in 16-bit lanes: entries k and k+1 per register
	const uint32_t mag01 = prmt(mag, 0u, 0x4140), mag23 = prmt(mag, 0u, 0x4342);	// a | b<<16, c | d<<16
	const uint32_t base2 = base * 0x00010001u, mult2 = mult * 0x00010001u;
	const uint32_t p01 = mag01 * mult, p23 = mag23 * mult;				// <= 15*15 per lane
	const uint32_t up = __vsub2(base2, mult2);					// base - mult (entries 4..7: + mag*mult)
	const uint32_t e01 = __vimin_s16x2_relu(__vsub2(base2, p01), 0x00FF00FFu);
	const uint32_t e23 = __vimin_s16x2_relu(__vsub2(base2, p23), 0x00FF00FFu);
	const uint32_t e45 = __viaddmin_s16x2_relu(up, p01, 0x00FF00FFu);
	const uint32_t e67 = __viaddmin_s16x2_relu(up, p23, 0x00FF00FFu);
	const uint32_t t_lo = prmt(e01, e23, 0x6420), t_hi = prmt(e45, e67, 0x6420);	// byte k = entry k
#pragma unroll
	for (int y = 0; y < 4; y++) {
		const uint32_t a4 = prmt(t_lo, t_hi, eac_row_selector_rev(hi, lo, y));	// byte 3-x = alpha of pixel x
		px[4 * y + 0] = prmt(px[4 * y + 0], a4, 0x7210);
		px[4 * y + 1] = prmt(px[4 * y + 1], a4, 0x6210);
		px[4 * y + 2] = prmt(px[4 * y + 2], a4, 0x5210);
		px[4 * y + 3] = prmt(px[4 * y + 3], a4, 0x4210);
	}
	return ok_colour && !(mult == 0u && (flags & FLAG_ENCODE));
}

// One 11-bit EAC channel -> the four pixel rows as pairs of words (two 16-bit pixels each).
// Unsigned: clamp(base*8+4 + mod*mult8, 0, 2047), replicated (v<<5)|(v>>6);
// signed: base int8 (-128 invalid), clamp(base*8 + mod*mult8, +-1023), magnitude
// replicated (m<<5)|(m>>5) then re-signed; multiplier 0 means x1 (Q14;
// decompress-eac.c:111-128, :167-202).  rows[y][0] = pixels x=0,1, rows[y][1] = x=2,3.
template <bool SIGNED>
__device__ __forceinline__ bool eac11_channel(uint32_t w0, uint32_t w1, uint32_t (&rows)[4][2]) {
	const uint32_t hi = bswap32(w0), lo = bswap32(w1);
	const uint32_t mag = __ldg(&eac_magnitudes[(hi >> 16) & 0xFu]);
	uint32_t mult8 = ((hi >> 20) & 0xFu) * 8u;
	if (mult8 == 0u)
		mult8 = 1u;
	bool ok = true;
	uint32_t e01, e23, e45, e67;			// the eight 16-bit results, two per register
	if (!SIGNED) {
		const uint32_t base2 = ((hi >> 24) * 8u + 4u) * 0x00010001u, mult2 = mult8 * 0x00010001u;
		const uint32_t mag01 = prmt(mag, 0u, 0x4140), mag23 = prmt(mag, 0u, 0x4342);
		const uint32_t p01 = mag01 * mult8, p23 = mag23 * mult8;			// <= 15*120 per lane
		const uint32_t up = __vsub2(base2, mult2);
		const uint32_t v01 = __vimin_s16x2_relu(__vsub2(base2, p01), 0x07FF07FFu);
		const uint32_t v23 = __vimin_s16x2_relu(__vsub2(base2, p23), 0x07FF07FFu);
		const uint32_t v45 = __viaddmin_s16x2_relu(up, p01, 0x07FF07FFu);
		const uint32_t v67 = __viaddmin_s16x2_relu(up, p23, 0x07FF07FFu);
		e01 = (v01 << 5) | ((v01 >> 6) & 0x001F001Fu);
		e23 = (v23 << 5) | ((v23 >> 6) & 0x001F001Fu);
		e45 = (v45 << 5) | ((v45 >> 6) & 0x001F001Fu);
		e67 = (v67 << 5) | ((v67 >> 6) & 0x001F001Fu);
	}
	else {
		const int base = (int)(int8_t)(hi >> 24);
		ok = base != -128;
		uint32_t e[8];
#pragma unroll
		for (int k = 0; k < 8; k++) {
			const int m = (int)((mag >> (8 * (k & 3))) & 0xFFu);
			int v = base * 8 + (k < 4 ? -m : m - 1) * (int)mult8;
			v = min(max(v, -1023), 1023);
			const int a = abs(v), rep = (a << 5) | (a >> 5);
			e[k] = (uint32_t)(v < 0 ? -rep : rep) & 0xFFFFu;
		}
		e01 = e[0] | (e[1] << 16);	e23 = e[2] | (e[3] << 16);
		e45 = e[4] | (e[5] << 16);	e67 = e[6] | (e[7] << 16);
	}
	// low-byte and high-byte tables, byte k = entry k
	const uint32_t l_lo = prmt(e01, e23, 0x6420), l_hi = prmt(e45, e67, 0x6420);
	const uint32_t h_lo = prmt(e01, e23, 0x7531), h_hi = prmt(e45, e67, 0x7531);
#pragma unroll
	for (int y = 0; y < 4; y++) {
		const uint32_t sel = eac_row_selector_rev(hi, lo, y);
		const uint32_t l4 = prmt(l_lo, l_hi, sel), h4 = prmt(h_lo, h_hi, sel);	// byte 3-x = pixel x
		rows[y][0] = prmt(l4, h4, 0x6273);		// pixel 0 | pixel 1 << 16
		rows[y][1] = prmt(l4, h4, 0x4051);		// pixel 2 | pixel 3 << 16
	}
	return ok;
}

// EAC_R11 / EAC_SIGNED_R11 -> R16 / SIGNED_R16 (8 words).
template <bool SIGNED>
__device__ __forceinline__ bool decode_eac_r11(const uint2 blk, uint32_t, uint32_t, uint32_t (&px)[8]) {
	uint32_t rows[4][2];
	const bool ok = eac11_channel<SIGNED>(blk.x, blk.y, rows);
#pragma unroll
	for (int y = 0; y < 4; y++) {
		px[2 * y] = rows[y][0];
		px[2 * y + 1] = rows[y][1];
	}
	return ok;
}

// EAC_RG11 / EAC_SIGNED_RG11 -> RG16 / SIGNED_RG16 (16 words, R in the low half).
template <bool SIGNED>
__device__ __forceinline__ bool decode_eac_rg11(const uint4 blk, uint32_t, uint32_t, uint32_t (&px)[16]) {
	uint32_t r[4][2], g[4][2];
	const bool ok_r = eac11_channel<SIGNED>(blk.x, blk.y, r);
	const bool ok_g = eac11_channel<SIGNED>(blk.z, blk.w, g);
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
