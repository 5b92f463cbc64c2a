// BPTC (BC7) block -> palette + selectors.  Same arithmetic as decode_bptc.cuh / reference
// decompress-bptc.c:354-512, split differently:
//
//   A BC7 block holds at most NS * 2^IB distinct colours (per subset, the two endpoints and the
//   interpolation steps between them, decompress-bptc.c:182-193), and for every mode but 0 and 4
//   that is <= 16.  Instead of interpolating 16 pixels -- each with a run-time subset and a
//   run-time weight -- the lane that owns a block computes the PALETTE: every entry has a
//   compile-time subset and a compile-time weight (two IMADs with an immediate and one PRMT), and
//   the per-pixel work shrinks to building sixteen SELECTOR bytes (entry number of pixel i in
//   byte i).  The texture kernel parks palette + selectors in shared memory and the copy-out
//   threads do the sixteen look-ups while they un-sort and store (kernels.cu,
//   decode_bc7_palette_kernel).
//
//   Modes 0 (24 colours) and 4 (two index sets of different width, 32 combinations) do not fit
//   16 entries: they are decoded to pixels as before and use the identity selector.
#pragma once
#include "decode_bptc.cuh"

namespace detexb200 {

// Subset number of every pixel as one byte per pixel (pixel i in byte i of the uint4), derived
// once per CTA from the bit-packed partition tables.
struct Bc7PaletteTables {
	BptcTables packed;
	uint4 sub2[64];
	uint4 sub3[64];
};

// entry e_ of the byte tables from the packed words (thread-parallel over the 128 entries)
__device__ __forceinline__ uint4 bc7_subset_bytes(uint32_t bits, int bits_per_pixel) {
	uint32_t w[4];
	for (int k = 0; k < 4; k++) {
		uint32_t v = 0;
		for (int c = 0; c < 4; c++) {
			const int i = 4 * k + c;
			v |= ((bits >> (i * bits_per_pixel)) & ((1u << bits_per_pixel) - 1u)) << (8 * c);
		}
		w[k] = v;
	}
	uint4 r;
	r.x = w[0]; r.y = w[1]; r.z = w[2]; r.w = w[3];
	return r;
}

// Number of palette entries the mode stores (16 = pixels themselves for modes 0 and 4).
__host__ __device__ constexpr int bc7_palette_entries(int m) {
	return (m == 0 || m == 4) ? 16 : (m == 5 ? 16 : bc7_mode_info(m).ns << bc7_mode_info(m).ib);
}

// Sixteen IB-bit indices (pixel i at bits i*IB of the stream) -> sixteen bytes.
// Two multiplies place the even and the odd fields of a group of four at byte positions; the
// copies never overlap in the bits that are kept (checked per width below).
template <int IB, typename Stream>
__device__ __forceinline__ void bc7_spread_indices(Stream s, uint32_t (&b)[4]) {
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const uint32_t sh = (uint32_t)(s >> (4 * IB * k));
		if constexpr (IB == 2) {
			// fields at 0,2,4,6 -> 0,8,16,24.  evens: copies at +0,+12 (bits 0-5 | 12-17);
			// odds: copies at +6,+18 (bits 8-13 | 20-25)
			const uint32_t a = (sh & 0x33u) * 0x1001u, o = (sh & 0xCCu) * 0x40040u;
			b[k] = (o & 0x03000300u) | (a & 0x00030003u);
		}
		else if constexpr (IB == 3) {
			// fields at 0,3,6,9.  evens: copies at +0,+10 (bits 0-8 | 10-18); odds: +5,+15
			// (bits 8-16 | 18-26)
			const uint32_t a = (sh & 0x1C7u) * 0x401u, o = (sh & 0xE38u) * 0x8020u;
			b[k] = (o & 0x07000700u) | (a & 0x00070007u);
		}
		else {
			static_assert(IB == 4, "index width");
			b[k] = prmt(sh & 0x0F0Fu, (sh >> 4) & 0x0F0Fu, 0x5140);
		}
	}
}

// Palette + selectors of a block in mode M (1, 2, 3, 6, 7).  swap_rb: POST_SWAP_RB folded into
// the gather (BGRA8 / BGRX8 targets, convert.c:37-52).
template <int M>
__device__ __forceinline__ void bc7_palette_mode(const uint4 &w, const Bc7PaletteTables &tab, bool swap_rb,
uint32_t (&pal)[16], uint32_t (&sel)[4]) {
	using L = Bc7Layout<M>;
	constexpr Bc7ModeInfo I = bc7_mode_info(M);
	constexpr int NS = L::NS, NE = L::NE, IB = L::IB, NW = 1 << IB;
	static_assert(L::IB2 == 0 && NS * NW <= 16, "mode does not fit a 16-entry palette");
	constexpr int W2[4] = { 0, 21, 43, 64 };
	constexpr int W3[8] = { 0, 9, 18, 27, 37, 46, 55, 64 };
	constexpr int W4[16] = { 0, 4, 9, 13, 17, 21, 26, 30, 34, 38, 43, 47, 51, 55, 60, 64 };	// bptc-tables.c:192-201

	uint32_t e_rb[NE], e_ga[NE];
	bc7_endpoints<M>(w, e_rb, e_ga);
	const uint32_t sx = swap_rb ? 0x0202u : 0u;		// R <-> B: exchange the selectors of bytes 0 and 2
	const uint32_t g_end = 0x6240u ^ sx;			// bytes 0 / 2 of (rb, ga)
	const uint32_t g_mid = 0x7351u ^ sx;			// bytes 1 / 3 of (T_rb, T_ga)
#pragma unroll
	for (int s = 0; s < NS; s++) {
		const uint32_t rb0 = e_rb[2 * s], rb1 = e_rb[2 * s + 1], ga0 = e_ga[2 * s], ga1 = e_ga[2 * s + 1];
		// A = 4*(64*E0 + 32) per 16-bit lane, K = 4*(E1 - E0) as a packed integer difference
		// (see decode_bptc.cuh: every lane of A + w*K stays in [128, 65408])
		const uint32_t a_rb = rb0 * 256u + 0x00800080u, k_rb = (rb1 - rb0) * 4u;
		const uint32_t a_ga = ga0 * 256u + 0x00800080u, k_ga = (ga1 - ga0) * 4u;
		pal[s * NW] = prmt(rb0, ga0, g_end);
		pal[s * NW + NW - 1] = prmt(rb1, ga1, g_end);
#pragma unroll
		for (int j = 1; j < NW - 1; j++) {
			const uint32_t wj = (uint32_t)(IB == 2 ? W2[j & 3] : (IB == 3 ? W3[j & 7] : W4[j & 15]));
			pal[s * NW + j] = prmt(a_rb + wj * k_rb, a_ga + wj * k_ga, g_mid);
		}
	}

	// ---- selectors: index stream with the anchors' missing bits re-inserted, spread to bytes,
	// plus 2^IB times the subset number of the pixel
	const uint32_t partition = field128<L::P_PART, I.pb>(w);
	using Stream1 = typename IndexWord<IB>::type;
	Stream1 s1 = index_window<L::P_IDX, IB>(w);
	s1 = insert_zero_bit(s1, IB - 1);
	if constexpr (NS == 2) {
		const uint32_t anchor_b = tab.packed.two[partition] >> 16;
		s1 = insert_zero_bit(s1, anchor_b * IB + (IB - 1));
	}
	else if constexpr (NS == 3) {
		const uint32_t t = tab.packed.three[partition].y;
		const uint32_t anchor_b = t & 0xFu, anchor_c = t >> 8;
		const uint32_t qa = min(anchor_b, anchor_c) * IB + (IB - 1);
		const uint32_t qb = max(anchor_b, anchor_c) * IB + (IB - 1);
		s1 = insert_zero_bit(insert_zero_bit(s1, qa), qb);
	}
	bc7_spread_indices<IB>(s1, sel);
	if constexpr (NS >= 2) {
		const uint4 sb = NS == 2 ? tab.sub2[partition] : tab.sub3[partition];
		sel[0] = mad_u32(sb.x, (uint32_t)NW, sel[0]);
		sel[1] = mad_u32(sb.y, (uint32_t)NW, sel[1]);
		sel[2] = mad_u32(sb.z, (uint32_t)NW, sel[2]);
		sel[3] = mad_u32(sb.w, (uint32_t)NW, sel[3]);
	}
}

// Mode 5: one subset, 2-bit colour indices and 2-bit alpha indices (decompress-bptc.c:422-480)
// -> 4 x 4 combinations, entry = colour index + 4 * alpha index.  The channel rotation (:497-508)
// is part of the gather, as in decode_bptc.cuh.
__device__ __forceinline__ void bc7_palette_mode5(const uint4 &w, bool swap_rb, uint32_t (&pal)[16], uint32_t (&sel)[4]) {
	using L = Bc7Layout<5>;
	constexpr uint32_t W2[4] = { 0, 21, 43, 64 };
	uint32_t e_rb[2], e_ga[2];
	bc7_endpoints<5>(w, e_rb, e_ga);
	const uint32_t rotation = field128<L::P_ROT, 2>(w);
	uint32_t gather = rotation == 0u ? 0x7351u : (rotation == 1u ? 0x1357u : (rotation == 2u ? 0x5371u : 0x3751u));
	if (swap_rb)		// R <-> B of the finished pixel: exchange the selectors of bytes 0 and 2
		gather = (gather & 0xF0F0u) | ((gather & 0xFu) << 8) | ((gather >> 8) & 0xFu);
	const uint32_t a_rb = e_rb[0] * 256u + 0x00800080u, k_rb = (e_rb[1] - e_rb[0]) * 4u;
	const uint32_t a_ga = e_ga[0] * 256u + 0x00800080u, k_ga = (e_ga[1] - e_ga[0]) * 4u;
	// K_ga split into its green lane (sign-extended) and its alpha lane: the two weights differ
	const uint32_t kg = (uint32_t)(int)(int16_t)(k_ga & 0xFFFFu), ka = k_ga - kg;
#pragma unroll
	for (int c = 0; c < 4; c++) {
		const uint32_t t_rb = a_rb + W2[c] * k_rb, t_g = a_ga + W2[c] * kg;
#pragma unroll
		for (int a = 0; a < 4; a++)
			pal[c + 4 * a] = prmt(t_rb, t_g + W2[a] * ka, gather);
	}
	// selectors: both 31-bit streams get the anchor's zero bit, are spread to one nibble per pixel,
	// merged (colour + 4 * alpha) and the nibbles spread to bytes
	const uint32_t s1 = insert_zero_bit(index_window<L::P_IDX, 2>(w), 1u);
	const uint32_t s2 = insert_zero_bit(index_window<L::P_IDX2, 2>(w), 1u);
	auto nibbles = [](uint32_t x) {			// eight 2-bit fields (16 bits) -> eight nibbles
		x = (x | (x << 8)) & 0x00FF00FFu;
		x = (x | (x << 4)) & 0x0F0F0F0Fu;
		return (x | (x << 2)) & 0x33333333u;
	};
	const uint32_t n_lo = nibbles(s1 & 0xFFFFu) + 4u * nibbles(s2 & 0xFFFFu);
	const uint32_t n_hi = nibbles(s1 >> 16) + 4u * nibbles(s2 >> 16);
	bc7_spread_indices<4>(((uint64_t)n_hi << 32) | n_lo, sel);
}

// Mode of a block (lowest set bit of byte 0, decompress-bptc.c:229-237) or -1.
__device__ __forceinline__ int bc7_mode_of(const uint4 &blk) { return __ffs((int)(blk.x & 0xFFu)) - 1; }

// Any valid block -> (palette, selectors), handed to emit(pal, sel, std::integral_constant<int, NE>)
// with the compile-time number of entries the mode fills.  `mode` must be bc7_mode_of(blk) in
// [0, 7] (warp-uniform in the sorted texture kernel); validity (mode mask, flags) is the caller's
// business.
template <typename Emit>
__device__ __forceinline__ void bc7_palette(const uint4 &blk, int mode, const Bc7PaletteTables &tab, bool swap_rb,
Emit &&emit) {
	uint32_t pal[16], sel[4];
	switch (mode) {
	case 1: bc7_palette_mode<1>(blk, tab, swap_rb, pal, sel); emit(pal, sel, std::integral_constant<int, bc7_palette_entries(1)>{}); return;
	case 2: bc7_palette_mode<2>(blk, tab, swap_rb, pal, sel); emit(pal, sel, std::integral_constant<int, bc7_palette_entries(2)>{}); return;
	case 3: bc7_palette_mode<3>(blk, tab, swap_rb, pal, sel); emit(pal, sel, std::integral_constant<int, bc7_palette_entries(3)>{}); return;
	case 5: bc7_palette_mode5(blk, swap_rb, pal, sel); emit(pal, sel, std::integral_constant<int, 16>{}); return;
	case 6: bc7_palette_mode<6>(blk, tab, swap_rb, pal, sel); emit(pal, sel, std::integral_constant<int, bc7_palette_entries(6)>{}); return;
	case 7: bc7_palette_mode<7>(blk, tab, swap_rb, pal, sel); emit(pal, sel, std::integral_constant<int, bc7_palette_entries(7)>{}); return;
	default: break;
	}
	// modes 0 and 4: the sixteen pixels themselves, identity selectors
	if (mode == 0)
		bc7_decode_mode<0>(blk, tab.packed, pal);
	else
		bc7_decode_mode<4>(blk, tab.packed, pal);
	if (swap_rb) {
#pragma unroll
		for (int k = 0; k < 16; k++)
			pal[k] = prmt(pal[k], 0u, 0x3012);
	}
	sel[0] = 0x03020100u; sel[1] = 0x07060504u; sel[2] = 0x0B0A0908u; sel[3] = 0x0F0E0D0Cu;
	emit(pal, sel, std::integral_constant<int, 16>{});
}

}  // namespace detexb200
