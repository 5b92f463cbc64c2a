// BPTC (BC7) block decoder.  Behaviour follows reference decompress-bptc.c:23-512
// (mode layout comment :23-42, per-mode constants :45-71,134,195-221,265-267).
//
// Design notes (B200-native, not the reference's byte-array loops):
//   * every mode is a template instantiation: all field positions are compile-time
//     constants, so extraction is a funnel shift + mask per field;
//   * endpoints are kept as two registers per endpoint with 16-bit lanes
//     (R | B<<16 and G | A<<16) and interpolation is one IMAD per register:
//     T = 4*((64*E0 + 32) + w*(E1 - E0)), evaluated on the packed word.  Each lane's true
//     value lies in [128, 65408], so the packed integer identity holds without
//     cross-lane borrows even though (E1 - E0) is negative in some lanes; the wanted
//     byte ((64*E0+32+w*(E1-E0)) >> 6) is then byte 1 of each lane, and one PRMT gathers
//     R,G,B,A (and applies the mode 4/5 channel rotation for free);
//   * the "anchor" pixels that carry one index bit less are handled once per block
//     by inserting a zero bit into the 64-bit index stream (S += S & ~low(q) shifts
//     everything from bit q upwards left by one), after which every pixel's index sits
//     at the compile-time position i*IB;
//   * partition sets come from bit-packed shared-memory tables (bptc_tables.cuh).
#pragma once
#include <type_traits>
#include "common.cuh"
#include "bptc_tables.cuh"

// Modes 2, 3, 4, 5 and 7 -- the ones with 2-bit colour indices -- decode through per-channel byte tables and PRMT
// look-ups (four pixels per instruction) instead of sixteen per-pixel interpolations.  Measured on B200 against
// the per-pixel path (profiles/README.md, round 2: 16.8 M blocks, LINEAR, distribution M 45.2 -> 47.9 G blocks/s,
// U 36.3 -> 37.9): the default since round 2.  -DDETEXB200_BC7_PERPIXEL builds the per-pixel path for all modes
// (A/B runs; tests/test_emu_decoders.py keeps both bit-exact).
#if !defined(DETEXB200_BC7_PERPIXEL) && !defined(DETEXB200_BC7_TABLE2)
#define DETEXB200_BC7_TABLE2 1
#endif

namespace detexb200 {

// Shared-memory copy of the packed partition tables (768 bytes).
struct BptcTables {
	uint32_t two[64];
	uint2 three[64];
};

__device__ __forceinline__ void bptc_tables_load(BptcTables &t) {
	for (int i = threadIdx.x; i < 64; i += blockDim.x) {
		t.two[i] = bptc_two_global[i];
		t.three[i] = bptc_three_global[i];
	}
}

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F &&f) {
	if constexpr (I < N) {
		f(std::integral_constant<int, I>{});
		static_for<I + 1, N>(f);
	}
}

// Insert a zero bit at position q of an index stream (bits >= q move up by one): adding the
// part above q to the stream doubles it, i.e. shifts it left by one.
__device__ __forceinline__ uint64_t insert_zero_bit(uint64_t s, uint32_t q) {
	return s + (s & ~((1ull << q) - 1ull));
}
__device__ __forceinline__ uint32_t insert_zero_bit(uint32_t s, uint32_t q) {
	return s + (s & ~((1u << q) - 1u));
}
// Sixteen BITS-bit indices fit one 32-bit word when BITS == 2, else they need 64 bits.
template <int BITS> struct IndexWord { using type = uint64_t; };
template <> struct IndexWord<2> { using type = uint32_t; };
template <int POS, int BITS>
__device__ __forceinline__ typename IndexWord<BITS>::type index_window(const uint4 &w) {
	if constexpr (BITS == 2)
		return field128<POS, (POS + 32 <= 128 ? 32 : 128 - POS)>(w);
	else {
		const uint2 v = window64<POS>(w);
		return ((uint64_t)v.y << 32) | v.x;
	}
}
// a*b + c, pinned to the integer multiply-add (FMA pipe) instead of a compiler-chosen select
__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c) {
#ifdef DETEXB200_HOST_EMU
	return a * b + c;
#else
	uint32_t d;
	asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	return d;
#endif
}

// Interpolation weights aWeight2/3/4 (bptc-tables.c:192-201) in closed form:
// {0,21,43,64} = 21i + (i>>1); {0,9,18,27,37,46,55,64} = 9i + (i>>2);
// {0,4,9,13,...,60,64} = (273i + 32) >> 6.  The 2- and 3-bit tables fit one / two registers
// and are read with a single byte permute (selector = index; the unused selector nibbles pick
// table byte 0, which is 0, so the result is the clean weight).
template <int BITS>
__device__ __forceinline__ uint32_t bptc_weight(uint32_t i) {
	if constexpr (BITS == 2)
		return prmt(0x402B1500u, 0u, i);
	else if constexpr (BITS == 3)
		return prmt(0x1B120900u, 0x40372E25u, i);
	else
		return (i * 273u + 32u) >> 6;
}

struct Bc7ModeInfo {
	int ns, pb, rb, isb, cb, ab, epb, spb, ib, ib2;
};
__host__ __device__ constexpr Bc7ModeInfo bc7_mode_info(int m) {
	constexpr Bc7ModeInfo t[8] = {
		/*    NS PB RB ISB CB AB EPB SPB IB IB2 */
		{ 3, 4, 0, 0, 4, 0, 1, 0, 3, 0 },
		{ 2, 6, 0, 0, 6, 0, 0, 1, 3, 0 },
		{ 3, 6, 0, 0, 5, 0, 0, 0, 2, 0 },
		{ 2, 6, 0, 0, 7, 0, 1, 0, 2, 0 },
		{ 1, 0, 2, 1, 5, 6, 0, 0, 2, 3 },
		{ 1, 0, 2, 0, 7, 8, 0, 0, 2, 2 },
		{ 1, 0, 0, 0, 7, 7, 1, 0, 4, 0 },
		{ 2, 6, 0, 0, 5, 5, 1, 0, 2, 0 },
	};
	return t[m];
}

// Widen an N-bit value to 8 bits (left-align, replicate MSBs; decompress-bptc.c:160-175).
template <int N>
__device__ __forceinline__ uint32_t widen8(uint32_t v) {
	if constexpr (N >= 8)
		return v;
	else
		return ((v << (8 - N)) | (v >> (2 * N - 8))) & 0xFFu;
}

// Compile-time field positions of mode M (layout comment decompress-bptc.c:23-42).
template <int M> struct Bc7Layout {
	static constexpr Bc7ModeInfo I = bc7_mode_info(M);
	static constexpr int NS = I.ns, NE = 2 * I.ns, CB = I.cb, AB = I.ab, IB = I.ib, IB2 = I.ib2;
	static constexpr int P_PART = M + 1;
	static constexpr int P_ROT = P_PART + I.pb;
	static constexpr int P_ISB = P_ROT + I.rb;
	static constexpr int P_COL = P_ISB + I.isb;
	static constexpr int P_ALPHA = P_COL + 3 * NE * CB;
	static constexpr int P_PBIT = P_ALPHA + NE * AB;
	static constexpr int P_IDX = P_PBIT + (I.epb ? NE : (I.spb ? NS : 0));
	static constexpr int P_IDX2 = P_IDX + 16 * IB - NS;
	static_assert(IB2 == 0 || NS == 1, "dual index sets only occur in single-subset modes");
	static_assert(P_IDX2 + (IB2 ? 16 * IB2 - 1 : 0) == 128, "mode layout must fill the block");
};

// Endpoints of mode M: component-major fields (decompress-bptc.c:82-87) -> packed lanes
// e_rb[e] = R | B << 16, e_ga[e] = G | A << 16, widened to 8 bits.  Fields are packed two to a
// register first so that the p-bit append and the widening to 8 bits (left-align + MSB
// replication, :160-175) run once per register.
template <int M>
__device__ __forceinline__ void bc7_endpoints(const uint4 &w, uint32_t (&e_rb)[2 * bc7_mode_info(M).ns],
uint32_t (&e_ga)[2 * bc7_mode_info(M).ns]) {
	using L = Bc7Layout<M>;
	constexpr int NE = L::NE;
	if constexpr (M == 0) {
		// Mode 0: six 4-bit fields per channel, contiguous, then six p-bits.  All six endpoints of a
		// channel are widened together, one byte each: even endpoints (= endpoint 0 of subsets
		// 0,1,2) in bytes 0-2 of one register, odd endpoints in another; byte 3 stays zero and
		// serves as the zero byte of the PRMT that forms the 16-bit lanes.
		const uint32_t xr = field128<L::P_COL, 24>(w), xg = field128<L::P_COL + 24, 24>(w), xb = field128<L::P_COL + 48, 24>(w);
		const uint32_t p = field128<L::P_PBIT, 6>(w);
		// p-bits 0,2,4 -> bit 0 of bytes 0,1,2: copies at +0,+6,+12 touch nine distinct bit positions
		const uint32_t pe = ((p & 0x15u) * 0x1041u) & 0x010101u, po = (((p >> 1) & 0x15u) * 0x1041u) & 0x010101u;
		auto widen = [](uint32_t nib, uint32_t pb) {		// 4 bits + p-bit -> 5 -> 8 (:160-175), three bytes at once
			const uint32_t v5 = nib * 2u + pb;
			return (v5 * 8u) | ((v5 >> 2) & 0x070707u);
		};
		const uint32_t re = widen(xr & 0x0F0F0Fu, pe), ro = widen((xr >> 4) & 0x0F0F0Fu, po);
		const uint32_t ge = widen(xg & 0x0F0F0Fu, pe), go = widen((xg >> 4) & 0x0F0F0Fu, po);
		const uint32_t be = widen(xb & 0x0F0F0Fu, pe), bo = widen((xb >> 4) & 0x0F0F0Fu, po);
#pragma unroll
		for (int s = 0; s < 3; s++) {
			e_rb[2 * s] = prmt(re, be, 0x3430u + 0x0101u * s);	// {re.s, 0, be.s, 0}
			e_rb[2 * s + 1] = prmt(ro, bo, 0x3430u + 0x0101u * s);
			e_ga[2 * s] = prmt(ge, 0xFFu, 0x3430u + s);		// {ge.s, 0, 255, 0}: alpha 255 (:176-179)
			e_ga[2 * s + 1] = prmt(go, 0xFFu, 0x3430u + s);
		}
		return;
	}
	static_for<0, NE>([&](auto e_) {
		constexpr int e = decltype(e_)::value;
		constexpr Bc7ModeInfo J = bc7_mode_info(M);
		constexpr int CB = J.cb, AB = J.ab, NEJ = 2 * J.ns;
		constexpr int P_COL = Bc7Layout<M>::P_COL, P_ALPHA = Bc7Layout<M>::P_ALPHA, P_PBIT = Bc7Layout<M>::P_PBIT;
		constexpr int CPJ = J.cb + ((J.epb || J.spb) ? 1 : 0), APJ = J.ab + (J.epb ? 1 : 0);
		const uint32_t r = field128<P_COL + (0 * NEJ + e) * CB, CB>(w);
		const uint32_t g = field128<P_COL + (1 * NEJ + e) * CB, CB>(w);
		const uint32_t b = field128<P_COL + (2 * NEJ + e) * CB, CB>(w);
		const uint32_t a = field128<P_ALPHA + e * AB, AB>(w);
		uint32_t rb = r | (b << 16), ga = g | (a << 16);
		if constexpr (J.epb != 0) {
			// Q6: mode 6 keeps its two p-bits at stream bits 63 and 64 but the
			// reference reads both from data0 >> 63, so the second is always 0
			// (decompress-bptc.c:142-158).
			const uint32_t p = (M == 6 && e == 1) ? 0u : field128<P_PBIT + e, 1>(w);
			const uint32_t p2 = p * 0x00010001u;
			rb = (rb << 1) | p2;
			ga = (ga << 1) | p2;
		}
		else if constexpr (J.spb != 0) {	// mode 1: one p-bit per subset, colour only (:297-306)
			const uint32_t p = field128<P_PBIT + e / 2, 1>(w);
			rb = (rb << 1) | (p * 0x00010001u);
			ga = (ga << 1) | p;
		}
		if constexpr (CPJ < 8)
			rb = ((rb << (8 - CPJ)) | (rb >> (2 * CPJ - 8))) & 0x00FF00FFu;
		if constexpr (J.ab == 0) {				// modes 0-3: alpha 255 (:176-179)
			if constexpr (CPJ < 8)
				ga = ((ga << (8 - CPJ)) | (ga >> (2 * CPJ - 8))) & 0x000000FFu;
			ga = (ga & 0xFFu) | 0x00FF0000u;
		}
		else if constexpr (APJ == CPJ) {
			if constexpr (CPJ < 8)
				ga = ((ga << (8 - CPJ)) | (ga >> (2 * CPJ - 8))) & 0x00FF00FFu;
		}
		else						// modes 4, 5: alpha has its own precision
			ga = widen8<CPJ>(ga & 0xFFFFu) | (widen8<APJ>(ga >> 16) << 16);
		e_rb[e] = rb;
		e_ga[e] = ga;
	});
}

template <int M>
__device__ __forceinline__ void bc7_decode_mode(const uint4 &w, const BptcTables &tab, uint32_t (&px)[16]) {
	using L = Bc7Layout<M>;
	constexpr Bc7ModeInfo I = bc7_mode_info(M);
	constexpr int NS = L::NS, NE = L::NE, IB = L::IB, IB2 = L::IB2;
	constexpr int P_PART = L::P_PART, P_ROT = L::P_ROT, P_ISB = L::P_ISB, P_IDX = L::P_IDX, P_IDX2 = L::P_IDX2;

	const uint32_t partition = field128<P_PART, I.pb>(w);
	const uint32_t rotation = field128<P_ROT, I.rb>(w);
	const uint32_t isb = field128<P_ISB, I.isb>(w);

	uint32_t e_rb[NE], e_ga[NE];
	bc7_endpoints<M>(w, e_rb, e_ga);
	// Per subset: A = 4*(64*E0 + 32) per lane, K = 4*(E1 - E0) (packed integer difference).
	uint32_t a_rb[NS], k_rb[NS], a_ga[NS], k_ga[NS];
#pragma unroll
	for (int s = 0; s < NS; s++) {
		a_rb[s] = e_rb[2 * s] * 256u + 0x00800080u;
		k_rb[s] = (e_rb[2 * s + 1] - e_rb[2 * s]) * 4u;
		a_ga[s] = e_ga[2 * s] * 256u + 0x00800080u;
		k_ga[s] = (e_ga[2 * s + 1] - e_ga[2 * s]) * 4u;
	}

	// ---- partition / anchors
	uint32_t subset_bits = 0u, anchor_b = 0u, anchor_c = 0u;
	if constexpr (NS == 2) {
		const uint32_t t = tab.two[partition];
		subset_bits = t & 0xFFFFu;
		anchor_b = t >> 16;
	}
	else if constexpr (NS == 3) {
		const uint2 t = tab.three[partition];
		subset_bits = t.x;
		anchor_b = t.y & 0xFu;
		anchor_c = t.y >> 8;
	}

	// ---- index streams, anchors expanded to full width (decompress-bptc.c:409-420).  With 2-bit
	// indices the expanded stream is exactly 32 bits, otherwise 48 or 64.
	using Stream1 = typename IndexWord<IB>::type;
	Stream1 s1 = index_window<P_IDX, IB>(w);
	s1 = insert_zero_bit(s1, IB - 1);
	if constexpr (NS == 2)
		s1 = insert_zero_bit(s1, anchor_b * IB + (IB - 1));
	else if constexpr (NS == 3) {
		const uint32_t qa = min(anchor_b, anchor_c) * IB + (IB - 1);
		const uint32_t qb = max(anchor_b, anchor_c) * IB + (IB - 1);
		s1 = insert_zero_bit(insert_zero_bit(s1, qa), qb);
	}
	using Stream2 = typename IndexWord<IB2 ? IB2 : 2>::type;
	Stream2 s2 = 0;
	if constexpr (IB2 > 0)
		s2 = insert_zero_bit(index_window<P_IDX2, IB2>(w), IB2 - 1);
	// Gather selector: byte 1 / 3 of T_rb (R, B) and of T_ga (G, A) -> R G B A; the rotation of
	// modes 4/5 (swap A with R / G / B after interpolation, :497-508) is folded into it.
	uint32_t gather = 0x7351u;
	if constexpr (I.rb > 0)
		gather = rotation == 0u ? 0x7351u : (rotation == 1u ? 0x1357u : (rotation == 2u ? 0x5371u : 0x3751u));
	// Subset membership as one bit per pixel (bit i of m1: pixel i is in subset 1; m2: subset 2),
	// and the per-subset differences of A and K: the endpoint pair of a pixel is then picked with
	// multiply-adds (X0 + m1*(X1-X0) + m2*(X2-X0)) that issue on the FMA pipe, leaving the ALU
	// pipe -- the busier one in this kernel -- to the bit-field work.
	uint32_t m1 = 0u, m2 = 0u;
	uint32_t d1[4] = { 0u, 0u, 0u, 0u }, d2[4] = { 0u, 0u, 0u, 0u };
	if constexpr (NS == 2)
		m1 = subset_bits;
	else if constexpr (NS == 3) {
		const uint32_t lo = subset_bits & 0x55555555u, hi = (subset_bits >> 1) & 0x55555555u;
		m1 = lo & ~hi;		// 2-bit fields: bit 2i
		m2 = hi & ~lo;
	}
	if constexpr (NS >= 2) {
		d1[0] = a_rb[1] - a_rb[0];	d1[1] = k_rb[1] - k_rb[0];	d1[2] = a_ga[1] - a_ga[0];	d1[3] = k_ga[1] - k_ga[0];
	}
	if constexpr (NS == 3) {
		d2[0] = a_rb[NS - 1] - a_rb[0];	d2[1] = k_rb[NS - 1] - k_rb[0];	d2[2] = a_ga[NS - 1] - a_ga[0];	d2[3] = k_ga[NS - 1] - k_ga[0];
	}
	// (Modes 0 and 1 -- 3-bit indices, eight candidates per subset and channel = two registers, every subset looked up
	// for every pixel row and merged through byte masks -- were tried as table look-ups too in round 2 and measured
	// SLOWER than the per-pixel path below: all blocks in mode 0: 43.6 -> 42.2 G blocks/s, mode 1: 51.7 -> 48.3; building
	// 3 x 6 table registers per block costs more than the sixteen per-pixel selections it saves.  Not kept.)
#ifdef DETEXB200_BC7_TABLE2
	// Modes 3 and 7 -- two subsets, 2-bit indices -- as a byte-table look-up.  A pixel
	// is one of 2 x 4 colours, so per channel the eight candidates fit two registers and one PRMT per channel
	// fetches FOUR pixels (selector nibble = index | subset << 2), as the ETC palette engine does; the 4x4 byte
	// transpose that follows costs less than sixteen per-pixel endpoint selections and interpolations.
	if constexpr (NS == 2 && IB == 2 && IB2 == 0) {
		uint32_t tab_r[2], tab_g[2], tab_b[2], tab_a[2];
#pragma unroll
		for (int s = 0; s < 2; s++) {
			const uint32_t t1_rb = a_rb[s] + 21u * k_rb[s], t2_rb = a_rb[s] + 43u * k_rb[s], t3_rb = a_rb[s] + 64u * k_rb[s];
			const uint32_t t1_ga = a_ga[s] + 21u * k_ga[s], t2_ga = a_ga[s] + 43u * k_ga[s], t3_ga = a_ga[s] + 64u * k_ga[s];
			const uint32_t rb01 = prmt(a_rb[s], t1_rb, 0x7351), rb23 = prmt(t2_rb, t3_rb, 0x7351);	// R0 R1 B0 B1 / R2 R3 B2 B3
			const uint32_t ga01 = prmt(a_ga[s], t1_ga, 0x7351), ga23 = prmt(t2_ga, t3_ga, 0x7351);
			tab_r[s] = prmt(rb01, rb23, 0x5410);
			tab_b[s] = prmt(rb01, rb23, 0x7632);
			tab_g[s] = prmt(ga01, ga23, 0x5410);
			tab_a[s] = prmt(ga01, ga23, 0x7632);
		}
		// 2-bit indices and 1-bit subset flags of eight pixels -> eight selector nibbles
		auto spread2 = [](uint32_t v) {		// 16 bits (8 x 2) -> 8 nibbles
			v = (v | (v << 8)) & 0x00FF00FFu;
			v = (v | (v << 4)) & 0x0F0F0F0Fu;
			return (v | (v << 2)) & 0x33333333u;
		};
		auto spread1 = [](uint32_t v) {		// 8 bits -> bit 2 of 8 nibbles
			v = (v | (v << 12)) & 0x000F000Fu;
			v = (v | (v << 6)) & 0x03030303u;
			return ((v | (v << 3)) & 0x11111111u) << 2;
		};
		const uint32_t sel01 = spread2(s1 & 0xFFFFu) | spread1(subset_bits & 0xFFu);	// pixel rows 0, 1
		const uint32_t sel23 = spread2(s1 >> 16) | spread1((subset_bits >> 8) & 0xFFu);	// pixel rows 2, 3
#pragma unroll
		for (int y = 0; y < 4; y++) {
			const uint32_t word = y < 2 ? sel01 : sel23, sel = (y & 1) ? (word >> 16) : word;
			const uint32_t rq = prmt(tab_r[0], tab_r[1], sel), gq = prmt(tab_g[0], tab_g[1], sel);
			const uint32_t bq = prmt(tab_b[0], tab_b[1], sel), aq = prmt(tab_a[0], tab_a[1], sel);
			const uint32_t rg01 = prmt(rq, gq, 0x5140), rg23 = prmt(rq, gq, 0x7362);
			const uint32_t ba01 = prmt(bq, aq, 0x5140), ba23 = prmt(bq, aq, 0x7362);
			px[4 * y + 0] = prmt(rg01, ba01, 0x5410);
			px[4 * y + 1] = prmt(rg01, ba01, 0x7632);
			px[4 * y + 2] = prmt(rg23, ba23, 0x5410);
			px[4 * y + 3] = prmt(rg23, ba23, 0x7632);
		}
		return;
	}
#endif
#ifdef DETEXB200_BC7_TABLE2
	// Mode 2: three subsets, 2-bit indices, opaque.  Subsets 0 and 1 share one look-up (selector
	// nibble = index | (subset & 1) << 2), subset 2 gets a second one, and a byte mask (0xFF where the pixel is
	// in subset 2) picks between them.
	if constexpr (M == 2) {
		uint32_t tab_r[3], tab_g[3], tab_b[3];
#pragma unroll
		for (int s = 0; s < 3; s++) {
			const uint32_t t1_rb = a_rb[s] + 21u * k_rb[s], t2_rb = a_rb[s] + 43u * k_rb[s], t3_rb = a_rb[s] + 64u * k_rb[s];
			const uint32_t t1_ga = a_ga[s] + 21u * k_ga[s], t2_ga = a_ga[s] + 43u * k_ga[s], t3_ga = a_ga[s] + 64u * k_ga[s];
			const uint32_t rb01 = prmt(a_rb[s], t1_rb, 0x7351), rb23 = prmt(t2_rb, t3_rb, 0x7351);
			tab_r[s] = prmt(rb01, rb23, 0x5410);
			tab_b[s] = prmt(rb01, rb23, 0x7632);
			tab_g[s] = prmt(prmt(a_ga[s], t1_ga, 0x5151), prmt(t2_ga, t3_ga, 0x5151), 0x5410);
		}
		auto spread2 = [](uint32_t v) {		// 16 bits (8 x 2) -> 8 nibbles
			v = (v | (v << 8)) & 0x00FF00FFu;
			v = (v | (v << 4)) & 0x0F0F0F0Fu;
			return (v | (v << 2)) & 0x33333333u;
		};
		const uint32_t sub1 = subset_bits & 0x55555555u;		// bit 2i: pixel i in subset 1 (subset values 0, 1, 2)
		const uint32_t sel01 = spread2(s1 & 0xFFFFu) | (spread2(sub1 & 0xFFFFu) << 2);
		const uint32_t sel23 = spread2(s1 >> 16) | (spread2(sub1 >> 16) << 2);
#pragma unroll
		for (int y = 0; y < 4; y++) {
			const uint32_t word = y < 2 ? sel01 : sel23, sel = (y & 1) ? (word >> 16) : word;
			// bits 8y+1, +3, +5, +7 of subset_bits: pixel in subset 2 -> 0xFF in its byte
			const uint32_t in2 = (subset_bits >> (8 * y + 1)) & 0x55u;
			const uint32_t mask2 = ((in2 * 0x41041u) & 0x01010101u) * 255u;
			const uint32_t rq = (prmt(tab_r[0], tab_r[1], sel) & ~mask2) | (prmt(tab_r[2], 0u, sel) & mask2);
			const uint32_t gq = (prmt(tab_g[0], tab_g[1], sel) & ~mask2) | (prmt(tab_g[2], 0u, sel) & mask2);
			const uint32_t bq = (prmt(tab_b[0], tab_b[1], sel) & ~mask2) | (prmt(tab_b[2], 0u, sel) & mask2);
			const uint32_t rg01 = prmt(rq, gq, 0x5140), rg23 = prmt(rq, gq, 0x7362);
			const uint32_t ba01 = prmt(bq, 0xFFFFFFFFu, 0x5140), ba23 = prmt(bq, 0xFFFFFFFFu, 0x7362);
			px[4 * y + 0] = prmt(rg01, ba01, 0x5410);
			px[4 * y + 1] = prmt(rg01, ba01, 0x7632);
			px[4 * y + 2] = prmt(rg23, ba23, 0x5410);
			px[4 * y + 3] = prmt(rg23, ba23, 0x7632);
		}
		return;
	}
#endif
	// mode 4/5: K_ga split into its green lane (sign-extended) and its alpha lane
	const uint32_t kg0 = (uint32_t)(int)(int16_t)(k_ga[0] & 0xFFFFu), ka0 = k_ga[0] - kg0;
#ifdef DETEXB200_BC7_TABLE2
	// Mode 5: one subset, 2-bit colour indices and 2-bit alpha indices.  Four candidates per
	// channel = one register per channel; R, G, B are fetched with the colour selector, A with the alpha
	// selector, and the rotation (swap A with R / G / B, :497-508) is folded into the selectors of the last
	// transposing PRMT.
	if constexpr (M == 5) {
		const uint32_t c1 = a_rb[0] + 21u * k_rb[0], c2 = a_rb[0] + 43u * k_rb[0], c3 = a_rb[0] + 64u * k_rb[0];
		const uint32_t g1 = a_ga[0] + 21u * kg0, g2 = a_ga[0] + 43u * kg0, g3 = a_ga[0] + 64u * kg0;	// byte 1 = green
		const uint32_t l1 = a_ga[0] + 21u * ka0, l2 = a_ga[0] + 43u * ka0, l3 = a_ga[0] + 64u * ka0;	// byte 3 = alpha
		const uint32_t rb01 = prmt(a_rb[0], c1, 0x7351), rb23 = prmt(c2, c3, 0x7351);
		const uint32_t tab_r = prmt(rb01, rb23, 0x5410), tab_b = prmt(rb01, rb23, 0x7632);
		const uint32_t tab_g = prmt(prmt(a_ga[0], g1, 0x5151), prmt(g2, g3, 0x5151), 0x5410);
		const uint32_t tab_a = prmt(prmt(a_ga[0], l1, 0x7373), prmt(l2, l3, 0x7373), 0x5410);
		auto spread2 = [](uint32_t v) {		// 16 bits (8 x 2) -> 8 nibbles
			v = (v | (v << 8)) & 0x00FF00FFu;
			v = (v | (v << 4)) & 0x0F0F0F0Fu;
			return (v | (v << 2)) & 0x33333333u;
		};
		const uint32_t selc01 = spread2(s1 & 0xFFFFu), selc23 = spread2(s1 >> 16);
		const uint32_t sela01 = spread2(s2 & 0xFFFFu), sela23 = spread2(s2 >> 16);
		// {R0 G0 R1 G1} x {B0 A0 B1 A1} -> pixel: byte indices R 0, G 1, B 4, A 5 (+2 for the odd pixel)
		const uint32_t even = rotation == 0u ? 0x5410u : (rotation == 1u ? 0x0415u : (rotation == 2u ? 0x1450u : 0x4510u));
		const uint32_t odd = even + 0x2222u;
#pragma unroll
		for (int y = 0; y < 4; y++) {
			const uint32_t wc = y < 2 ? selc01 : selc23, wa = y < 2 ? sela01 : sela23;
			const uint32_t sc = (y & 1) ? (wc >> 16) : wc, sa = (y & 1) ? (wa >> 16) : wa;
			const uint32_t rq = prmt(tab_r, 0u, sc), gq = prmt(tab_g, 0u, sc), bq = prmt(tab_b, 0u, sc), aq = prmt(tab_a, 0u, sa);
			const uint32_t rg01 = prmt(rq, gq, 0x5140), rg23 = prmt(rq, gq, 0x7362);
			const uint32_t ba01 = prmt(bq, aq, 0x5140), ba23 = prmt(bq, aq, 0x7362);
			px[4 * y + 0] = prmt(rg01, ba01, even);
			px[4 * y + 1] = prmt(rg01, ba01, odd);
			px[4 * y + 2] = prmt(rg23, ba23, even);
			px[4 * y + 3] = prmt(rg23, ba23, odd);
		}
		return;
	}
	// Mode 4: one subset, a 2-bit and a 3-bit index set; the index selection bit says which of
	// them drives the colour (:380-385).  Eight candidates per channel (two registers; the 2-bit case fills
	// entries 4-7 with copies of 0-3), their weights picked per block from the 2- or the 3-bit table.
	if constexpr (M == 4) {
		constexpr uint32_t W2[8] = { 0, 21, 43, 64, 0, 21, 43, 64 }, W3[8] = { 0, 9, 18, 27, 37, 46, 55, 64 };
		uint32_t c[8], g[8], l[8];		// byte 1 / 3 of c: R / B; byte 1 of g: G; byte 3 of l: A
#pragma unroll
		for (int j = 0; j < 8; j++) {
			const uint32_t wc = isb ? W3[j] : W2[j], wa = isb ? W2[j] : W3[j];
			c[j] = a_rb[0] + wc * k_rb[0];
			g[j] = a_ga[0] + wc * kg0;
			l[j] = a_ga[0] + wa * ka0;
		}
		uint32_t tab_r[2], tab_g[2], tab_b[2], tab_a[2];
#pragma unroll
		for (int h = 0; h < 2; h++) {
			const uint32_t rb01 = prmt(c[4 * h], c[4 * h + 1], 0x7351), rb23 = prmt(c[4 * h + 2], c[4 * h + 3], 0x7351);
			tab_r[h] = prmt(rb01, rb23, 0x5410);
			tab_b[h] = prmt(rb01, rb23, 0x7632);
			tab_g[h] = prmt(prmt(g[4 * h], g[4 * h + 1], 0x5151), prmt(g[4 * h + 2], g[4 * h + 3], 0x5151), 0x5410);
			tab_a[h] = prmt(prmt(l[4 * h], l[4 * h + 1], 0x7373), prmt(l[4 * h + 2], l[4 * h + 3], 0x7373), 0x5410);
		}
		auto spread2 = [](uint32_t v) {		// 16 bits (8 x 2) -> 8 nibbles
			v = (v | (v << 8)) & 0x00FF00FFu;
			v = (v | (v << 4)) & 0x0F0F0F0Fu;
			return (v | (v << 2)) & 0x33333333u;
		};
		auto spread3 = [](uint32_t v) {		// 24 bits (8 x 3) -> 8 nibbles
			v = (v & 0x00000FFFu) | ((v & 0x00FFF000u) << 4);
			v = (v & 0x003F003Fu) | ((v & 0x0FC00FC0u) << 2);
			return (v & 0x07070707u) | ((v & 0x38383838u) << 1);
		};
		const uint32_t two01 = spread2(s1 & 0xFFFFu), two23 = spread2(s1 >> 16);
		const uint32_t three01 = spread3((uint32_t)s2 & 0x00FFFFFFu), three23 = spread3((uint32_t)(s2 >> 24) & 0x00FFFFFFu);
		const uint32_t selc01 = isb ? three01 : two01, selc23 = isb ? three23 : two23;
		const uint32_t sela01 = isb ? two01 : three01, sela23 = isb ? two23 : three23;
		const uint32_t even = rotation == 0u ? 0x5410u : (rotation == 1u ? 0x0415u : (rotation == 2u ? 0x1450u : 0x4510u));
		const uint32_t odd = even + 0x2222u;
#pragma unroll
		for (int y = 0; y < 4; y++) {
			const uint32_t wc = y < 2 ? selc01 : selc23, wa = y < 2 ? sela01 : sela23;
			const uint32_t sc = (y & 1) ? (wc >> 16) : wc, sa = (y & 1) ? (wa >> 16) : wa;
			const uint32_t rq = prmt(tab_r[0], tab_r[1], sc), gq = prmt(tab_g[0], tab_g[1], sc);
			const uint32_t bq = prmt(tab_b[0], tab_b[1], sc), aq = prmt(tab_a[0], tab_a[1], sa);
			const uint32_t rg01 = prmt(rq, gq, 0x5140), rg23 = prmt(rq, gq, 0x7362);
			const uint32_t ba01 = prmt(bq, aq, 0x5140), ba23 = prmt(bq, aq, 0x7362);
			px[4 * y + 0] = prmt(rg01, ba01, even);
			px[4 * y + 1] = prmt(rg01, ba01, odd);
			px[4 * y + 2] = prmt(rg23, ba23, even);
			px[4 * y + 3] = prmt(rg23, ba23, odd);
		}
		return;
	}
#endif

#ifdef DETEXB200_BC7_TABLE2
	constexpr bool TABLE_MODE = M == 2 || M == 3 || M == 4 || M == 5 || M == 7;	// handled above
#else
	constexpr bool TABLE_MODE = false;
#endif
	if constexpr (!TABLE_MODE) {
#pragma unroll
	for (int i = 0; i < 16; i++) {
		uint32_t arb = a_rb[0], krb = k_rb[0], aga = a_ga[0], kga = k_ga[0];
		if constexpr (NS >= 2) {
			constexpr int STEP = NS == 2 ? 1 : 2;
			const uint32_t b1 = (m1 >> (STEP * i)) & 1u;
			arb = mad_u32(b1, d1[0], arb);	krb = mad_u32(b1, d1[1], krb);
			aga = mad_u32(b1, d1[2], aga);	kga = mad_u32(b1, d1[3], kga);
			if constexpr (NS == 3) {
				const uint32_t b2 = (m2 >> (2 * i)) & 1u;
				arb = mad_u32(b2, d2[0], arb);	krb = mad_u32(b2, d2[1], krb);
				aga = mad_u32(b2, d2[2], aga);	kga = mad_u32(b2, d2[3], kga);
			}
		}
		const uint32_t i1 = (uint32_t)(s1 >> (i * IB)) & ((1u << IB) - 1u);
		uint32_t t_rb, t_ga;
		if constexpr (IB2 == 0) {
			const uint32_t wgt = bptc_weight<IB>(i1);
			t_rb = arb + wgt * krb;
			t_ga = aga + wgt * kga;
		}
		else {
			// Two index sets: colour and alpha weights may differ; in mode 4 the index
			// selection bit swaps which set drives which (:380-385, :422-480).
			const uint32_t i2 = (uint32_t)(s2 >> (i * IB2)) & ((1u << IB2) - 1u);
			const uint32_t w1 = bptc_weight<IB>(i1), w2 = bptc_weight<IB2>(i2);
			const uint32_t wc = isb ? w2 : w1, wa = isb ? w1 : w2;
			t_rb = arb + wc * krb;
			t_ga = aga + wc * kg0 + wa * ka0;
		}
		px[i] = prmt(t_rb, t_ga, gather);
	}
	}
}

// BPTC -> RGBA8.  decompress-bptc.c:354-512.
__device__ __forceinline__ bool decode_bptc(const uint4 blk, uint32_t mode_mask, uint32_t flags,
const BptcTables &tab, uint32_t (&px)[16]) {
	const int mode = __ffs((int)(blk.x & 0xFFu)) - 1;	// lowest set bit of byte 0; -1 -> invalid (Q7)
	switch (mode) {
	case 0: bc7_decode_mode<0>(blk, tab, px); break;
	case 1: bc7_decode_mode<1>(blk, tab, px); break;
	case 2: bc7_decode_mode<2>(blk, tab, px); break;
	case 3: bc7_decode_mode<3>(blk, tab, px); break;
	case 4: bc7_decode_mode<4>(blk, tab, px); break;
	case 5: bc7_decode_mode<5>(blk, tab, px); break;
	case 6: bc7_decode_mode<6>(blk, tab, px); break;
	case 7: bc7_decode_mode<7>(blk, tab, px); break;
	default: return false;
	}
	bool ok = (mode_mask >> mode) & 1u;					// :364
	ok = ok && !(mode >= 4 && (flags & FLAG_OPAQUE_ONLY));			// :366-369
	ok = ok && !(mode < 4 && (flags & FLAG_NON_OPAQUE_ONLY));
	return ok;
}

}  // namespace detexb200
