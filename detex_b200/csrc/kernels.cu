// Texture-level decode kernels for sm_100a and their host-side launcher.
//
// Replaces the two loops of reference texture.c: detexDecompressTextureTiled
// (:77-98) and detexDecompressTextureLinear (:105-145).  The work is a pure map over
// independent 8/16-byte blocks and is bound by HBM traffic (in + out bytes per block,
// DESIGN.md) or, for the heavy decoders, by instruction issue, so the kernels are organised
// around the memory system and around keeping warps convergent:
//   * one block per lane, fetched with one 64/128-bit non-coherent streaming load,
//     32 lanes reading 256/512 contiguous bytes;
//   * all 16 pixels of the block are produced in registers;
//   * LINEAR layout: lane L of a warp owns block column bx0+L, so for every pixel row
//     the warp's 32 row segments form one contiguous span of the image row and go out
//     as full-width vector stores;
//   * TILED layout: a block's pixels are contiguous (16*ps bytes); every lane writes its own
//     tile with 256-bit stores (whole 32-byte sectors; 16-byte tiles as one 128-bit store of a
//     warp-contiguous span).  The older shared-memory transposition is kept for A/B runs;
//   * BC7 and BC6H, whose modes are different code: a CTA counting-sorts tiles of blocks by
//     mode in shared memory so that a warp decodes 32 blocks of one mode (decode_sorted_kernel);
//     outputs whose rows or tiles would be partial sectors when written by lanes in sorted
//     order are staged in shared memory and copied out in image order;
//   * conversions to the pixel format the caller asked for run in the store epilogue
//     ("post op" / "post class" below), including BC6H's half -> 16-bit -> 8-bit chain (tone.cuh);
//   * invalid blocks become zero pixels and raise one flag word (Q16).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <cuda.h>		// CUtensorMap (the encoder is fetched through cudaGetDriverEntryPoint: no -lcuda)

#include "common.cuh"
#include "decode_s3tc.cuh"
#include "decode_etc.cuh"
#include "decode_bptc.cuh"
#ifdef DETEXB200_EXPERIMENTS
#include "decode_bptc_palette.cuh"
#endif
#include "decode_bptc_float.cuh"
#include "tone.cuh"
#include "launch.h"

namespace detexb200 {

// ---- per-format traits ---------------------------------------------------------------
template <int F> struct FormatTraits;
#define DETEX_TRAITS(F, BW, OW, TAB) \
	template <> struct FormatTraits<F> { \
		static constexpr int BLOCK_WORDS = BW;	/* compressed block size in 32-bit words */ \
		static constexpr int OUT_WORDS = OW;	/* decoded tile size in 32-bit words */ \
		static constexpr bool USES_TABLES = TAB; \
	};
DETEX_TRAITS(F_BC1, 2, 16, false)		DETEX_TRAITS(F_BC1A, 2, 16, false)
DETEX_TRAITS(F_BC2, 4, 16, false)		DETEX_TRAITS(F_BC3, 4, 16, false)
DETEX_TRAITS(F_RGTC1, 2, 4, false)		DETEX_TRAITS(F_SIGNED_RGTC1, 2, 8, false)
DETEX_TRAITS(F_RGTC2, 4, 8, false)		DETEX_TRAITS(F_SIGNED_RGTC2, 4, 16, false)
DETEX_TRAITS(F_BPTC_FLOAT, 4, 32, true)		DETEX_TRAITS(F_BPTC_SIGNED_FLOAT, 4, 32, true)
DETEX_TRAITS(F_BPTC, 4, 16, true)		DETEX_TRAITS(F_ETC1, 2, 16, false)
DETEX_TRAITS(F_ETC2, 2, 16, false)		DETEX_TRAITS(F_ETC2_PUNCHTHROUGH, 2, 16, false)
DETEX_TRAITS(F_ETC2_EAC, 4, 16, false)		DETEX_TRAITS(F_EAC_R11, 2, 8, false)
DETEX_TRAITS(F_EAC_SIGNED_R11, 2, 8, false)	DETEX_TRAITS(F_EAC_RG11, 4, 16, false)
DETEX_TRAITS(F_EAC_SIGNED_RG11, 4, 16, false)
#undef DETEX_TRAITS

template <int BW> struct BlockReg;
template <> struct BlockReg<2> {
	using type = uint2;
	static __device__ __forceinline__ uint2 load(const void *p) { return ld_stream_v2(p); }
};
template <> struct BlockReg<4> {
	using type = uint4;
	static __device__ __forceinline__ uint4 load(const void *p) { return ld_stream_v4(p); }
};

// One block: registers in, registers out.  Same role as the reference's function
// table decompress_function[] (texture.c:27-48), resolved at compile time.
template <int F, int PATHS = ETC_PATHS_ALL>
__device__ __forceinline__ bool decode_block(const typename BlockReg<FormatTraits<F>::BLOCK_WORDS>::type &blk,
uint32_t mode_mask, uint32_t flags, const BptcTables *tab, uint32_t (&px)[FormatTraits<F>::OUT_WORDS]) {
	if constexpr (F == F_BC1) return decode_bc1(blk, mode_mask, flags, px);
	else if constexpr (F == F_BC1A) return decode_bc1a(blk, mode_mask, flags, px);
	else if constexpr (F == F_BC2) return decode_bc2(blk, mode_mask, flags, px);
	else if constexpr (F == F_BC3) return decode_bc3(blk, mode_mask, flags, px);
	else if constexpr (F == F_RGTC1) return decode_rgtc1(blk, mode_mask, flags, px);
	else if constexpr (F == F_SIGNED_RGTC1) return decode_signed_rgtc1(blk, mode_mask, flags, px);
	else if constexpr (F == F_RGTC2) return decode_rgtc2(blk, mode_mask, flags, px);
	else if constexpr (F == F_SIGNED_RGTC2) return decode_signed_rgtc2(blk, mode_mask, flags, px);
	else if constexpr (F == F_BPTC_FLOAT) return decode_bptc_float<false>(blk, mode_mask, flags, *tab, px);
	else if constexpr (F == F_BPTC_SIGNED_FLOAT) return decode_bptc_float<true>(blk, mode_mask, flags, *tab, px);
	else if constexpr (F == F_BPTC) return decode_bptc(blk, mode_mask, flags, *tab, px);
	else if constexpr (F == F_ETC1) return decode_etc1(blk, mode_mask, flags, px);
	else if constexpr (F == F_ETC2) return decode_etc2<PATHS>(blk, mode_mask, flags, px);
	else if constexpr (F == F_ETC2_PUNCHTHROUGH) return decode_etc2_punchthrough<PATHS>(blk, mode_mask, flags, px);
	else if constexpr (F == F_ETC2_EAC) return decode_etc2_eac<PATHS>(blk, mode_mask, flags, px);
	else if constexpr (F == F_EAC_R11) return decode_eac_r11<false>(blk, mode_mask, flags, px);
	else if constexpr (F == F_EAC_SIGNED_R11) return decode_eac_r11<true>(blk, mode_mask, flags, px);
	else if constexpr (F == F_EAC_RG11) return decode_eac_rg11<false>(blk, mode_mask, flags, px);
	else return decode_eac_rg11<true>(blk, mode_mask, flags, px);
}

// Output-format conversion fused into the store epilogue.
//   post_op (run-time, uniform over the launch) edits the decoded pixels in place:
//     POST_SWAP_RB  R <-> B of 32-bit RGBA8/RGBX8 pixels (-> BGRA8/BGRX8, reference convert.c:37-52)
//                   or of 64-bit FLOAT_RGBX16 pixels (-> FLOAT_BGRX16, convert.c:54-70);
//     POST_FLIP16   signed <-> unsigned 16-bit components (+-32768, convert.c:139-186);
//     POST_FLIP8    signed <-> unsigned 8-bit components (+-128, convert.c:89-136).
//   The post class (compile-time, it changes the size of the stored tile) repacks them:
//     PC_PACK3      drop the fourth component: RGBX8/RGBA8 -> RGB8 (convert.c:671-690),
//                   FLOAT_RGBX16 -> FLOAT_RGB16 (convert.c:700-716);
//     PC_NARROW     16-bit -> 8-bit components (v + 127) * 255 / 65535: R16 -> R8, RG16 -> RG8
//                   (convert.c:263-290);
//     tone classes  BC6H only (tone.cuh): half -> 16-bit integer -> 8 bit.  Class =
//                   PC_TONE_FIRST + 3 * mode + kind, mode = TONE_PLAIN / TONE_RANGE / TONE_GAMMA (how the
//                   half becomes an integer), kind = TONE_RGBX8 (FLOAT_RGBX16 -> RGBX8 / RGBA8; POST_SWAP_RB
//                   then acts on the 8-bit pixels -> BGRX8 / BGRA8), TONE_RGB8 (the same, fourth component
//                   dropped) or TONE_RGBX16 (the 16-bit stage alone, same tile size as native).
constexpr uint32_t POST_NONE = 0, POST_SWAP_RB = 1, POST_FLIP16 = 2, POST_FLIP8 = 4;
constexpr int PC_NATIVE = 0, PC_PACK3 = 1, PC_NARROW = 2, PC_TONE_FIRST = 3, PC_TONE_LAST = 11;
constexpr int TONE_RGBX8 = 0, TONE_RGB8 = 1, TONE_RGBX16 = 2;
constexpr bool is_tone_any(int pc) { return pc >= PC_TONE_FIRST && pc <= PC_TONE_LAST; }
constexpr int tone_kind(int pc) { return (pc - PC_TONE_FIRST) % 3; }
constexpr int tone_mode(int pc) { return (pc - PC_TONE_FIRST) / 3; }
constexpr bool is_tone_class(int pc) { return is_tone_any(pc) && tone_kind(pc) != TONE_RGBX16; }	// the 8-bit ones
constexpr bool is_tone16_class(int pc) { return is_tone_any(pc) && tone_kind(pc) == TONE_RGBX16; }
constexpr bool tone_packs3(int pc) { return is_tone_any(pc) && tone_kind(pc) == TONE_RGB8; }

template <int OW>
__device__ __forceinline__ void apply_post_op(uint32_t (&px)[OW], uint32_t post_op) {
	if (post_op & POST_SWAP_RB) {
		if constexpr (OW == 16) {
#pragma unroll
			for (int k = 0; k < 16; k++)
				px[k] = prmt(px[k], 0u, 0x3012);
		}
		else if constexpr (OW == 32) {
#pragma unroll
			for (int k = 0; k < 16; k++) {
				const uint32_t lo = px[2 * k], hi = px[2 * k + 1];		// R|G<<16, B|X<<16
				px[2 * k] = prmt(lo, hi, 0x3254);				// B|G<<16
				px[2 * k + 1] = prmt(lo, hi, 0x7610);				// R|X<<16
			}
		}
	}
	if constexpr (OW <= 16) {		// one- and two-component formats (the API sets these bits only for them)
		if (post_op & (POST_FLIP16 | POST_FLIP8)) {
			const uint32_t top = (post_op & POST_FLIP16) ? 0x80008000u : 0x80808080u;
#pragma unroll
			for (int k = 0; k < OW; k++)
				px[k] ^= top;
		}
	}
}

// words per decoded tile after the post class
template <int OW, int PC> constexpr int post_words() {
	return PC == PC_PACK3 ? OW * 3 / 4 : (tone_packs3(PC) ? OW * 3 / 8 : (PC == PC_NARROW || is_tone_class(PC) ? OW / 2 : OW));
}
// which (format, class) pairs exist: PACK3 for the four-component 8-bit formats and BC6H unsigned,
// NARROW for the 16-bit one- and two-component formats
template <int F, int PC> constexpr bool post_class_valid() {
	constexpr int OW = FormatTraits<F>::OUT_WORDS;
	if (PC == PC_NATIVE)
		return true;
	if (PC == PC_PACK3)
		return (OW == 16 && F != F_EAC_RG11 && F != F_EAC_SIGNED_RG11 && F != F_SIGNED_RGTC2) || F == F_BPTC_FLOAT;
	if (is_tone_any(PC))
		return F == F_BPTC_FLOAT;
	return F == F_EAC_R11 || F == F_EAC_SIGNED_R11 || F == F_EAC_RG11 || F == F_EAC_SIGNED_RG11 ||
		F == F_SIGNED_RGTC1 || F == F_SIGNED_RGTC2;
}

// (v + 127) * 255 / 65535 == (v + 127) / 257 == ((v + 127) * 65281) >> 24 for every 16-bit v
// (65281 * 257 = 2^24 + 1; checked exhaustively, the product stays below 2^32)
__device__ __forceinline__ uint32_t narrow16(uint32_t v) { return ((v + 127u) * 65281u) >> 24; }
// two 16-bit components of one word -> two bytes
__device__ __forceinline__ uint32_t narrow_pair(uint32_t w) { return narrow16(w & 0xFFFFu) | (narrow16(w >> 16) << 8); }

template <int PC, int OW>
__device__ __forceinline__ void repack(const uint32_t (&px)[OW], uint32_t (&q)[post_words<OW, PC>()]) {
	if constexpr (PC == PC_PACK3 && OW == 16) {
#pragma unroll
		for (int r = 0; r < 4; r++) {			// R0G0B0R1 | G1B1R2G2 | B2R3G3B3
			q[3 * r] = prmt(px[4 * r], px[4 * r + 1], 0x4210);
			q[3 * r + 1] = prmt(px[4 * r + 1], px[4 * r + 2], 0x5421);
			q[3 * r + 2] = prmt(px[4 * r + 2], px[4 * r + 3], 0x6542);
		}
	}
	else if constexpr (PC == PC_PACK3 && OW == 32) {
#pragma unroll
		for (int h = 0; h < 8; h++) {			// pixel pairs: R0G0 | B0R1 | G1B1
			const uint32_t lo0 = px[4 * h], hi0 = px[4 * h + 1], lo1 = px[4 * h + 2], hi1 = px[4 * h + 3];
			q[3 * h] = lo0;
			q[3 * h + 1] = prmt(hi0, lo1, 0x5410);
			q[3 * h + 2] = prmt(lo1, hi1, 0x5432);
		}
	}
	else if constexpr (PC == PC_NARROW && OW == 8) {	// R16: two pixels per word -> four per word
#pragma unroll
		for (int r = 0; r < 4; r++)
			q[r] = narrow_pair(px[2 * r]) | (narrow_pair(px[2 * r + 1]) << 16);
	}
	else {							// RG16: one pixel per word -> two per word
		static_assert(PC == PC_NARROW && OW == 16, "unsupported post class");
#pragma unroll
		for (int k = 0; k < 8; k++)
			q[k] = narrow_pair(px[2 * k]) | (narrow_pair(px[2 * k + 1]) << 16);
	}
}

// What the tone classes need besides the pixels: the (uniform) tone parameters and whether the block
// was valid -- an invalid block is zero in the target format or the conversion of zero pixels,
// ToneArgs::zero_target.  Empty for every other class.
template <int PC, bool TONE = is_tone_any(PC)> struct EmitCtxImpl { __device__ __forceinline__ EmitCtxImpl(const ToneArgs &) {} };
template <int PC> struct EmitCtxImpl<PC, true> {
	Tone tone;
	bool zero_target;
	__device__ __forceinline__ EmitCtxImpl(const ToneArgs &a) : tone(tone_setup<tone_mode(PC)>(a)), zero_target(a.zero_target != 0u) {}
};
template <int PC> using EmitCtx = EmitCtxImpl<PC>;

// Tone classes: the decoder's pixel sink has already produced 16 RGBX8 pixels in px[0..15]
// (decode_px below).  What is left: the invalid rule and the R <-> B swap.
template <int PC>
__device__ __forceinline__ void tone_tile(const uint32_t (&px)[32], uint32_t (&t)[16], uint32_t post_op, const EmitCtx<PC> &ctx, bool ok) {
	const uint32_t fill = ctx.zero_target ? 0u : tone8_of_zero<tone_mode(PC)>(ctx.tone);
#pragma unroll
	for (int k = 0; k < 16; k++)
		t[k] = ok ? px[k] : fill;
	apply_post_op<16>(t, post_op & POST_SWAP_RB);
}

// decode_block for the class: the tone classes decode BC6H through the 8-bit pixel sink (tone.cuh).
template <int F, int PC, int PATHS = ETC_PATHS_ALL>
__device__ __forceinline__ bool decode_px(const typename BlockReg<FormatTraits<F>::BLOCK_WORDS>::type &blk,
uint32_t mode_mask, uint32_t flags, const BptcTables *tab, const EmitCtx<PC> &ctx, uint32_t (&px)[FormatTraits<F>::OUT_WORDS]) {
	if constexpr (is_tone_class(PC))
		return decode_bptc_float_tone8<tone_mode(PC)>(blk, mode_mask, *tab, ctx.tone, px);
	else if constexpr (is_tone16_class(PC))
		return decode_bptc_float_tone16<tone_mode(PC)>(blk, mode_mask, *tab, ctx.tone, px);
	else
		return decode_block<F, PATHS>(blk, mode_mask, flags, tab, px);
}

// The second half of the split BC6H decoder (bc6h_prepare -> record -> here) for the class.
template <int F, int PC>
__device__ __forceinline__ void pixels_px(const Bc6hPrepared &p, const EmitCtx<PC> &ctx, uint32_t (&px)[32]) {
	if constexpr (is_tone_class(PC))
		bc6h_pixels<false>(p, Bc6hTone8Sink<32, tone_mode(PC)>{ px, ctx.tone });
	else if constexpr (is_tone16_class(PC))
		bc6h_pixels<false>(p, Bc6hTone16Sink<tone_mode(PC)>{ px, ctx.tone, tone16_of_zero<tone_mode(PC)>(ctx.tone) });
	else
		bc6h_pixels<F == F_BPTC_SIGNED_FLOAT>(p, Bc6hHalfSink<F == F_BPTC_SIGNED_FLOAT>{ px });
}

// post_op, then hand the words that go to memory to store(words): the decoded array itself for
// the native class, a repacked one otherwise.  ok: the block was valid (its pixels are already
// filled with invalid_fill() otherwise).
template <int PC, int OW, typename Store>
__device__ __forceinline__ void emit(uint32_t (&px)[OW], uint32_t post_op, const EmitCtx<PC> &ctx, bool ok, Store &&store) {
	if constexpr (is_tone_class(PC)) {
		uint32_t t[16];
		tone_tile<PC>(px, t, post_op, ctx, ok);
		if constexpr (!tone_packs3(PC))
			store(t);
		else {
			uint32_t q[12];
			repack<PC_PACK3, 16>(t, q);
			store(q);
		}
	}
	else if constexpr (is_tone16_class(PC)) {
		// the pixel sink wrote RGBX16 pixels; an invalid block is zero or the conversion of zero pixels
		if (!ok) {
			const uint32_t z = ctx.zero_target ? 0u : tone16_of_zero<tone_mode(PC)>(ctx.tone);
#pragma unroll
			for (int k = 0; k < OW; k++)
				px[k] = z | (z << 16);
		}
		store(px);
	}
	else {
		apply_post_op<OW>(px, post_op);
		if constexpr (PC == PC_NATIVE)
			store(px);
		else {
			uint32_t q[post_words<OW, PC>()];
			repack<PC, OW>(px, q);
			store(q);
		}
	}
}

// Pixels of an invalid block are zero IN THE TARGET FORMAT (texture.c:90-93,125-128).  Swap,
// drop-X and narrowing keep zero at zero; a sign flip does not, so the fill value is the flip's
// pre-image of zero.
__device__ __forceinline__ uint32_t invalid_fill(uint32_t post_op) {
	return (post_op & POST_FLIP16) ? 0x80008000u : ((post_op & POST_FLIP8) ? 0x80808080u : 0u);
}
// Optional per-block record of invalid blocks (one bit per block), for the decode + conversion
// pass path: the conversion of a zero pixel is not always zero, so the tiles of invalid blocks are
// cleared again after the conversion (zero_invalid_kernel).
__device__ __forceinline__ void mark_invalid(uint32_t *invalid_map, uint32_t block) {
	if (invalid_map != nullptr)
		atomicOr(&invalid_map[block >> 5], 1u << (block & 31u));
}

constexpr int CTA_THREADS = 256;
constexpr int CTA_WARPS = CTA_THREADS / 32;

// Raise the texture-level "some block was invalid" flag: one read-mostly probe and at most one
// plain store per warp that met an invalid block (every writer stores the same value).
__device__ __forceinline__ void report_invalid(bool any_bad_in_thread, uint32_t *invalid_flag) {
	if (any_bad_in_thread && invalid_flag != nullptr) {
		// one probe per warp that saw an invalid block; no CTA barrier on the common path
		const unsigned peers = __activemask();
		if ((threadIdx.x & 31) == __ffs((int)peers) - 1 && *(volatile uint32_t *)invalid_flag == 0u)
			*(volatile uint32_t *)invalid_flag = 1u;
	}
}

// ---- rare-mode deferral ------------------------------------------------------------------------
// The ETC2 family has five modes.  Individual and differential (the ETC1 subset) are 91 % of uniformly random blocks and
// share one short code path; T, H (paint colours and a distance) and planar (a bilinear ramp) are rare and run
// different code.  With one block per lane almost every warp contains a rare block, and a decoder that can meet all
// five modes either diverges or -- as round 1's unified palette engine did -- computes both variants of everything
// for every block.  Instead a lane that meets a T / H / planar block only queues the block's index in shared memory;
// after a CTA barrier the queued blocks are decoded together, 32 per warp pass (ETC_PATHS_RARE), and the main pass
// carries the individual / differential code only (ETC_PATHS_SIMPLE).
#ifndef DETEXB200_DEFER_ITERS
#define DETEXB200_DEFER_ITERS 8
#endif
constexpr int DEFER_ITERS = DETEXB200_DEFER_ITERS;	// blocks per thread between two drains of the queue
// Which blocks are deferred depends on the layout.  TILED: T, H and planar (measured, round 2, 33.5 M blocks: ETC2 67.9 ->
// 74.7 G blocks/s, punch-through 62.9 -> 65.2, ETC2_EAC 53.0 -> 57.6).  LINEAR: planar only, as in round 1 -- there the
// deferred pass writes a block's four 16-byte row segments long after its row neighbours were written, the main pass
// leaves 9 % of its 512-byte warp rows with holes, and the drain (three barriers, one latency-bound block per thread)
// grows with the number of deferred blocks: ETC2 LINEAR 57.5 -> 39.9 with T / H deferred, DRAM reads doubled by the
// partial-sector fills (ncu).  -DDETEXB200_DEFER_PLANAR_ONLY / -DDETEXB200_DEFER_TH_LINEAR force one split for both.
template <int LAYOUT> constexpr bool defer_th() {
#ifdef DETEXB200_DEFER_PLANAR_ONLY
	return false;
#elif defined(DETEXB200_DEFER_TH_LINEAR)
	return true;
#else
	return LAYOUT == LAYOUT_TILED;
#endif
}
template <int LAYOUT> constexpr int etc_main_paths() { return defer_th<LAYOUT>() ? ETC_PATHS_SIMPLE : ETC_PATHS_NO_PLANAR; }
template <int LAYOUT> constexpr int etc_deferred_paths() { return defer_th<LAYOUT>() ? ETC_PATHS_RARE : ETC_PATHS_PLANAR_ONLY; }
template <int F> struct DeferTraits { static constexpr bool ENABLED = false; };
struct DeferEtc2 {
	static constexpr bool ENABLED = true;
	// word holding bytes 0..3 of the colour block; ETC2: the rare modes need the differential bit set
	template <int LAYOUT>
	static __device__ __forceinline__ bool rare(uint32_t w0, bool needs_diff_bit) {
		const uint32_t mode = etc2_mode(w0);
		const bool hit = defer_th<LAYOUT>() ? mode != ETC_DIFFERENTIAL : mode == ETC_PLANAR;
		return (!needs_diff_bit || ((w0 >> 24) & 2u)) && hit;
	}
};
template <> struct DeferTraits<F_ETC2> : DeferEtc2 {
	template <int LAYOUT> static __device__ __forceinline__ bool is_rare(const uint2 &b) { return rare<LAYOUT>(b.x, true); }
};
template <> struct DeferTraits<F_ETC2_PUNCHTHROUGH> : DeferEtc2 {
	template <int LAYOUT> static __device__ __forceinline__ bool is_rare(const uint2 &b) { return rare<LAYOUT>(b.x, false); }
};
template <> struct DeferTraits<F_ETC2_EAC> : DeferEtc2 {
	template <int LAYOUT> static __device__ __forceinline__ bool is_rare(const uint4 &b) { return rare<LAYOUT>(b.z, true); }
};

// Measured tuning of the LINEAR kernel (profiles/README.md, "LINEAR kernel tuning"): block rows
// per thread, and whether the warps of a CTA leave together.  Two rows per thread (both loads
// issued up front) gain 5-8 % on every format that is not already at the HBM limit; BC1/BC1A sit
// at the limit and run best with one row per thread and a closing CTA barrier.
template <int F> struct LinearTune {
	static constexpr int ITERS = DeferTraits<F>::ENABLED ? DEFER_ITERS : 2;
	static constexpr bool EXIT_BARRIER = false;
};
// Resident CTAs per SM the plain kernels are compiled for.  The light decoders are latency-bound on
// HBM and want all 64 warps of an SM (8 CTAs -> at most 32 registers); the heavy ones keep their
// registers.  (BC1 LINEAR lost 4 % when its kernel grew from 30 to 37 registers and 8 -> 6 CTAs.)
#ifndef DETEXB200_LIGHT_MIN_CTAS
#define DETEXB200_LIGHT_MIN_CTAS 8
#endif
template <int F> constexpr int min_ctas() {
	// 0: no constraint.  BC3, SIGNED_RGTC1 and EAC_SIGNED_R11 measured 2-4 % faster unconstrained.
	return (F == F_BPTC || F == F_BPTC_FLOAT || F == F_BPTC_SIGNED_FLOAT || F == F_ETC2 ||
		F == F_ETC2_PUNCHTHROUGH || F == F_ETC2_EAC || F == F_BC3 || F == F_SIGNED_RGTC1 ||
		F == F_EAC_SIGNED_R11) ? 0 : DETEXB200_LIGHT_MIN_CTAS;
}
struct LinearTuneAtLimit { static constexpr int ITERS = 1; static constexpr bool EXIT_BARRIER = true; };
template <> struct LinearTune<F_BC1> : LinearTuneAtLimit {};
template <> struct LinearTune<F_BC1A> : LinearTuneAtLimit {};
template <> struct LinearTune<F_BC2> { static constexpr int ITERS = 1; static constexpr bool EXIT_BARRIER = false; };
// TILED kernel: tiles per warp pass.  Two gain 3-9 % on the formats that are not at the HBM limit.
template <int F> struct TiledTune { static constexpr int ITERS = DeferTraits<F>::ENABLED ? 1 : 2; };
template <> struct TiledTune<F_BC1> { static constexpr int ITERS = 1; };
template <> struct TiledTune<F_BC1A> { static constexpr int ITERS = 1; };
template <> struct TiledTune<F_BC2> { static constexpr int ITERS = 1; };

// pixel rows of one block into a LINEAR image (fast path: rows are whole vector stores)
template <int OW>
__device__ __forceinline__ void store_linear_rows(uint8_t *dst, size_t pitch, int bx, int by, int height, const uint32_t (&px)[OW]) {
	constexpr int ROW_WORDS = OW / 4, ROW_BYTES = OW;
	uint8_t *out = dst + (size_t)by * 4 * pitch + (size_t)bx * ROW_BYTES;
	const int rows = min(4, height - by * 4);		// clip the bottom edge (texture.c:117-120)
	// 32-byte rows go out as one 256-bit store when image base and pitch keep them 32-byte aligned
	const bool wide_ok = ROW_WORDS == 8 && ((((uintptr_t)dst) | pitch) & 31u) == 0;
#pragma unroll
	for (int r = 0; r < 4; r++) {
		if (r < rows) {
			uint8_t *p = out + r * pitch;
			if constexpr (ROW_WORDS == 1)
				st_stream_u32(p, px[r]);
			else if constexpr (ROW_WORDS == 2)
				st_stream_v2(p, px[2 * r], px[2 * r + 1]);
			else if constexpr (ROW_WORDS == 3) {		// 12-byte RGB8 rows
				st_stream_u32(p, px[3 * r]);
				st_stream_u32(p + 4, px[3 * r + 1]);
				st_stream_u32(p + 8, px[3 * r + 2]);
			}
			else if constexpr (ROW_WORDS == 4)
				st_stream_v4(p, px[4 * r], px[4 * r + 1], px[4 * r + 2], px[4 * r + 3]);
			else if constexpr (ROW_WORDS == 6) {		// 24-byte FLOAT_RGB16 rows: 16 + 8 or 8 + 16 bytes
				if ((((uintptr_t)p) & 15u) == 0) {
					st_stream_v4(p, px[6 * r], px[6 * r + 1], px[6 * r + 2], px[6 * r + 3]);
					st_stream_v2(p + 16, px[6 * r + 4], px[6 * r + 5]);
				}
				else {
					st_stream_v2(p, px[6 * r], px[6 * r + 1]);
					st_stream_v4(p + 8, px[6 * r + 2], px[6 * r + 3], px[6 * r + 4], px[6 * r + 5]);
				}
			}
			else if (wide_ok)
				st_stream_v8(p, px[8 * r], px[8 * r + 1], px[8 * r + 2], px[8 * r + 3],
					px[8 * r + 4], px[8 * r + 5], px[8 * r + 6], px[8 * r + 7]);
			else {
				st_stream_v4(p, px[8 * r], px[8 * r + 1], px[8 * r + 2], px[8 * r + 3]);
				st_stream_v4(p + 16, px[8 * r + 4], px[8 * r + 5], px[8 * r + 6], px[8 * r + 7]);
			}
		}
	}
}

// ---- TILED --------------------------------------------------------------------------------
// DIRECT: every lane writes its own tile with 256-bit stores (tiles of >= 32 bytes on a 32-byte
// aligned destination) instead of transposing the warp's tiles through shared memory.
template <int F, bool DIRECT, int PC>
__global__ void __launch_bounds__(CTA_THREADS, min_ctas<F>())
decode_tiled_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, uint32_t n_blocks,
uint32_t mode_mask, uint32_t flags, uint32_t post_op, uint32_t *invalid_flag, uint32_t *invalid_map, ToneArgs tone_args) {
	using T = FormatTraits<F>;
	using D = DeferTraits<F>;
	const EmitCtx<PC> ctx(tone_args);
	constexpr int BW = T::BLOCK_WORDS, OW = T::OUT_WORDS, OWO = post_words<OW, PC>(), CHUNKS = OWO / 4;
	static_assert(PC == PC_NATIVE || DIRECT, "repacked tiles are stored directly");
	__shared__ BptcTables tables;
	__shared__ uint4 stage[CHUNKS > 1 && !DIRECT ? CTA_WARPS * 32 * CHUNKS : 1];
	__shared__ uint32_t queue[D::ENABLED ? CTA_THREADS * DEFER_ITERS : 1];
	__shared__ int queue_n;
	if constexpr (T::USES_TABLES)
		bptc_tables_load(tables);
	if (threadIdx.x == 0)
		queue_n = 0;
	if constexpr (T::USES_TABLES || D::ENABLED)
		__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t n_tiles = (n_blocks + 31u) / 32u;		// a tile = 32 consecutive blocks
	bool bad = false;
	int it = 0;
	// the loop bound is uniform over the CTA (needed for the barriers of the deferral pass)
	// a warp takes TI consecutive tiles per pass, all of their loads issued before the first decode
	constexpr int TI = TiledTune<F>::ITERS;
	const uint32_t n_groups = (n_tiles + TI - 1) / TI;
	// deferral formats (TI == 1, a rolled pass loop): the block of the NEXT pass is requested before this pass is
	// decoded, so that its HBM latency hides behind the decode (see decode_linear_kernel)
	using BlockT = typename BlockReg<BW>::type;
	BlockT blk_next{};
	if constexpr (D::ENABLED) {
		const uint32_t i0 = ((uint32_t)blockIdx.x * CTA_WARPS + warp) * 32u + lane;
		if (blockIdx.x * CTA_WARPS < n_groups && (blockIdx.x * CTA_WARPS + warp) < n_tiles && i0 < n_blocks)
			blk_next = BlockReg<BW>::load(src + (size_t)i0 * (BW * 4));
	}
	for (uint32_t tile_base = blockIdx.x * CTA_WARPS; tile_base < n_groups; tile_base += gridDim.x * CTA_WARPS, it++) {
#pragma unroll
		for (int k = 0; k < TI; k++) {
			const uint32_t tile = (tile_base + warp) * TI + k;
			const uint32_t first = tile * 32u, i = first + lane;
			const bool active = tile < n_tiles && i < n_blocks;
			uint32_t px[OW];
			bool ok = true, rare = false;
			BlockT blk = blk_next;
			if constexpr (D::ENABLED) {
				static_assert(!D::ENABLED || TI == 1, "the prefetch assumes one tile per pass");
				const uint32_t tile_n = tile + gridDim.x * CTA_WARPS, i_n = tile_n * 32u + lane;
				if (tile_base + gridDim.x * CTA_WARPS < n_groups && tile_n < n_tiles && i_n < n_blocks)
					blk_next = BlockReg<BW>::load(src + (size_t)i_n * (BW * 4));
			}
			if (active) {
				if constexpr (!D::ENABLED)
					blk = BlockReg<BW>::load(src + (size_t)i * (BW * 4));
				if constexpr (D::ENABLED)
					rare = D::template is_rare<LAYOUT_TILED>(blk);
				if (!rare)
					ok = decode_px<F, PC, D::ENABLED ? etc_main_paths<LAYOUT_TILED>() : ETC_PATHS_ALL>(blk, mode_mask, flags, &tables, ctx, px);
			}
			if (!ok) {
				bad = true;
				mark_invalid(invalid_map, i);
#pragma unroll
				for (int w = 0; w < OW; w++)
					px[w] = invalid_fill(post_op);
			}
			const unsigned rare_mask = D::ENABLED ? __ballot_sync(0xFFFFFFFFu, rare) : 0u;
			emit<PC>(px, post_op, ctx, ok, [&](auto &w) {
				if (tile >= n_tiles)
					return;
				uint8_t *tile_out = dst + (size_t)first * (OWO * 4);
				if constexpr (CHUNKS == 1) {
					if (active && !rare)
						st_stream_v4(tile_out + lane * 16, w[0], w[1], w[2], w[3]);
				}
				else if constexpr (DIRECT) {
					if (active && !rare) {
						uint8_t *mine = tile_out + lane * (OWO * 4);
						if constexpr (CHUNKS % 2 == 0) {
#pragma unroll
							for (int c = 0; c < CHUNKS; c += 2)
								st_stream_v8(mine + 16 * c, w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3],
									w[4 * c + 4], w[4 * c + 5], w[4 * c + 6], w[4 * c + 7]);
						}
						else {
#pragma unroll
							for (int c = 0; c < CHUNKS; c++)
								st_stream_v4(mine + 16 * c, w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
						}
					}
				}
				else {
					// Transpose through shared memory: lane L parks chunk c of its tile at
					// L*CHUNKS + (c ^ f(L)), f(L) = (L / (8/CHUNKS)) % CHUNKS, which makes both the
					// 128-bit stores (fixed c, 8 consecutive lanes) and the 128-bit loads below
					// (8 consecutive logical chunks) hit eight distinct 16-byte bank groups.
					uint4 *buf = stage + warp * 32 * CHUNKS;
					constexpr int GROUP = 8 / CHUNKS;
					const int swz = (lane / GROUP) % CHUNKS;
#pragma unroll
					for (int c = 0; c < CHUNKS; c++)
						buf[lane * CHUNKS + (c ^ swz)] = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
					__syncwarp();
					const uint32_t valid_chunks = min(32u, n_blocks - first) * CHUNKS;
#pragma unroll
					for (int j = 0; j < CHUNKS; j++) {
						const uint32_t q = j * 32 + lane;		// logical 16-byte chunk of the warp's output
						const uint32_t b = q / CHUNKS, c = q % CHUNKS;
						const uint4 v = buf[b * CHUNKS + (c ^ ((b / GROUP) % CHUNKS))];
						if (q < valid_chunks && !((rare_mask >> b) & 1u))
							st_stream_v4(tile_out + (size_t)q * 16, v.x, v.y, v.z, v.w);
					}
					__syncwarp();
				}
			});
			if constexpr (D::ENABLED) {
				if (rare)
					queue[atomicAdd(&queue_n, 1)] = i;
			}
		}
		if constexpr (D::ENABLED) {
			// drain the queue every DEFER_ITERS iterations and after the last one
			const bool last = tile_base + gridDim.x * CTA_WARPS >= n_groups;
			if ((it + 1) % DEFER_ITERS != 0 && !last)
				continue;
			__syncthreads();
			const int n_rare = queue_n;
			for (int e = threadIdx.x; e < n_rare; e += CTA_THREADS) {
				const uint32_t j = queue[e];
				const auto blk = BlockReg<BW>::load(src + (size_t)j * (BW * 4));
				uint32_t qx[OW];
				if (!decode_block<F, etc_deferred_paths<LAYOUT_TILED>()>(blk, mode_mask, flags, &tables, qx)) {
					bad = true;
					mark_invalid(invalid_map, j);
#pragma unroll
					for (int k = 0; k < OW; k++)
						qx[k] = invalid_fill(post_op);
				}
				emit<PC>(qx, post_op, ctx, true, [&](auto &w) {	// (deferral exists for the ETC2 family only)
					uint8_t *out = dst + (size_t)j * (OWO * 4);
#pragma unroll
					for (int c = 0; c < CHUNKS; c++)
						st_stream_v4(out + 16 * c, w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
				});
			}
			__syncthreads();
			if (threadIdx.x == 0)
				queue_n = 0;
			__syncthreads();
		}
	}
	report_invalid(bad, invalid_flag);
}

// ---- LINEAR, fast path: width % 4 == 0 and 16-byte aligned output ---------------------------
// blockDim = (TX, TY): thread (x, y) decodes block (blockIdx.x*TX + x, by) with by
// grid-strided over block rows.
template <int F, int PC>
__global__ void __launch_bounds__(CTA_THREADS, min_ctas<F>())
decode_linear_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int width, int height,
int wib, int hib, uint32_t mode_mask, uint32_t flags, uint32_t post_op, uint32_t *invalid_flag, uint32_t *invalid_map,
ToneArgs tone_args) {
	using T = FormatTraits<F>;
	using D = DeferTraits<F>;
	const EmitCtx<PC> ctx(tone_args);
	constexpr int BW = T::BLOCK_WORDS, OW = T::OUT_WORDS, ROW_BYTES = post_words<OW, PC>();
	__shared__ BptcTables tables;
	__shared__ uint32_t queue[D::ENABLED ? CTA_THREADS * DEFER_ITERS : 1];
	__shared__ int queue_n;
	const int tid = threadIdx.y * blockDim.x + threadIdx.x;
	if constexpr (T::USES_TABLES) {
		for (int i = tid; i < 64; i += CTA_THREADS) {
			tables.two[i] = bptc_two_global[i];
			tables.three[i] = bptc_three_global[i];
		}
	}
	if (tid == 0)
		queue_n = 0;
	if constexpr (T::USES_TABLES || D::ENABLED)
		__syncthreads();
	const int bx = blockIdx.x * blockDim.x + threadIdx.x;
	const size_t pitch = (size_t)width * (ROW_BYTES / 4);		// bytes per image row
	bool bad = false;
	// CTA (bx tile, g) owns the ITERS strips of TY block rows starting at g*ITERS*TY; g grid-strides.
	constexpr int ITERS = LinearTune<F>::ITERS;
	const int rows_per_group = ITERS * blockDim.y;
	for (int g = blockIdx.y; g * rows_per_group < hib; g += gridDim.y) {
		// Formats with a deferral queue run DEFER_ITERS strips per barrier interval in a rolled loop (their decoders are
		// too large to unroll eight times); the next strip's block is requested before the current one is decoded, or
		// every iteration would start by waiting out the full HBM latency of its one load (ncu, round 2: one third of
		// all warp samples of the ETC2 kernels sat on the first use of the loaded block).
		using BlockT = typename BlockReg<BW>::type;
		BlockT blk_next{};
		if constexpr (D::ENABLED) {
			const int by0 = g * rows_per_group + threadIdx.y;
			if (bx < wib && by0 < hib)
				blk_next = BlockReg<BW>::load(src + ((size_t)by0 * wib + bx) * (BW * 4));
		}
#pragma unroll(D::ENABLED ? 1 : ITERS)
		for (int k = 0; k < ITERS; k++) {
			const int by = g * rows_per_group + k * blockDim.y + threadIdx.y;
			BlockT blk = blk_next;
			if constexpr (D::ENABLED) {
				const int by_n = by + blockDim.y;
				if (k + 1 < ITERS && bx < wib && by_n < hib)
					blk_next = BlockReg<BW>::load(src + ((size_t)by_n * wib + bx) * (BW * 4));
			}
			if (bx < wib && by < hib) {
				if constexpr (!D::ENABLED)
					blk = BlockReg<BW>::load(src + ((size_t)by * wib + bx) * (BW * 4));
				bool rare = false;
				if constexpr (D::ENABLED)
					rare = D::template is_rare<LAYOUT_LINEAR>(blk);
				if (rare)
					queue[atomicAdd(&queue_n, 1)] = (uint32_t)by * (uint32_t)wib + (uint32_t)bx;
				else {
					uint32_t px[OW];
					const bool ok = decode_px<F, PC, D::ENABLED ? etc_main_paths<LAYOUT_LINEAR>() : ETC_PATHS_ALL>(blk, mode_mask, flags, &tables, ctx, px);
					if (!ok) {
						bad = true;
						mark_invalid(invalid_map, (uint32_t)by * (uint32_t)wib + (uint32_t)bx);
#pragma unroll
						for (int j = 0; j < OW; j++)
							px[j] = invalid_fill(post_op);
					}
					emit<PC>(px, post_op, ctx, ok, [&](auto &w) { store_linear_rows(dst, pitch, bx, by, height, w); });
				}
			}
		}
		if constexpr (D::ENABLED) {
			// drain the queue of rare (planar) blocks collected over the last ITERS strips
			__syncthreads();
			const int n_rare = queue_n;
			for (int e = tid; e < n_rare; e += CTA_THREADS) {
				const uint32_t j = queue[e];
				const int qy = (int)(j / (uint32_t)wib), qx_ = (int)(j - (uint32_t)qy * (uint32_t)wib);
				const auto blk = BlockReg<BW>::load(src + (size_t)j * (BW * 4));
				uint32_t qp[OW];
				if (!decode_block<F, etc_deferred_paths<LAYOUT_LINEAR>()>(blk, mode_mask, flags, &tables, qp)) {
					bad = true;
					mark_invalid(invalid_map, j);
#pragma unroll
					for (int jj = 0; jj < OW; jj++)
						qp[jj] = invalid_fill(post_op);
				}
				emit<PC>(qp, post_op, ctx, true, [&](auto &w) { store_linear_rows(dst, pitch, qx_, qy, height, w); });
			}
			__syncthreads();
			if (tid == 0)
				queue_n = 0;
			__syncthreads();
		}
	}
	if constexpr (LinearTune<F>::EXIT_BARRIER)
		bad = __syncthreads_or(bad ? 1 : 0) != 0;	// (measured: the reducing barrier, not a plain one)
	report_invalid(bad, invalid_flag);
}

// ---- LINEAR, general path: any width / alignment, byte-granular clipped stores -------------
template <int F, int PC>
__global__ void __launch_bounds__(CTA_THREADS)
decode_linear_clipped_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int width, int height,
int wib, int hib, uint32_t mode_mask, uint32_t flags, uint32_t post_op, uint32_t *invalid_flag, uint32_t *invalid_map,
ToneArgs tone_args) {
	using T = FormatTraits<F>;
	const EmitCtx<PC> ctx(tone_args);
	constexpr int BW = T::BLOCK_WORDS, OW = T::OUT_WORDS, PS = post_words<OW, PC>() / 4;	// PS = bytes per stored pixel
	__shared__ BptcTables tables;
	if constexpr (T::USES_TABLES) {
		bptc_tables_load(tables);
		__syncthreads();
	}
	const uint32_t n_blocks = (uint32_t)wib * (uint32_t)hib;
	bool bad = false;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_blocks; i += gridDim.x * blockDim.x) {
		const int by = (int)(i / (uint32_t)wib), bx = (int)(i % (uint32_t)wib);
		const auto blk = BlockReg<BW>::load(src + (size_t)i * (BW * 4));
		uint32_t px[OW];
		const bool ok = decode_px<F, PC>(blk, mode_mask, flags, &tables, ctx, px);
		if (!ok) {
			bad = true;
			mark_invalid(invalid_map, i);
#pragma unroll
			for (int k = 0; k < OW; k++)
				px[k] = invalid_fill(post_op);
		}
		const int rows = min(4, height - by * 4), cols = min(4, width - bx * 4);
		emit<PC>(px, post_op, ctx, ok, [&](auto &w) {
#pragma unroll
			for (int r = 0; r < 4; r++)
#pragma unroll
				for (int x = 0; x < 4; x++)
					if (r < rows && x < cols) {
						uint8_t *p = dst + ((size_t)(by * 4 + r) * width + (size_t)(bx * 4 + x)) * PS;
#pragma unroll
						for (int b = 0; b < PS; b++) {
							const int byte_index = (r * 4 + x) * PS + b;
							p[b] = (uint8_t)(w[byte_index >> 2] >> (8 * (byte_index & 3)));
						}
					}
		});
	}
	report_invalid(bad, invalid_flag);
}

// ---- mode-sorted kernel for the multi-mode formats (BC7, BC6H) ---------------------------------
// These decoders are instruction-bound, not HBM-bound, and every mode runs different code:
// with one block per lane in index order a warp serialises over every mode present in its 32
// blocks (8 for BC7 distribution M).  Here a CTA takes a tile of 1024 consecutive blocks,
// counting-sorts them by mode through shared memory (warp-aggregated shared atomics, bins
// padded to whole warps), and each warp then decodes 32 blocks of ONE mode: the mode switch
// inside decode_block<F> becomes warp-uniform.  Results are stored straight to each block's own
// place (16-byte row segments; L2 merges the half sectors of horizontally adjacent blocks).
template <int F> struct SortTraits { static constexpr bool ENABLED = false; static constexpr int NUM_BINS = 1; };
// Bin of BC7 mode m = nibble m.  The sorted groups of a tile are dealt round-robin to the warps of the CTA, so with
// equally frequent modes (distribution M: block i has mode i % 8) the bin ORDER decides which modes meet in one warp:
// eight warps and a 512-block tile give warp w one group of bin w / 2 and one of bin w / 2 + 4, a 1024-block tile the
// even or the odd bins.  The modes cost between 121 and 173 us per 8.4 M blocks (profiles/r02_sweep_bc7_single_modes.log);
// in mode order the heaviest warp carries modes 0 + 4 = 308 against a mean of 278, in the order m0 m1 m3 m2 | m5 m6 m4 m7
// (the best of the 8! orders for both tile sizes) 294.  -DDETEXB200_BC7_BIN_NIBBLES=0x76543210u builds the mode order.
#ifndef DETEXB200_BC7_BIN_NIBBLES
#define DETEXB200_BC7_BIN_NIBBLES 0x75462310u
#endif
// the inverse: mode of a BC7 bin (bin 8 = invalid stays 8)
__host__ __device__ constexpr int bc7_mode_of_bin(int bin) {
	for (int m = 0; m < 8; m++)
		if ((int)((DETEXB200_BC7_BIN_NIBBLES >> (4 * m)) & 7u) == bin)
			return m;
	return 8;
}
static_assert(bc7_mode_of_bin(0) + bc7_mode_of_bin(1) + bc7_mode_of_bin(2) + bc7_mode_of_bin(3) + bc7_mode_of_bin(4) +
	bc7_mode_of_bin(5) + bc7_mode_of_bin(6) + bc7_mode_of_bin(7) == 28, "DETEXB200_BC7_BIN_NIBBLES must be a permutation of 0-7");
template <> struct SortTraits<F_BPTC> {
	static constexpr bool ENABLED = true;
	static constexpr int NUM_BINS = 9;		// modes 0-7 + invalid
	// every BC7 mode is separate code: bins are padded to whole warps so that no warp mixes modes
	static constexpr bool PAD_BINS = true;
	// 64-byte tiles: pixels are staged in shared memory and copied out coalesced
	static constexpr bool STAGE_OUTPUT = true;
	static constexpr int TILE_BLOCKS = 1024;
	static __device__ __forceinline__ int bin_of(const uint4 &b, uint32_t mode_mask, uint32_t flags) {
		const int m = __ffs((int)(b.x & 0xFFu)) - 1;
		const bool ok = m >= 0 && ((mode_mask >> m) & 1u) && !(m >= 4 && (flags & FLAG_OPAQUE_ONLY)) &&
			!(m < 4 && (flags & FLAG_NON_OPAQUE_ONLY));
		return ok ? (int)((DETEXB200_BC7_BIN_NIBBLES >> (4 * m)) & 7u) : 8;
	}
};
struct SortTraitsBc6h {
	static constexpr bool ENABLED = true;
	static constexpr int NUM_BINS = 15;		// detex modes 0-13 + invalid
	static constexpr bool PAD_BINS = true;
	// 128-byte tiles whose pixel rows are 32 bytes = one full sector: a lane's direct stores are
	// already sector-complete in both layouts, so no staging buffer is needed and the shared memory
	// goes into a 2048-block tile instead (15 bins padded to whole warps waste 43 % of the lanes
	// on a 512-block tile with distribution M, 9 % on 2048 blocks).
	static constexpr bool STAGE_OUTPUT = false;
	static constexpr int TILE_BLOCKS = 2048;
	static __device__ __forceinline__ int bin_of(const uint4 &b, uint32_t mode_mask, uint32_t) {
		const int m = bc6h_mode(b.x);
		return (m >= 0 && ((mode_mask >> m) & 1u)) ? m : 14;
	}
};
template <> struct SortTraits<F_BPTC_FLOAT> : SortTraitsBc6h {};
template <> struct SortTraits<F_BPTC_SIGNED_FLOAT> : SortTraitsBc6h {};

// A staged tile of the tone classes is 16 RGBX8 / BGRX8 words per block (the toned pixels are what is
// parked); every other class stages the decoder's native tile.  A kernel that stages needs 64 bytes of
// shared memory per block on top of the double-buffered input, which fits two CTAs per SM with tiles
// of 1024 blocks.
// BC6H -> FLOAT_RGB16 stages the packed 96-byte tile (24-byte rows are partial sectors whichever lane
// writes them) in tiles of 512 blocks, which keeps two CTAs per SM.
template <int F, int PC> constexpr bool stages_packed_half3() { return F == F_BPTC_FLOAT && PC == PC_PACK3; }
template <int F, int PC> constexpr int stage_words() {
	return is_tone_class(PC) ? 16 : (stages_packed_half3<F, PC>() ? 24 : FormatTraits<F>::OUT_WORDS);
}
// Tile size.  Kernels that stage their output: 1024 blocks (512 for the packed half tiles), two CTAs per
// SM.  Kernels that store directly have the shared memory for more (2 x 48 KiB of double-buffered input
// per CTA at 3072 blocks): fewer barriers and fewer padded bins per block.  Measured, G blocks/s:
//   BC7 TILED native, 1024 -> 2048 -> 3072: distribution U 39.8 -> 45.4 -> 47.5 (8.4 M blocks; 38.2 -> 43.1 at
//     4.2 M), M 43.5 -> 44.9 -> 46.3: adopted;
//   BC7 TILED -> RGB8 (48-byte tiles = partial sectors whose halves come from different warps): 34.3 at
//     1024, 29.0 at 3072 -- the further apart in time the halves are written, the worse: stays at 1024;
//   BC6H, 2048 -> 3072: TILED +3-9 % and LINEAR -4 ... -13 % at 8.4 M blocks, +-2 % at 4.2 M: stays at 2048
//     (SortTraits::TILE_BLOCKS).
// Threads per CTA (1024 threads = 32 warps per SM in every configuration, 64 registers each).  BC7: four CTAs of 256
// threads -- more independent CTAs to run while one sits at a barrier or waits for its staged tile to drain; measured
// against two CTAs of 512 (round 2, 16.8 M blocks, M LINEAR / TILED 48.7 / 48.9 -> 51.7 / 51.4 G blocks/s, U LINEAR
// 38.5 -> 40.1; 4.2 M blocks: M 45.7 / 43.7 -> 48.7 / 48.0).  BC6H loses 6-13 % that way (its 15 bins want the larger
// tiles that only 512-thread CTAs have the shared memory for) and keeps two CTAs of 512.
// -DDETEXB200_SORT_1024 builds the round-1 experiment "ONE CTA of 1024 threads per SM, staged tiles of 2048 blocks":
// measured slower then (BC7 LINEAR M 41.8 -> 36.9), kept for A/B runs.
template <int F> constexpr int sort_threads() {
#ifdef DETEXB200_SORT_1024
	return 1024;
#elif defined(DETEXB200_SORT_512)
	return 512;
#else
	return F == F_BPTC ? 256 : 512;
#endif
}
template <int F, int PC, bool STAGE, int LAYOUT, bool SPLIT = false> constexpr int sort_tile() {
	constexpr int THREADS = sort_threads<F>();
	if (SPLIT)		// 48-byte records instead of staged pixels: two blocks per thread
		return 2 * THREADS;
	if (THREADS == 1024 && STAGE)
		return stages_packed_half3<F, PC>() ? 1024 : 2048;
	if (THREADS == 256)		// at most 56 KiB of shared memory per CTA
		return STAGE ? (stages_packed_half3<F, PC>() ? 256 : 512) : 1024;
	if (STAGE)
		return stages_packed_half3<F, PC>() ? 512 : 1024;
	if (F == F_BPTC)
		return PC == PC_NATIVE ? 3072 : 1024;
	return SortTraits<F>::TILE_BLOCKS;
}

constexpr uint32_t SLOT_EMPTY = 0xFFFFu;		// sorted slot that is bin padding

// Physical 16-byte slot of chunk c of the block with tile-local index id in the staging buffer:
// plane c, position id with its low three bits XOR-hashed by the next three bits of id and by
// a chunk-dependent constant.  Both access patterns are then bank-conflict free (eight lanes
// of a 128-bit access must hit eight distinct 16-byte bank groups):
//   * decode writes: fixed c, lanes hold arbitrary ids -- including the adversarial case of
//     distribution M where all ids of a group are congruent mod 8 (block i has mode i % 8);
//   * copy-out reads: consecutive logical chunks (TILED: one or two ids x all c; LINEAR:
//     consecutive ids / halves of one pixel row).
//
// TMA variant (LINEAR, 16-byte rows): the tile is staged as TB / 64 boxes of 64 consecutive blocks, each box 4 KiB =
// 4 pixel rows x 8 lines of 128 bytes (a line = one pixel row of eight blocks), and chunk (id & 7) of a line sits at
// (id & 7) ^ (line & 7) = (id ^ (id >> 3)) & 7 -- which is precisely the layout the hardware's SWIZZLE_128B mode gives a
// box of {128 bytes, 8 groups of eight blocks, 4 pixel rows}.  The decode writes stay conflict-free as above and
// one cp.async.bulk.tensor instruction per box replaces the whole copy-out.
template <int TB, int CHUNKS, bool TMA = false>
__device__ __forceinline__ uint32_t stage_slot(uint32_t id, uint32_t c) {
	if (TMA)
		return (id >> 6) * 256u + c * 64u + ((id & 63u) ^ ((id >> 3) & 7u));
	const uint32_t fc = CHUNKS == 8 ? (((c & 1u) << 2) | (c >> 1)) : (2u * c);
	return c * TB + (id ^ (((id >> 3) ^ fc) & 7u));
}

// ---- TMA store of a staged LINEAR tile ------------------------------------------------------------------------
// The image is described to the TMA unit as a rank-3 tensor of bytes {128, pitch / 128, height} (strides 128 and
// pitch): coordinate 1 counts groups of eight blocks along a pixel row.  With wib % 64 == 0 every run of 64
// consecutive blocks lies inside one block row, so box b of a tile that starts at block t0 is the box {128, 8, 4}
// at (0, bx / 8, 4 * by) with (by, bx) the block coordinates of block t0 + 64 b: always entirely inside the image
// (measured on B200, tools/debug/tma_test.cu: a box that sticks out at the high end of a dimension is clipped as
// documented -- the last block row of an image whose height is not a multiple of four relies on it -- but a
// negative start coordinate raises an illegal-instruction fault on this store path, so row wrap-around is not left
// to the hardware).
__device__ __forceinline__ void fence_proxy_async_smem() {
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, const void *smem_src, int c0, int c1, int c2) {
	const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_src);
	asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
		:: "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(s) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the shared-memory source of every committed store has been read (the buffer may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// mbarrier with one expected arrival per phase, in the CTA's shared memory
__device__ __forceinline__ void mbar_init1(void *bar) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void *bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, uint32_t parity) {
	asm volatile("{\n\t.reg .pred p;\n\tMBAR_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra MBAR_DONE;\n\t"
		"bra MBAR_WAIT;\n\tMBAR_DONE:\n\t}" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}

// whether a staged LINEAR kernel can hand its tile to the TMA unit: 16-byte rows staged as they are stored
template <int F, int PC, bool STAGE, int LAYOUT> constexpr bool tma_capable() {
	if (STAGE && LAYOUT == LAYOUT_LINEAR && stages_packed_half3<F, PC>())
		return true;		// 24-byte rows: boxes of {192, 8, 4} bytes, no swizzle (the inner extent exceeds 128 bytes)
	return STAGE && LAYOUT == LAYOUT_LINEAR && stage_words<F, PC>() == 16 && !stages_packed_half3<F, PC>() &&
		!((PC == PC_PACK3 && FormatTraits<F>::OUT_WORDS == 16) || tone_packs3(PC));
}
// bytes of one stored pixel row of one block in a TMA-stored stage
template <int F, int PC> constexpr int tma_row_bytes() { return stages_packed_half3<F, PC>() ? 24 : 16; }

template <int F, int PC, bool STAGE, int LAYOUT, bool SPLIT = false>
struct SortedSmem {
	static constexpr int TB = sort_tile<F, PC, STAGE, LAYOUT, SPLIT>(), NB = SortTraits<F>::NUM_BINS, CAP = TB + NB * 32;
	static constexpr int CHUNKS = stage_words<F, PC>() / 4;		// 16-byte chunks per staged tile
	uint4 out[SPLIT ? TB * 3 : (STAGE ? TB * CHUNKS : 1)];	// staged pixels (stage_slot layout); SPLIT: prepared blocks, three planes
	uint4 raw[2][TB];		// double-buffered compressed tile, filled by cp.async
	uint16_t id[2][CAP];		// tile-local block index in each sorted slot, double-buffered across tiles
	int count[2][16];		// per-mode histogram, double-buffered across tiles
	unsigned long long stage_bar;	// mbarrier: the TMA unit has read the staged tile (TMA variants)
	uint8_t bin_lut[256];		// byte 0 of a block -> bin under this launch's mode_mask / flags
	BptcTables tables;
};

__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gmem_src) {
	const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_and_wait_all() {
	asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
// slot of block id inside a raw tile buffer: low three bits hashed with the next three, so that
// a group whose ids are all congruent mod 8 (distribution M) still reads eight distinct banks
__device__ __forceinline__ uint32_t raw_slot(uint32_t id) { return id ^ ((id >> 3) & 7u); }

// Persistent CTAs (two per SM), each looping over tiles of TB consecutive blocks:
//   top     wait for this tile's cp.async data, barrier, start the NEXT tile's cp.async into the
//           other buffer (in flight during everything below)
//   sort    classify own blocks (mode = f(first word)), warp-aggregated histogram, barrier, every
//           warp derives the padded bin starts with a shuffle scan (kept in registers: lane b
//           holds start[b]), ids scattered to their sorted slots, barrier
//   decode  each warp takes groups of 32 same-mode blocks -> pixels parked in shared memory
//   store   barrier, cooperative copy-out, 512 contiguous bytes per warp store instruction
// PAD: bins are padded to a multiple of PAD sorted slots.  32 keeps every warp in one mode; 16 lets a
// warp hold two half-warps of different modes (BC6H modes share everything after the bit scatter, so a
// mixed warp costs about a quarter more) and halves the lanes lost to padding, which matters for the
// 1024-block tiles of the staged BC6H variants (14 modes: 31 % of the slots are padding at 32).
// TMA: the staged LINEAR tile leaves through cp.async.bulk.tensor stores issued by one thread (tmap describes the image,
// see tma_store_3d above) instead of the cooperative copy-out: no LDS / STG / address arithmetic per block, and one
// barrier less per tile because the store of tile t is issued at the top of tile t + 1 and runs under its sort.
// SPLIT (BC6H): the decode step only PREPARES the blocks of a group -- the mode-dependent half of the decoder, down to a
// 48-byte record that looks the same for every mode (decode_bptc_float.cuh) -- and parks the records in shared memory;
// after the second barrier every thread takes blocks in IMAGE order, runs the mode-independent half (interpolation,
// finish, output conversion: most of the work) and stores the pixels itself.  No lane idles in padded bins during that
// half, the lanes of a warp hold consecutive blocks of the image (a LINEAR row of 32 blocks is one contiguous run of
// every store instruction, whatever the row size), and what is staged is 48 bytes per block instead of 64-128.
template <int F, int LAYOUT, bool STAGE, int PC, int PAD = 32, bool TMA = false, bool SPLIT = false>
__global__ void __launch_bounds__(sort_threads<F>(), 1024 / sort_threads<F>())
decode_sorted_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, uint32_t n_blocks, int width,
int height, int wib, uint32_t mode_mask, uint32_t flags, uint32_t post_op, uint32_t *invalid_flag, uint32_t *invalid_map,
ToneArgs tone_args, uint32_t stagger_ns, const __grid_constant__ CUtensorMap tmap) {
	using T = FormatTraits<F>;
	using S = SortTraits<F>;
	using SM = SortedSmem<F, PC, STAGE, LAYOUT, SPLIT>;
	const EmitCtx<PC> ctx(tone_args);
	constexpr int SORT_THREADS = sort_threads<F>(), SORT_WARPS = SORT_THREADS / 32;
	constexpr int OW = T::OUT_WORDS, NB = S::NUM_BINS, TB = SM::TB, PER = TB / SORT_THREADS, CHUNKS = SM::CHUNKS;
	static_assert(!SPLIT || (!STAGE && !TMA && (F == F_BPTC_FLOAT || F == F_BPTC_SIGNED_FLOAT)), "SPLIT: BC6H, direct stores");
	constexpr int CPR = CHUNKS / 4;			// 16-byte chunks per pixel row of one block
	static_assert(T::BLOCK_WORDS == 4, "sorted kernel handles 16-byte blocks");
	static_assert(TB % SORT_THREADS == 0 && NB <= 16, "tile / bin limits");
	static_assert(!TMA || (tma_capable<F, PC, STAGE, LAYOUT>() && TB % 64 == 0 && TB / 8 <= 256), "TMA store variant");
	extern __shared__ __align__(16) unsigned char smem_raw[];
	// (the staged tile is the first member: the hardware swizzle wants it 1024-byte aligned in the shared window)
	SM &sm = *reinterpret_cast<SM *>(smem_raw + (TMA ? ((0u - (uint32_t)__cvta_generic_to_shared(smem_raw)) & 1023u) : 0u));
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t n_tiles = (n_blocks + TB - 1) / TB;
	if (tid < 32)
		(&sm.count[0][0])[tid] = 0;
	if (TMA && tid == 0)
		mbar_init1(&sm.stage_bar);
	bptc_tables_load(sm.tables);
	// the bin of a block depends on its first byte only (BC7: lowest set bit; BC6H: 2 / 5 mode bits)
	for (int v = tid; v < 256; v += SORT_THREADS) {
		uint4 probe;
		probe.x = (uint32_t)v; probe.y = probe.z = probe.w = 0u;
		sm.bin_lut[v] = (uint8_t)S::bin_of(probe, mode_mask, flags);
	}

	// The blocks a thread loads, classifies and scatters.  Kernels that stage their output: PER consecutive ones (thread t
	// owns blocks t * PER ...).  A warp then contributes RUNS of consecutive ids to a bin -- for distribution M, where
	// block i has mode i % 8, runs whose (id >> 3) & 7 are all different, which is what the XOR swizzle of the stage
	// needs to keep the decode's 128-bit shared stores free of bank conflicts (ncu, round 2, BC7 M LINEAR: 2.9 -> 1.0
	// extra wavefronts per store against the interleaved assignment id = k * THREADS + t, whose runs are four ids long
	// and collide pairwise; same speed).  Kernels that store directly keep the interleaved assignment: its cp.async
	// requests are coalesced, and with four blocks per thread the blocked one measured 5-6 % slower (BC6H, BC7 TILED).
	auto own_id = [&](int k) { return STAGE ? (uint32_t)(tid * PER + k) : (uint32_t)(k * SORT_THREADS + tid); };
	auto issue_tile_load = [&](uint32_t tile, int buf) {
#pragma unroll
		for (int k = 0; k < PER; k++) {
			const uint32_t id = own_id(k), i = tile * TB + id;
			if (i < n_blocks)
				cp_async_16(&sm.raw[buf][raw_slot(id)], src + (size_t)i * 16);
		}
	};
	// the staged tile that starts at block t0 -> image: one bulk tensor store per box of 64 blocks (thread 0)
	auto tma_store_tile = [&](uint32_t t0, uint32_t n_in_tile) {
		uint32_t by = t0 / (uint32_t)wib, bx = t0 - by * (uint32_t)wib;
		for (uint32_t b = 0; 64u * b < n_in_tile; b++) {
			tma_store_3d(&tmap, &sm.out[64u * CHUNKS * b], 0, (int)(bx >> 3), (int)(4u * by));	// a box = 64 blocks x CHUNKS x 16 bytes
			bx += 64u;
			if (bx >= (uint32_t)wib) {
				bx = 0u;
				by++;
			}
		}
		tma_store_commit();
	};
	// ---- sort of one tile, split in two so that the halves can sit on either side of a barrier ----------------------
	// classify: every thread looks at the blocks IT loaded (its own cp.async copies: no barrier needed, only its own
	// wait), takes their bin from the first byte, and the warp adds its counts to the tile's histogram.
	int bin[PER], off[PER];
	int inv_start = 0;		// first sorted slot of the invalid bin of the tile being decoded (set by scatter)
	auto classify = [&](uint32_t t, int buf) {
		const uint32_t n_in_tile = min((uint32_t)TB, n_blocks - t * TB);
#pragma unroll
		for (int k = 0; k < PER; k++) {
			const uint32_t id = own_id(k);
			bin[k] = -1;
			if (id < n_in_tile)
				bin[k] = sm.bin_lut[reinterpret_cast<const uint8_t *>(&sm.raw[buf][raw_slot(id)])[0]];
		}
#pragma unroll
		for (int k = 0; k < PER; k++) {
			const unsigned peers = __match_any_sync(0xFFFFFFFFu, bin[k]);
			const int leader = __ffs((int)peers) - 1;
			int base = 0;
			if (lane == leader && bin[k] >= 0)
				base = atomicAdd(&sm.count[buf][bin[k]], __popc(peers));
			base = __shfl_sync(0xFFFFFFFFu, base, leader);
			off[k] = base + __popc(peers & ((1u << lane) - 1u));
		}
	};
	// scatter (after a barrier: the histogram is complete): padded exclusive scan of the bin counts, redundantly in
	// every warp (lane b <- start[b]), then the ids of the classified blocks go to their sorted slots.  Returns the
	// padded total.
	auto scatter = [&](int buf) -> int {
		const int my_count = lane < NB ? sm.count[buf][lane] : 0;
		const int padded = S::PAD_BINS ? ((my_count + (PAD - 1)) & ~(PAD - 1)) : my_count;
		int start = padded;
#pragma unroll
		for (int d = 1; d < 16; d *= 2) {
			const int t = __shfl_up_sync(0xFFFFFFFFu, start, d);
			if (lane >= d)
				start += t;
		}
		start -= padded;		// lane NB holds the padded total
		// the padding slots at the end of every bin get the sentinel (one warp per bin); the occupied
		// slots are written below, so the array needs no clearing pass
		if (S::PAD_BINS) {
#pragma unroll
			for (int b = warp; b < NB; b += SORT_WARPS) {
				const int c = __shfl_sync(0xFFFFFFFFu, my_count, b), st = __shfl_sync(0xFFFFFFFFu, start, b);
				if (lane < ((-c) & (PAD - 1)))
					sm.id[buf][st + c + lane] = (uint16_t)SLOT_EMPTY;
			}
		}
#pragma unroll
		for (int k = 0; k < PER; k++) {
			const int st = __shfl_sync(0xFFFFFFFFu, start, bin[k] & 31);
			if (bin[k] >= 0)
				sm.id[buf][st + off[k]] = (uint16_t)own_id(k);
		}
		inv_start = __shfl_sync(0xFFFFFFFFu, start, NB - 1);
		return __shfl_sync(0xFFFFFFFFu, start, NB);
	};

	// ---- software pipeline over the tiles of this CTA (two barriers per tile) --------------------------------------
	//   B1 | decode(t): sorted groups -> pixels (stored, or parked in the stage)
	//      | classify(t+1): own blocks of the next tile (its cp.async data was requested one iteration ago) -> histogram
	//   B2 | thread 0: TMA store of the staged tile t      | cp.async of tile t+2 into the buffer decode(t) just left
	//      | scatter(t+1) -> sorted ids of the next tile   | (no TMA: cooperative copy-out of the stage)
	//      | thread 0: wait until the TMA unit has read the stage
	// The latency-bound steps of the sort (shared atomics, shuffles, dependent shared loads) of tile t+1 share their
	// barrier intervals with the arithmetic of tile t instead of having two intervals of their own.
	bool bad = false;
	uint32_t tile = blockIdx.x;
	if (tile >= n_tiles)
		return;			// (uniform over the CTA; the launcher never starts more CTAs than tiles)
	// the words of one finished block (block id of the tile that starts at block t0) -> image, from the thread that holds them
	auto store_direct = [&](uint32_t t0, uint32_t id, auto &w) {
		constexpr int OWO = post_words<OW, PC>(), CH = OWO / 4;
		if (LAYOUT == LAYOUT_TILED) {
			uint8_t *out = dst + (size_t)(t0 + id) * (OWO * 4);
			if (CH % 2 == 0 && (((uintptr_t)dst) & 31u) == 0) {
#pragma unroll
				for (int c = 0; c + 1 < CH; c += 2)
					st_stream_v8(out + 16 * c, w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3],
						w[4 * c + 4], w[4 * c + 5], w[4 * c + 6], w[4 * c + 7]);
			}
			else {
#pragma unroll
				for (int c = 0; c < CH; c++)
					st_stream_v4(out + 16 * c, w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
			}
		}
		else {
			const uint32_t i = t0 + id;
			const uint32_t by = i / (uint32_t)wib, bx = i - by * (uint32_t)wib;
			store_linear_rows(dst, (size_t)width * (OWO / 4), (int)bx, (int)by, height, w);
		}
	};
	// Start-up stagger.  All CTAs start together and every tile costs about the same, so without it the whole chip
	// alternates between a phase in which every CTA decodes (the memory system idles) and a phase in which every CTA
	// drains its tile (the SMs idle at the barrier behind the drain): ncu showed 35 % of all warp samples at that one
	// barrier.  CTA i therefore sleeps frac(i * golden ratio) of roughly one tile period before its first tile; the
	// phases stay spread because the tiles are uniform.
	if (stagger_ns != 0u) {
		const uint32_t frac = (blockIdx.x * 0x9E3779B9u) >> 16;		// 16-bit fraction, low discrepancy in blockIdx.x
		uint32_t ns = (frac * stagger_ns) >> 16;
		while (ns > 0u) {
			const uint32_t step = ns > 500000u ? 500000u : ns;
			__nanosleep(step);
			ns -= step;
		}
	}
	issue_tile_load(tile, 0);
	cp_async_commit_and_wait_all();
	__syncthreads();		// tables, bin table, zeroed histograms; (everyone's copies of the first tile)
	classify(tile, 0);
	if (tile + gridDim.x < n_tiles)
		issue_tile_load(tile + gridDim.x, 1);
	asm volatile("cp.async.commit_group;" ::: "memory");
	__syncthreads();		// histogram of the first tile complete
	int total = scatter(0);
	// TMA variants: the store of tile t is issued after B2 and the TMA unit needs about a microsecond to read the stage.
	// Nobody waits for that at a CTA barrier (ncu, round 2, BC7 M LINEAR, with thread 0 waiting before B1: 23 % of all
	// warp samples sat at B1): every warp first decodes a group of tile t + 1 into registers; thread 0 then collects the
	// bulk group -- normally complete by then -- and arrives on an mbarrier, which the others pass before their first
	// write to the stage.
	bool drain_pending = false;
	uint32_t drain_parity = 0u;
	auto await_drain = [&]() {
#ifdef DETEXB200_TMA_EARLY_WAIT
		(void)drain_parity;
#else
		if (tid == 0) {
			tma_store_wait_read();
			mbar_arrive(&sm.stage_bar);
		}
		mbar_wait(&sm.stage_bar, drain_parity);
		drain_parity ^= 1u;
#endif
		drain_pending = false;
	};
	for (int cur = 0;; cur ^= 1) {
		const uint32_t tile0 = tile * TB, n_here = min((uint32_t)TB, n_blocks - tile0);
		const uint32_t next = tile + gridDim.x;
		const bool has_next = next < n_tiles;
		__syncthreads();	// B1: id[cur] scattered, raw[cur] landed (classified one iteration ago), stage free
		// ---- decode: groups of 32 same-mode blocks; padding slots hold the 0xFFFF sentinel.
		// The next group's ids are fetched before the current group is decoded.  Groups are dealt round-robin to the
		// warps; drawing them from a shared ticket counter instead (heaviest bins first, the ticket drawn one group
		// ahead) was measured in both rounds and lost both times (round 2: BC7 M LINEAR 51.2 -> 48.8 G blocks/s, signed
		// BC6H 20.5 -> 18.7): with a ticket all warps of the CTA work through the same mode at the same time.
		// (Also measured and dropped in round 2: dealing odd rounds in reverse warp order so that no warp collects the most
		// expensive group of every round -- no gain for BC7, -3 % for BC6H M.)
		int g_next = warp * 32;
		// (with PAD < 32 the last group may end inside the warp: slots from total on are not initialised)
		uint32_t id_next = g_next + (PAD < 32 ? lane : 0) < total ? sm.id[cur][g_next + lane] : SLOT_EMPTY;
		while (g_next < total) {
			const uint32_t id = id_next;
			// BC7: the sort has already judged the block (bin_of: mode, mode_mask, flags) -- every block behind inv_start is
			// invalid and no other is -- so the decoder runs without the per-block mask tests (M LINEAR / TILED 53.0 / 51.5
			// -> 53.3 / 52.2 G blocks/s, U 39.0 / 49.1 -> 39.4 / 50.2).  BC6H keeps its one-shift test: measured the same
			// for the unsigned format and -3 % for signed TILED.  -DDETEXB200_SORT_RECHECK_MASK: the tests everywhere (A/B).
#ifdef DETEXB200_SORT_RECHECK_MASK
			constexpr bool SORT_DECIDES = false;
#else
			constexpr bool SORT_DECIDES = F == F_BPTC;
#endif
			const bool in_valid_bin = !SORT_DECIDES || g_next + lane < inv_start;
			const uint32_t dec_mask = SORT_DECIDES ? 0xFFFFFFFFu : mode_mask, dec_flags = SORT_DECIDES ? 0u : flags;
			g_next += SORT_WARPS * 32;
			id_next = g_next + (PAD < 32 ? lane : 0) < total ? sm.id[cur][g_next + lane] : SLOT_EMPTY;
			const bool live = id != SLOT_EMPTY;
			uint32_t px[SPLIT ? 1 : OW];
			bool ok = true;
			if (live) {
				const uint4 w = sm.raw[cur][raw_slot(id)];
				if constexpr (SPLIT) {
					// the mode-dependent half; the record waits in shared memory for the image-order half below
					Bc6hPrepared p;
					ok = bc6h_prepare<F == F_BPTC_SIGNED_FLOAT>(w, dec_mask, sm.tables, p) && in_valid_bin;
					if (!in_valid_bin)
						p.q0.w &= 0xFFFFu;	// (the record's valid flag)
					const uint32_t slot = raw_slot(id);
					sm.out[slot] = p.q0;
					sm.out[TB + slot] = p.q1;
					sm.out[2 * TB + slot] = p.q2;
				}
				else
					ok = decode_px<F, PC>(w, dec_mask, dec_flags, &sm.tables, ctx, px) && in_valid_bin;
				if (!ok) {
					bad = true;
					mark_invalid(invalid_map, tile0 + id);
					if constexpr (!SPLIT) {
#pragma unroll
						for (int k = 0; k < OW; k++)
							px[k] = invalid_fill(post_op);
					}
				}
			}
			if constexpr (TMA) {
				if (drain_pending)
					await_drain();		// (warp-uniform) the pixels are in registers; the stage is written next
			}
			if (live) {
				if constexpr (SPLIT) {
				}
				else if constexpr (STAGE && is_tone_class(PC)) {
					// the toned 8-bit pixels are what is staged; dropping X is left to the copy-out
					uint32_t t[16];
					tone_tile<PC>(px, t, post_op, ctx, ok);
#pragma unroll
					for (int c = 0; c < 4; c++)
						sm.out[stage_slot<TB, 4, TMA>(id, c)] = make_uint4(t[4 * c], t[4 * c + 1], t[4 * c + 2], t[4 * c + 3]);
				}
				else if constexpr (STAGE && stages_packed_half3<F, PC>()) {
					apply_post_op<OW>(px, post_op);
					uint32_t q[24];
					repack<PC_PACK3, 32>(px, q);
					if constexpr (TMA) {
						// box of 64 blocks = 4 pixel rows x 8 lines of 192 bytes (eight blocks x 24 bytes), unswizzled
						uint2 *line = reinterpret_cast<uint2 *>(sm.out) + (id >> 6) * 768u + ((id >> 3) & 7u) * 24u + (id & 7u) * 3u;
#pragma unroll
						for (int r = 0; r < 4; r++)
#pragma unroll
							for (int j = 0; j < 3; j++)
								line[r * 192 + j] = make_uint2(q[6 * r + 2 * j], q[6 * r + 2 * j + 1]);
					}
					else {
#pragma unroll
						for (int c = 0; c < 6; c++)
							sm.out[stage_slot<TB, 6>(id, c)] = make_uint4(q[4 * c], q[4 * c + 1], q[4 * c + 2], q[4 * c + 3]);
					}
				}
				else if constexpr (STAGE) {
					// staged in the native layout; a post class is applied by the copy-out below
					apply_post_op<OW>(px, post_op);
#pragma unroll
					for (int c = 0; c < CHUNKS; c++)
						sm.out[stage_slot<TB, CHUNKS, TMA>(id, c)] = make_uint4(px[4 * c], px[4 * c + 1], px[4 * c + 2], px[4 * c + 3]);
				}
				else {
					emit<PC>(px, post_op, ctx, ok, [&](auto &w) { store_direct(tile0, id, w); });
				}
			}
		}
		if constexpr (TMA) {
			if (drain_pending)
				await_drain();		// (a warp without a group in this tile)
		}
		if (has_next) {
			asm volatile("cp.async.wait_group 0;" ::: "memory");	// my copies of the next tile have landed
			classify(next, cur ^ 1);
		}
		if constexpr (TMA)
			fence_proxy_async_smem();	// my writes to the staged tile -> visible to the TMA unit
		__syncthreads();	// B2: tile decoded (stage complete); histogram of the next tile complete; raw[cur] free
		if constexpr (TMA) {
			if (tid == 0)
				tma_store_tile(tile0, n_here);
		}
		if (has_next && next + gridDim.x < n_tiles)
			issue_tile_load(next + gridDim.x, cur);
		asm volatile("cp.async.commit_group;" ::: "memory");
		if (tid < 16)
			sm.count[cur][tid] = 0;		// (consumed by scatter one iteration ago; next used after B1)
		if constexpr (STAGE && !TMA) {
			// ---- store
			constexpr bool PACK_ROWS = (PC == PC_PACK3 && OW == 16) || tone_packs3(PC);	// RGBX8 stage -> RGB8 rows
			static_assert(!STAGE || PC == PC_NATIVE || (is_tone_class(PC) && LAYOUT == LAYOUT_LINEAR) ||
				PACK_ROWS || (stages_packed_half3<F, PC>() && LAYOUT == LAYOUT_LINEAR),
				"staged copy-out repacks RGB8 rows only");
			if constexpr (stages_packed_half3<F, PC>() && LAYOUT == LAYOUT_LINEAR) {
				// 24-byte rows out of packed 96-byte tiles, in 8-byte pieces: piece q = id * 3 + j of pixel row r
				// lies at byte 24 r + 8 j of its tile and at row_base + 8 q of the image row, so a warp's 32
				// pieces are 256 contiguous bytes of the image
				const size_t pitch = (size_t)width * 6;
				const uint32_t by0 = tile0 / (uint32_t)wib, bx0 = tile0 - by0 * (uint32_t)wib;	// uniform
				uint8_t *const row0 = dst + (size_t)by0 * 4 * pitch;
				const uint8_t *const stage_bytes = reinterpret_cast<const uint8_t *>(sm.out);
	#pragma unroll
				for (int k = 0; k < TB * 3 / SORT_THREADS; k++) {
					const uint32_t q = k * SORT_THREADS + tid, id = q / 3u, j = q - id * 3u;
					if (id < n_here) {
						uint32_t bx = bx0 + id, by = by0;
						while (bx >= (uint32_t)wib) {	// tiles may wrap over block rows
							bx -= (uint32_t)wib;
							by++;
						}
						uint8_t *out = row0 + (size_t)(by - by0) * 4 * pitch + bx * 24u + j * 8u;
						const int rows = height - (int)by * 4;	// >= 4 except in a clipped last block row
	#pragma unroll
						for (int r = 0; r < 4; r++) {
							if (r < rows) {
								const uint32_t byte = 24u * r + 8u * j;
								const uint2 v = *reinterpret_cast<const uint2 *>(stage_bytes +
									(size_t)stage_slot<TB, 6>(id, byte >> 4) * 16 + (byte & 8u));
								st_stream_v2(out, v.x, v.y);
							}
							out += pitch;
						}
					}
				}
			}
			else if constexpr (LAYOUT == LAYOUT_TILED && PACK_ROWS) {
				// 48-byte RGB8 tiles out of staged RGBX8 tiles: thread <-> block in image order, three 128-bit
				// stores per tile; a warp completes 1536 contiguous bytes with three back-to-back instructions
				uint8_t *out = dst + (size_t)tile0 * 48;
	#pragma unroll
				for (int k = 0; k < TB / SORT_THREADS; k++) {
					const uint32_t id = k * SORT_THREADS + tid;
					if (id < n_here) {
						uint32_t t[16], q[12];
	#pragma unroll
						for (int r = 0; r < 4; r++) {
							const uint4 v = sm.out[stage_slot<TB, CHUNKS>(id, r)];
							t[4 * r] = v.x; t[4 * r + 1] = v.y; t[4 * r + 2] = v.z; t[4 * r + 3] = v.w;
						}
						repack<PC_PACK3, 16>(t, q);
						uint8_t *o = out + (size_t)id * 48;
	#pragma unroll
						for (int c = 0; c < 3; c++)
							st_stream_v4(o + 16 * c, q[4 * c], q[4 * c + 1], q[4 * c + 2], q[4 * c + 3]);
					}
				}
			}
			else if (LAYOUT == LAYOUT_TILED) {
				uint8_t *out = dst + (size_t)tile0 * (CHUNKS * 16);
				for (uint32_t q = tid; q < n_here * CHUNKS; q += SORT_THREADS) {
					const uint4 v = sm.out[stage_slot<TB, CHUNKS>(q / CHUNKS, q % CHUNKS)];
					st_stream_v4(out + (size_t)q * 16, v.x, v.y, v.z, v.w);
				}
			}
			else {
				constexpr int ROW_BYTES = post_words<OW, PC>();	// bytes of one stored pixel row of one block
				const size_t pitch = (size_t)width * (ROW_BYTES / 4);
				const uint32_t by0 = tile0 / (uint32_t)wib, bx0 = tile0 - by0 * (uint32_t)wib;	// uniform
				uint8_t *const row0 = dst + (size_t)by0 * 4 * pitch;
				// thread t owns row positions p = t, t+512, ... (p = id*CPR + half) for all 4 pixel rows
	#pragma unroll
				for (int j = 0; j < TB * CPR / SORT_THREADS; j++) {
					const uint32_t p = j * SORT_THREADS + tid, id = p / CPR, half = p % CPR;
					if (id < n_here) {
						uint32_t bx = bx0 + id, by = by0;
						while (bx >= (uint32_t)wib) {	// tiles may wrap over block rows (0-1 turns unless wib < TB)
							bx -= (uint32_t)wib;
							by++;
						}
						uint8_t *out = row0 + (size_t)(by - by0) * 4 * pitch + bx * (uint32_t)ROW_BYTES + half * 16;
						const int rows = height - (int)by * 4;	// >= 4 except in a clipped last block row
	#pragma unroll
						for (int r = 0; r < 4; r++) {
							if (r < rows) {
								const uint4 v = sm.out[stage_slot<TB, CHUNKS>(id, r * CPR + half)];
								if constexpr (PACK_ROWS) {		// four RGBX8 pixels -> 12 bytes
									st_stream_u32(out, prmt(v.x, v.y, 0x4210));
									st_stream_u32(out + 4, prmt(v.y, v.z, 0x5421));
									st_stream_u32(out + 8, prmt(v.z, v.w, 0x6542));
								}
								else
									st_stream_v4(out, v.x, v.y, v.z, v.w);
							}
							out += pitch;
						}
					}
				}
			}
		}
		if constexpr (SPLIT) {
			// ---- the mode-independent half, in image order: lane <-> consecutive blocks
#pragma unroll 1
			for (int k = 0; k < PER; k++) {
				const uint32_t id = k * SORT_THREADS + tid;
				if (id < n_here) {
					const uint32_t slot = raw_slot(id);
					Bc6hPrepared p;
					p.q0 = sm.out[slot];
					p.q1 = sm.out[TB + slot];
					p.q2 = sm.out[2 * TB + slot];
					const bool ok = bc6h_prepared_valid(p);
					uint32_t px[OW];
					pixels_px<F, PC>(p, ctx, px);
					if (!ok) {
#pragma unroll
						for (int j = 0; j < OW; j++)
							px[j] = invalid_fill(post_op);
					}
					emit<PC>(px, post_op, ctx, ok, [&](auto &w) { store_direct(tile0, id, w); });
				}
			}
		}
		if (has_next)
			total = scatter(cur ^ 1);
		if constexpr (TMA) {
#ifdef DETEXB200_TMA_EARLY_WAIT
			const bool wait_now = true;		// A/B: thread 0 waits for the TMA unit before B1
#else
			const bool wait_now = !has_next;	// (shared memory must outlive the reads of the last tile)
#endif
			if (wait_now) {
				if (tid == 0)
					tma_store_wait_read();	// the stage may be overwritten
			}
			else
				drain_pending = true;
		}
		if (!has_next)
			break;
		tile = next;
	}
	report_invalid(bad, invalid_flag);
}

#ifdef DETEXB200_EXPERIMENTS
// ---- mode-sorted kernel, second structure: one sorter warp, ticketed groups, one barrier per tile ---------------
// EXPERIMENT (round 2), measured slower than decode_sorted_kernel and therefore only in -DDETEXB200_EXPERIMENTS builds,
// selected at run time with DETEXB200_SORT_KERNEL=stream: BC7 16.8 M blocks M LINEAR 30.7 vs 47.9 G blocks/s, TILED 38.2 vs
// 48.6 (U TILED 45.2 vs 48.2); bit-exact (the GPU parity tests pass under it).  Why it loses: the sort is ~12 % of the
// decode's instructions whoever executes it, and ONE warp gets one eighth of its scheduler's issue slots while the
// seven decoding warps beside it are busy -- the sorter becomes the critical path and the decoders wait for it at the
// tile barrier (ncu: 44-56 % of all warp samples at that barrier).  The idea behind it:
// What ncu showed for decode_sorted_kernel on BC7 (distribution M, round 2): the decode itself -- two thirds of the
// instructions -- takes a quarter of the warp samples; the counting sort (shared atomics, shuffles, dependent shared
// loads: latency-bound, executed by ALL warps between barriers of its own), the loop head and the barriers take the
// rest.  Here the sort of a tile is the job of ONE warp, which needs no atomics and no CTA barrier for it, and it runs
// while the other warps decode the previous tile:
//   * the sorter warp requests tile t+1 (cp.async), and streams through it 32 blocks at a time: match the bins of the
//     32 blocks, append them to the open group of their bin (a running count per bin), open a new group whenever a bin
//     reaches a multiple of 32.  Groups of 32 same-mode ids come out in one pass, no histogram / prefix / scatter; the
//     open groups left at the end of the tile stay partly filled (the rest of their slots holds the sentinel);
//   * every warp -- the sorter too, once it is done -- draws groups of tile t from a ticket counter (drawn one group
//     ahead, so the atomic's latency is hidden) and decodes them;
//   * ONE __syncthreads per tile hands tile t+1 to the decoders and the buffers of tile t back to the sorter;
//   * staged LINEAR output (BC7, toned BC6H): thread 0 of the sorter warp hands the staged tile t-1 to the TMA unit
//     right after the barrier and waits until it has been read; the decoders meet it on a named barrier before their
//     FIRST write to the stage -- i.e. after a whole group has been decoded into registers, which covers the drain.
constexpr int STREAM_THREADS = 512, STREAM_WARPS = STREAM_THREADS / 32, STREAM_CTAS = 2;
template <int F, int PC, bool STAGE, int LAYOUT> constexpr int stream_tile() {
	if (STAGE)
		return 1024;
	return F == F_BPTC ? 3072 : 2048;
}
template <int F, int PC, bool STAGE, int LAYOUT>
struct StreamSmem {
	static constexpr int TB = stream_tile<F, PC, STAGE, LAYOUT>(), NB = SortTraits<F>::NUM_BINS;
	static constexpr int MAX_GROUPS = TB / 32 + NB;
	uint4 out[STAGE ? TB * 4 : 1];		// staged pixels (stage_slot<.., TMA> layout), first: 1024-byte aligned
	uint4 raw[2][TB];			// double-buffered compressed tile, filled by cp.async
	uint16_t ids[2][MAX_GROUPS * 32];	// groups of 32 tile-local block indices (sentinel = unused slot)
	int n_groups[2];
	int ticket[2];
	int delta[8][16];			// sorter: blocks per bin in each chunk of the current batch
	int gmap[16][16];			// sorter: group number of group q of a bin's stream, at [bin][q % 16]
	uint8_t bin_lut[256];
	BptcTables tables;
};

template <int F, int LAYOUT, bool STAGE, int PC>
__global__ void __launch_bounds__(STREAM_THREADS, STREAM_CTAS)
decode_stream_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, uint32_t n_blocks, int width,
int height, int wib, uint32_t mode_mask, uint32_t flags, uint32_t post_op, uint32_t *invalid_flag, uint32_t *invalid_map,
ToneArgs tone_args, const __grid_constant__ CUtensorMap tmap) {
	using T = FormatTraits<F>;
	using S = SortTraits<F>;
	using SM = StreamSmem<F, PC, STAGE, LAYOUT>;
	const EmitCtx<PC> ctx(tone_args);
	constexpr int OW = T::OUT_WORDS, NB = S::NUM_BINS, TB = SM::TB;
	constexpr bool TMA = STAGE;
	static_assert(T::BLOCK_WORDS == 4 && NB <= 15 && TB % 64 == 0, "16-byte blocks, at most 15 bins");
	static_assert(!STAGE || tma_capable<F, PC, true, LAYOUT>(), "the only staged output of this kernel is the TMA one");
	extern __shared__ __align__(16) unsigned char smem_raw[];
	SM &sm = *reinterpret_cast<SM *>(smem_raw + ((0u - (uint32_t)__cvta_generic_to_shared(smem_raw)) & 1023u));
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const bool sorter = warp == STREAM_WARPS - 1;
	const unsigned lt_mask = (1u << lane) - 1u;
	const uint32_t n_tiles = (n_blocks + TB - 1) / TB;
	uint32_t tile = blockIdx.x;
	if (tile >= n_tiles)
		return;
	bptc_tables_load(sm.tables);
	for (int v = tid; v < 256; v += STREAM_THREADS) {
		uint4 probe;
		probe.x = (uint32_t)v; probe.y = probe.z = probe.w = 0u;
		sm.bin_lut[v] = (uint8_t)S::bin_of(probe, mode_mask, flags);
	}
	if (tid < 2) {
		sm.n_groups[tid] = 0;
		sm.ticket[tid] = 0;
	}
	__syncthreads();

	// ---- sorter warp ---------------------------------------------------------------------------------------------
	auto request_tile = [&](uint32_t t, int buf) {		// (whole warp) cp.async of tile t into raw[buf]
		const uint32_t n_in_tile = min((uint32_t)TB, n_blocks - t * TB);
		for (uint32_t id = lane; id < n_in_tile; id += 32)
			cp_async_16(&sm.raw[buf][raw_slot(id)], src + ((size_t)t * TB + id) * 16);
		asm volatile("cp.async.commit_group;" ::: "memory");
	};
	// (whole warp) raw[buf] -> ids[buf], n_groups[buf].  Eight chunks of 32 blocks are in flight together so that the
	// shared-memory and shuffle latencies of a chunk overlap those of its neighbours: per batch there are six
	// warp-wide dependent steps instead of four per chunk.  Lane b < 16 keeps the running count of bin b in a register.
	auto sort_tile_into = [&](uint32_t t, int buf) {
		const uint32_t n_in_tile = min((uint32_t)TB, n_blocks - t * TB);
		uint32_t *ids32 = reinterpret_cast<uint32_t *>(sm.ids[buf]);
		for (int i = lane; i < SM::MAX_GROUPS * 16; i += 32)
			ids32[i] = 0xFFFFFFFFu;
		int my_cnt = 0, n_groups = 0;
		asm volatile("cp.async.wait_group 0;" ::: "memory");
		__syncwarp();
		constexpr int BATCH = 8;
		for (uint32_t c0 = 0; c0 * 32 < n_in_tile; c0 += BATCH) {
			int bin[BATCH], rank[BATCH], pos[BATCH];
			reinterpret_cast<uint4 *>(&sm.delta[0][0])[lane] = make_uint4(0u, 0u, 0u, 0u);	// 8 x 16 counts
#pragma unroll
			for (int k = 0; k < BATCH; k++) {
				const uint32_t id = (c0 + k) * 32 + lane;
				bin[k] = 15;				// lanes beyond the tile: a bin of their own, never emitted
				if (id < n_in_tile)
					bin[k] = sm.bin_lut[reinterpret_cast<const uint8_t *>(&sm.raw[buf][raw_slot(id)])[0]];
			}
			__syncwarp();
#pragma unroll
			for (int k = 0; k < BATCH; k++) {
				const unsigned peers = __match_any_sync(0xFFFFFFFFu, bin[k]);
				rank[k] = __popc(peers & lt_mask);
				if (rank[k] == 0)
					sm.delta[k][bin[k]] = __popc(peers);	// blocks of this bin in chunk k
			}
			__syncwarp();
			// lane b: count of bin b before each chunk of the batch
#pragma unroll
			for (int k = 0; k < BATCH; k++) {
				pos[k] = my_cnt;
				my_cnt += sm.delta[k][lane & 15];
			}
			unsigned openers[BATCH];
#pragma unroll
			for (int k = 0; k < BATCH; k++) {
				pos[k] = __shfl_sync(0xFFFFFFFFu, pos[k], bin[k]) + rank[k];	// position of the block in its bin's stream
				openers[k] = __ballot_sync(0xFFFFFFFFu, bin[k] != 15 && (pos[k] & 31) == 0);
			}
#pragma unroll
			for (int k = 0; k < BATCH; k++) {
				if ((openers[k] >> lane) & 1u)		// first block of a new group of its bin
					sm.gmap[bin[k]][(pos[k] >> 5) & 15] = n_groups + __popc(openers[k] & lt_mask);
				n_groups += __popc(openers[k]);
			}
			__syncwarp();
#pragma unroll
			for (int k = 0; k < BATCH; k++)
				if (bin[k] != 15)
					sm.ids[buf][sm.gmap[bin[k]][(pos[k] >> 5) & 15] * 32 + (pos[k] & 31)] = (uint16_t)((c0 + k) * 32 + lane);
			__syncwarp();
		}
		if (lane == 0) {
			sm.n_groups[buf] = n_groups;
			sm.ticket[buf] = 0;
		}
	};
	// the staged tile that starts at block t0 -> image: one bulk tensor store per box of 64 blocks (one thread)
	auto tma_store_tile = [&](uint32_t t0, uint32_t n_in_tile) {
		uint32_t by = t0 / (uint32_t)wib, bx = t0 - by * (uint32_t)wib;
		for (uint32_t b = 0; 64u * b < n_in_tile; b++) {
			tma_store_3d(&tmap, &sm.out[256u * b], 0, (int)(bx >> 3), (int)(4u * by));
			bx += 64u;
			if (bx >= (uint32_t)wib) {
				bx = 0u;
				by++;
			}
		}
		tma_store_commit();
	};

	if (sorter) {
		request_tile(tile, 0);
		sort_tile_into(tile, 0);
	}
	bool bad = false;
	for (int it = 0;; it++) {
		const int cur = it & 1;
		const uint32_t tile0 = tile * TB;
		const uint32_t next = tile + gridDim.x;
		const bool has_next = next < n_tiles;
		__syncthreads();	// tile `tile` is sorted (ids[cur], n_groups[cur]); everything of the previous tile is done
		bool stage_open = !(TMA && it > 0);	// may this thread write to the stage yet?
		if (sorter) {
			if constexpr (TMA) {
				if (it > 0 && lane == 0)
					tma_store_tile(tile0 - gridDim.x * TB, TB);	// (only the last tile of the texture can be partial)
			}
			if (has_next)
				request_tile(next, cur ^ 1);
			if constexpr (TMA) {
				if (it > 0) {
					if (lane == 0)
						tma_store_wait_read();
					__syncwarp();
					asm volatile("bar.arrive 1, %0;" :: "n"(STREAM_THREADS) : "memory");	// the stage is free
					stage_open = true;
				}
			}
			if (has_next)
				sort_tile_into(next, cur ^ 1);
		}
		// ---- decode: groups drawn from the ticket, one ahead
		const int n_groups = *(volatile int *)&sm.n_groups[cur];
		int g_next = 0;
		if (lane == 0)
			g_next = atomicAdd(&sm.ticket[cur], 1);
		g_next = __shfl_sync(0xFFFFFFFFu, g_next, 0);
		uint32_t id_next = g_next < n_groups ? sm.ids[cur][g_next * 32 + lane] : SLOT_EMPTY;
		while (g_next < n_groups) {
			const uint32_t id = id_next;
			if (lane == 0)
				g_next = atomicAdd(&sm.ticket[cur], 1);
			g_next = __shfl_sync(0xFFFFFFFFu, g_next, 0);
			id_next = g_next < n_groups ? sm.ids[cur][g_next * 32 + lane] : SLOT_EMPTY;
			const bool live = id != SLOT_EMPTY;
			uint32_t px[OW];
			bool ok = true;
			const uint4 w = sm.raw[cur][raw_slot(live ? id : 0u)];
			if (live) {
				ok = decode_px<F, PC>(w, mode_mask, flags, &sm.tables, ctx, px);
				if (!ok) {
					bad = true;
					mark_invalid(invalid_map, tile0 + id);
#pragma unroll
					for (int k = 0; k < OW; k++)
						px[k] = invalid_fill(post_op);
				}
			}
			if constexpr (TMA) {
				if (!stage_open) {		// (uniform over the warp) the staged previous tile has been read
					asm volatile("bar.sync 1, %0;" :: "n"(STREAM_THREADS) : "memory");
					stage_open = true;
				}
			}
			if (live) {
				if constexpr (STAGE && is_tone_class(PC)) {
					uint32_t t[16];
					tone_tile<PC>(px, t, post_op, ctx, ok);
#pragma unroll
					for (int c = 0; c < 4; c++)
						sm.out[stage_slot<TB, 4, true>(id, c)] = make_uint4(t[4 * c], t[4 * c + 1], t[4 * c + 2], t[4 * c + 3]);
				}
				else if constexpr (STAGE) {
					apply_post_op<OW>(px, post_op);
#pragma unroll
					for (int c = 0; c < 4; c++)
						sm.out[stage_slot<TB, 4, true>(id, c)] = make_uint4(px[4 * c], px[4 * c + 1], px[4 * c + 2], px[4 * c + 3]);
				}
				else {
					emit<PC>(px, post_op, ctx, ok, [&](auto &q) {
						constexpr int OWO = post_words<OW, PC>(), CH = OWO / 4;
						if (LAYOUT == LAYOUT_TILED) {
							uint8_t *out = dst + (size_t)(tile0 + id) * (OWO * 4);
							if (CH % 2 == 0 && (((uintptr_t)dst) & 31u) == 0) {
#pragma unroll
								for (int c = 0; c + 1 < CH; c += 2)
									st_stream_v8(out + 16 * c, q[4 * c], q[4 * c + 1], q[4 * c + 2], q[4 * c + 3],
										q[4 * c + 4], q[4 * c + 5], q[4 * c + 6], q[4 * c + 7]);
							}
							else {
#pragma unroll
								for (int c = 0; c < CH; c++)
									st_stream_v4(out + 16 * c, q[4 * c], q[4 * c + 1], q[4 * c + 2], q[4 * c + 3]);
							}
						}
						else {
							const uint32_t i = tile0 + id;
							const uint32_t by = i / (uint32_t)wib, bx = i - by * (uint32_t)wib;
							store_linear_rows(dst, (size_t)width * (OWO / 4), (int)bx, (int)by, height, q);
						}
					});
				}
			}
		}
		if constexpr (TMA) {
			if (!stage_open)		// a warp that drew no group still owes the named barrier its arrival
				asm volatile("bar.arrive 1, %0;" :: "n"(STREAM_THREADS) : "memory");
			fence_proxy_async_smem();	// my writes to the staged tile -> visible to the TMA unit
		}
		if (!has_next)
			break;
		tile = next;
	}
	if constexpr (TMA) {
		__syncthreads();		// the last tile is staged
		if (tid == 0) {
			tma_store_tile(tile * TB, min((uint32_t)TB, n_blocks - tile * TB));
			tma_store_wait_read();
		}
	}
	report_invalid(bad, invalid_flag);
}

#endif  // DETEXB200_EXPERIMENTS

#ifdef DETEXB200_EXPERIMENTS
// (a measured-and-dropped design, kept out of the default build: -DDETEXB200_EXPERIMENTS compiles it in and
// DETEXB200_BC7_KERNEL=flow selects it at run time; profiles/README.md, "BC7 palette / dataflow experiment")
// ---- BC7: mode-sorted, palette-staged, barrier-free dataflow kernel ---------------------------
// One persistent CTA of 1024 threads per SM works through tiles of 1024 consecutive blocks.  As in
// decode_sorted_kernel the blocks of a tile are counting-sorted by mode so that a warp decodes 32
// blocks of ONE mode, but
//   * what a lane parks in shared memory is not the 16 pixels of its block: it is the block's
//     PALETTE (<= 16 colours, each computed with a compile-time subset and weight,
//     decode_bptc_palette.cuh) plus 16 selector bytes; the look-ups happen in the copy-out, whose
//     lanes map to blocks in image order so that their stores are contiguous;
//   * there is no CTA-wide barrier after start-up.  The cost of a decode group differs 6x between
//     modes, so any scheme that makes all warps meet once per tile leaves most of them waiting for
//     the mode-0 groups (measured: 24 % of all warp time).  Instead
//       - four BOOKKEEPER warps run ahead through the tiles: cp.async load (one tile ahead),
//         classify, histogram, bin starts, scatter of the sorted ids -- synchronising among
//         themselves with a 128-thread named barrier -- and publish `book_done`;
//       - every other warp (and the bookkeepers once they run out of tiles) draws WORK ITEMS from one
//         monotonic ticket: per tile FLOW_DEC decode slots (group r of tile j; slots beyond the
//         tile's group count are empty) followed by FLOW_OUT copy-out chunks of tile j-1;
//       - dependencies are monotonic counters in shared memory (decode slots finished per stage
//         buffer, copy-out chunks finished per compressed-tile buffer, tiles scattered), waited for
//         with a short polling loop.  Every wait is for work with smaller ticket numbers (or for the
//         bookkeepers, whose own waits are for strictly older tiles), so the scheme cannot deadlock.
//
// Buffers: stage[tile & 1], raw[tile & 3] (compressed blocks, later the selectors), id[tile & 1].
//   decode(j) needs: tile j scattered; stage[j & 1] free = copy-out(j-2) complete
//   copy-out(j) needs: decode(j) complete
//   load(j) needs: raw[j & 3] free = copy-out(j-4) complete
//   scatter(j) needs: id[j & 1] free = decode(j-2) complete
//
// Stage layout: word (chunk, e, p) at chunk*1024 + e*64 + p with chunk = id / 64 and
// p = (id % 64) ^ ((id / 32) % 8): the bank of an entry is p % 32 whatever e is, so
//   * the copy-out look-ups of 32 consecutive ids never conflict (id / 32 is warp-uniform),
//   * the decode stores of a same-mode group conflict only by chance (the XOR spreads the
//     adversarial distribution M, where all ids of a group are congruent mod 8).
// A stage is 4 KiB-aligned in the shared window and an entry row is 256 bytes, so the address of
// entry e of block id is base(id) | e << 8: with the selector bytes pre-biased by byte 1 of
// base(id), ONE byte permute builds the address of a pixel (bytes 0, 2, 3 from base(id), byte 1
// from the selector word).  The selectors overwrite the block's own slot of the compressed tile.
constexpr int FLOW_TB = 1024, FLOW_NT = 1024, FLOW_NB = 9, FLOW_CAP = FLOW_TB + FLOW_NB * 32;
constexpr int FLOW_BOOK = 4, FLOW_BT = FLOW_BOOK * 32, FLOW_PER = FLOW_TB / FLOW_BT;
constexpr int FLOW_DEC = (FLOW_TB + FLOW_NB * 31 + 31) / 32, FLOW_OUT = FLOW_TB / 32, FLOW_ITEMS = FLOW_DEC + FLOW_OUT;
constexpr int FLOW_SPIN_LIMIT = 1 << 22;	// polls before a wait gives up and flags the launch as failed
struct Bc7FlowSmem {
	uint32_t stage[2][FLOW_TB * 16];	// [tile & 1]
	uint4 raw[4][FLOW_TB];			// [tile & 3] compressed blocks, later the selectors
	uint16_t id[2][FLOW_CAP];		// [tile & 1] tile-local block index per sorted slot
	int start[2][16];			// [tile & 1] first sorted slot of bin b; [NB] = padded total
	int count[2][16];			// [tile & 1] per-mode histogram
	int ticket;				// next work item
	int book_done;				// tiles scattered
	int dec_cnt[2];				// decode slots finished, per stage buffer (monotonic)
	int out_cnt[4];				// copy-out chunks finished, per compressed-tile buffer (monotonic)
	int error;
	uint8_t bin_lut[256];			// byte 0 of a block -> bin (mode, or 8 = invalid under mode_mask / flags)
	Bc7PaletteTables tables;
};
constexpr size_t FLOW_SMEM_BYTES = sizeof(Bc7FlowSmem) + 4096;	// + alignment slack

__device__ __forceinline__ void st_shared_u32(uint32_t addr, uint32_t v) {
	asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}
// shared-window byte address of entry 0 of block id (tile-local index)
__device__ __forceinline__ uint32_t pal_base(uint32_t stage_base, uint32_t id) {
	return stage_base + ((id >> 6) << 12) + ((((id & 63u) ^ ((id >> 5) & 7u))) << 2);
}
// acquire side of a counter dependency: poll until *counter >= target
__device__ __forceinline__ void flow_wait(const int *counter, int target, int *error) {
	const volatile int *c = counter;
	if (*c < target) {
		int spins = 0;
		while (*c < target && *(volatile int *)error == 0) {	// (the error flag is sticky: a failed launch ends quickly)
			__nanosleep(64);
			if (++spins > FLOW_SPIN_LIMIT)
				*(volatile int *)error = 1;
		}
	}
	__threadfence_block();
}
// release side: the warp's shared-memory writes become visible, then one lane bumps the counter
__device__ __forceinline__ void flow_signal(int *counter, int lane) {
	__threadfence_block();
	__syncwarp();
	if (lane == 0)
		atomicAdd(counter, 1);
}

template <int LAYOUT, int PC>
__global__ void __launch_bounds__(FLOW_NT, 1)
decode_bc7_flow_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, uint32_t n_blocks, int width,
int height, int wib, uint32_t mode_mask, uint32_t flags, uint32_t post_op, uint32_t *invalid_flag, uint32_t *invalid_map) {
	using S = SortTraits<F_BPTC>;
	using SM = Bc7FlowSmem;
	constexpr int NB = FLOW_NB, TB = FLOW_TB, BT = FLOW_BT, PER = FLOW_PER;
	static_assert(NB == S::NUM_BINS, "bins");
	static_assert(PC == PC_NATIVE || PC == PC_PACK3, "BC7 post classes");
	constexpr int OWO = post_words<16, PC>();		// words of a stored tile
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const uint32_t window = (uint32_t)__cvta_generic_to_shared(smem_raw);
	SM &sm = *reinterpret_cast<SM *>(smem_raw + ((0u - window) & 4095u));
	const uint32_t stage_base0 = (uint32_t)__cvta_generic_to_shared(sm.stage);	// multiple of 4096
	const int tid = threadIdx.x, lane = tid & 31;
	const uint32_t n_tiles = (n_blocks + TB - 1) / TB;
	// tiles of this CTA: tile(j) = blockIdx.x + j * gridDim.x, j = 0 .. n_local-1
	const int n_local = blockIdx.x < n_tiles ? (int)((n_tiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;
	const bool swap_rb = (post_op & POST_SWAP_RB) != 0;
	const size_t pitch = (size_t)width * (OWO / 4);
	if (tid < 32)
		(&sm.count[0][0])[tid] = 0;
	if (tid < 16)
		(&sm.ticket)[tid < 9 ? tid : 0] = 0;		// ticket, book_done, dec_cnt[2], out_cnt[4], error
	bptc_tables_load(sm.tables.packed);
	if (tid < 64) {
		sm.tables.sub2[tid] = bc7_subset_bytes(bptc_two_global[tid] & 0xFFFFu, 1);
		sm.tables.sub3[tid] = bc7_subset_bytes(bptc_three_global[tid].x, 2);
	}
	if (tid < 256) {
		uint4 probe;
		probe.x = (uint32_t)tid; probe.y = probe.z = probe.w = 0u;
		sm.bin_lut[tid] = (uint8_t)S::bin_of(probe, mode_mask, flags);
	}
	__syncthreads();
	auto tile_of = [&](int j) { return blockIdx.x + (uint32_t)j * gridDim.x; };
	auto blocks_in = [&](int j) { return min((uint32_t)TB, n_blocks - tile_of(j) * TB); };

	// ================= bookkeepers: load, classify, histogram, bin starts, scatter =================
	if (tid < BT) {
		auto issue_load = [&](int j) {
			const uint32_t tile0 = tile_of(j) * TB;
#pragma unroll
			for (int k = 0; k < PER; k++) {
				const uint32_t id = k * BT + tid, i = tile0 + id;
				if (i < n_blocks)
					cp_async_16(&sm.raw[j & 3][raw_slot(id)], src + (size_t)i * 16);
			}
			asm volatile("cp.async.commit_group;" ::: "memory");
		};
		if (n_local > 0)
			issue_load(0);
		for (int b = 0; b < n_local; b++) {
			// one tile ahead: its buffer held tile b-3, whose selectors the copy-out must have consumed
			if (b + 1 < n_local) {
				if (b + 1 >= 4)
					flow_wait(&sm.out_cnt[(b + 1) & 3], FLOW_OUT * ((b - 3) / 4 + 1), &sm.error);
				issue_load(b + 1);
				asm volatile("cp.async.wait_group 1;" ::: "memory");
			}
			else
				asm volatile("cp.async.wait_group 0;" ::: "memory");
			// id[b & 1], start[b & 1] are free once every decode slot of tile b-2 has finished
			if (b >= 2)
				flow_wait(&sm.dec_cnt[b & 1], FLOW_DEC * ((b - 2) / 2 + 1), &sm.error);
			const uint32_t n_here = blocks_in(b);
			int *count = sm.count[b & 1];
			int bin[PER], off[PER];
#pragma unroll
			for (int k = 0; k < PER; k++) {
				const uint32_t id = k * BT + tid;
				bin[k] = -1;
				if (id < n_here)
					bin[k] = sm.bin_lut[reinterpret_cast<const uint8_t *>(&sm.raw[b & 3][raw_slot(id)])[0]];
			}
#pragma unroll
			for (int k = 0; k < PER; k++) {
				const unsigned peers = __match_any_sync(0xFFFFFFFFu, bin[k]);
				const int leader = __ffs((int)peers) - 1;
				int base = 0;
				if (lane == leader && bin[k] >= 0)
					base = atomicAdd(&count[bin[k]], __popc(peers));
				base = __shfl_sync(0xFFFFFFFFu, base, leader);
				off[k] = base + __popc(peers & ((1u << lane) - 1u));
			}
			asm volatile("bar.sync 1, %0;" :: "n"(FLOW_BT) : "memory");	// histogram complete; everyone's copies landed
			const int my_count = lane < NB ? count[lane] : 0;
			const int padded = (my_count + 31) & ~31;
			int st = padded;
#pragma unroll
			for (int d = 1; d < 16; d *= 2) {
				const int t = __shfl_up_sync(0xFFFFFFFFu, st, d);
				if (lane >= d)
					st += t;
			}
			st -= padded;			// lane b: first sorted slot of bin b; lane NB: the padded total
			if (tid < 16)
				sm.start[b & 1][tid] = st;
			uint16_t *ids = sm.id[b & 1];
			// padding slots at the end of every bin get the sentinel (one warp per bin)
			for (int q = tid >> 5; q < NB; q += FLOW_BOOK) {
				const int c = __shfl_sync(0xFFFFFFFFu, my_count, q), s0 = __shfl_sync(0xFFFFFFFFu, st, q);
				if (lane < ((-c) & 31))
					ids[s0 + c + lane] = (uint16_t)SLOT_EMPTY;
			}
#pragma unroll
			for (int k = 0; k < PER; k++) {
				const int s0 = __shfl_sync(0xFFFFFFFFu, st, bin[k] & 31);
				if (bin[k] >= 0)
					ids[s0 + off[k]] = (uint16_t)(k * BT + tid);
			}
			asm volatile("bar.sync 1, %0;" :: "n"(FLOW_BT) : "memory");	// scatter complete, histogram consumed
			if (tid < 16)
				count[tid] = 0;			// for tile b+2
			if (tid == 0) {
				__threadfence_block();
				*(volatile int *)&sm.book_done = b + 1;
			}
		}
	}

	// ================= workers: decode groups and copy-out chunks from the ticket =================
	bool bad = false;
	auto draw = [&]() {
		int t = 0;
		if (lane == 0)
			t = atomicAdd(&sm.ticket, 1);
		return __shfl_sync(0xFFFFFFFFu, t, 0);
	};
	int t_next = draw();
	while (true) {
		const int t = t_next;
		const int j = t / FLOW_ITEMS, r = t - j * FLOW_ITEMS;	// tile of the decode slots; copy-out chunks belong to tile j-1
		if (j > n_local)
			break;
		t_next = draw();
		if (r < FLOW_DEC) {
			if (j >= n_local)
				continue;
			// ---- decode slot r of tile j
			flow_wait(&sm.book_done, j + 1, &sm.error);
			const int start_cur = sm.start[j & 1][lane & 15];
			const int n_dec = __shfl_sync(0xFFFFFFFFu, start_cur, NB);
			const int g = r * 32;
			if (g < n_dec) {
				if (j >= 2)		// stage[j & 1] held tile j-2
					flow_wait(&sm.out_cnt[(j - 2) & 3], FLOW_OUT * ((j - 2) / 4 + 1), &sm.error);
				const uint32_t id = sm.id[j & 1][g + lane];
				// the group's bin from its position: number of bins that start at or before g (uniform)
				const int bin_g = __popc(__ballot_sync(0xFFFFFFFFu, lane < NB && start_cur <= g)) - 1;
				int mode = 8;
#pragma unroll
				for (int b = 0; b < 8; b++)
					if (bin_g == b)
						mode = bc7_mode_of_bin(b);
				if (id != SLOT_EMPTY) {
					uint4 *raw_dec = sm.raw[j & 3];
					const uint32_t rs = raw_slot(id);
					const uint4 w = raw_dec[rs];
					const uint32_t base = pal_base(stage_base0 + (uint32_t)(j & 1) * (TB * 64u), id);
					const uint32_t bias = prmt(base, 0u, 0x1111);	// byte 1 of the address in every byte
					if (mode == 8) {
						bad = true;
						mark_invalid(invalid_map, tile_of(j) * TB + id);
						st_shared_u32(base, invalid_fill(post_op));
						raw_dec[rs] = make_uint4(bias, bias, bias, bias);
					}
					else {
						bc7_palette(w, mode, sm.tables, swap_rb, [&](uint32_t (&pal)[16], uint32_t (&sel)[4], auto ne_) {
							constexpr int NE = decltype(ne_)::value;
							if (post_op & (POST_FLIP16 | POST_FLIP8)) {	// (never set for RGBA8 targets by the API)
								const uint32_t top = (post_op & POST_FLIP16) ? 0x80008000u : 0x80808080u;
#pragma unroll
								for (int e = 0; e < NE; e++)
									pal[e] ^= top;
							}
#pragma unroll
							for (int e = 0; e < NE; e++)
								st_shared_u32(base + e * 256, pal[e]);
							raw_dec[rs] = make_uint4(sel[0] + bias, sel[1] + bias, sel[2] + bias, sel[3] + bias);
						});
					}
				}
			}
			flow_signal(&sm.dec_cnt[j & 1], lane);
		}
		else {
			if (j < 1)
				continue;
			// ---- copy-out chunk of tile j-1: 32 consecutive blocks, lane <-> block
			const int jj = j - 1, g = (r - FLOW_DEC) * 32;
			flow_wait(&sm.dec_cnt[jj & 1], FLOW_DEC * (jj / 2 + 1), &sm.error);
			const uint32_t n_here = blocks_in(jj), tile0 = tile_of(jj) * TB;
			const uint32_t cid = (uint32_t)g + lane;
			// block coordinates of the chunk's first block (uniform), then per lane
			const uint32_t i0 = tile0 + (uint32_t)g;
			const uint32_t by0 = LAYOUT == LAYOUT_LINEAR ? i0 / (uint32_t)wib : 0u, bx0 = i0 - by0 * (uint32_t)wib;
			if (cid < n_here) {
				const uint4 s4 = sm.raw[jj & 3][raw_slot(cid)];
				const uint32_t base = pal_base(stage_base0 + (uint32_t)(jj & 1) * (TB * 64u), cid);
				const uint32_t sw[4] = { s4.x, s4.y, s4.z, s4.w };
				uint32_t px[16];
#pragma unroll
				for (int q = 0; q < 4; q++) {
					px[4 * q + 0] = ld_shared_u32(prmt(sw[q], base, 0x7604));
					px[4 * q + 1] = ld_shared_u32(prmt(sw[q], base, 0x7614));
					px[4 * q + 2] = ld_shared_u32(prmt(sw[q], base, 0x7624));
					px[4 * q + 3] = ld_shared_u32(prmt(sw[q], base, 0x7634));
				}
				uint32_t q[OWO];
				if constexpr (PC == PC_PACK3)
					repack<PC_PACK3, 16>(px, q);
				else {
#pragma unroll
					for (int k = 0; k < 16; k++)
						q[k] = px[k];
				}
				if (LAYOUT == LAYOUT_TILED) {
					uint8_t *out = dst + (size_t)(tile0 + cid) * (OWO * 4);
					if (PC == PC_NATIVE && (((uintptr_t)dst) & 31u) == 0) {
						st_stream_v8(out, q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7]);
						st_stream_v8(out + 32, q[OWO - 8], q[OWO - 7], q[OWO - 6], q[OWO - 5], q[OWO - 4], q[OWO - 3], q[OWO - 2], q[OWO - 1]);
					}
					else {
#pragma unroll
						for (int c = 0; c < OWO / 4; c++)
							st_stream_v4(out + 16 * c, q[4 * c], q[4 * c + 1], q[4 * c + 2], q[4 * c + 3]);
					}
				}
				else {
					uint32_t bx = bx0 + lane, by = by0;
					while (bx >= (uint32_t)wib) {	// a chunk may wrap over block rows
						bx -= (uint32_t)wib;
						by++;
					}
					uint8_t *pp = dst + (size_t)by * 4 * pitch + bx * (uint32_t)OWO;	// OWO bytes per pixel row of a block
					const int rows = height - (int)by * 4;	// >= 4 except in a clipped last block row (texture.c:117-120)
#pragma unroll
					for (int rr = 0; rr < 4; rr++) {
						if (rr < rows) {
							if constexpr (PC == PC_PACK3) {
								st_stream_u32(pp, q[3 * rr]);
								st_stream_u32(pp + 4, q[3 * rr + 1]);
								st_stream_u32(pp + 8, q[3 * rr + 2]);
							}
							else
								st_stream_v4(pp, q[4 * rr], q[4 * rr + 1], q[4 * rr + 2], q[4 * rr + 3]);
						}
						pp += pitch;
					}
				}
			}
			flow_signal(&sm.out_cnt[jj & 3], lane);
		}
	}
	if (*(volatile int *)&sm.error && invalid_flag != nullptr)
		*(volatile uint32_t *)invalid_flag = 0x7FFFFFFFu;	// a dependency wait gave up: the output is not trustworthy
	report_invalid(bad, invalid_flag);
}

#endif  // DETEXB200_EXPERIMENTS

// ---- host-side launcher ---------------------------------------------------------------------
static std::atomic<uint64_t> g_launches{0};

uint64_t launch_count() { return g_launches.load(); }
void count_launch() { g_launches.fetch_add(1); }

static std::atomic<int> g_sm_count{0};
int sm_count() {
	int n = g_sm_count.load();
	if (n == 0) {
		int dev = 0;
		if (cudaGetDevice(&dev) != cudaSuccess ||
		    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
			n = 148;
		g_sm_count.store(n);
	}
	return n;
}

// DETEXB200_UNSORTED=1 in the environment forces the plain one-block-per-lane kernels
// (A/B measurements of the mode sort).
static bool force_unsorted() {
	static const bool v = [] { const char *e = getenv("DETEXB200_UNSORTED"); return e && e[0] == '1'; }();
	return v;
}

// TILED output of tiles >= 32 bytes is written by per-lane 256-bit stores; DETEXB200_TILED_DIRECT=0
// selects the older shared-memory transposition / staging instead (A/B measurements).
static bool tiled_direct() {
	static const bool v = [] { const char *e = getenv("DETEXB200_TILED_DIRECT"); return !(e && e[0] == '0'); }();
	return v;
}

// Bin padding of the sorted kernels (template parameter PAD of decode_sorted_kernel).  BC7 modes are
// separate code, so a warp must not mix them: 32.  BC6H modes share everything after the bit scatter:
// the unsigned direct-store kernels run unpadded (measured, G blocks/s LINEAR / TILED, padded -> unpadded:
// distribution M 25.2 / 24.4 -> 26.5 / 25.9, U 27.3 / 26.5 -> 27.0 / 26.2); the signed format loses
// 5-12 % without padding and the staged tone kernels 6 % (1024-block tiles: 20.2 -> 18.9), so they keep 32.
// The staged FLOAT_RGB16 kernel (512-block tiles: 14 padded bins would leave 43 % of the lanes idle) also
// runs unpadded.
template <int F, int PC, bool STAGE> constexpr int sort_pad() {
	return (F == F_BPTC_FLOAT && (!STAGE || stages_packed_half3<F, PC>())) ? 1 : 32;
}

// One instantiation of the sorted kernel: opt in to > 48 KiB of dynamic shared memory once per
// device, then launch.
static ToneArgs tone_args_of(const DecodeArgs &a) {
	ToneArgs t;
	t.range_min = a.tone_min;
	t.range_max = a.tone_max;
	t.inv_gamma = a.tone_inv_gamma;
	t.gamma_lo = a.tone_gamma_lo;
	t.gamma_factor = a.tone_gamma_factor;
	t.zero_target = a.tone_zero_target;
	return t;
}

// Tensor map of a LINEAR image for tma_store_3d: bytes {128, pitch / 128, height}, box {128, 8, 4}, SWIZZLE_128B.
// cuTensorMapEncodeTiled is a driver-API function; it is fetched through the runtime so that the library keeps
// linking against cudart only.  Returns false when the driver entry point is missing or refuses the shape (the
// caller then launches the cooperative copy-out variant).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
	const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
	static const EncodeTiledFn fn = [] {
		void *p = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
		    q != cudaDriverEntryPointSuccess) {
			(void)cudaGetLastError();
			p = nullptr;
		}
		return (EncodeTiledFn)p;
	}();
	return fn;
}
// DETEXB200_TMA_STORE=0 in the environment selects the cooperative copy-out of the staged LINEAR kernels (A/B runs)
static bool tma_store_enabled() {
	static const bool v = [] { const char *e = getenv("DETEXB200_TMA_STORE"); return !(e && e[0] == '0'); }();
	return v;
}
// row_bytes = bytes of one pixel row of one block: 16 (line of eight blocks = 128 bytes, SWIZZLE_128B) or 24 (192-byte lines,
// no swizzle: the swizzle modes need an inner extent of at most 128 bytes)
static bool make_linear_tensor_map(CUtensorMap *map, void *dst, size_t pitch, int height, int row_bytes) {
	const EncodeTiledFn encode = encode_tiled_fn();
	const size_t line = 8 * (size_t)row_bytes;
	if (encode == nullptr || pitch % (8 * line) != 0 || height < 1 || (((uintptr_t)dst) & 15u) != 0)
		return false;
	const cuuint64_t dims[3] = { (cuuint64_t)line, (cuuint64_t)(pitch / line), (cuuint64_t)height };
	const cuuint64_t strides[2] = { (cuuint64_t)line, (cuuint64_t)pitch };		// bytes, dimensions 1 and 2
	const cuuint32_t box[3] = { (cuuint32_t)line, 8, 4 };			// 64 blocks x 4 pixel rows
	const cuuint32_t elem[3] = { 1, 1, 1 };
	return encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, dst, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
		row_bytes == 16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
		CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Start-up stagger of the persistent sorted kernels (see the kernel): about one tile period.  DETEXB200_STAGGER_NS
// overrides it (0 switches it off; A/B runs).
static uint32_t stagger_ns() {
	static const uint32_t v = [] { const char *e = getenv("DETEXB200_STAGGER_NS"); return e ? (uint32_t)atoi(e) : 0u; }();
	return v;
}

template <int F, int LAYOUT, bool STAGE, int PC, int PAD, bool TMA, bool SPLIT = false>
static void launch_sorted_variant(const uint8_t *src, uint8_t *dst, uint32_t n_blocks, const DecodeArgs &a, cudaStream_t stream,
const CUtensorMap &tmap) {
	using SM = SortedSmem<F, PC, STAGE, LAYOUT, SPLIT>;
	const size_t smem = sizeof(SM) + (TMA ? 1024 : 0);		// + alignment slack for the swizzled stage
	const uint64_t tiles = ((uint64_t)n_blocks + SM::TB - 1) / SM::TB;
	constexpr int SORT_THREADS = sort_threads<F>();
	const uint64_t resident = (uint64_t)sm_count() * (1024 / SORT_THREADS);	// persistent: 1024 threads per SM
	const unsigned grid = (unsigned)(tiles < resident ? tiles : resident);
	static std::atomic<bool> configured[64];
	int dev = 0;
	cudaGetDevice(&dev);
	if (!configured[dev & 63].load()) {
		cudaFuncSetAttribute(decode_sorted_kernel<F, LAYOUT, STAGE, PC, PAD, TMA, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		configured[dev & 63].store(true);
	}
	decode_sorted_kernel<F, LAYOUT, STAGE, PC, PAD, TMA, SPLIT><<<grid, SORT_THREADS, smem, stream>>>(src, dst, n_blocks,
		a.width, a.height, a.wib, a.mode_mask, a.flags, a.post_op, a.invalid, a.invalid_map, tone_args_of(a),
		tiles > 2 * resident ? stagger_ns() : 0u, tmap);
}

// One instantiation of the sorted kernel: opt in to > 48 KiB of dynamic shared memory once per
// device, then launch.  Staged LINEAR tiles of 16-byte rows leave through the TMA unit when the image allows it
// (whole boxes of 64 blocks per block row, 16-byte aligned base).
#ifdef DETEXB200_EXPERIMENTS
// DETEXB200_SORT_KERNEL=stream selects decode_stream_kernel where it can run (experiment builds only).
static bool stream_kernel_enabled() {
	static const bool v = [] { const char *e = getenv("DETEXB200_SORT_KERNEL"); return e && e[0] == 's'; }();
	return v;
}

template <int F, int LAYOUT, bool STAGE, int PC>
static void launch_stream(const uint8_t *src, uint8_t *dst, uint32_t n_blocks, const DecodeArgs &a, cudaStream_t stream,
const CUtensorMap &tmap) {
	using SM = StreamSmem<F, PC, STAGE, LAYOUT>;
	const size_t smem = sizeof(SM) + 1024;		// + slack to align the struct (the swizzled stage) to 1024 bytes
	const uint64_t tiles = ((uint64_t)n_blocks + SM::TB - 1) / SM::TB;
	const uint64_t resident = (uint64_t)sm_count() * STREAM_CTAS;		// persistent: two CTAs per SM
	const unsigned grid = (unsigned)(tiles < resident ? tiles : resident);
	static std::atomic<bool> configured[64];
	int dev = 0;
	cudaGetDevice(&dev);
	if (!configured[dev & 63].load()) {
		cudaFuncSetAttribute(decode_stream_kernel<F, LAYOUT, STAGE, PC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		configured[dev & 63].store(true);
	}
	decode_stream_kernel<F, LAYOUT, STAGE, PC><<<grid, STREAM_THREADS, smem, stream>>>(src, dst, n_blocks,
		a.width, a.height, a.wib, a.mode_mask, a.flags, a.post_op, a.invalid, a.invalid_map, tone_args_of(a), tmap);
}

#endif  // DETEXB200_EXPERIMENTS

// Which BC6H launches take the SPLIT structure of decode_sorted_kernel: the LINEAR ones whose rows are narrower than a
// sector pair -- the 8-bit tone classes (16- and 12-byte rows) and FLOAT_RGB16 (24-byte rows) -- which otherwise stage
// their pixels.  Measured on B200, 4.19 M blocks, G blocks/s, staged (TMA store where it applies) -> SPLIT:
//   -> FLOAT_RGB16 17.8 -> 20.5, -> RGBX8 20.1 -> 22.7, -> BGRA8 19.9 -> 22.4, -> RGB8 20.2 -> 22.4, hdr -> RGBX8 18.4 -> 21.3.
// Everything else loses with it: native LINEAR 26.0 -> 22.4 (7 % more instructions, and with every warp of a CTA in the
// arithmetic half at the same time the integer pipe throttles more: issue slots 65 % -> 59 % busy), RGBX16 LINEAR
// 23.2 -> 20.6, TILED 26.1 -> 18.9 (32 sectors in 32 lines per store instruction, issued in bursts: mio_throttle +
// short_scoreboard 1.2 -> 5.3 warps per issue), signed native 22.3 -> 18.2.  DETEXB200_BC6H_SPLIT=1 / 0 in the
// environment forces it on / off for every BC6H launch (A/B runs, and the parity tests run under both).
template <int F, int PC>
static bool bc6h_split_selected(int layout) {
	static const int forced = [] { const char *e = getenv("DETEXB200_BC6H_SPLIT"); return e ? (e[0] == '1' ? 1 : 0) : -1; }();
	if (forced >= 0)
		return forced == 1;
	return layout == LAYOUT_LINEAR && F == F_BPTC_FLOAT && (is_tone_class(PC) || stages_packed_half3<F, PC>());
}

// One instantiation of the sorted kernel: opt in to > 48 KiB of dynamic shared memory once per device, then launch.
// Staged LINEAR tiles of 16-byte rows leave through the TMA unit when the image allows it (whole boxes of 64 blocks per
// block row, 16-byte aligned base); otherwise, and for the tiles that are repacked on the way out, the CTA copies them out.
template <int F, int LAYOUT, bool STAGE, int PC, int PAD = 32>
static void launch_sorted(const uint8_t *src, uint8_t *dst, uint32_t n_blocks, const DecodeArgs &a, cudaStream_t stream) {
	CUtensorMap tmap;
	memset(&tmap, 0, sizeof(tmap));
#ifdef DETEXB200_EXPERIMENTS
	if constexpr (!STAGE) {
		if (stream_kernel_enabled()) {
			launch_stream<F, LAYOUT, false, PC>(src, dst, n_blocks, a, stream, tmap);
			return;
		}
	}
#endif
	if constexpr (tma_capable<F, PC, STAGE, LAYOUT>()) {
		constexpr int ROW_BYTES = tma_row_bytes<F, PC>();
		// whole boxes of 64 blocks per block row (pixel rows beyond the image height are clipped by the hardware)
		if (tma_store_enabled() && a.wib % 64 == 0 && a.height > 4 * (a.hib - 1) &&
		    make_linear_tensor_map(&tmap, dst, (size_t)a.wib * ROW_BYTES, a.height, ROW_BYTES)) {
#ifdef DETEXB200_EXPERIMENTS
			if (stream_kernel_enabled()) {
				launch_stream<F, LAYOUT, true, PC>(src, dst, n_blocks, a, stream, tmap);
				return;
			}
#endif
			launch_sorted_variant<F, LAYOUT, STAGE, PC, PAD, true>(src, dst, n_blocks, a, stream, tmap);
			return;
		}
		memset(&tmap, 0, sizeof(tmap));
	}
	launch_sorted_variant<F, LAYOUT, STAGE, PC, PAD, false>(src, dst, n_blocks, a, stream, tmap);
}

#ifdef DETEXB200_EXPERIMENTS
// DETEXB200_BC7_KERNEL=flow selects the palette-staged dataflow kernel for BC7 instead of
// decode_sorted_kernel.  It is bit-exact (tests/test_gpu_parity.py runs the BC7 cases under it too)
// but measured slower (profiles/README.md, "BC7 palette / dataflow experiment"), so it is not the default.
static bool bc7_flow_kernel() {
	static const bool v = [] { const char *e = getenv("DETEXB200_BC7_KERNEL"); return e && e[0] == 'f'; }();
	return v;
}
template <int LAYOUT, int PC>
static void launch_bc7_flow(const uint8_t *src, uint8_t *dst, uint32_t n_blocks, const DecodeArgs &a, cudaStream_t stream) {
	static std::atomic<bool> configured[64];
	int dev = 0;
	cudaGetDevice(&dev);
	if (!configured[dev & 63].load()) {
		cudaFuncSetAttribute(decode_bc7_flow_kernel<LAYOUT, PC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FLOW_SMEM_BYTES);
		configured[dev & 63].store(true);
	}
	const uint64_t tiles = ((uint64_t)n_blocks + FLOW_TB - 1) / FLOW_TB;
	const uint64_t resident = (uint64_t)sm_count();		// persistent: one CTA of 1024 threads per SM
	const unsigned grid = (unsigned)(tiles < resident ? tiles : resident);
	decode_bc7_flow_kernel<LAYOUT, PC><<<grid, FLOW_NT, FLOW_SMEM_BYTES, stream>>>(src, dst, n_blocks,
		a.width, a.height, a.wib, a.mode_mask, a.flags, a.post_op, a.invalid, a.invalid_map);
}

#endif  // DETEXB200_EXPERIMENTS

template <int F, int PC>
static cudaError_t launch_format(const DecodeArgs &a, cudaStream_t stream) {
	using T = FormatTraits<F>;
	const uint64_t n_blocks = (uint64_t)a.wib * (uint64_t)a.hib;
	if (n_blocks == 0)
		return cudaSuccess;
	const uint8_t *src = (const uint8_t *)a.blocks;
	uint8_t *dst = (uint8_t *)a.pixels;
	const bool aligned_out = (((uintptr_t)dst & 15u) == 0);
	const bool aligned32 = (((uintptr_t)dst & 31u) == 0);
	if (PC != PC_NATIVE && !aligned32)
		return cudaErrorMisalignedAddress;	// the caller falls back to decode + conversion pass
	if constexpr (SortTraits<F>::ENABLED) {
		// multi-mode formats: mode-sorted kernel whenever there is enough work to sort
		// (LINEAR fast paths have no column clipping: the image must be exactly wib blocks wide)
		const bool linear_ok = a.layout == LAYOUT_TILED || (a.width == a.wib * 4);
		constexpr int TB = SortTraits<F>::TILE_BLOCKS;
		if (n_blocks >= (uint64_t)TB && aligned_out && linear_ok && !force_unsorted()) {
#ifdef DETEXB200_EXPERIMENTS
			if constexpr (F == F_BPTC && (PC == PC_NATIVE || PC == PC_PACK3)) {
				if (bc7_flow_kernel()) {
					if (a.layout == LAYOUT_TILED)
						launch_bc7_flow<LAYOUT_TILED, PC>(src, dst, (uint32_t)n_blocks, a, stream);
					else
						launch_bc7_flow<LAYOUT_LINEAR, PC>(src, dst, (uint32_t)n_blocks, a, stream);
					count_launch();
					return cudaGetLastError();
				}
			}
#endif
			if constexpr (F == F_BPTC_FLOAT || F == F_BPTC_SIGNED_FLOAT) {
				if (bc6h_split_selected<F, PC>(a.layout) && (a.layout == LAYOUT_LINEAR || PC != PC_NATIVE || (tiled_direct() && aligned32))) {
					CUtensorMap none;
					memset(&none, 0, sizeof(none));
					if (a.layout == LAYOUT_TILED)
						launch_sorted_variant<F, LAYOUT_TILED, false, PC, 1, false, true>(src, dst, (uint32_t)n_blocks, a, stream, none);
					else
						launch_sorted_variant<F, LAYOUT_LINEAR, false, PC, 1, false, true>(src, dst, (uint32_t)n_blocks, a, stream, none);
					count_launch();
					return cudaGetLastError();
				}
			}
			// LINEAR rows of 16 bytes or less per block are half sectors when written by lanes in sorted
			// order: those tiles (BC7, the toned BC6H classes) are staged and copied out in image order
			// (measured for the toned classes: staged 20.2 G blocks/s, direct 15.3)
			constexpr bool ST = SortTraits<F>::STAGE_OUTPUT;
			// 48-byte RGB8 tiles written by lanes in sorted order are partial sectors too: BC7 -> RGB8 TILED is
			// staged like its LINEAR twin
			constexpr bool STT = F == F_BPTC && PC == PC_PACK3;
			if (a.layout == LAYOUT_TILED && (PC != PC_NATIVE || (tiled_direct() && aligned32)))
				launch_sorted<F, LAYOUT_TILED, STT, PC, sort_pad<F, PC, STT>()>(src, dst, (uint32_t)n_blocks, a, stream);
			else if (a.layout == LAYOUT_TILED)
				launch_sorted<F, LAYOUT_TILED, ST, PC_NATIVE, sort_pad<F, PC_NATIVE, ST>()>(src, dst, (uint32_t)n_blocks, a, stream);
			else {
				// A/B: DETEXB200_BC7_LINEAR_DIRECT=1 stores BC7's 16-byte rows straight from the decoding lanes
				// (half sectors, merged in L2) instead of staging the tile
				static const bool bc7_direct = [] { const char *e = getenv("DETEXB200_BC7_LINEAR_DIRECT"); return e && e[0] == '1'; }();
				if constexpr (F == F_BPTC && PC == PC_NATIVE) {
					if (bc7_direct) {
						launch_sorted<F, LAYOUT_LINEAR, false, PC, 32>(src, dst, (uint32_t)n_blocks, a, stream);
						count_launch();
						return cudaGetLastError();
					}
				}
				constexpr bool STL = ST || is_tone_class(PC) || stages_packed_half3<F, PC>();
				static const bool stage_half3 = [] { const char *e = getenv("DETEXB200_HALF3_STAGE"); return !(e && e[0] == '0'); }();
				if constexpr (stages_packed_half3<F, PC>()) {
					if (!stage_half3) {	// A/B: direct 24-byte row stores (partial sectors)
						launch_sorted<F, LAYOUT_LINEAR, false, PC, 1>(src, dst, (uint32_t)n_blocks, a, stream);
						count_launch();
						return cudaGetLastError();
					}
				}
				launch_sorted<F, LAYOUT_LINEAR, STL, PC, sort_pad<F, PC, STL>()>(src, dst, (uint32_t)n_blocks, a, stream);
			}
			count_launch();
			return cudaGetLastError();
		}
	}
	if (a.layout == LAYOUT_TILED) {
		const uint64_t tiles = (n_blocks + 31) / 32;
		// enough CTAs to cover the work once, capped at a few waves of resident CTAs
		const uint64_t per_cta = (uint64_t)CTA_WARPS * (DeferTraits<F>::ENABLED ? DEFER_ITERS : TiledTune<F>::ITERS);
		const uint64_t want = (tiles + per_cta - 1) / per_cta;
		const uint64_t cap = (uint64_t)sm_count() * 8 * 16;
		const unsigned grid = (unsigned)(want < cap ? want : cap);
		constexpr bool CAN_DIRECT = post_words<T::OUT_WORDS, PC>() >= 8;
		bool launched = false;
		if constexpr (PC != PC_NATIVE || CAN_DIRECT) {
			if (PC != PC_NATIVE || (aligned32 && tiled_direct())) {
				decode_tiled_kernel<F, true, PC><<<grid, CTA_THREADS, 0, stream>>>(src, dst, (uint32_t)n_blocks,
					a.mode_mask, a.flags, a.post_op, a.invalid, a.invalid_map, tone_args_of(a));
				launched = true;
			}
		}
		if constexpr (PC == PC_NATIVE) {
			if (!launched)
				decode_tiled_kernel<F, false, PC_NATIVE><<<grid, CTA_THREADS, 0, stream>>>(src, dst, (uint32_t)n_blocks,
					a.mode_mask, a.flags, a.post_op, a.invalid, a.invalid_map, tone_args_of(a));
		}
	}
	else {
		// store_linear_rows clips rows only: the fast path needs the image to be exactly wib blocks wide
		// (a narrower image -- width <= (wib - 1) * 4 is a legal detexTexture -- takes the clipped kernel)
		const bool fast = (a.width == a.wib * 4) && aligned_out;
		if (fast) {
			int tx = 32;
			while (tx < CTA_THREADS && tx < a.wib)
				tx *= 2;
			const int ty = CTA_THREADS / tx;
			dim3 block(tx, ty);
			const int rows_per_cta = ty * LinearTune<F>::ITERS;
			unsigned gy = (unsigned)((a.hib + rows_per_cta - 1) / rows_per_cta);
			if (gy > 65535u)
				gy = 65535u;
			dim3 grid((unsigned)((a.wib + tx - 1) / tx), gy);
			decode_linear_kernel<F, PC><<<grid, block, 0, stream>>>(src, dst, a.width, a.height, a.wib, a.hib,
				a.mode_mask, a.flags, a.post_op, a.invalid, a.invalid_map, tone_args_of(a));
		}
		else {
			const uint64_t want = (n_blocks + CTA_THREADS - 1) / CTA_THREADS;
			const uint64_t cap = (uint64_t)sm_count() * 8 * 16;
			const unsigned grid = (unsigned)(want < cap ? want : cap);
			decode_linear_clipped_kernel<F, PC><<<grid, CTA_THREADS, 0, stream>>>(src, dst, a.width, a.height,
				a.wib, a.hib, a.mode_mask, a.flags, a.post_op, a.invalid, a.invalid_map, tone_args_of(a));
		}
	}
	count_launch();
	return cudaGetLastError();
}

// post class -> instantiation, for the (format, class) pairs that exist
template <int F>
static cudaError_t launch_classes(const DecodeArgs &a, cudaStream_t stream) {
	if (a.post_class == PC_NATIVE)
		return launch_format<F, PC_NATIVE>(a, stream);
	if constexpr (post_class_valid<F, PC_PACK3>()) {
		if (a.post_class == PC_PACK3)
			return launch_format<F, PC_PACK3>(a, stream);
	}
	if constexpr (post_class_valid<F, PC_NARROW>()) {
		if (a.post_class == PC_NARROW)
			return launch_format<F, PC_NARROW>(a, stream);
	}
	if constexpr (post_class_valid<F, PC_TONE_FIRST>()) {
		switch (a.post_class) {
#define DETEX_TONE_CASE(K) case PC_TONE_FIRST + K: return launch_format<F, PC_TONE_FIRST + K>(a, stream);
		DETEX_TONE_CASE(0) DETEX_TONE_CASE(1) DETEX_TONE_CASE(2) DETEX_TONE_CASE(3) DETEX_TONE_CASE(4)
		DETEX_TONE_CASE(5) DETEX_TONE_CASE(6) DETEX_TONE_CASE(7) DETEX_TONE_CASE(8)
#undef DETEX_TONE_CASE
		default: break;
		}
	}
	return cudaErrorInvalidValue;
}

cudaError_t launch_decode(int format_index, const DecodeArgs &a, cudaStream_t stream) {
	switch (format_index) {
#define DETEX_CASE(F) case F: return launch_classes<F>(a, stream);
	DETEX_CASE(F_BC1) DETEX_CASE(F_BC1A) DETEX_CASE(F_BC2) DETEX_CASE(F_BC3)
	DETEX_CASE(F_RGTC1) DETEX_CASE(F_SIGNED_RGTC1) DETEX_CASE(F_RGTC2) DETEX_CASE(F_SIGNED_RGTC2)
	DETEX_CASE(F_BPTC_FLOAT) DETEX_CASE(F_BPTC_SIGNED_FLOAT) DETEX_CASE(F_BPTC)
	DETEX_CASE(F_ETC1) DETEX_CASE(F_ETC2) DETEX_CASE(F_ETC2_PUNCHTHROUGH) DETEX_CASE(F_ETC2_EAC)
	DETEX_CASE(F_EAC_R11) DETEX_CASE(F_EAC_SIGNED_R11) DETEX_CASE(F_EAC_RG11) DETEX_CASE(F_EAC_SIGNED_RG11)
#undef DETEX_CASE
	default: return cudaErrorInvalidValue;
	}
}

}  // namespace detexb200
