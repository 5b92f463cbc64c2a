// Shared device helpers for the sm_100a block decoders.
//
// Conventions used by every decoder in this directory:
//   * one compressed block is handed to the decoder in registers (uint2 for the
//     8-byte formats, uint4 for the 16-byte formats), little-endian words exactly
//     as they sit in memory;
//   * the 16 decoded pixels come back in registers as OUT_WORDS 32-bit words in
//     the reference's row-major tile order (pixel y*4+x, detex.h:435-531):
//     4 B/px formats -> 16 words, 8 B/px -> 32 words (lo, hi per pixel),
//     2 B/px -> 8 words (two pixels per word), 1 B/px -> 4 words (one row each);
//   * the decoder returns false for an invalid / masked-out block; the caller
//     zero-fills the tile (Q16, reference texture.c:90-93,125-128).
#pragma once
#include <cstdint>
#ifdef DETEXB200_HOST_EMU
// tests/emu/host_emu.h: the decoders' register arithmetic compiled for the CPU so that tests can
// compare it with the oracle without a GPU (test infrastructure only; never part of the library)
#include "host_emu.h"
#else
#include <cuda_runtime.h>
#endif

namespace detexb200 {

// ---- format indices: texture_format >> 24, reference detex.h:575-612 ----------
enum FormatIndex : int {
	F_BC1 = 1, F_BC1A, F_BC2, F_BC3, F_RGTC1, F_SIGNED_RGTC1, F_RGTC2, F_SIGNED_RGTC2,
	F_BPTC_FLOAT, F_BPTC_SIGNED_FLOAT, F_BPTC, F_ETC1, F_ETC2, F_ETC2_PUNCHTHROUGH,
	F_ETC2_EAC, F_EAC_R11, F_EAC_SIGNED_R11, F_EAC_RG11, F_EAC_SIGNED_RG11,
	F_COUNT
};

// Decompression flags (reference detex.h:398-411).
constexpr uint32_t FLAG_ENCODE = 1, FLAG_OPAQUE_ONLY = 2, FLAG_NON_OPAQUE_ONLY = 4;

#ifndef DETEXB200_HOST_EMU
// ---- streaming global memory access --------------------------------------------
// Inputs are read exactly once: non-coherent path, do not allocate in L1.
__device__ __forceinline__ uint2 ld_stream_v2(const void *p) {
	uint2 r;
	asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];"
		: "=r"(r.x), "=r"(r.y) : "l"(p));
	return r;
}
__device__ __forceinline__ uint4 ld_stream_v4(const void *p) {
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
		: "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
	return r;
}
// Outputs are written once and never re-read by the kernel: streaming stores.
__device__ __forceinline__ void st_stream_v4(void *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
	asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};"
		:: "l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// 256-bit streaming store (sm_100: STG.E.ENL2.256): one full 32-byte sector per lane.
// The address must be 32-byte aligned.
__device__ __forceinline__ void st_stream_v8(void *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d,
uint32_t e, uint32_t f, uint32_t g, uint32_t h) {
	asm volatile("st.global.cs.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
		:: "l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f), "r"(g), "r"(h) : "memory");
}
__device__ __forceinline__ void st_stream_v2(void *p, uint32_t a, uint32_t b) {
	asm volatile("st.global.cs.v2.u32 [%0], {%1, %2};" :: "l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void st_stream_u32(void *p, uint32_t a) {
	asm volatile("st.global.cs.u32 [%0], %1;" :: "l"(p), "r"(a) : "memory");
}

#endif  // DETEXB200_HOST_EMU

// ---- bit helpers ---------------------------------------------------------------
__device__ __forceinline__ int bfe_s(uint32_t v, int pos, int len) {	// sign-extending
	return ((int)(v << (32 - pos - len))) >> (32 - len);
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
	return __byte_perm(a, b, sel);
}
__device__ __forceinline__ uint32_t bswap32(uint32_t v) {
	return __byte_perm(v, 0, 0x0123);
}

// n-bit field at compile-time position POS of a 128-bit little-endian stream held
// in four words (reference bits.c:22-45 does this one bit at a time).
template <int POS, int LEN>
__device__ __forceinline__ uint32_t field128(const uint4 &w) {
	static_assert(LEN >= 0 && LEN <= 32 && POS >= 0 && POS + LEN <= 128, "field out of range");
	if constexpr (LEN == 0)
		return 0u;
	constexpr int W = POS >> 5, S = POS & 31;
	const uint32_t lo = W == 0 ? w.x : (W == 1 ? w.y : (W == 2 ? w.z : w.w));
	uint32_t v;
	if constexpr (S + LEN <= 32)
		v = lo >> S;
	else {
		const uint32_t hi = W == 0 ? w.y : (W == 1 ? w.z : w.w);
		v = __funnelshift_r(lo, hi, S);
	}
	if constexpr (LEN < 32)
		v &= (1u << LEN) - 1u;
	return v;
}

// 64-bit window starting at compile-time position POS (bits beyond 127 read as 0).
template <int POS>
__device__ __forceinline__ uint2 window64(const uint4 &w) {
	constexpr int W = POS >> 5, S = POS & 31;
	const uint32_t a[6] = { w.x, w.y, w.z, w.w, 0u, 0u };
	uint2 r;
	if constexpr (S == 0) {
		r.x = a[W];
		r.y = a[W + 1];
	}
	else {
		r.x = __funnelshift_r(a[W], a[W + 1], S);
		r.y = __funnelshift_r(a[W + 1], a[W + 2], S);
	}
	return r;
}

}  // namespace detexb200
