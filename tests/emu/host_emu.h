/*
 * TEST INFRASTRUCTURE ONLY.  Lets the register-level arithmetic of the CUDA decoders
 * (detex_b200/csrc/decode_*.cuh) compile for the CPU with a plain C++ compiler: the CUDA
 * keywords become no-ops and the handful of intrinsics the decoders use are restated with
 * their PTX semantics.  tests/test_emu_bptc.py runs the result against the oracle, so that a
 * change to a decoder can be checked bit-for-bit in a container without a GPU.  Nothing here
 * is part of, or reachable from, libdetex_b200.so.
 */
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
#include <cfenv>
#include <algorithm>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __restrict__

struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{ x, y, z, w }; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{ x, y }; }

/* prmt.b32 (default mode): result byte i = byte (sel nibble i & 7) of {b:a}; nibble bit 3
   replicates that byte's sign bit */
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
	const uint64_t v = ((uint64_t)b << 32) | a;
	uint32_t r = 0;
	for (int i = 0; i < 4; i++) {
		const uint32_t n = (sel >> (4 * i)) & 0xFu;
		uint32_t byte = (uint32_t)(v >> (8 * (n & 7u))) & 0xFFu;
		if (n & 8u)
			byte = (byte & 0x80u) ? 0xFFu : 0u;
		r |= byte << (8 * i);
	}
	return r;
}
/* shf.r.wrap: low word of {hi:lo} >> (shift & 31) */
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t shift) {
	const uint64_t v = ((uint64_t)hi << 32) | lo;
	return (uint32_t)(v >> (shift & 31u));
}
static inline int __ffs(int v) { return v == 0 ? 0 : __builtin_ctz((unsigned)v) + 1; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
using std::min;
using std::max;
/* single-thread stand-ins for the launch geometry (table loaders loop over them) */
struct EmuDim3 { unsigned x, y, z; };
static const EmuDim3 threadIdx = { 0, 0, 0 }, blockDim = { 1, 1, 1 }, blockIdx = { 0, 0, 0 }, gridDim = { 1, 1, 1 };
static inline uint32_t __brev(uint32_t v) {
	uint32_t r = 0;
	for (int i = 0; i < 32; i++)
		r |= ((v >> i) & 1u) << (31 - i);
	return r;
}
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline int __mulhi(int a, int b) { return (int)(((int64_t)a * b) >> 32); }

/* float intrinsics with an explicit rounding mode (tone.cuh): evaluated through volatile operands
   under fesetround so that the host compiler can neither fold nor reorder them */
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
template <typename Op> static inline float emu_rounded(int mode, Op op) {
	const int old = fegetround();
	fesetround(mode);
	volatile float r = op();
	fesetround(old);
	return r;
}
static inline float __fmul_rn(float a, float b) { volatile float x = a, y = b; return emu_rounded(FE_TONEAREST, [&] { return x * y; }); }
static inline float __fmul_rd(float a, float b) { volatile float x = a, y = b; return emu_rounded(FE_DOWNWARD, [&] { return x * y; }); }
static inline float __fadd_rd(float a, float b) { volatile float x = a, y = b; return emu_rounded(FE_DOWNWARD, [&] { return x + y; }); }
static inline float __fsub_rn(float a, float b) { volatile float x = a, y = b; return emu_rounded(FE_TONEAREST, [&] { return x - y; }); }
static inline float __fsub_rd(float a, float b) { volatile float x = a, y = b; return emu_rounded(FE_DOWNWARD, [&] { return x - y; }); }
static inline float __fdiv_rd(float a, float b) { volatile float x = a, y = b; return emu_rounded(FE_DOWNWARD, [&] { return x / y; }); }
static inline float __fmaf_rd(float a, float b, float c) { volatile float x = a, y = b, z = c; return emu_rounded(FE_DOWNWARD, [&] { return fmaf(x, y, z); }); }
static inline float __fadd_rn(float a, float b) { volatile float x = a, y = b; return emu_rounded(FE_TONEAREST, [&] { return x + y; }); }
static inline float __fmaf_rn(float a, float b, float c) { volatile float x = a, y = b, z = c; return emu_rounded(FE_TONEAREST, [&] { return fmaf(x, y, z); }); }
static inline float __int2float_rn(int v) { return (float)v; }
/* the .sat modifier of mul.f32: clamp to [0, 1], NaN -> +0 */
static inline float emu_sat(float f) { return f != f ? 0.0f : (f < 0.0f ? 0.0f : (f > 1.0f ? 1.0f : f)); }

/* ---- S3TC / ETC / EAC decoders (decode_s3tc.cuh, decode_etc.cuh) ---- */
#define __align__(n) alignas(n)
template <typename T> static inline T __ldg(const T *p) { return *p; }
/* per-halfword add / subtract, wrapping (vadd2 / vsub2) */
static inline uint32_t __vadd2(uint32_t a, uint32_t b) {
	return ((a + b) & 0xFFFFu) | ((((a >> 16) + (b >> 16)) & 0xFFFFu) << 16);
}
static inline uint32_t __vsub2(uint32_t a, uint32_t b) {
	return ((a - b) & 0xFFFFu) | ((((a >> 16) - (b >> 16)) & 0xFFFFu) << 16);
}
/* per signed halfword max(min(a + b, c), 0) and max(min(a, b), 0) (VIADDMNMX.S16x2.RELU / VIMNMX.S16x2.RELU) */
static inline uint32_t __viaddmin_s16x2_relu(uint32_t a, uint32_t b, uint32_t c) {
	uint32_t r = 0;
	for (int k = 0; k < 2; k++) {
		const int s = (int16_t)(uint16_t)((a >> (16 * k)) + (b >> (16 * k)));
		const int lim = (int16_t)(uint16_t)(c >> (16 * k));
		r |= (uint32_t)(uint16_t)std::max(std::min(s, lim), 0) << (16 * k);
	}
	return r;
}
static inline uint32_t __vimin_s16x2_relu(uint32_t a, uint32_t b) {
	uint32_t r = 0;
	for (int k = 0; k < 2; k++) {
		const int x = (int16_t)(uint16_t)(a >> (16 * k)), y = (int16_t)(uint16_t)(b >> (16 * k));
		r |= (uint32_t)(uint16_t)std::max(std::min(x, y), 0) << (16 * k);
	}
	return r;
}

/* ---- pixel-format conversion steps (convert_steps.cuh) ---- */
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline int __float2int_rd(float f) { return (int)floorf(f); }
