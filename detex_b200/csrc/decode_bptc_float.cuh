// BPTC_FLOAT / BPTC_SIGNED_FLOAT (BC6H) block decoder.  Behaviour follows reference
// decompress-bptc-float.c:23-644.  The output is the half-float BIT PATTERN built
// with integer arithmetic only (no floating point anywhere), so the result is
// bit-exact, not merely within 1 ULP.
//
// Design: the 14 irregular per-mode bit scatters are generated at compile time from
// the layout strings of the format specification (constexpr parser below; the
// reference spells the same layouts out as 14 hand-written switch cases,
// decompress-bptc-float.c:128-485).  Only that scatter is mode-specific; delta
// decoding, unquantisation, index expansion and interpolation are shared code with
// the mode's constants passed as values, so a warp with mixed modes re-converges
// after ~1/4 of the work.
#pragma once
#include "common.cuh"
#include "decode_bptc.cuh"

namespace detexb200 {

// ---- compile-time layout description ------------------------------------------------
// "x[a:b]": component bit b sits at the lowest stream position, walking towards a;
// a < b therefore denotes a bit-reversed field (modes 12 and 13).  m = mode bits,
// d = partition.  r/g/b + endpoint number 0..3.
struct Bc6hPiece {
	int comp;	// 0 r, 1 g, 2 b, 3 partition, 4 mode (skipped)
	int ep;		// endpoint 0..3
	int lo;		// lowest component bit written
	int len;
	int reversed;	// 1: ascending stream position = descending component bit
	int pos;	// stream position of the field
};
struct Bc6hLayout {
	Bc6hPiece piece[28];
	int n;
	int end;	// stream position after the last field (= start of the indices)
};

__host__ __device__ constexpr int parse_int(const char *s, int &i) {
	int v = 0;
	while (s[i] >= '0' && s[i] <= '9')
		v = v * 10 + (s[i++] - '0');
	return v;
}

__host__ __device__ constexpr Bc6hLayout bc6h_parse_layout(const char *s) {
	Bc6hLayout L{};
	int i = 0, pos = 0;
	while (s[i]) {
		if (s[i] == ' ') {
			i++;
			continue;
		}
		Bc6hPiece p{};
		const char c = s[i++];
		p.comp = c == 'r' ? 0 : (c == 'g' ? 1 : (c == 'b' ? 2 : (c == 'd' ? 3 : 4)));
		if (p.comp <= 2)
			p.ep = s[i++] - '0';
		i++;				// '['
		const int a = parse_int(s, i);
		int b = a;
		if (s[i] == ':') {
			i++;
			b = parse_int(s, i);
		}
		i++;				// ']'
		p.reversed = a < b ? 1 : 0;
		p.lo = a < b ? a : b;
		p.len = (a < b ? b - a : a - b) + 1;
		p.pos = pos;
		pos += p.len;
		L.piece[L.n++] = p;
	}
	L.end = pos;
	return L;
}

__host__ __device__ constexpr const char *bc6h_layout_string(int mode) {
	switch (mode) {
	case 0: return "m[1:0] g2[4] b2[4] b3[4] r0[9:0] g0[9:0] b0[9:0] r1[4:0] g3[4] g2[3:0] g1[4:0] b3[0] g3[3:0] b1[4:0] b3[1] b2[3:0] r2[4:0] b3[2] r3[4:0] b3[3] d[4:0]";
	case 1: return "m[1:0] g2[5] g3[4] g3[5] r0[6:0] b3[0] b3[1] b2[4] g0[6:0] b2[5] b3[2] g2[4] b0[6:0] b3[3] b3[5] b3[4] r1[5:0] g2[3:0] g1[5:0] g3[3:0] b1[5:0] b2[3:0] r2[5:0] r3[5:0] d[4:0]";
	case 2: return "m[4:0] r0[9:0] g0[9:0] b0[9:0] r1[4:0] r0[10] g2[3:0] g1[3:0] g0[10] b3[0] g3[3:0] b1[3:0] b0[10] b3[1] b2[3:0] r2[4:0] b3[2] r3[4:0] b3[3] d[4:0]";
	case 3: return "m[4:0] r0[9:0] g0[9:0] b0[9:0] r1[3:0] r0[10] g3[4] g2[3:0] g1[4:0] g0[10] g3[3:0] b1[3:0] b0[10] b3[1] b2[3:0] r2[3:0] b3[0] b3[2] r3[3:0] g2[4] b3[3] d[4:0]";
	case 4: return "m[4:0] r0[9:0] g0[9:0] b0[9:0] r1[3:0] r0[10] b2[4] g2[3:0] g1[3:0] g0[10] b3[0] g3[3:0] b1[4:0] b0[10] b2[3:0] r2[3:0] b3[1] b3[2] r3[3:0] b3[4] b3[3] d[4:0]";
	case 5: return "m[4:0] r0[8:0] b2[4] g0[8:0] g2[4] b0[8:0] b3[4] r1[4:0] g3[4] g2[3:0] g1[4:0] b3[0] g3[3:0] b1[4:0] b3[1] b2[3:0] r2[4:0] b3[2] r3[4:0] b3[3] d[4:0]";
	case 6: return "m[4:0] r0[7:0] g3[4] b2[4] g0[7:0] b3[2] g2[4] b0[7:0] b3[3] b3[4] r1[5:0] g2[3:0] g1[4:0] b3[0] g3[3:0] b1[4:0] b3[1] b2[3:0] r2[5:0] r3[5:0] d[4:0]";
	case 7: return "m[4:0] r0[7:0] b3[0] b2[4] g0[7:0] g2[5] g2[4] b0[7:0] g3[5] b3[4] r1[4:0] g3[4] g2[3:0] g1[5:0] g3[3:0] b1[4:0] b3[1] b2[3:0] r2[4:0] b3[2] r3[4:0] b3[3] d[4:0]";
	case 8: return "m[4:0] r0[7:0] b3[1] b2[4] g0[7:0] b2[5] g2[4] b0[7:0] b3[5] b3[4] r1[4:0] g3[4] g2[3:0] g1[4:0] b3[0] g3[3:0] b1[5:0] b2[3:0] r2[4:0] b3[2] r3[4:0] b3[3] d[4:0]";
	case 9: return "m[4:0] r0[5:0] g3[4] b3[0] b3[1] b2[4] g0[5:0] g2[5] b2[5] b3[2] g2[4] b0[5:0] g3[5] b3[3] b3[5] b3[4] r1[5:0] g2[3:0] g1[5:0] g3[3:0] b1[5:0] b2[3:0] r2[5:0] r3[5:0] d[4:0]";
	case 10: return "m[4:0] r0[9:0] g0[9:0] b0[9:0] r1[9:0] g1[9:0] b1[9:0]";
	case 11: return "m[4:0] r0[9:0] g0[9:0] b0[9:0] r1[8:0] r0[10] g1[8:0] g0[10] b1[8:0] b0[10]";
	case 12: return "m[4:0] r0[9:0] g0[9:0] b0[9:0] r1[7:0] r0[10:11] g1[7:0] g0[10:11] b1[7:0] b0[10:11]";
	default: return "m[4:0] r0[9:0] g0[9:0] b0[9:0] r1[3:0] r0[10:15] g1[3:0] g0[10:15] b1[3:0] b0[10:15]";
	}
}

// Endpoint precision and per-component delta widths by detex mode number
// (decompress-bptc-float.c:42-43 and the tail of each switch case); modes 9 and 10
// store all endpoints directly (delta width 0 here).
struct Bc6hModeInfo {
	int epb, dr, dg, db;
};
__host__ __device__ constexpr Bc6hModeInfo bc6h_mode_info(int m) {
	constexpr Bc6hModeInfo t[14] = {
		{ 10, 5, 5, 5 }, { 7, 6, 6, 6 }, { 11, 5, 4, 4 }, { 11, 4, 5, 4 }, { 11, 4, 4, 5 },
		{ 9, 5, 5, 5 }, { 8, 6, 5, 5 }, { 8, 5, 6, 5 }, { 8, 5, 5, 6 }, { 6, 0, 0, 0 },
		{ 10, 0, 0, 0 }, { 11, 9, 9, 9 }, { 12, 8, 8, 8 }, { 16, 4, 4, 4 },
	};
	return t[m];
}

// Raw endpoint fields of one block: ep[e][c], c = r, g, b.
struct Bc6hRaw {
	uint32_t ep[4][3];
	uint32_t partition;
};

template <int M>
__device__ __forceinline__ void bc6h_scatter(const uint4 &w, Bc6hRaw &raw) {
	constexpr Bc6hLayout L = bc6h_parse_layout(bc6h_layout_string(M));
	static_assert(L.end == (M >= 10 ? 65 : 82), "layout must end where the indices start");
#pragma unroll
	for (int e = 0; e < 4; e++)
#pragma unroll
		for (int c = 0; c < 3; c++)
			raw.ep[e][c] = 0u;
	raw.partition = 0u;
	static_for<0, L.n>([&](auto k_) {
		constexpr int k = decltype(k_)::value;
		constexpr Bc6hPiece P = L.piece[k];
		if constexpr (P.comp != 4) {
			uint32_t v = field128<P.pos, P.len>(w);
			if constexpr (P.reversed)
				v = __brev(v) >> (32 - P.len);
			// Q9: mode 12 reads b0[11] (stream bit 63) through detexGetBits64(data0, 63, 63),
			// whose mask is the undefined 1<<64 (bits.h:31); the reference build (gcc 13.3,
			// -Ofast) yields 0 for it (decompress-bptc-float.c:462).
			if constexpr (M == 12 && P.comp == 2 && P.ep == 0 && P.reversed)
				v &= 1u;		// keep b0[10], drop b0[11]
			if constexpr (P.comp == 3)
				raw.partition |= v << P.lo;
			else
				raw.ep[P.ep][P.comp] |= v << P.lo;
		}
	});
}

__device__ __forceinline__ int sext_runtime(uint32_t v, int bits) {
	return ((int)(v << (32 - bits))) >> (32 - bits);
}

// Shared tail: delta decode, unquantise, expand indices, interpolate, emit.
// NS = 2 for detex modes 0-9 (3-bit indices from stream bit 82), 1 for modes 10-13
// (4-bit indices from bit 65).
// What becomes of a pixel: sink(i, v_r, v_g, v_b) receives the three interpolated 16-bit values of
// pixel i (InterpolateFloat's result, before the final scaling to a half-float pattern).  The native
// sink below finishes them into FLOAT_RGBX16 / SIGNED_FLOAT_RGBX16 words; tone.cuh has a sink that
// goes on to 8-bit pixels.
// Two ways to hand a pixel to a sink:
//   * integer (FLOAT_PIPE = false): sink(i, v_r, v_g, v_b) with v = InterpolateFloat's 16-bit result;
//   * float (FLOAT_PIPE = true): sink.emit(i, y_r, y_g, y_b) with y = BIAS + v held in a float whose unit in the last
//     place is 1 (BIAS = 2^23 unsigned, 1.5 * 2^23 signed), i.e. the integer v sits in the mantissa field.  The
//     interpolation and the finish then run on the FP32 multiply-add pipe, which this integer-bound decoder leaves
//     idle: the integer ALU pipe issues a warp instruction every other cycle, the FMA pipe every cycle (ncu, round 2:
//     ALU pipe 72-77 % busy, FMA 18-23 %).  Every step below is exact -- all operands are integers or integers / 64
//     below 2^24, sums and fused products included -- so the floors are the only roundings and they are asked for
//     explicitly (round-down adds against the bias): bit-exact with the integer statement, which
//     tests/test_emu_decoders.py checks against the oracle on the CPU.
__device__ __forceinline__ float sat01(float f) {
#ifdef DETEXB200_HOST_EMU
	return emu_sat(f);
#else
	return __saturatef(f);
#endif
}

template <bool SIGNED>
struct Bc6hHalfSink {
	// Measured (B200, 4.19 M blocks, G blocks/s LINEAR U / M, integer -> float pipe): signed 19.0 / 18.0 -> 22.4 / 21.1
	// (the sign-magnitude finish is six integer-pipe instructions per component and three on the FMA pipe, 162 M ->
	// 146 M warp instructions); unsigned 25.1 / 25.1 -> 23.8 / 24.3 (115 M -> 127 M instructions: the kernel is
	// issue-bound while it decodes, so the better pipe balance does not pay for the extra instructions).  Hence:
#ifdef DETEXB200_BC6H_FLOAT_PIPE_ALL
	static constexpr bool FLOAT_PIPE = true;
#else
	static constexpr bool FLOAT_PIPE = SIGNED;
#endif
	uint32_t (&px)[32];
	__device__ __forceinline__ void operator()(int i, int v0, int v1, int v2) const {
		const int v[3] = { v0, v1, v2 };
		uint32_t h[3];
#pragma unroll
		for (int c = 0; c < 3; c++) {
			if (SIGNED) {					// :577-613
				const int mag = ((v[c] < 0 ? -v[c] : v[c]) * 31) >> 5;
				h[c] = (uint32_t)mag | ((v[c] < 0 && mag != 0) ? 0x8000u : 0u);
			}
			else
				h[c] = (uint32_t)((v[c] * 31) >> 6);	// :616-621 (v >= 0)
		}
		px[2 * i] = (h[0] & 0xFFFFu) | (h[1] << 16);		// Q10: R | G<<16 | B<<32, X = 0
		px[2 * i + 1] = h[2] & 0xFFFFu;
	}
	// y = BIAS + v  ->  a float whose low 16 mantissa bits are the half-float pattern
	static __device__ __forceinline__ uint32_t finish(float y) {
		if (SIGNED) {
			// sign-magnitude (:577-613): mag = (|v| * 31) >> 5, sign bit when v < 0 and mag != 0, i.e. v <= -2
			const float v = __fsub_rn(y, 12582912.0f);
			const float z = __fmaf_rd(fabsf(v), 0.96875f, 8388608.0f);		// 2^23 + floor(|v| * 31 / 32)
			const float neg = sat01(__fsub_rn(-1.0f, v));				// 1 for v <= -2, 0 for v >= -1
			return __float_as_uint(__fmaf_rn(neg, 32768.0f, z));
		}
		// (v * 31) >> 6 (:616-621): y * 31/64 - 2^23 * 31/64 = v * 31 / 64 exactly, then the floor
		return __float_as_uint(__fadd_rd(__fmaf_rn(y, 0.484375f, -4063232.0f), 8388608.0f));
	}
	__device__ __forceinline__ void emit(int i, float y0, float y1, float y2) const {
		px[2 * i] = prmt(finish(y0), finish(y1), 0x5410);	// Q10: R | G<<16 | B<<32, X = 0
		px[2 * i + 1] = finish(y2) & 0xFFFFu;
	}
	// unsigned, integer pipe: t = 4 * (interpolation sum), i.e. the 16-bit value v = (sum >> 6) is bytes 1-2 of t;
	// (v * 31) >> 6 = (v * 124) >> 8 = bytes 1-2 of v * 124 (< 2^23)
#ifdef DETEXB200_BC6H_NO_SCALED
	static constexpr bool TAKES_SCALED = false;
#else
	static constexpr bool TAKES_SCALED = !SIGNED;
#endif
	__device__ __forceinline__ void emit_scaled(int i, uint32_t t0, uint32_t t1, uint32_t t2) const {
		const uint32_t p0 = prmt(t0, 0u, 0x4421) * 124u, p1 = prmt(t1, 0u, 0x4421) * 124u, p2 = prmt(t2, 0u, 0x4421) * 124u;
		px[2 * i] = prmt(p0, p1, 0x6521);			// Q10: R | G<<16 | B<<32, X = 0
		px[2 * i + 1] = prmt(p2, 0u, 0x4421);
	}
};

// Delta decode and unquantisation of the endpoints: e[c][k] = unquantised value of endpoint k, component c
// (16 bits: 0 .. 0xFFFF, signed formats -0x8000 .. 0x7FFF).
template <bool SIGNED, int NS>
__device__ __forceinline__ void bc6h_unquantized(const Bc6hRaw &raw, int epb, int dr, int dg, int db, int (&out)[3][2 * NS]) {
	constexpr int NE = 2 * NS;
	const uint32_t epmask = epb >= 32 ? 0xFFFFFFFFu : ((1u << epb) - 1u);
	const int delta_bits[3] = { dr, dg, db };
#pragma unroll
	for (int c = 0; c < 3; c++) {
		int e[NE];
		e[0] = SIGNED ? sext_runtime(raw.ep[0][c], epb) : (int)raw.ep[0][c];	// :491-495
#pragma unroll
		for (int k = 1; k < NE; k++) {
			uint32_t v = raw.ep[k][c];
			if (delta_bits[c] != 0)		// transformed endpoints (:496-511)
				v = ((uint32_t)e[0] + (uint32_t)sext_runtime(v, delta_bits[c])) & epmask;
			e[k] = SIGNED ? sext_runtime(v, epb) : (int)v;
		}
#pragma unroll
		for (int k = 0; k < NE; k++) {
			int x = e[k], unq;
			if (SIGNED) {			// UnquantizeSigned, :64-86
				const bool neg = x < 0;
				const int ax = neg ? -x : x;
				if (epb >= 16)
					unq = x;
				else {
					// ((ax << 15) + 0x4000) >> (epb - 1) == (2*ax + 1) << (15 - epb) for epb <= 15
					unq = (2 * ax + 1) << (15 - epb);
					if (ax == 0)
						unq = 0;
					if (ax >= (1 << (epb - 1)) - 1)
						unq = 0x7FFF;
					if (neg)
						unq = -unq;
				}
			}
			else {				// Unquantize, :52-62 (mode 13 <=> epb 16)
				// ((x << 15) + 0x4000) >> (epb - 1) == (2x + 1) << (15 - epb) for epb <= 15
				unq = (int)((2u * (uint32_t)x + 1u) << ((15 - epb) & 31));
				if (x == 0)
					unq = 0;
				if ((uint32_t)x == epmask)
					unq = 0xFFFF;
				if (epb >= 16)
					unq = x;
			}
			out[c][k] = unq;
		}
	}
}

// Interpolation constants of a block from its unquantised endpoints: A = 64*e0 + 32, K = e1 - e0 per subset and
// component (integer pipe), or the same divided by 64 as floats (float pipe).
template <int NS, bool FP>
struct Bc6hLines {
	int a_c[NS][3], k_c[NS][3];
	float a_f[NS][3], k_f[NS][3];
	__device__ __forceinline__ void set(const int (&e)[3][2 * NS]) {
#pragma unroll
		for (int c = 0; c < 3; c++)
#pragma unroll
			for (int s = 0; s < NS; s++) {
				if constexpr (FP) {
					const float e0 = __int2float_rn(e[c][2 * s]), e1 = __int2float_rn(e[c][2 * s + 1]);	// exact: |e| < 2^17
					a_f[s][c] = __fadd_rn(e0, 0.5f);				// (64 e0 + 32) / 64
					k_f[s][c] = __fmul_rn(__fsub_rn(e1, e0), 0.015625f);		// (e1 - e0) / 64
				}
				else {
					a_c[s][c] = 64 * e[c][2 * s] + 32;
					k_c[s][c] = e[c][2 * s + 1] - e[c][2 * s];
				}
			}
	}
};

// The 16 pixels of a block: weight(i) = interpolation weight of pixel i, subset_bits bit i = subset of pixel i (NS = 2).
template <bool SIGNED, int NS, bool FP, typename Sink, typename Weight>
__device__ __forceinline__ void bc6h_pixel_loop(const Bc6hLines<NS, FP> &ln, uint32_t subset_bits, const Sink &sink, Weight &&weight) {
	if constexpr (FP) {
		// endpoint pair of a pixel: X0 + bit * (X1 - X0), interpolation ((64 - w) e0 + w e1 + 32) / 64 = a + w k, floor
		constexpr float BIAS = SIGNED ? 12582912.0f : 8388608.0f;
		float da[3] = { 0.0f, 0.0f, 0.0f }, dk[3] = { 0.0f, 0.0f, 0.0f };
		if constexpr (NS == 2) {
#pragma unroll
			for (int c = 0; c < 3; c++) {
				da[c] = __fsub_rn(ln.a_f[NS - 1][c], ln.a_f[0][c]);
				dk[c] = __fsub_rn(ln.k_f[NS - 1][c], ln.k_f[0][c]);
			}
		}
#pragma unroll
		for (int i = 0; i < 16; i++) {
			const uint32_t wgt = weight(i);
			const float wf = __fsub_rn(__uint_as_float(0x4B000000u | wgt), 8388608.0f);	// 2^23 + w - 2^23
			// 1.0f or 0.0f from the subset bit: an integer multiply-add makes the bit pattern (FMA pipe)
			const float bit = NS == 2 ? __uint_as_float(mad_u32((subset_bits >> i) & 1u, 0x3F800000u, 0u)) : 0.0f;
			float y[3];
#pragma unroll
			for (int c = 0; c < 3; c++) {
				float a = ln.a_f[0][c], k = ln.k_f[0][c];
				if constexpr (NS == 2) {
					a = __fmaf_rn(bit, da[c], a);
					k = __fmaf_rn(bit, dk[c], k);
				}
				y[c] = __fadd_rd(__fmaf_rn(wf, k, a), BIAS);		// BIAS + floor((a64 + w k64) / 64), :97-108 (Q18)
			}
			sink.emit(i, y[0], y[1], y[2]);
		}
		return;
	}
	else {
		// endpoint pair of a pixel: X0 + bit*(X1 - X0) with multiply-adds (FMA pipe), as in the BC7 decoder
		uint32_t da[3] = { 0u, 0u, 0u }, dk[3] = { 0u, 0u, 0u };
		if constexpr (NS == 2) {
#pragma unroll
			for (int c = 0; c < 3; c++) {
				da[c] = (uint32_t)(ln.a_c[NS - 1][c] - ln.a_c[0][c]);
				dk[c] = (uint32_t)(ln.k_c[NS - 1][c] - ln.k_c[0][c]);
			}
		}
		// Unsigned format, sinks that can take it: everything is carried scaled by 4, so that the two ">> 6" of the
		// reference (InterpolateFloat :97-108, the finish :616-621) become ">> 8" -- byte extractions that ride on the
		// PRMT which assembles the pixel anyway, instead of a shift each on the integer ALU pipe.
		constexpr bool SCALED = !SIGNED && Sink::TAKES_SCALED;
		constexpr int SC = SCALED ? 4 : 1;
#pragma unroll
		for (int i = 0; i < 16; i++) {
			const uint32_t bit = NS == 2 ? ((subset_bits >> i) & 1u) : 0u;
			const int wgt = (int)weight(i);
			int v[3];
#pragma unroll
			for (int c = 0; c < 3; c++) {
				int a = ln.a_c[0][c] * SC, k = ln.k_c[0][c] * SC;
				if constexpr (NS == 2) {
					a = (int)mad_u32(bit, da[c] * SC, (uint32_t)a);
					k = (int)mad_u32(bit, dk[c] * SC, (uint32_t)k);
				}
				if constexpr (SCALED)
					v[c] = a + wgt * k;			// 4 * (a + w k) < 2^25: the interpolated value is bytes 1-2
				else
					v[c] = (a + wgt * k) >> 6;		// InterpolateFloat, :97-108 (Q18)
			}
			if constexpr (SCALED)
				sink.emit_scaled(i, (uint32_t)v[0], (uint32_t)v[1], (uint32_t)v[2]);
			else
				sink(i, v[0], v[1], v[2]);
		}
	}
}

// Index stream of a block with the anchor pixels widened to IB bits (:544-564) and the subset bit of every pixel.
template <int NS>
__device__ __forceinline__ uint64_t bc6h_index_stream(const uint4 &w, uint32_t partition, const BptcTables &tab, uint32_t &subset_bits) {
	constexpr int IB = NS == 1 ? 4 : 3;
	const uint2 sw = window64<NS == 1 ? 65 : 82>(w);
	uint64_t s = insert_zero_bit(((uint64_t)sw.y << 32) | sw.x, IB - 1);
	subset_bits = 0u;
	if constexpr (NS == 2) {
		const uint32_t t = tab.two[partition];
		subset_bits = t & 0xFFFFu;
		s = insert_zero_bit(s, (t >> 16) * IB + (IB - 1));
	}
	return s;
}

template <typename Sink> constexpr bool bc6h_float_pipe() {
#ifdef DETEXB200_BC6H_INT_PIPE
	return false;			// A/B: the integer statement for every sink
#else
	return Sink::FLOAT_PIPE;
#endif
}

template <bool SIGNED, int NS, typename Sink>
__device__ __forceinline__ void bc6h_finish(const uint4 &w, const Bc6hRaw &raw, int epb, int dr, int dg, int db,
const BptcTables &tab, const Sink &sink) {
	constexpr int IB = NS == 1 ? 4 : 3;
	constexpr bool FP = bc6h_float_pipe<Sink>();
	int e[3][2 * NS];
	bc6h_unquantized<SIGNED, NS>(raw, epb, dr, dg, db, e);
	Bc6hLines<NS, FP> ln;
	ln.set(e);
	uint32_t subset_bits;
	const uint64_t s = bc6h_index_stream<NS>(w, raw.partition, tab, subset_bits);
	bc6h_pixel_loop<SIGNED, NS, FP>(ln, subset_bits, sink,
		[&](int i) { return bptc_weight<IB>((uint32_t)(s >> (i * IB)) & ((1u << IB) - 1u)); });
}

// 5-bit mode code -> detex mode number, -1 for the reserved codes (Q8,
// decompress-bptc-float.c:23-33).
__device__ __forceinline__ int bc6h_mode(uint32_t w0) {
	if ((w0 & 2u) == 0u)
		return (int)(w0 & 1u);
	// codes with bit 1 set: 2,6,10,14,18,22,26,30 -> 2..9; 3,7,11,15 -> 10..13
	const uint32_t code = w0 & 0x1Fu;
	if ((code & 1u) == 0u)
		return 2 + (int)(code >> 2);
	return code < 16u ? 10 + (int)(code >> 2) : -1;
}

template <bool SIGNED, typename Sink>
__device__ __forceinline__ bool decode_bptc_float_to(const uint4 blk, uint32_t mode_mask, const BptcTables &tab, const Sink &sink) {
	const int mode = bc6h_mode(blk.x);
	Bc6hRaw raw;
	Bc6hModeInfo info = bc6h_mode_info(0);
	switch (mode) {
#define BC6H_CASE(M) case M: bc6h_scatter<M>(blk, raw); info = bc6h_mode_info(M); break;
	BC6H_CASE(0) BC6H_CASE(1) BC6H_CASE(2) BC6H_CASE(3) BC6H_CASE(4) BC6H_CASE(5) BC6H_CASE(6)
	BC6H_CASE(7) BC6H_CASE(8) BC6H_CASE(9) BC6H_CASE(10) BC6H_CASE(11) BC6H_CASE(12) BC6H_CASE(13)
#undef BC6H_CASE
	default: return false;
	}
	if (mode >= 10)
		bc6h_finish<SIGNED, 1>(blk, raw, info.epb, info.dr, info.dg, info.db, tab, sink);
	else
		bc6h_finish<SIGNED, 2>(blk, raw, info.epb, info.dr, info.dg, info.db, tab, sink);
	return (mode_mask >> mode) & 1u;		// :121
}

// ---- the decoder in two halves ---------------------------------------------------------------------------------
// bc6h_prepare: everything that depends on the mode (bit scatter, delta decode, unquantisation, index widths) ->
// a 48-byte record that looks the same for all 14 modes; bc6h_pixels: record -> 16 pixels, one instruction stream
// for every block.  The mode-sorted kernels run the first half on groups of same-mode blocks and the second half
// in image order (kernels.cu, SPLIT), so that a warp's stores are consecutive blocks of the image.
//   q0 = { (e0 | e1 << 16) of subset 0 for r, g, b;  subset bit of every pixel | valid << 16 }
//   q1 = { the same of subset 1 (one-subset modes: a copy of subset 0);  0 }
//   q2 = the interpolation weight of every pixel, one byte each
struct Bc6hPrepared {
	uint4 q0, q1, q2;
};

template <bool SIGNED, int NS>
__device__ __forceinline__ void bc6h_prepare_tail(const uint4 &w, const Bc6hRaw &raw, const Bc6hModeInfo &info,
const BptcTables &tab, Bc6hPrepared &p) {
	constexpr int IB = NS == 1 ? 4 : 3;
	int e[3][2 * NS];
	bc6h_unquantized<SIGNED, NS>(raw, info.epb, info.dr, info.dg, info.db, e);
	uint32_t subset_bits;
	const uint64_t s = bc6h_index_stream<NS>(w, raw.partition, tab, subset_bits);
	p.q0 = make_uint4(prmt((uint32_t)e[0][0], (uint32_t)e[0][1], 0x5410), prmt((uint32_t)e[1][0], (uint32_t)e[1][1], 0x5410),
		prmt((uint32_t)e[2][0], (uint32_t)e[2][1], 0x5410), subset_bits);
	if constexpr (NS == 2)
		p.q1 = make_uint4(prmt((uint32_t)e[0][2], (uint32_t)e[0][3], 0x5410), prmt((uint32_t)e[1][2], (uint32_t)e[1][3], 0x5410),
			prmt((uint32_t)e[2][2], (uint32_t)e[2][3], 0x5410), 0u);
	else
		p.q1 = make_uint4(p.q0.x, p.q0.y, p.q0.z, 0u);
	uint32_t wq[4];
#pragma unroll
	for (int j = 0; j < 4; j++) {
		uint32_t b[4];
#pragma unroll
		for (int k = 0; k < 4; k++)
			b[k] = bptc_weight<IB>((uint32_t)(s >> ((4 * j + k) * IB)) & ((1u << IB) - 1u));
		wq[j] = prmt(prmt(b[0], b[1], 0x0040), prmt(b[2], b[3], 0x0040), 0x5410);
	}
	p.q2 = make_uint4(wq[0], wq[1], wq[2], wq[3]);
}

// Returns whether the block is valid; the record carries the same flag (bit 16 of q0.w).
template <bool SIGNED>
__device__ __forceinline__ bool bc6h_prepare(const uint4 blk, uint32_t mode_mask, const BptcTables &tab, Bc6hPrepared &p) {
	const int mode = bc6h_mode(blk.x);
	Bc6hRaw raw;
	Bc6hModeInfo info = bc6h_mode_info(0);
	p.q0 = p.q1 = p.q2 = make_uint4(0u, 0u, 0u, 0u);
	switch (mode) {
#define BC6H_CASE(M) case M: bc6h_scatter<M>(blk, raw); info = bc6h_mode_info(M); break;
	BC6H_CASE(0) BC6H_CASE(1) BC6H_CASE(2) BC6H_CASE(3) BC6H_CASE(4) BC6H_CASE(5) BC6H_CASE(6)
	BC6H_CASE(7) BC6H_CASE(8) BC6H_CASE(9) BC6H_CASE(10) BC6H_CASE(11) BC6H_CASE(12) BC6H_CASE(13)
#undef BC6H_CASE
	default: return false;
	}
	if (mode >= 10)
		bc6h_prepare_tail<SIGNED, 1>(blk, raw, info, tab, p);
	else
		bc6h_prepare_tail<SIGNED, 2>(blk, raw, info, tab, p);
	const bool ok = (mode_mask >> mode) & 1u;		// :121
	p.q0.w |= ok ? 0x10000u : 0u;
	return ok;
}

__device__ __forceinline__ bool bc6h_prepared_valid(const Bc6hPrepared &p) { return (p.q0.w >> 16) & 1u; }

template <bool SIGNED, typename Sink>
__device__ __forceinline__ void bc6h_pixels(const Bc6hPrepared &p, const Sink &sink) {
	constexpr bool FP = bc6h_float_pipe<Sink>();
	const uint32_t pair[2][3] = { { p.q0.x, p.q0.y, p.q0.z }, { p.q1.x, p.q1.y, p.q1.z } };
	int e[3][4];
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int s = 0; s < 2; s++) {
			e[c][2 * s] = SIGNED ? (int)(int16_t)(pair[s][c] & 0xFFFFu) : (int)(pair[s][c] & 0xFFFFu);
			e[c][2 * s + 1] = SIGNED ? ((int)pair[s][c] >> 16) : (int)(pair[s][c] >> 16);
		}
	Bc6hLines<2, FP> ln;
	ln.set(e);
	const uint32_t wq[4] = { p.q2.x, p.q2.y, p.q2.z, p.q2.w };
	bc6h_pixel_loop<SIGNED, 2, FP>(ln, p.q0.w & 0xFFFFu, sink, [&](int i) { return prmt(wq[i >> 2], 0u, 0x4440u | (uint32_t)(i & 3)); });
}

template <bool SIGNED>
__device__ __forceinline__ bool decode_bptc_float(const uint4 blk, uint32_t mode_mask, uint32_t,
const BptcTables &tab, uint32_t (&px)[32]) {
	return decode_bptc_float_to<SIGNED>(blk, mode_mask, tab, Bc6hHalfSink<SIGNED>{ px });
}

}  // namespace detexb200
