/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's pixel-format conversion
 * (convert.c, half-float.c, hdr.c), used by tests/ as the checker for the device conversion
 * kernel.  Never linked into the product library.
 *
 * Written buffer-at-a-time like the reference's driver (convert.c:1082-1166): a chain of at most
 * four steps found by a search over a table of (source, target) pairs, each step run over the
 * whole buffer into a fresh temporary.  Pinned against the compiled reference
 * (oracle/_ref/libdetex_ref.so) by tests/test_convert_oracle.py for every pair the search can
 * connect; see that file for the three pairs where the reference's own behaviour is undefined.
 */
#include <fenv.h>
#include <math.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

void oracle_set_error(const char *text);

/* pixel formats, detex.h:83-379 */
enum {
	RGBA8 = 0x334, BGRA8 = 0x33C, RGBX8 = 0x320, BGRX8 = 0x328, RGB8 = 0x220,
	R8 = 0x0, SR8 = 0x1000, RG8 = 0x110, SRG8 = 0x1110,
	R16 = 0x101, SR16 = 0x1101, RG16 = 0x311, SRG16 = 0x1311,
	RGB16 = 0x521, RGBX16 = 0x721, RGBA16 = 0x735,
	FR16 = 0x2101, FR16H = 0x6101, FRG16 = 0x2311, FRG16H = 0x6311,
	FRGBX16 = 0x2721, FRGBX16H = 0x6721, FRGBA16 = 0x6735,
	FRGB16 = 0x2521, FRGB16H = 0x6521, FBGRX16 = 0x2729,
	FR32 = 0x2302, FR32H = 0x6302, FRG32 = 0x2712, FRG32H = 0x6712,
	FRGB32 = 0x2B22, FRGB32H = 0x6B22, FRGBX32 = 0x2F22, FRGBX32H = 0x6F22
};

/* kinds of step; the letter codes keep the table below readable */
enum {
	K_SAME = 's',		/* bytes unchanged / leading components kept */
	K_SWAP = 'x',		/* red <-> blue */
	K_BGRX = 'b',		/* RGB8 -> BGRX8 */
	K_SIGN = 'g',		/* +-128 or +-32768 */
	K_GROW = 'p',		/* add components */
	K_16TO8 = 'n', K_8TO16 = 'w',
	K_U16H = 'h',		/* 16-bit integer -> half */
	K_HU16 = 'u',		/* half -> 16-bit integer */
	K_HDRU16 = 'U',		/* HDR half -> 16-bit integer */
	K_FH = 'f',		/* float -> half */
	K_FU16 = 'i',		/* float -> 16-bit integer */
	K_HF = 'F',		/* half -> float */
	K_HDRF = 'D'		/* HDR float -> float */
};

typedef struct { uint16_t from, to; char kind; } Pair;

/* order of detex_conversion_table, convert.c:765-864 */
static const Pair table[] = {
	{ RGBX8, RGBA8, 's' }, { RGBA8, RGBX8, 's' }, { BGRX8, BGRA8, 's' }, { BGRA8, BGRX8, 's' },
	{ RGBX8, BGRX8, 'x' }, { BGRX8, RGBX8, 'x' }, { RGBA8, BGRA8, 'x' }, { BGRA8, RGBA8, 'x' },
	{ FRGBX16, FBGRX16, 'x' }, { FBGRX16, FRGBX16, 'x' },
	{ RGB8, BGRX8, 'b' },
	{ R8, SR8, 'g' }, { RG8, SRG8, 'g' }, { SR8, R8, 'g' }, { SRG8, RG8, 'g' },
	{ R16, SR16, 'g' }, { RG16, SRG16, 'g' }, { SR16, R16, 'g' }, { SRG16, RG16, 'g' },
	{ RGBA8, R8, 's' }, { RGBA8, RG8, 's' }, { RGB8, R8, 's' }, { RGB8, RG8, 's' },
	{ R8, RGBX8, 'p' }, { RG8, RGBX8, 'p' },
	{ R16, R8, 'n' }, { RG16, RG8, 'n' }, { RGB16, RGB8, 'n' }, { RGBX16, RGBX8, 'n' }, { RGBA16, RGBA8, 'n' },
	{ R8, R16, 'w' }, { RG8, RG16, 'w' }, { RGB8, RGB16, 'w' }, { RGBX8, RGBX16, 'w' }, { RGBA8, RGBA16, 'w' },
	{ R16, FR16, 'h' }, { RG16, FRG16, 'h' }, { RGB16, FRGB16, 'h' }, { RGBX16, FRGBX16, 'h' },
	{ FR16, R16, 'u' }, { FRG16, RG16, 'u' }, { FRGB16, RGB16, 'u' }, { FRGBX16, RGBX16, 'u' }, { FRGBA16, RGBA16, 'u' },
	{ FR16H, R16, 'U' }, { FRG16H, RG16, 'U' }, { FRGBX16H, RGBX16, 'U' },
	{ FR32, FR16, 'f' }, { FRG32, FRG16, 'f' }, { FRGB32, FRGB16, 'f' }, { FRGBX32, FRGBX16, 'f' },
	{ FR32, R16, 'i' }, { FRG32, RG16, 'i' }, { FRGB32, RGB16, 'i' }, { FRGBX32, RGBX16, 'i' },
	{ FR16, FR32, 'F' }, { FRG16, FRG32, 'F' }, { FRGB16, FRGB32, 'F' }, { FRGBX16, FRGBX32, 'F' },
	{ FR32H, FR32, 'D' }, { FRG32H, FRG32, 'D' }, { FRGB32H, FRGB32, 'D' }, { FRGBX32H, FRGBX32, 'D' },
	{ RGB8, RGBX8, 'p' }, { RGBX8, RGB8, 's' },
	{ FRGB16, FRGBX16, 'p' }, { FRGBX16, FRGB16, 's' }, { FRGB16H, FRGBX16H, 'p' }, { FRGBX16H, FRGB16H, 's' },
	{ FRGB32, FRGBX32, 'p' }, { FRGBX32, FRGB32, 's' }, { FRGB32H, FRGBX32H, 'p' }, { FRGBX32H, FRGB32H, 's' },
};
#define N_PAIRS ((int)(sizeof(table) / sizeof(table[0])))

static int px_bytes(uint32_t f) { return 1 + ((f >> 8) & 0xF); }		/* detex.h:879 */
static int comp_bytes(uint32_t f) { return 1 << (f & 3); }			/* detex.h:889 */
static int comp_count(uint32_t f) { return 1 + ((f >> 4) & 3); }		/* detex.h:884 */
static int comp_bits(uint32_t f) {						/* detex.h:894 */
	int b = comp_bytes(f) * 8;
	if (f & 0x2000)
		b -= comp_bytes(f) == 4 ? 10 : 5;
	return b;
}

/* thread-local range / gamma, hdr.c:26-38 */
static __thread float hdr_gamma = 1.0f, hdr_lo = 0.0f, hdr_hi = 1.0f;
void detexSetHDRParameters(float gamma, float range_min, float range_max) {
	hdr_gamma = gamma;
	hdr_lo = range_min;
	hdr_hi = range_max;
}

/* ---- scalar building blocks ------------------------------------------------------------- */

/* single -> half, half-float.c:93-161: flush tiny values to signed zero, round by adding the
 * highest dropped mantissa bit. */
static uint16_t half_of_float(float value) {
	uint32_t w;
	memcpy(&w, &value, 4);
	uint16_t sign = (uint16_t)((w >> 16) & 0x8000u);
	uint32_t mag = w & 0x7FFFFFFFu;
	if (mag == 0)
		return sign;
	int biased = (int)(mag >> 23);
	uint32_t frac = mag & 0x7FFFFFu;
	if (biased == 0)
		return sign;					/* single denormal */
	if (biased == 255)
		return frac ? 0xFE00u : (uint16_t)(sign | 0x7C00u);
	int e = biased - 112;					/* half exponent field */
	if (e >= 31)
		return (uint16_t)(sign | 0x7C00u);
	if (e >= 1) {
		uint16_t h = (uint16_t)(sign | (e << 10) | (frac >> 13));
		return (frac & 0x1000u) ? (uint16_t)(h + 1) : h;
	}
	int drop = 14 - e;					/* 14 .. */
	if (drop > 24)
		return sign;
	frac |= 0x800000u;
	uint16_t m = (uint16_t)(frac >> drop);
	if ((frac >> (drop - 1)) & 1u)
		m++;
	return (uint16_t)(sign | m);
}

/* half -> single, half-float.c:179-256 */
static float float_of_half(uint16_t h) {
	uint32_t sign = (uint32_t)(h & 0x8000u) << 16, w;
	int e = (h >> 10) & 31;
	uint32_t m = h & 0x3FFu;
	if (e == 0 && m == 0)
		w = sign;
	else if (e == 31)
		w = m ? 0xFFC00000u : (sign | 0x7F800000u);
	else {
		int ex = e + 112;
		if (e == 0) {					/* renormalise a half denormal */
			ex = 113;
			while (!(m & 0x400u)) {
				m <<= 1;
				ex--;
			}
			m &= 0x3FFu;
		}
		w = sign | ((uint32_t)ex << 23) | (m << 13);
	}
	float f;
	memcpy(&f, &w, 4);
	return f;
}

/* detex.h:941-948 as the shipped -Ofast build evaluates it (max / min instructions): anything
 * that is not greater than zero, -0.0 included, becomes +0.0. */
static float unit_clamp(float f) { return f > 0.0f ? (f > 1.0f ? 1.0f : f) : 0.0f; }

/* lrintf(clamp(f) * 65535 + 0.5) evaluated entirely in FE_DOWNWARD (half-float.c:304-322).
 * The reference leaves the thread in that mode afterwards; the oracle restores the caller's. */
static uint16_t u16_of_unit_float(float f) {
	volatile float c = unit_clamp(f), scaled, shifted;
	int mode = fegetround();
	fesetround(FE_DOWNWARD);
	scaled = c * 65535.0f;
	shifted = scaled + 0.5f;
	long r = lrintf(shifted);
	fesetround(mode);
	return (uint16_t)r;
}

static float pow_keep_sign(float f, float e) { return f >= 0.0f ? powf(f, e) : -powf(-f, e); }	/* hdr.c:46-49 */

static uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static void wr16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }

/* ---- one table step over a whole buffer ---------------------------------------------------- */
static void run_step(const Pair *t, const uint8_t *in, uint8_t *out, uint32_t n) {
	const int si = px_bytes(t->from), so = px_bytes(t->to), cb = comp_bytes(t->from), cbo = comp_bytes(t->to);
	const int ni = si / cb, no = so / cbo;			/* stored components */
	for (uint32_t i = 0; i < n; i++) {
		const uint8_t *s = in + (size_t)i * si;
		uint8_t *d = out + (size_t)i * so;
		switch (t->kind) {
		case K_SAME:					/* also drops trailing components */
			memcpy(d, s, so);
			break;
		case K_SWAP:					/* convert.c:37-70 */
			memcpy(d, s, so);
			memcpy(d, s + 2 * cb, cb);
			memcpy(d + 2 * cb, s, cb);
			break;
		case K_BGRX:					/* convert.c:73-86 */
			d[0] = s[2]; d[1] = s[1]; d[2] = s[0]; d[3] = 0xFF;
			break;
		case K_SIGN:					/* convert.c:89-186 */
			for (int k = 0; k < ni; k++) {
				if (cb == 1)
					d[k] = (uint8_t)(s[k] + 128);
				else
					wr16(d + 2 * k, (uint32_t)(rd16(s + 2 * k) + 32768) & 0xFFFFu);
			}
			break;
		case K_GROW:					/* convert.c:237-260, 660-669, 692-698, 719-735 */
			memset(d, 0, so);
			memcpy(d, s, si);
			if (cbo == 1)
				d[3] = 0xFF;
			else if (cbo == 2)
				wr16(d + 6, 0x3C00);		/* half 1.0 */
			else {
				float one = 1.0f;
				memcpy(d + 12, &one, 4);
			}
			break;
		case K_16TO8:					/* convert.c:263-331 */
			for (int k = 0; k < ni; k++)
				d[k] = (uint8_t)(((uint32_t)rd16(s + 2 * k) + 127) * 255 / 65535);
			if (t->to == RGBX8)
				d[3] = 0xFF;
			break;
		case K_8TO16:					/* convert.c:333-404 */
			for (int k = 0; k < ni; k++)
				wr16(d + 2 * k, (uint32_t)s[k] * 65535 / 255);
			if (t->to == RGBX16)
				wr16(d + 6, 0xFFFF);
			break;
		case K_U16H:					/* convert.c:473-586 */
			for (int k = 0; k < ni; k++) {
				uint16_t v = rd16(s + 2 * k);
				/* the four-component variant reads int16_t (:565-567) */
				float f = (t->from == RGBX16 ? (float)(int16_t)v : (float)v) * (1.0f / 65535.0f);
				wr16(d + 2 * k, half_of_float(f));
			}
			if (t->from == RGBX16)
				wr16(d + 6, half_of_float(1.0f));	/* :574 */
			break;
		case K_HU16:					/* half-float.c:304-312 */
			for (int k = 0; k < ni; k++)
				wr16(d + 2 * k, u16_of_unit_float(float_of_half(rd16(s + 2 * k))));
			break;
		case K_FH:					/* convert.c:407-425 */
			for (int k = 0; k < ni; k++) {
				float f;
				memcpy(&f, s + 4 * k, 4);
				wr16(d + 2 * k, half_of_float(f));
			}
			break;
		case K_FU16:					/* half-float.c:315-322 */
			for (int k = 0; k < ni; k++) {
				float f;
				memcpy(&f, s + 4 * k, 4);
				wr16(d + 2 * k, u16_of_unit_float(f));
			}
			break;
		case K_HF:					/* convert.c:449-467 */
			for (int k = 0; k < ni; k++) {
				float f = float_of_half(rd16(s + 2 * k));
				memcpy(d + 4 * k, &f, 4);
			}
			break;
		case K_HDRU16: case K_HDRF: {			/* hdr.c:119-214 */
			const bool from_half = t->kind == K_HDRU16;
			for (int k = 0; k < ni; k++) {
				float f;
				if (from_half)
					f = float_of_half(rd16(s + 2 * k));
				else
					memcpy(&f, s + 4 * k, 4);
				if (hdr_gamma == 1.0f) {
					if (!(hdr_lo == 0.0f && hdr_hi == 1.0f)) {
						/* range arithmetic runs after fesetround(FE_DOWNWARD), hdr.c:122-137 */
						int mode = fegetround();
						fesetround(FE_DOWNWARD);
						volatile float span = hdr_hi - hdr_lo;
						volatile float factor = 1.0f / span;
						volatile float shifted = f - hdr_lo;
						volatile float scaled = shifted * factor;
						f = scaled;
						fesetround(mode);
					}
				}
				else {
					/* Special gamma (hdr.c:139-163, :192-212).  This path never selects a rounding
					 * mode itself; it inherits whatever an earlier conversion left in the thread.
					 * In round-to-nearest lrintf(65535.5) gives 65536 and full white wraps to 0, so
					 * the oracle restates the path as it behaves once any float -> uint16
					 * conversion has run in the thread (FE_DOWNWARD, half-float.c:306). */
					int mode = fegetround();
					fesetround(FE_DOWNWARD);
					volatile float e = 1.0f / hdr_gamma;
					volatile float lo = pow_keep_sign(hdr_lo, e), hi = pow_keep_sign(hdr_hi, e);
					volatile float span = hi - lo;
					volatile float factor = 1.0f / span;
					volatile float v = f;
					if (from_half)		/* only the half path gamma-corrects the value, hdr.c:40-50 vs :199-207 */
						v = pow_keep_sign(f, e);
					volatile float shifted = v - lo;
					volatile float scaled = shifted * factor;
					f = scaled;
					fesetround(mode);
				}
				if (from_half)
					wr16(d + 2 * k, u16_of_unit_float(f));
				else {
					f = unit_clamp(f);
					memcpy(d + 4 * k, &f, 4);
				}
			}
			break;
		}
		}
		(void)no;
	}
}

/* ---- search, convert.c:887-1045 --------------------------------------------------------------- */
static int find_chain(uint32_t from, uint32_t to, int chain[4]) {
	int need_c = comp_count(from) < comp_count(to) ? comp_count(from) : comp_count(to);
	int need_b = comp_bits(from) < comp_bits(to) ? comp_bits(from) : comp_bits(to);
#define LOSSLESS(f) (comp_count(f) >= need_c && comp_bits(f) >= need_b)
	for (int a = 0; a < N_PAIRS; a++)
		if (table[a].from == from && table[a].to == to) {
			chain[0] = a;
			return 1;
		}
	for (int z = 0; z < N_PAIRS; z++) {			/* last step first, then a matching first step */
		if (table[z].to != to || !LOSSLESS(table[z].from))
			continue;
		for (int a = 0; a < N_PAIRS; a++)
			if (table[a].from == from && table[a].to == table[z].from) {
				chain[0] = a;
				chain[1] = z;
				return 2;
			}
	}
	for (int a = 0; a < N_PAIRS; a++) {
		if (table[a].from != from || !LOSSLESS(table[a].to))
			continue;
		for (int z = 0; z < N_PAIRS; z++) {
			if (table[z].to != to || !LOSSLESS(table[z].from))
				continue;
			for (int m = 0; m < N_PAIRS; m++)
				if (table[m].from == table[a].to && table[m].to == table[z].from) {
					chain[0] = a;
					chain[1] = m;
					chain[2] = z;
					return 3;
				}
		}
	}
	for (int a = 0; a < N_PAIRS; a++) {
		if (table[a].from != from || !LOSSLESS(table[a].to))
			continue;
		for (int z = 0; z < N_PAIRS; z++) {
			if (table[z].to != to || !LOSSLESS(table[z].from))
				continue;
			for (int b = 0; b < N_PAIRS; b++) {
				if (table[b].from != table[a].to || !LOSSLESS(table[b].to))
					continue;
				for (int c = 0; c < N_PAIRS; c++)
					if (table[c].from == table[b].to && table[c].to == table[z].from) {
						chain[0] = a;
						chain[1] = b;
						chain[2] = c;
						chain[3] = z;
						return 4;
					}
			}
		}
	}
#undef LOSSLESS
	return -1;
}

/* formats_out[0..steps]: the chain of formats (test helper, mirrors detexb200ConversionPlan) */
int oracle_conversion_chain(uint32_t from, uint32_t to, uint32_t formats_out[5]) {
	int chain[4];
	if (from == to) {
		formats_out[0] = from;
		return 0;
	}
	int n = find_chain(from, to, chain);
	if (n < 0)
		return -1;
	formats_out[0] = from;
	for (int i = 0; i < n; i++)
		formats_out[i + 1] = table[chain[i]].to;
	return n;
}

/* convert.c:1082-1166 */
bool detexConvertPixels(uint8_t *src, uint32_t n, uint32_t from, uint8_t *dst, uint32_t to) {
	if (from == to) {
		if (dst != NULL && dst != src)
			memcpy(dst, src, (size_t)n * px_bytes(from));
		return true;
	}
	int chain[4];
	int steps = find_chain(from, to, chain);
	if (steps < 0) {
		oracle_set_error("detexConvertPixels: Unable to find conversion path");
		return false;
	}
	if (dst == NULL) {					/* in-place request: every step must keep the size */
		for (int i = 0; i < steps; i++)
			if (px_bytes(table[chain[i]].from) != px_bytes(table[chain[i]].to)) {
				oracle_set_error("Unable to find in-place conversion path");
				return false;
			}
		dst = src;
	}
	uint8_t *cur = malloc((size_t)n * px_bytes(from) + 16);
	memcpy(cur, src, (size_t)n * px_bytes(from));
	for (int i = 0; i < steps; i++) {
		uint8_t *next = malloc((size_t)n * px_bytes(table[chain[i]].to) + 16);
		run_step(&table[chain[i]], cur, next, n);
		free(cur);
		cur = next;
	}
	memcpy(dst, cur, (size_t)n * px_bytes(to));
	free(cur);
	return true;
}

bool detexConvertPixelsInPlace(uint8_t *buf, uint32_t n, uint32_t from, uint32_t to) {
	return detexConvertPixels(buf, n, from, NULL, to);
}
