/*
 * TEST INFRASTRUCTURE ONLY: the BC7 palette decoder of detex_b200/csrc/decode_bptc_palette.cuh
 * compiled for the CPU (host_emu.h) behind a C entry point, for tests/test_emu_bptc.py.
 */
#define DETEXB200_HOST_EMU 1
#include "decode_bptc_palette.cuh"
#include "decode_bptc_float.cuh"
#include "tone.cuh"
#include <cstring>

using namespace detexb200;

static Bc7PaletteTables g_tab;
static bool g_tab_ready = false;

static void tables_init() {
	if (g_tab_ready)
		return;
	for (int i = 0; i < 64; i++) {
		g_tab.packed.two[i] = bptc_two_global[i];
		g_tab.packed.three[i] = bptc_three_global[i];
		g_tab.sub2[i] = bc7_subset_bytes(bptc_two_global[i] & 0xFFFFu, 1);
		g_tab.sub3[i] = bc7_subset_bytes(bptc_three_global[i].x, 2);
	}
	g_tab_ready = true;
}

template <bool SIGNED>
static bool bc6h_split(const uint4 &w, uint32_t (&px)[32]) {
	Bc6hPrepared p;
	const bool ok = bc6h_prepare<SIGNED>(w, 0xFFFFFFFFu, g_tab.packed, p);
	if (ok != bc6h_prepared_valid(p))
		abort();
	/* the record travels through memory between the halves */
	uint8_t bytes[48];
	memcpy(bytes, &p, 48);
	Bc6hPrepared r;
	memcpy(&r, bytes, 48);
	if (ok)
		bc6h_pixels<SIGNED>(r, Bc6hHalfSink<SIGNED>{ px });
	return ok;
}

extern "C" {

/* the variant the host picks (api.cu): plain, or HDR range when the range is not [0, 1] */
static int tone_mode_of(int hdr, float range_min, float range_max) {
	return (hdr && !(range_min == 0.0f && range_max == 1.0f)) ? TONE_RANGE : TONE_PLAIN;
}
static Tone tone_of(int hdr, float range_min, float range_max) {
	ToneArgs a;
	a.range_min = range_min;
	a.range_max = range_max;
	a.inv_gamma = 1.0f;
	a.gamma_lo = 0.0f;
	a.gamma_factor = 1.0f;
	a.zero_target = 0;
	return tone_mode_of(hdr, range_min, range_max) == TONE_RANGE ? tone_setup<TONE_RANGE>(a) : tone_setup<TONE_PLAIN>(a);
}

/* tone.cuh, gamma variant: the host side prepares the corrected range exactly as api.cu does (powf, sign
   kept, hdr.c:149-158); the per-component power is exp2f(e * log2f(f)) here and MUFU on the device */
void emu_tone8_pixels_gamma(const uint8_t *pixels, int n, float gamma, float range_min, float range_max, uint8_t *out) {
	auto spow = [](float v, float ex) { return v >= 0.0f ? powf(v, ex) : -powf(-v, ex); };
	ToneArgs a;
	a.range_min = range_min;
	a.range_max = range_max;
	a.inv_gamma = 1.0f / gamma;
	a.gamma_lo = spow(range_min, a.inv_gamma);
	a.gamma_factor = 1.0f / (spow(range_max, a.inv_gamma) - a.gamma_lo);
	a.zero_target = 0;
	const Tone t = tone_setup<TONE_GAMMA>(a);
	for (int i = 0; i < n; i++) {
		uint16_t h[4];
		memcpy(h, pixels + 8 * (size_t)i, 8);
		const uint32_t q = tone8_pack(tone8_top<TONE_GAMMA>((uint32_t)h[0] << 13, t), tone8_top<TONE_GAMMA>((uint32_t)h[1] << 13, t),
			tone8_top<TONE_GAMMA>((uint32_t)h[2] << 13, t));
		memcpy(out + 4 * (size_t)i, &q, 4);
	}
}

/* tone.cuh, the per-component chain: n FLOAT_RGBX16 pixels (non-negative finite halves) -> n RGBX8
   pixels with the given range parameters (hdr = 0: plain normalised half) */
void emu_tone8_pixels(const uint8_t *pixels, int n, int hdr, float range_min, float range_max, uint8_t *out) {
	const Tone t = tone_of(hdr, range_min, range_max);
	const bool rescale = tone_mode_of(hdr, range_min, range_max) == TONE_RANGE;
	for (int i = 0; i < n; i++) {
		uint16_t h[4];
		memcpy(h, pixels + 8 * (size_t)i, 8);
		const uint32_t q = rescale
			? tone8_pack(tone8_top<TONE_RANGE>((uint32_t)h[0] << 13, t), tone8_top<TONE_RANGE>((uint32_t)h[1] << 13, t),
				tone8_top<TONE_RANGE>((uint32_t)h[2] << 13, t))
			: tone8_pack(tone8_top<TONE_PLAIN>((uint32_t)h[0] << 13, t), tone8_top<TONE_PLAIN>((uint32_t)h[1] << 13, t),
				tone8_top<TONE_PLAIN>((uint32_t)h[2] << 13, t));
		memcpy(out + 4 * (size_t)i, &q, 4);
	}
}

/* tone.cuh, the decoder with the 16-bit pixel sink: n BPTC_FLOAT blocks -> n tiles of 16 RGBX16 pixels; an
   invalid block gets valid[i] = 0 and the conversion of zero pixels */
void emu_bc6h_decode_tone16(const uint8_t *blocks, int n, int hdr, float range_min, float range_max, uint8_t *out,
uint8_t *valid) {
	tables_init();
	const Tone t = tone_of(hdr, range_min, range_max);
	const bool rescale = tone_mode_of(hdr, range_min, range_max) == TONE_RANGE;
	for (int i = 0; i < n; i++) {
		uint4 w;
		memcpy(&w, blocks + 16 * (size_t)i, 16);
		uint32_t q[32];
		const bool ok = rescale ? decode_bptc_float_tone16<TONE_RANGE>(w, 0xFFFFFFFFu, g_tab.packed, t, q)
					  : decode_bptc_float_tone16<TONE_PLAIN>(w, 0xFFFFFFFFu, g_tab.packed, t, q);
		if (!ok) {
			const uint32_t z = rescale ? tone16_of_zero<TONE_RANGE>(t) : tone16_of_zero<TONE_PLAIN>(t);
			for (int k = 0; k < 32; k++)
				q[k] = z | (z << 16);
		}
		memcpy(out + 128 * (size_t)i, q, 128);
		valid[i] = ok ? 1 : 0;
	}
}

/* tone.cuh, the decoder with the 8-bit pixel sink: n BPTC_FLOAT blocks -> n tiles of 16 RGBX8 pixels;
   an invalid block gets valid[i] = 0 and the conversion of zero pixels */
void emu_bc6h_decode_tone8(const uint8_t *blocks, int n, int hdr, float range_min, float range_max, uint8_t *out,
uint8_t *valid) {
	tables_init();
	const Tone t = tone_of(hdr, range_min, range_max);
	const bool rescale = tone_mode_of(hdr, range_min, range_max) == TONE_RANGE;
	for (int i = 0; i < n; i++) {
		uint4 w;
		memcpy(&w, blocks + 16 * (size_t)i, 16);
		uint32_t q[16];
		const bool ok = rescale ? decode_bptc_float_tone8<TONE_RANGE>(w, 0xFFFFFFFFu, g_tab.packed, t, q)
					  : decode_bptc_float_tone8<TONE_PLAIN>(w, 0xFFFFFFFFu, g_tab.packed, t, q);
		if (!ok)
			for (int k = 0; k < 16; k++)
				q[k] = rescale ? tone8_of_zero<TONE_RANGE>(t) : tone8_of_zero<TONE_PLAIN>(t);
		memcpy(out + 64 * (size_t)i, q, 64);
		valid[i] = ok ? 1 : 0;
	}
}

/* BC6H: n blocks -> n tiles of 16 (SIGNED_)FLOAT_RGBX16 pixels (128 bytes, TILED); valid[i] = 0 and a
   zero tile for an invalid block.  split != 0: through the two halves the SPLIT kernels run (bc6h_prepare ->
   48-byte record -> bc6h_pixels) instead of the one-piece decoder */
int emu_bc6h_decode(const uint8_t *blocks, int n, int is_signed, int split, uint8_t *pixels, uint8_t *valid) {
	tables_init();
	for (int i = 0; i < n; i++) {
		uint4 w;
		memcpy(&w, blocks + 16 * (size_t)i, 16);
		uint32_t px[32];
		bool ok;
		if (split)
			ok = is_signed ? bc6h_split<true>(w, px) : bc6h_split<false>(w, px);
		else
			ok = is_signed ? decode_bptc_float<true>(w, 0xFFFFFFFFu, 0u, g_tab.packed, px)
				       : decode_bptc_float<false>(w, 0xFFFFFFFFu, 0u, g_tab.packed, px);
		if (!ok)
			memset(px, 0, sizeof px);
		memcpy(pixels + 128 * (size_t)i, px, 128);
		valid[i] = ok ? 1 : 0;
	}
	return 0;
}

/* n blocks -> n tiles of 16 RGBA8 pixels (TILED layout); valid[i] = 0 for an invalid block (tile
   zero-filled like texture.c:90-93).  path 0: palette + selectors, path 1: the per-pixel decoder
   (decode_bptc).  entries_out (optional): palette entries per block. */
int emu_bc7_decode(const uint8_t *blocks, int n, int swap_rb, int path, uint8_t *pixels, uint8_t *valid,
int *entries_out) {
	tables_init();
	for (int i = 0; i < n; i++) {
		uint4 w;
		memcpy(&w, blocks + 16 * (size_t)i, 16);
		uint32_t px[16];
		bool ok;
		int entries = 16;
		if (path == 1) {
			ok = decode_bptc(w, 0xFFFFFFFFu, 0u, g_tab.packed, px);
			if (ok && swap_rb)
				for (int k = 0; k < 16; k++)
					px[k] = prmt(px[k], 0u, 0x3012);
		}
		else {
			const int mode = bc7_mode_of(w);
			ok = mode >= 0;
			if (ok) {
				int rc = 0;
				bc7_palette(w, mode, g_tab, swap_rb != 0, [&](uint32_t (&pal)[16], uint32_t (&sel)[4], auto ne_) {
					constexpr int NE = decltype(ne_)::value;
					entries = NE;
					uint32_t stored[16];
					for (int k = 0; k < 16; k++)	/* only the first NE entries reach the stage */
						stored[k] = k < NE ? pal[k] : 0xDEADBEEFu;
					for (int k = 0; k < 16; k++) {
						const uint32_t s = (sel[k >> 2] >> (8 * (k & 3))) & 0xFFu;
						if ((int)s >= NE)
							rc = -(i + 1);	/* selector out of range */
						px[k] = stored[s & 15u];
					}
				});
				if (rc)
					return rc;
			}
		}
		if (!ok)
			memset(px, 0, sizeof px);
		memcpy(pixels + 64 * (size_t)i, px, 64);
		valid[i] = ok ? 1 : 0;
		if (entries_out)
			entries_out[i] = entries;
	}
	return 0;
}

}
