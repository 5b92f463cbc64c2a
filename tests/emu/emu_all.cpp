/*
 * TEST INFRASTRUCTURE ONLY: every per-block decoder of detex_b200/csrc/decode_*.cuh compiled for the
 * CPU (host_emu.h) behind one C entry point, for tests/test_emu_decoders.py.  The dispatch mirrors
 * decode_block<F> of kernels.cu (the counterpart of the reference's function table, texture.c:27-48).
 */
#define DETEXB200_HOST_EMU 1
#include "decode_s3tc.cuh"
#include "decode_etc.cuh"
#include "decode_bptc.cuh"
#include "decode_bptc_float.cuh"
#include <cstring>

using namespace detexb200;

static BptcTables g_tab;
static bool g_ready = false;

template <int BW, int OW, typename Fn>
static void run(const uint8_t *blocks, int n, uint8_t *pixels, uint8_t *valid, Fn fn) {
	for (int i = 0; i < n; i++) {
		uint32_t w[4] = { 0, 0, 0, 0 }, px[OW];
		memcpy(w, blocks + (size_t)i * BW * 4, BW * 4);
		const bool ok = fn(w, px);
		if (!ok)
			memset(px, 0, sizeof px);	/* zero tile, texture.c:90-93 */
		memcpy(pixels + (size_t)i * OW * 4, px, OW * 4);
		valid[i] = ok ? 1 : 0;
	}
}
static uint2 b2(const uint32_t *w) { return uint2{ w[0], w[1] }; }
static uint4 b4(const uint32_t *w) { return uint4{ w[0], w[1], w[2], w[3] }; }

extern "C" {

/* format_index = texture_format >> 24 (1..19); n blocks -> n tiles of 16 native pixels (TILED).
   Returns the tile size in bytes, or -1 for an unknown format. */
int emu_decode(int format_index, const uint8_t *blocks, int n, uint32_t mode_mask, uint32_t flags,
uint8_t *pixels, uint8_t *valid) {
	if (!g_ready) {
		for (int i = 0; i < 64; i++) {
			g_tab.two[i] = bptc_two_global[i];
			g_tab.three[i] = bptc_three_global[i];
		}
		g_ready = true;
	}
	const uint32_t mm = mode_mask, fl = flags;
	switch (format_index) {
	case F_BC1: run<2, 16>(blocks, n, pixels, valid, [&](const uint32_t *w, uint32_t (&px)[16]) { return decode_bc1(b2(w), mm, fl, px); }); return 64;
	case F_BC1A: run<2, 16>(blocks, n, pixels, valid, [&](const uint32_t *w, uint32_t (&px)[16]) { return decode_bc1a(b2(w), mm, fl, px); }); return 64;
	case F_BC2: run<4, 16>(blocks, n, pixels, valid, [&](const uint32_t *w, uint32_t (&px)[16]) { return decode_bc2(b4(w), mm, fl, px); }); return 64;
	case F_BC3: run<4, 16>(blocks, n, pixels, valid, [&](const uint32_t *w, uint32_t (&px)[16]) { return decode_bc3(b4(w), mm, fl, px); }); return 64;
	case F_RGTC1: run<2, 4>(blocks, n, pixels, valid, [&](const uint32_t *w, uint32_t (&px)[4]) { return decode_rgtc1(b2(w), mm, fl, px); }); return 16;
	case F_SIGNED_RGTC1: run<2, 8>(blocks, n, pixels, valid, [&](const uint32_t *w, uint32_t (&px)[8]) { return decode_signed_rgtc1(b2(w), mm, fl, px); }); return 32;
	case F_RGTC2: run<4, 8>(blocks, n, pixels, valid, [&](const uint32_t *w, uint32_t (&px)[8]) { return decode_rgtc2(b4(w), mm, fl, px); }); return 32;
	case F_SIGNED_RGTC2: run<4, 16>(blocks, n, pixels, valid, [&](const uint32_t *w, uint32_t (&px)[16]) { return decode_signed_rgtc2(b4(w), mm, fl, px); }); return 64;
	case F_BPTC_FLOAT: run<4, 32>(blocks, n, pixels, valid, [&](const uint32_t *w, uint32_t (&px)[32]) { return decode_bptc_float<false>(b4(w), mm, fl, g_tab, px); }); return 128;
	case F_BPTC_SIGNED_FLOAT: run<4, 32>(blocks, n, pixels, valid, [&](const uint32_t *w, uint32_t (&px)[32]) { return decode_bptc_float<true>(b4(w), mm, fl, g_tab, px); }); return 128;
	case F_BPTC: run<4, 16>(blocks, n, pixels, valid, [&](const uint32_t *w, uint32_t (&px)[16]) { return decode_bptc(b4(w), mm, fl, g_tab, px); }); return 64;
	case F_ETC1: run<2, 16>(blocks, n, pixels, valid, [&](const uint32_t *w, uint32_t (&px)[16]) { return decode_etc1(b2(w), mm, fl, px); }); return 64;
	case F_ETC2: run<2, 16>(blocks, n, pixels, valid, [&](const uint32_t *w, uint32_t (&px)[16]) { return decode_etc2<ETC_PATHS_ALL>(b2(w), mm, fl, px); }); return 64;
	case F_ETC2_PUNCHTHROUGH: run<2, 16>(blocks, n, pixels, valid, [&](const uint32_t *w, uint32_t (&px)[16]) { return decode_etc2_punchthrough<ETC_PATHS_ALL>(b2(w), mm, fl, px); }); return 64;
	case F_ETC2_EAC: run<4, 16>(blocks, n, pixels, valid, [&](const uint32_t *w, uint32_t (&px)[16]) { return decode_etc2_eac<ETC_PATHS_ALL>(b4(w), mm, fl, px); }); return 64;
	case F_EAC_R11: run<2, 8>(blocks, n, pixels, valid, [&](const uint32_t *w, uint32_t (&px)[8]) { return decode_eac_r11<false>(b2(w), mm, fl, px); }); return 32;
	case F_EAC_SIGNED_R11: run<2, 8>(blocks, n, pixels, valid, [&](const uint32_t *w, uint32_t (&px)[8]) { return decode_eac_r11<true>(b2(w), mm, fl, px); }); return 32;
	case F_EAC_RG11: run<4, 16>(blocks, n, pixels, valid, [&](const uint32_t *w, uint32_t (&px)[16]) { return decode_eac_rg11<false>(b4(w), mm, fl, px); }); return 64;
	case F_EAC_SIGNED_RG11: run<4, 16>(blocks, n, pixels, valid, [&](const uint32_t *w, uint32_t (&px)[16]) { return decode_eac_rg11<true>(b4(w), mm, fl, px); }); return 64;
	default: return -1;
	}
}

}
