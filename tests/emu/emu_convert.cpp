/*
 * TEST INFRASTRUCTURE ONLY: the conversion table, the chain search and the per-pixel step functions of
 * detex_b200/csrc/convert_steps.cuh compiled for the CPU (host_emu.h) behind a C entry point, for
 * tests/test_emu_convert.py.  The loop below is convert_kernel (convert.cu) without its shared-memory
 * staging: load_pixel, the steps of the chain in order, store_pixel.
 */
#define DETEXB200_HOST_EMU 1
#include "convert_steps.cuh"
#include <cstring>

using namespace detexb200;

extern "C" {

/* n pixels of format spf -> format tpf with the given HDR parameters.  Returns the number of steps of the
   chain (0 = identical formats, copied), or -1 when the formats are not connected.  chain_out[5] receives
   the formats the chain goes through. */
int emu_convert(const uint8_t *src, uint64_t n, uint32_t spf, uint8_t *dst, uint32_t tpf, float gamma,
float range_min, float range_max, uint32_t *chain_out) {
	int steps[4];
	const int n_steps = convert_plan_search(spf, tpf, steps);
	if (n_steps < 0)
		return -1;
	ConvertProgram prog;
	memset(&prog, 0, sizeof prog);
	prog.n_steps = n_steps;
	prog.format[0] = (uint16_t)spf;
	for (int i = 0; i < n_steps; i++) {
		prog.op[i] = conversion_table[steps[i]].op;
		prog.format[i + 1] = (uint16_t)conversion_table[steps[i]].target;
	}
	prog.gamma = gamma;
	prog.range_min = range_min;
	prog.range_max = range_max;
	for (int i = 0; i <= n_steps; i++)
		chain_out[i] = prog.format[i];
	const uint32_t sf = prog.format[0], tf = prog.format[n_steps];
	const uint32_t ps_in = 1u + ((sf & 0xF00u) >> 8), ps_out = 1u + ((tf & 0xF00u) >> 8);
	const uint32_t cs_in = 1u << (sf & 3u), cs_out = 1u << (tf & 3u);
	const int n_in = (int)(ps_in / cs_in), n_out = (int)(ps_out / cs_out);
	if (n_steps == 0) {
		memcpy(dst, src, n * ps_in);
		return 0;
	}
	alignas(16) uint8_t in[16], out[16];
	for (uint64_t p = 0; p < n; p++) {
		memcpy(in, src + p * ps_in, ps_in);
		uint32_t c[4] = { 0u, 0u, 0u, 0u };
		load_pixel(in, ps_in, cs_in, n_in, c);
		for (int s = 0; s < n_steps; s++)
			apply_step(prog.op[s], prog.format[s], prog.format[s + 1], c, prog);
		store_pixel(out, ps_out, cs_out, n_out, c);
		memcpy(dst + p * ps_out, out, ps_out);
	}
	return n_steps;
}

}
