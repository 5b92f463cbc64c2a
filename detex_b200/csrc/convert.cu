// Pixel-format conversion on device: the callers' side of the decode path.
//
// The reference converts by chaining up to four single-purpose loops found by a search over a
// table of (source format, target format, function) triples (convert.c:765-864 table,
// :887-1045 search, :1082-1166 driver with temporary buffers between the steps).  Every step
// works on one pixel at a time, so a chain is a composition of per-pixel functions.  Here
//   * the table is restated as data (same order -- the search result depends on it) with an
//     operation code per entry instead of a function pointer;
//   * the search is restated on the host (plan());
//   * ONE kernel executes the whole chain: a CTA stages a tile of source pixels in shared memory
//     with coalesced loads, each thread takes pixels apart into up to four component registers,
//     applies the steps in order, and the converted tile leaves through shared memory again.
// No intermediate buffer ever reaches HBM.  All integer steps are bit-exact; the float <-> half
// steps follow the reference's integer bit manipulation (half-float.c:93-275) and its
// round-down float -> uint16 (half-float.c:304-322) with explicit rounding-mode intrinsics.
#include "common.cuh"
#include "convert_steps.cuh"
#include "launch.h"
#include <cstring>
#include <cstdlib>

namespace detexb200 {

int conversion_table_size() { return TABLE_SIZE; }
void conversion_table_entry(int i, uint32_t *source, uint32_t *target) {
	*source = conversion_table[i].source;
	*target = conversion_table[i].target;
}

int convert_plan(uint32_t s, uint32_t d, int steps[4]) { return convert_plan_search(s, d, steps); }

bool convert_supported(uint32_t s, uint32_t d) {
	int steps[4];
	return convert_plan(s, d, steps) >= 0;
}

// ---- HDR parameters (reference hdr.c:26-38: thread-local gamma and range) -----------------------
struct HdrParams { float gamma, range_min, range_max; };
static thread_local HdrParams g_hdr = { 1.0f, 0.0f, 1.0f };
void set_hdr_parameters(float gamma, float range_min, float range_max) { g_hdr = { gamma, range_min, range_max }; }
void get_hdr_parameters(float *gamma, float *range_min, float *range_max) {
	*gamma = g_hdr.gamma;
	*range_min = g_hdr.range_min;
	*range_max = g_hdr.range_max;
}

// Which conversions the decode kernels absorb into their store epilogue (kernels.cu, "post op" and
// "post class"): relabelling steps anywhere, at most one R <-> B swap of four-component pixels, at
// most one sign flip, and -- last -- at most one step that shrinks the pixel: dropping the fourth
// component, or narrowing one / two 16-bit components to 8 bits.
// BC6H's way to integer pixels is a class of its own (tone.cuh): FLOAT_RGBX16[_HDR] -> RGBX16 (class 5),
// -> RGBX8 (class 3), after which relabelling, one R <-> B swap and dropping X (class 4) may still
// follow.  With the HDR source format the parameters of the calling thread apply (tone_hdr; the
// caller adds 3 * mode to the class).
// DETEXB200_TONE_FUSION=0 in the environment sends the BC6H -> 8-bit targets through the separate
// conversion pass instead (A/B measurements).
static bool tone_fusion() {
	static const bool v = [] { const char *e = getenv("DETEXB200_TONE_FUSION"); return !(e && e[0] == '0'); }();
	return v;
}

FusedConversion convert_fused(uint32_t s, uint32_t d) {
	FusedConversion f = { false, 0u, 0, false };
	if (s == d) {
		f.ok = true;
		return f;
	}
	int steps[4];
	const int n = convert_plan(s, d, steps);
	if (n <= 0)
		return f;
	bool shrunk = false, toned = false;
	for (int i = 0; i < n; i++) {
		const TableEntry &e = conversion_table[steps[i]];
		const int ps_in = pixel_bytes(e.source), ps_out = pixel_bytes(e.target), cs = component_bytes(e.source);
		if (toned) {					// on RGBX8 pixels, after the tone step
			if (e.op == ST_KEEP && ps_in == ps_out)
				continue;
			if (e.op == ST_SWAP_RB && f.post_class == 3 && !(f.post_op & 1u)) {
				f.post_op |= 1u;
				continue;
			}
			if (e.op == ST_KEEP && e.source == PF_RGBX8 && e.target == PF_RGB8 && f.post_class == 3) {
				f.post_class = 4;
				continue;
			}
			return f;
		}
		if (shrunk)
			return f;				// nothing may follow the size change
		if ((e.op == ST_HALF_TO_INT || e.op == ST_HALF_HDR_TO_INT) && e.target == PF_RGBX16 && f.post_op == 0u && tone_fusion() &&
		    i + 1 == n && f.post_class == 0) {	// the 16-bit stage on its own: target RGBX16
			if (e.op == ST_HALF_HDR_TO_INT)
				f.tone_hdr = true;
			f.post_class = 5;
			shrunk = true;
			continue;
		}
		if ((e.op == ST_HALF_TO_INT || e.op == ST_HALF_HDR_TO_INT) && e.target == PF_RGBX16 && f.post_op == 0u && tone_fusion() &&
		    i + 1 < n && conversion_table[steps[i + 1]].op == ST_NARROW && conversion_table[steps[i + 1]].target == PF_RGBX8) {
			if (e.op == ST_HALF_HDR_TO_INT)
				f.tone_hdr = true;
			f.post_class = 3;
			toned = true;
			i++;
			continue;
		}
		if (e.op == ST_KEEP && ps_in == ps_out)
			continue;
		if (e.op == ST_SWAP_RB && ps_in / cs == 4 && !(f.post_op & 1u)) {
			f.post_op |= 1u;
			continue;
		}
		if (e.op == ST_FLIP_SIGN && !(f.post_op & 6u)) {
			f.post_op |= cs == 2 ? 2u : 4u;
			continue;
		}
		if (e.op == ST_KEEP && ps_in / cs == 4 && ps_out * 4 == ps_in * 3 && (cs == 1 || cs == 2)) {
			f.post_class = 1;			// RGBX8 -> RGB8, FLOAT_RGBX16 -> FLOAT_RGB16
			shrunk = true;
			continue;
		}
		if (e.op == ST_NARROW && ps_in / cs <= 2) {
			f.post_class = 2;			// R16 -> R8, RG16 -> RG8
			shrunk = true;
			continue;
		}
		return f;
	}
	f.ok = true;
	return f;
}

constexpr int CONVERT_THREADS = 256, CONVERT_TILE = 1024;		// pixels per CTA pass

// Coalesced copy of n bytes between global and shared memory: 16-byte vectors when both sides
// allow it, bytes otherwise (odd pixel sizes at a tile tail, unaligned caller buffers).
template <bool TO_SHARED>
__device__ __forceinline__ void tile_copy(uint8_t *smem, uint8_t *gmem, uint32_t n) {
	uint32_t done = 0;
	if ((((uintptr_t)gmem) & 15u) == 0) {
		const uint32_t vecs = n >> 4;
		for (uint32_t i = threadIdx.x; i < vecs; i += CONVERT_THREADS) {
			if (TO_SHARED)
				reinterpret_cast<uint4 *>(smem)[i] = *reinterpret_cast<const uint4 *>(gmem + 16 * i);
			else {
				const uint4 v = reinterpret_cast<const uint4 *>(smem)[i];
				st_stream_v4(gmem + 16 * i, v.x, v.y, v.z, v.w);
			}
		}
		done = vecs << 4;
	}
	for (uint32_t i = done + threadIdx.x; i < n; i += CONVERT_THREADS) {
		if (TO_SHARED)
			smem[i] = gmem[i];
		else
			gmem[i] = smem[i];
	}
}

__global__ void __launch_bounds__(CONVERT_THREADS)
convert_kernel(const uint8_t *src, uint8_t *dst, uint64_t n_pixels, ConvertProgram prog) {
	__shared__ __align__(16) uint8_t tile_in[CONVERT_TILE * 16];
	__shared__ __align__(16) uint8_t tile_out[CONVERT_TILE * 16];
	const uint32_t sf = prog.format[0], tf = prog.format[prog.n_steps];
	const uint32_t ps_in = 1u + ((sf & 0xF00u) >> 8), ps_out = 1u + ((tf & 0xF00u) >> 8);
	const uint32_t cs_in = 1u << (sf & 3u), cs_out = 1u << (tf & 3u);
	const int n_in = (int)(ps_in / cs_in), n_out = (int)(ps_out / cs_out);
	const uint64_t n_tiles = (n_pixels + CONVERT_TILE - 1) / CONVERT_TILE;
	for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
		const uint64_t first = t * CONVERT_TILE;
		const uint32_t here = (uint32_t)(n_pixels - first < (uint64_t)CONVERT_TILE ? n_pixels - first : CONVERT_TILE);
		tile_copy<true>(tile_in, const_cast<uint8_t *>(src) + first * ps_in, here * ps_in);
		__syncthreads();
		for (uint32_t px = threadIdx.x; px < here; px += CONVERT_THREADS) {
			uint32_t c[4] = { 0u, 0u, 0u, 0u };
			load_pixel(tile_in + px * ps_in, ps_in, cs_in, n_in, c);
			for (int s = 0; s < prog.n_steps; s++)
				apply_step(prog.op[s], prog.format[s], prog.format[s + 1], c, prog);
			store_pixel(tile_out + px * ps_out, ps_out, cs_out, n_out, c);
		}
		__syncthreads();
		tile_copy<false>(tile_out, dst + first * ps_out, here * ps_out);
		__syncthreads();
	}
}

// One thread per block: clear the 4x4 pixels (clipped to the image for LINEAR) of marked blocks.
__global__ void __launch_bounds__(256)
zero_invalid_kernel(const uint32_t *__restrict__ invalid_map, uint8_t *pixels, int width, int height, int wib,
uint32_t n_blocks, int layout, int ps) {
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_blocks; i += gridDim.x * blockDim.x) {
		if (!((invalid_map[i >> 5] >> (i & 31u)) & 1u))
			continue;
		if (layout == LAYOUT_TILED) {
			uint8_t *p = pixels + (size_t)i * 16 * ps;
			for (int b = 0; b < 16 * ps; b++)
				p[b] = 0;
		}
		else {
			const int by = (int)(i / (uint32_t)wib), bx = (int)(i % (uint32_t)wib);
			const int rows = min(4, height - by * 4), cols = min(4, width - bx * 4);
			for (int r = 0; r < rows; r++) {
				uint8_t *p = pixels + ((size_t)(by * 4 + r) * width + (size_t)bx * 4) * ps;
				for (int b = 0; b < cols * ps; b++)
					p[b] = 0;
			}
		}
	}
}

cudaError_t launch_zero_invalid(const uint32_t *invalid_map, void *pixels, int width, int height, int wib, int hib,
int layout, int ps, cudaStream_t stream) {
	const uint64_t n = (uint64_t)wib * (uint64_t)hib;
	if (n == 0)
		return cudaSuccess;
	const uint64_t want = (n + 255) / 256;
	zero_invalid_kernel<<<(unsigned)(want < 148ull * 32 ? want : 148ull * 32), 256, 0, stream>>>(invalid_map,
		(uint8_t *)pixels, width, height, wib, (uint32_t)n, layout, ps);
	count_launch();
	return cudaGetLastError();
}

cudaError_t launch_convert(const void *src, uint64_t n, uint32_t s, void *dst, uint32_t d, cudaStream_t stream) {
	if (n == 0)
		return cudaSuccess;
	if (s == d) {
		if (src == dst)
			return cudaSuccess;
		return cudaMemcpyAsync(dst, src, (size_t)n * pixel_bytes(s), cudaMemcpyDeviceToDevice, stream);
	}
	int steps[4];
	const int n_steps = convert_plan(s, d, steps);
	if (n_steps < 0)
		return cudaErrorNotSupported;
	ConvertProgram prog;
	memset(&prog, 0, sizeof(prog));
	prog.n_steps = n_steps;
	prog.format[0] = (uint16_t)s;
	for (int i = 0; i < n_steps; i++) {
		prog.op[i] = conversion_table[steps[i]].op;
		prog.format[i + 1] = (uint16_t)conversion_table[steps[i]].target;
	}
	prog.gamma = g_hdr.gamma;
	prog.range_min = g_hdr.range_min;
	prog.range_max = g_hdr.range_max;
	const uint64_t tiles = (n + CONVERT_TILE - 1) / CONVERT_TILE;
	const uint64_t cap = 148ull * 6 * 8;
	convert_kernel<<<(unsigned)(tiles < cap ? tiles : cap), CONVERT_THREADS, 0, stream>>>(
		(const uint8_t *)src, (uint8_t *)dst, n, prog);
	count_launch();
	return cudaGetLastError();
}

}  // namespace detexb200
