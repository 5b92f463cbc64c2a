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
//     (cp.async, double-buffered), each thread takes pixels apart into up to four component registers,
//     applies the steps in order, and the converted tile leaves through shared memory again;
//   * chains that only shuffle 8-bit components are folded into one byte permutation on the host and
//     run register to register (byte_map_kernel).
// No intermediate buffer ever reaches HBM.  All integer steps are bit-exact; the float <-> half
// steps follow the reference's integer bit manipulation (half-float.c:93-275) and its
// round-down float -> uint16 (half-float.c:304-322) with explicit rounding-mode intrinsics.
#include "common.cuh"
#include "convert_steps.cuh"
#include "launch.h"
#include <atomic>
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

// A CTA converts tiles of pixels: the source bytes of tile t + 1 stream into shared memory (cp.async, double-buffered)
// while tile t is converted and stored, so that every resident CTA always has one tile of loads in flight -- with the
// synchronous load -> convert -> store loop of round 1 the kernel was bound by the latency of those loads (4 KiB in
// flight per CTA: 0.6-2.7 TB/s).  A tile is as many pixels as fit CONVERT_TILE_BYTES on its larger side.
constexpr int CONVERT_THREADS = 256, CONVERT_TILE_BYTES = 16384;
// pixels per thread and pass: more for small pixels, whose conversion is little work against the dispatch of a pass
#ifndef DETEXB200_CONVERT_PIXELS_SMALL
#define DETEXB200_CONVERT_PIXELS_SMALL 8
#endif
constexpr int CONVERT_PIXELS_LARGE = 4, CONVERT_PIXELS_SMALL = DETEXB200_CONVERT_PIXELS_SMALL;

__device__ __forceinline__ void convert_cp_async_16(void *smem_dst, const void *gmem_src) {
	const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gmem_src) : "memory");
}

// Source bytes of one tile -> shared memory.  ASYNC: 16-byte cp.async requests (source and tile start 16-byte aligned;
// the tail bytes of the last tile are copied synchronously); otherwise a plain byte / vector copy.
template <bool ASYNC>
__device__ __forceinline__ void tile_load(uint8_t *smem, const uint8_t *gmem, uint32_t n) {
	uint32_t done = 0;
	if (ASYNC) {
		const uint32_t vecs = n >> 4;
		for (uint32_t i = threadIdx.x; i < vecs; i += CONVERT_THREADS)
			convert_cp_async_16(smem + 16 * i, gmem + 16 * i);
		done = vecs << 4;
	}
	for (uint32_t i = done + threadIdx.x; i < n; i += CONVERT_THREADS)
		smem[i] = gmem[i];
}

// Converted tile -> global memory: 16-byte vectors when the destination allows it, bytes otherwise (odd pixel sizes at
// a tile tail, unaligned caller buffers).
__device__ __forceinline__ void tile_store(const uint8_t *smem, uint8_t *gmem, uint32_t n) {
	uint32_t done = 0;
	if ((((uintptr_t)gmem) & 15u) == 0) {
		const uint32_t vecs = n >> 4;
		for (uint32_t i = threadIdx.x; i < vecs; i += CONVERT_THREADS) {
			const uint4 v = reinterpret_cast<const uint4 *>(smem)[i];
			st_stream_v4(gmem + 16 * i, v.x, v.y, v.z, v.w);
		}
		done = vecs << 4;
	}
	for (uint32_t i = done + threadIdx.x; i < n; i += CONVERT_THREADS)
		gmem[i] = smem[i];
}

template <bool ASYNC, int CONVERT_PIXELS>
__global__ void __launch_bounds__(CONVERT_THREADS, 4)
convert_kernel(const uint8_t *src, uint8_t *dst, uint64_t n_pixels, uint32_t tile_px, ConvertProgram prog) {
	extern __shared__ __align__(16) uint8_t convert_smem[];
	uint8_t *const tile_in0 = convert_smem, *const tile_in1 = convert_smem + CONVERT_TILE_BYTES;
	uint8_t *const tile_out = convert_smem + 2 * CONVERT_TILE_BYTES;
	const uint32_t sf = prog.format[0], tf = prog.format[prog.n_steps];
	const uint32_t ps_in = 1u + ((sf & 0xF00u) >> 8), ps_out = 1u + ((tf & 0xF00u) >> 8);
	const uint32_t cs_in = 1u << (sf & 3u), cs_out = 1u << (tf & 3u);
	const uint64_t n_tiles = (n_pixels + tile_px - 1) / tile_px;
	auto pixels_of = [&](uint64_t t) { const uint64_t left = n_pixels - t * tile_px; return (uint32_t)(left < tile_px ? left : tile_px); };
	uint64_t t = blockIdx.x;
	if (t >= n_tiles)
		return;
	tile_load<ASYNC>(tile_in0, src + t * tile_px * ps_in, pixels_of(t) * ps_in);
	asm volatile("cp.async.commit_group;" ::: "memory");
	for (int cur = 0;; cur ^= 1) {
		uint8_t *const tile_in = cur ? tile_in1 : tile_in0;
		const uint64_t next = t + gridDim.x;
		if (next < n_tiles)		// (its buffer was read two barriers ago)
			tile_load<ASYNC>(cur ? tile_in0 : tile_in1, src + next * tile_px * ps_in, pixels_of(next) * ps_in);
		asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 1;" ::: "memory");
		__syncthreads();	// tile t has landed (everyone's copies); the previous tile has left tile_out
		const uint32_t here = pixels_of(t);
		// CONVERT_PIXELS pixels per thread and pass, CONVERT_THREADS apart (consecutive lanes hold consecutive pixels); the
		// run-time dispatch on the steps' operations is paid once per pass, not per pixel
		for (uint32_t px0 = threadIdx.x; px0 < here; px0 += CONVERT_THREADS * CONVERT_PIXELS) {
			uint32_t c[CONVERT_PIXELS][4], at[CONVERT_PIXELS];
#pragma unroll
			for (int k = 0; k < CONVERT_PIXELS; k++) {
#pragma unroll
				for (int j = 0; j < 4; j++)
					c[k][j] = 0u;
				// (a pixel behind the end of the tile: the last pixel is converted again and not stored)
				at[k] = min(px0 + k * CONVERT_THREADS, here - 1u);
			}
			// the pixel geometry is uniform too: one dispatch per pass on (pixel size, component size), then straight-line
			// code with compile-time sizes
#define DETEX_GEOMETRIES(X) X(1, 1) X(2, 1) X(3, 1) X(4, 1) X(2, 2) X(4, 2) X(6, 2) X(8, 2) X(4, 4) X(8, 4) X(12, 4) X(16, 4)
			switch (ps_in * 8u + cs_in) {
#define DETEX_LOAD(PS, CS) case PS * 8 + CS: \
				_Pragma("unroll") for (int k = 0; k < CONVERT_PIXELS; k++) load_pixel(tile_in + at[k] * PS, PS, CS, PS / CS, c[k]); \
				break;
			DETEX_GEOMETRIES(DETEX_LOAD)
#undef DETEX_LOAD
			}
			for (int s = 0; s < prog.n_steps; s++)
				apply_step_k<CONVERT_PIXELS>(prog.op[s], prog.format[s], prog.format[s + 1], c, prog);
			switch (ps_out * 8u + cs_out) {
#define DETEX_STORE(PS, CS) case PS * 8 + CS: \
				_Pragma("unroll") for (int k = 0; k < CONVERT_PIXELS; k++) { \
					const uint32_t px = px0 + k * CONVERT_THREADS; \
					if (px < here) store_pixel(tile_out + px * PS, PS, CS, PS / CS, c[k]); \
				} \
				break;
			DETEX_GEOMETRIES(DETEX_STORE)
#undef DETEX_STORE
			}
#undef DETEX_GEOMETRIES
		}
		__syncthreads();	// tile converted; tile_in may be refilled by the next iteration's loads
		tile_store(tile_out, dst + t * tile_px * ps_out, here * ps_out);
		if (next >= n_tiles)
			break;
		t = next;
	}
	asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ---- byte-shuffle conversions ---------------------------------------------------------------------------------------
// Chains that only move, drop, pad or sign-flip 8-bit components (RGBA8 <-> BGRA8, RGBX8 -> RGB8, RGB8 -> BGRX8, R8 ->
// RGBX8, SIGNED_RG8 -> RG8 ...: convert.c:37-136, :237-260, :660-690) are one byte permutation with constants per pixel.
// The plan is folded on the host into a PRMT selector and an XOR mask; a thread then converts 16 pixels from registers
// to registers -- whole 16-byte vectors in and out, about four instructions per pixel against 54 in the interpreter
// (ncu, RGBA8 -> BGRA8: 56.7 M warp instructions for 33.5 M pixels, issue-bound at 0.49 of the HBM peak).
struct ByteMap {
	uint32_t sel, xor_mask;
	int ps_in, ps_out;
};
static bool plan_byte_map(const ConvertProgram &prog, ByteMap *m) {
	if (component_bytes(prog.format[0]) != 1)
		return false;
	int from[4];
	uint32_t x[4] = { 0u, 0u, 0u, 0u };
	const int n0 = pixel_bytes(prog.format[0]);
	for (int j = 0; j < 4; j++)
		from[j] = j < n0 ? j : -1;		// (registers behind the last stored component start as zero)
	for (int s = 0; s < prog.n_steps; s++) {
		const uint32_t a = prog.format[s], b = prog.format[s + 1];
		if (component_bytes(a) != 1 || component_bytes(b) != 1)
			return false;
		switch (prog.op[s]) {		// the same edits as apply_step_k, on symbols
		case ST_KEEP:
			break;
		case ST_SWAP_RB: case ST_RGB_TO_BGRX: {
			const int t = from[0];
			from[0] = from[2];
			from[2] = t;
			const uint32_t u = x[0];
			x[0] = x[2];
			x[2] = u;
			if (prog.op[s] == ST_RGB_TO_BGRX) {
				from[3] = -1;
				x[3] = 0xFFu;
			}
			break;
		}
		case ST_FLIP_SIGN:
			for (int j = 0; j < 4; j++)
				x[j] ^= 0x80u;
			break;
		case ST_PAD:
			for (int j = pixel_bytes(a); j < 4; j++) {
				from[j] = -1;
				x[j] = j == 3 ? 0xFFu : 0u;
			}
			break;
		default:
			return false;
		}
	}
	m->ps_in = n0;
	m->ps_out = pixel_bytes(prog.format[prog.n_steps]);
	m->sel = 0u;
	m->xor_mask = 0u;
	for (int j = 0; j < 4; j++) {
		const bool stored = j < m->ps_out;
		m->sel |= (uint32_t)((stored && from[j] >= 0) ? from[j] : 4) << (4 * j);	// 4 = byte 0 of the zero operand
		m->xor_mask |= (stored ? (x[j] & 0xFFu) : 0u) << (8 * j);
	}
	return true;
}

template <int PS_IN, int PS_OUT>
__global__ void __launch_bounds__(256)
byte_map_kernel(const uint8_t *src, uint8_t *dst, uint64_t n_chunks, uint32_t sel, uint32_t xor_mask) {
	for (uint64_t i = (uint64_t)blockIdx.x * 256u + threadIdx.x; i < n_chunks; i += (uint64_t)gridDim.x * 256u) {
		uint32_t in[4 * PS_IN + 1], out[4 * PS_OUT + 1];
#pragma unroll
		for (int v = 0; v < PS_IN; v++) {
			const uint4 q = ld_stream_v4(src + (i * PS_IN + v) * 16);
			in[4 * v] = q.x; in[4 * v + 1] = q.y; in[4 * v + 2] = q.z; in[4 * v + 3] = q.w;
		}
		in[4 * PS_IN] = 0u;
#pragma unroll
		for (int w = 0; w <= 4 * PS_OUT; w++)
			out[w] = 0u;
#pragma unroll
		for (int p = 0; p < 16; p++) {
			constexpr uint32_t KEEP = PS_OUT == 4 ? 0xFFFFFFFFu : ((1u << (8 * (PS_OUT & 3))) - 1u);
			const int off = p * PS_IN, at = p * PS_OUT;		// byte positions of pixel p (compile-time after unrolling)
			const uint32_t v = (off & 3) == 0 ? in[off >> 2] : __funnelshift_r(in[off >> 2], in[(off >> 2) + 1], 8 * (off & 3));
			const uint32_t r = (prmt(v, 0u, sel) ^ xor_mask) & KEEP;
			out[at >> 2] |= r << (8 * (at & 3));
			if ((at & 3) + PS_OUT > 4)
				out[(at >> 2) + 1] |= r >> (8 * (4 - (at & 3)));
		}
#pragma unroll
		for (int v = 0; v < PS_OUT; v++)
			st_stream_v4(dst + (i * PS_OUT + v) * 16, out[4 * v], out[4 * v + 1], out[4 * v + 2], out[4 * v + 3]);
	}
}

template <int PS_IN>
static void launch_byte_map_in(const ByteMap &m, const uint8_t *src, uint8_t *dst, uint64_t n_chunks, unsigned grid, cudaStream_t stream) {
	switch (m.ps_out) {
	case 1: byte_map_kernel<PS_IN, 1><<<grid, 256, 0, stream>>>(src, dst, n_chunks, m.sel, m.xor_mask); break;
	case 2: byte_map_kernel<PS_IN, 2><<<grid, 256, 0, stream>>>(src, dst, n_chunks, m.sel, m.xor_mask); break;
	case 3: byte_map_kernel<PS_IN, 3><<<grid, 256, 0, stream>>>(src, dst, n_chunks, m.sel, m.xor_mask); break;
	default: byte_map_kernel<PS_IN, 4><<<grid, 256, 0, stream>>>(src, dst, n_chunks, m.sel, m.xor_mask); break;
	}
}
static void launch_byte_map(const ByteMap &m, const uint8_t *src, uint8_t *dst, uint64_t n_chunks, cudaStream_t stream) {
	const uint64_t want = (n_chunks + 255) / 256, cap = (uint64_t)sm_count() * 32;
	const unsigned grid = (unsigned)(want < cap ? want : cap);
	switch (m.ps_in) {
	case 1: launch_byte_map_in<1>(m, src, dst, n_chunks, grid, stream); break;
	case 2: launch_byte_map_in<2>(m, src, dst, n_chunks, grid, stream); break;
	case 3: launch_byte_map_in<3>(m, src, dst, n_chunks, grid, stream); break;
	default: launch_byte_map_in<4>(m, src, dst, n_chunks, grid, stream); break;
	}
	count_launch();
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
	// byte shuffles: whole chunks of 16 pixels through the register kernel, the last few pixels through the interpreter
	static const bool byte_maps = [] { const char *e = getenv("DETEXB200_CONVERT_BYTE_MAP"); return !(e && e[0] == '0'); }();
	ByteMap bm;
	if (byte_maps && n >= 16 && ((((uintptr_t)src) | ((uintptr_t)dst)) & 15u) == 0 && plan_byte_map(prog, &bm)) {
		const uint64_t n_chunks = n / 16;
		launch_byte_map(bm, (const uint8_t *)src, (uint8_t *)dst, n_chunks, stream);
		const cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess)
			return e;
		src = (const uint8_t *)src + n_chunks * 16 * (uint64_t)bm.ps_in;
		dst = (uint8_t *)dst + n_chunks * 16 * (uint64_t)bm.ps_out;
		n -= n_chunks * 16;
		if (n == 0)
			return cudaSuccess;
	}
	const uint32_t ps_in = pixel_bytes(s), ps_out = pixel_bytes(d), ps_max = ps_in > ps_out ? ps_in : ps_out;
	const uint32_t tile_px = ((uint32_t)CONVERT_TILE_BYTES / ps_max) & ~255u;	// 1024 (16-byte pixels) ... 16384 (1 byte)
	const uint64_t tiles = (n + tile_px - 1) / tile_px;
	const size_t smem = 3 * (size_t)CONVERT_TILE_BYTES;
	const uint64_t resident = (uint64_t)sm_count() * 4;		// persistent: four CTAs per SM (48 KiB of shared memory each)
	const unsigned grid = (unsigned)(tiles < resident ? tiles : resident);
	// tile starts are multiples of 256 pixels: 16-byte aligned whenever the source itself is
	const bool async = (((uintptr_t)src) & 15u) == 0;
	static std::atomic<bool> configured[64];
	int dev = 0;
	cudaGetDevice(&dev);
	if (!configured[dev & 63].load()) {
		cudaFuncSetAttribute(convert_kernel<true, CONVERT_PIXELS_LARGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		cudaFuncSetAttribute(convert_kernel<false, CONVERT_PIXELS_LARGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		cudaFuncSetAttribute(convert_kernel<true, CONVERT_PIXELS_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		cudaFuncSetAttribute(convert_kernel<false, CONVERT_PIXELS_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		configured[dev & 63].store(true);
	}
	const bool small = ps_max <= 4;		// (tiles of 4096 pixels and more)
	auto *kernel = async ? (small ? convert_kernel<true, CONVERT_PIXELS_SMALL> : convert_kernel<true, CONVERT_PIXELS_LARGE>)
			     : (small ? convert_kernel<false, CONVERT_PIXELS_SMALL> : convert_kernel<false, CONVERT_PIXELS_LARGE>);
	kernel<<<grid, CONVERT_THREADS, smem, stream>>>((const uint8_t *)src, (uint8_t *)dst, n, tile_px, prog);
	count_launch();
	return cudaGetLastError();
}

}  // namespace detexb200
