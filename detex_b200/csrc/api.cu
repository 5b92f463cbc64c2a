// C-ABI layer of libdetex_b200.so.
//
//  * include/detex_b200.h: device-pointer entry points (decode, convert, generator, hash);
//  * include/detex.h:      the reference's host-pointer API for this path
//                          (reference texture.c:55-145, convert.c:1082-1171, misc.c:75-93),
//                          implemented as pinned/pageable host <-> HBM copies pipelined
//                          around the same kernels.  There is no CPU decode path: when no
//                          CUDA device is usable every call fails with an error message.
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/detex.h"
#include "../../include/detex_b200.h"
#include "launch.h"

using namespace detexb200;

namespace {

// ---- thread-local error message (reference misc.c:75-93) --------------------------------
thread_local std::string g_error;

void set_error(const char *fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	g_error = buf;
}

}  // namespace
namespace detexb200 {
void set_error_text(const char *text) { g_error = text; }
}
namespace {

int cuda_fail(const char *what, cudaError_t e) {
	set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
	return DETEXB200_ECUDA;
}

inline int pixel_size(uint32_t pf) { return 1 + ((pf & 0xF00) >> 8); }
inline int block_bytes(uint32_t tf) { return 8 + ((tf & DETEX_TEXTURE_FORMAT_128BIT_BLOCK_BIT) >> 20); }

// ---- per-thread device context: streams + grow-only buffers -------------------------------
constexpr int N_STREAMS = 4;

struct Context {
	bool ready = false, failed = false;
	int device = -1;			// the context (streams, buffers) belongs to this device
	cudaStream_t streams[N_STREAMS] = {};
	void *d_in = nullptr, *d_out = nullptr, *d_tmp = nullptr;
	size_t cap_in = 0, cap_out = 0, cap_tmp = 0;
	uint32_t *d_flag = nullptr;
	uint32_t *h_flag = nullptr;	// pinned
	void *h_small = nullptr;		// pinned scratch for single-block calls (512 bytes)

	bool init() {
		if (ready) {
			int cur = -1;
			if (cudaGetDevice(&cur) == cudaSuccess && cur == device)
				return true;
			set_error("detex_b200: calling thread switched CUDA device (%d -> %d); the host-pointer API keeps "
				"one device per thread", device, cur);
			return false;
		}
		if (failed)
			return false;
		int n = 0;
		cudaError_t e = cudaGetDeviceCount(&n);
		if (e != cudaSuccess || n <= 0) {
			set_error("detex_b200: no usable CUDA device (%s); this library has no CPU fallback",
				e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
			failed = true;
			return false;
		}
		cudaGetDevice(&device);
		for (int i = 0; i < N_STREAMS; i++)
			if ((e = cudaStreamCreateWithFlags(&streams[i], cudaStreamNonBlocking)) != cudaSuccess) {
				cuda_fail("cudaStreamCreate", e);
				failed = true;
				return false;
			}
		if ((e = cudaMalloc((void **)&d_flag, 4)) != cudaSuccess ||
		    (e = cudaMallocHost((void **)&h_flag, 4)) != cudaSuccess ||
		    (e = cudaMallocHost(&h_small, 512)) != cudaSuccess) {
			cuda_fail("context allocation", e);
			failed = true;
			return false;
		}
		ready = true;
		return true;
	}

	bool reserve(void **p, size_t *cap, size_t need) {
		if (need <= *cap)
			return true;
		if (*p)
			cudaFree(*p);
		*p = nullptr;
		*cap = 0;
		const size_t want = need + need / 8 + 4096;
		cudaError_t e = cudaMalloc(p, want);
		if (e != cudaSuccess) {
			cuda_fail("cudaMalloc", e);
			return false;
		}
		*cap = want;
		return true;
	}

	~Context() {
		// Process teardown order is unknowable (the CUDA runtime may already be gone);
		// leak the handful of per-thread resources rather than risk calling into it.
	}
};

thread_local Context g_ctx;

bool is_compressed_format(uint32_t tf) {
	const uint32_t idx = tf >> 24;
	return idx >= 1 && idx <= 19;
}

// Enqueue decode (+ optional conversion through d_tmp) of a block range on a stream.
// hdr: the decoded pixels are taken as the _HDR twin of the native format (detexb200DecodeHDRDevice):
// the conversion applies the calling thread's HDR parameters, and an invalid block leaves the
// conversion of zero pixels, as the reference's decode + detexConvertPixels sequence does.
constexpr uint32_t PIXEL_FORMAT_HDR_BIT = 0x4000;	// detex.h:123
int enqueue_decode(uint32_t tf, const void *d_blocks, int width, int height, int wib, int hib, void *d_pixels,
uint32_t pf, int layout, uint32_t mode_mask, uint32_t flags, uint32_t *d_invalid, void *d_tmp, uint32_t *d_map,
cudaStream_t stream, bool hdr = false) {
	const uint32_t native = hdr ? ((tf & 0xFFFF) | PIXEL_FORMAT_HDR_BIT) : (tf & 0xFFFF);
	if (hdr)
		d_map = nullptr;
	DecodeArgs a;
	a.blocks = d_blocks;
	a.width = width;
	a.height = height;
	a.wib = wib;
	a.hib = hib;
	a.layout = layout;
	a.mode_mask = mode_mask;
	a.flags = flags;
	a.invalid = d_invalid;
	a.post_op = 0;
	a.post_class = 0;
	a.invalid_map = nullptr;
	a.tone_min = 0.0f;
	a.tone_max = 1.0f;
	a.tone_inv_gamma = 1.0f;
	a.tone_gamma_lo = 0.0f;
	a.tone_gamma_factor = 1.0f;
	a.tone_zero_target = hdr ? 0u : 1u;
	cudaError_t e;
	// Conversions the kernels do while storing: R<->B swap, sign flip, dropping the fourth component,
	// 16 -> 8 bit.  Repacked tiles need a 32-byte aligned target (else: separate conversion pass).
	const FusedConversion fuse = convert_fused(native, pf);
	if (fuse.ok && (fuse.post_class == 0 || (((uintptr_t)d_pixels) & 31u) == 0)) {
		a.post_op = fuse.post_op;
		a.post_class = fuse.post_class;
		if (fuse.tone_hdr) {
			// which variant of the tone kernels (tone.cuh): plain, HDR range, HDR range + gamma
			float gamma;
			get_hdr_parameters(&gamma, &a.tone_min, &a.tone_max);
			if (gamma != 1.0f) {
				// corrected range as hdr.c:149-158 computes it (host powf, sign kept)
				auto spow = [](float v, float ex) { return v >= 0.0f ? powf(v, ex) : -powf(-v, ex); };
				a.tone_inv_gamma = 1.0f / gamma;
				a.tone_gamma_lo = spow(a.tone_min, a.tone_inv_gamma);
				a.tone_gamma_factor = 1.0f / (spow(a.tone_max, a.tone_inv_gamma) - a.tone_gamma_lo);
				a.post_class += 6;
			}
			else if (!(a.tone_min == 0.0f && a.tone_max == 1.0f))	// hdr.c:124-131
				a.post_class += 3;
		}
		a.pixels = d_pixels;
		if ((e = launch_decode((int)(tf >> 24), a, stream)) != cudaSuccess)
			return cuda_fail("decode kernel launch", e);
		return DETEXB200_OK;
	}
	if (!convert_supported(native, pf) || d_tmp == nullptr) {
		set_error("detexConvertPixels: Unable to find conversion path (0x%04X -> 0x%04X on device)", native, pf);
		return DETEXB200_EUNSUPPORTED;
	}
	// Separate conversion pass.  The conversion of a zero pixel need not be zero (sign flips, forced
	// alpha), but an invalid block leaves zeros in the TARGET format (texture.c:90-93): the decode
	// kernels record invalid blocks in a bitmap and their tiles are cleared after the conversion.
	const uint64_t n_blocks = (uint64_t)wib * hib;
	a.pixels = d_tmp;
	a.invalid_map = d_map;
	if (d_map && (e = cudaMemsetAsync(d_map, 0, (size_t)((n_blocks + 31) / 32) * 4, stream)) != cudaSuccess)
		return cuda_fail("cudaMemsetAsync", e);
	if ((e = launch_decode((int)(tf >> 24), a, stream)) != cudaSuccess)
		return cuda_fail("decode kernel launch", e);
	const uint64_t n_px = layout == DETEXB200_LAYOUT_LINEAR ? (uint64_t)width * height : n_blocks * 16;
	if ((e = launch_convert(d_tmp, n_px, native, d_pixels, pf, stream)) != cudaSuccess)
		return cuda_fail("convert kernel launch", e);
	if (d_map && (e = launch_zero_invalid(d_map, d_pixels, width, height, wib, hib, layout, pixel_size(pf), stream)) != cudaSuccess)
		return cuda_fail("zero_invalid kernel launch", e);
	return DETEXB200_OK;
}

bool validate_decode_args(const char *who, uint32_t tf, const void *in, const void *out, int wib, int hib,
int width, int height, int layout) {
	if (!is_compressed_format(tf)) {
		set_error("%s: texture format 0x%08X is not a supported compressed format", who, tf);
		return false;
	}
	if (wib < 0 || hib < 0 || (uint64_t)wib * (uint64_t)hib > 0xFFFFFFFFull) {
		set_error("%s: invalid size in blocks %d x %d", who, wib, hib);
		return false;
	}
	if (layout != DETEXB200_LAYOUT_TILED && layout != DETEXB200_LAYOUT_LINEAR) {
		set_error("%s: invalid layout %d", who, layout);
		return false;
	}
	if (layout == DETEXB200_LAYOUT_LINEAR && (width < 0 || height < 0 || width > wib * 4 || height > hib * 4)) {
		set_error("%s: image size %d x %d does not fit %d x %d blocks", who, width, height, wib, hib);
		return false;
	}
	if ((uint64_t)wib * hib != 0 && (in == nullptr || out == nullptr)) {
		set_error("%s: null buffer", who);
		return false;
	}
	return true;
}

// Host-pointer texture decode, pipelined over block-row slabs on N_STREAMS streams:
// slab k: H2D(blocks) -> kernel -> D2H(pixels), all asynchronous on stream k % N.
// Pinned host buffers are DMA'd directly; pageable ones go through the driver's staging.
bool decode_texture_host(const detexTexture *t, uint8_t *pixel_buffer, uint32_t pf, int layout, const char *who) {
	Context &c = g_ctx;
	if (!c.init())
		return false;
	const uint32_t tf = t->format;
	const int wib = t->width_in_blocks, hib = t->height_in_blocks;
	if (!validate_decode_args(who, tf, t->data, pixel_buffer, wib, hib, t->width, t->height, layout))
		return false;
	const uint64_t n_blocks = (uint64_t)wib * hib;
	if (n_blocks == 0)
		return true;
	const uint32_t native = tf & 0xFFFF;
	if (pf != native && !convert_supported(native, pf)) {
		set_error("detexConvertPixels: Unable to find conversion path (0x%04X -> 0x%04X)", native, pf);
		return false;
	}
	const int bb = block_bytes(tf), ps = pixel_size(pf), ps_native = pixel_size(native);
	const size_t in_bytes = n_blocks * bb;
	const size_t out_bytes = layout == DETEXB200_LAYOUT_LINEAR ? (size_t)t->width * t->height * ps : n_blocks * 16 * ps;
	const FusedConversion fuse = convert_fused(native, pf);
	const size_t tmp_bytes = (fuse.ok && fuse.post_class == 0) ? 0 : ((n_blocks * 16 * ps_native + 255) & ~(size_t)255);
	// behind the scratch pixels: the invalid-block bitmaps of the slabs (slab k of blocks [b, b + n)
	// uses the words from b / 32 + k on; consecutive slabs never overlap)
	const size_t map_bytes = tmp_bytes ? (n_blocks / 32 + n_blocks / 1024 + 4096) * 4 : 0;
	if (!c.reserve(&c.d_in, &c.cap_in, in_bytes) || !c.reserve(&c.d_out, &c.cap_out, out_bytes + 16) ||
	    (tmp_bytes && !c.reserve(&c.d_tmp, &c.cap_tmp, tmp_bytes + map_bytes)))
		return false;
	uint32_t *const map_base = tmp_bytes ? (uint32_t *)((uint8_t *)c.d_tmp + tmp_bytes) : nullptr;
	cudaError_t e;
	if ((e = cudaMemsetAsync(c.d_flag, 0, 4, c.streams[0])) != cudaSuccess ||
	    (e = cudaStreamSynchronize(c.streams[0])) != cudaSuccess) {
		cuda_fail(who, e);
		return false;
	}
	// Slabs of about 16 MiB of output each: whole block rows for LINEAR (at least one row), any
	// run of consecutive blocks for TILED (a multiple of 1024 so that slabs stay 16-byte aligned
	// and tile-aligned for the sorted kernels).
	int rc = DETEXB200_OK;
	int k = 0;
	if (layout == DETEXB200_LAYOUT_LINEAR) {
		const size_t out_row_bytes = (size_t)t->width * 4 * ps;
		int rows_per_slab = (int)((16u << 20) / (out_row_bytes ? out_row_bytes : 1));
		if (rows_per_slab < 1)
			rows_per_slab = 1;
		for (int r0 = 0; r0 < hib && rc == DETEXB200_OK; r0 += rows_per_slab, k++) {
			const int rows = (hib - r0) < rows_per_slab ? (hib - r0) : rows_per_slab;
			cudaStream_t s = c.streams[k % N_STREAMS];
			const size_t in_off = (size_t)r0 * wib * bb, in_len = (size_t)rows * wib * bb;
			int slab_h = t->height - r0 * 4;
			if (slab_h > rows * 4)
				slab_h = rows * 4;
			if (slab_h < 0)
				slab_h = 0;
			const size_t out_off = (size_t)r0 * 4 * t->width * ps, out_len = (size_t)slab_h * t->width * ps;
			const size_t tmp_off = (size_t)r0 * 4 * t->width * ps_native;
			if ((e = cudaMemcpyAsync((uint8_t *)c.d_in + in_off, t->data + in_off, in_len, cudaMemcpyHostToDevice, s)) != cudaSuccess) {
				rc = cuda_fail("host to device copy", e);
				break;
			}
			rc = enqueue_decode(tf, (uint8_t *)c.d_in + in_off, t->width, slab_h, wib, rows, (uint8_t *)c.d_out + out_off, pf,
				layout, DETEX_MODE_MASK_ALL, 0, c.d_flag, tmp_bytes ? (uint8_t *)c.d_tmp + tmp_off : nullptr,
				map_base ? map_base + ((size_t)r0 * wib) / 32 + k : nullptr, s);
			if (rc != DETEXB200_OK)
				break;
			if (out_len && (e = cudaMemcpyAsync(pixel_buffer + out_off, (uint8_t *)c.d_out + out_off, out_len,
					cudaMemcpyDeviceToHost, s)) != cudaSuccess) {
				rc = cuda_fail("device to host copy", e);
				break;
			}
		}
	}
	else {
		uint64_t per_slab = ((16u << 20) / (16u * ps)) & ~1023ull;
		if (per_slab < 1024)
			per_slab = 1024;
		for (uint64_t b0 = 0; b0 < n_blocks && rc == DETEXB200_OK; b0 += per_slab, k++) {
			const uint64_t count = (n_blocks - b0) < per_slab ? (n_blocks - b0) : per_slab;
			cudaStream_t s = c.streams[k % N_STREAMS];
			const size_t in_off = (size_t)b0 * bb, out_off = (size_t)b0 * 16 * ps, tmp_off = (size_t)b0 * 16 * ps_native;
			if ((e = cudaMemcpyAsync((uint8_t *)c.d_in + in_off, t->data + in_off, (size_t)count * bb, cudaMemcpyHostToDevice, s)) != cudaSuccess) {
				rc = cuda_fail("host to device copy", e);
				break;
			}
			rc = enqueue_decode(tf, (uint8_t *)c.d_in + in_off, (int)count * 4, 4, (int)count, 1, (uint8_t *)c.d_out + out_off, pf,
				layout, DETEX_MODE_MASK_ALL, 0, c.d_flag, tmp_bytes ? (uint8_t *)c.d_tmp + tmp_off : nullptr,
				map_base ? map_base + b0 / 32 + k : nullptr, s);
			if (rc != DETEXB200_OK)
				break;
			if ((e = cudaMemcpyAsync(pixel_buffer + out_off, (uint8_t *)c.d_out + out_off, (size_t)count * 16 * ps,
					cudaMemcpyDeviceToHost, s)) != cudaSuccess) {
				rc = cuda_fail("device to host copy", e);
				break;
			}
		}
	}
	for (int i = 0; i < N_STREAMS; i++)
		if ((e = cudaStreamSynchronize(c.streams[i])) != cudaSuccess && rc == DETEXB200_OK)
			rc = cuda_fail(who, e);
	if (rc != DETEXB200_OK)
		return false;
	if ((e = cudaMemcpyAsync(c.h_flag, c.d_flag, 4, cudaMemcpyDeviceToHost, c.streams[0])) != cudaSuccess ||
	    (e = cudaStreamSynchronize(c.streams[0])) != cudaSuccess) {
		cuda_fail(who, e);
		return false;
	}
	if (*c.h_flag) {
		// Same text as the reference leaves behind for the last failing block (texture.c:63-64).
		set_error("detexDecompressBlock: Decompress function for format 0x%08X returned error", tf);
		return false;
	}
	return true;
}

// One block through the device (per-block API).  Output undefined-by-contract on failure;
// this implementation leaves zero pixels there.
bool decode_one_block(uint32_t tf, const uint8_t *bitstring, uint32_t mode_mask, uint32_t flags,
uint8_t *pixel_buffer, uint32_t pf, bool set_message) {
	Context &c = g_ctx;
	if (!c.init())
		return false;
	const uint32_t native = tf & 0xFFFF;
	if (pf != native && !convert_supported(native, pf)) {
		set_error("detexConvertPixels: Unable to find conversion path (0x%04X -> 0x%04X)", native, pf);
		return false;
	}
	const int bb = block_bytes(tf), out_bytes = 16 * pixel_size(pf);
	if (!c.reserve(&c.d_in, &c.cap_in, 64) || !c.reserve(&c.d_out, &c.cap_out, 512) || !c.reserve(&c.d_tmp, &c.cap_tmp, 1024))
		return false;
	cudaStream_t s = c.streams[0];
	uint8_t *hs = (uint8_t *)c.h_small;
	memcpy(hs, bitstring, bb);
	cudaError_t e;
	if ((e = cudaMemsetAsync(c.d_flag, 0, 4, s)) != cudaSuccess ||
	    (e = cudaMemcpyAsync(c.d_in, hs, bb, cudaMemcpyHostToDevice, s)) != cudaSuccess) {
		cuda_fail("detexDecompressBlock", e);
		return false;
	}
	if (enqueue_decode(tf, c.d_in, 4, 4, 1, 1, c.d_out, pf, DETEXB200_LAYOUT_TILED, mode_mask, flags, c.d_flag, c.d_tmp,
			(uint32_t *)((uint8_t *)c.d_tmp + 512), s) != DETEXB200_OK)
		return false;
	if ((e = cudaMemcpyAsync(hs + 64, c.d_out, out_bytes, cudaMemcpyDeviceToHost, s)) != cudaSuccess ||
	    (e = cudaMemcpyAsync(c.h_flag, c.d_flag, 4, cudaMemcpyDeviceToHost, s)) != cudaSuccess ||
	    (e = cudaStreamSynchronize(s)) != cudaSuccess) {
		cuda_fail("detexDecompressBlock", e);
		return false;
	}
	memcpy(pixel_buffer, hs + 64, out_bytes);
	if (*c.h_flag) {
		if (set_message)
			set_error("detexDecompressBlock: Decompress function for format 0x%08X returned error", tf);
		return false;
	}
	return true;
}

}  // namespace

// =============================================================================================
// include/detex_b200.h
// =============================================================================================
extern "C" {

static int decode_device(const char *who, uint32_t tf, const void *d_blocks, int width, int height, int wib, int hib,
void *d_pixels, uint32_t pf, int layout, uint32_t mode_mask, uint32_t flags, uint32_t *d_invalid, void *stream, bool hdr) {
	if (!validate_decode_args(who, tf, d_blocks, d_pixels, wib, hib, width, height, layout))
		return DETEXB200_EINVAL;
	if (((uintptr_t)d_blocks & 15u) != 0) {
		set_error("%s: block array must be 16-byte aligned", who);
		return DETEXB200_EINVAL;
	}
	const uint32_t native = hdr ? ((tf & 0xFFFF) | PIXEL_FORMAT_HDR_BIT) : (tf & 0xFFFF);
	void *tmp = nullptr;
	cudaStream_t s = (cudaStream_t)stream;
	if (pf != native && !convert_supported(native, pf)) {
		set_error("detexConvertPixels: Unable to find conversion path (0x%04X -> 0x%04X on device)", native, pf);
		return DETEXB200_EUNSUPPORTED;
	}
	const FusedConversion fuse = convert_fused(native, pf);
	if (fuse.ok && (fuse.post_class == 0 || (((uintptr_t)d_pixels) & 31u) == 0))
		return enqueue_decode(tf, d_blocks, width, height, wib, hib, d_pixels, pf, layout, mode_mask, flags, d_invalid, nullptr, nullptr, s, hdr);
	// Decode + conversion pass, slab by slab through ONE scratch buffer small enough to stay in the
	// 126 MB L2: the native-format pixels of a slab are read back by the conversion kernel before
	// they are evicted, so HBM sees the compressed input and the converted output only.
	const int bb = block_bytes(tf), ps = pixel_size(pf), ps_native = pixel_size(native);
	const size_t slab_budget = 24u << 20;
	const size_t pixel_region = (slab_budget + (size_t)wib * 64 * ps_native + 1024 * 16 * ps_native + 255) & ~(size_t)255;
	const size_t map_region = (pixel_region / (16 * ps_native) / 32 + 64) * 4;	// one bit per block of a slab
	cudaError_t e = cudaMallocAsync(&tmp, pixel_region + map_region, s);
	if (e != cudaSuccess)
		return cuda_fail("cudaMallocAsync", e);
	uint32_t *const map = (uint32_t *)((uint8_t *)tmp + pixel_region);
	int rc = DETEXB200_OK;
	if (layout == DETEXB200_LAYOUT_LINEAR) {
		const size_t row_bytes = (size_t)width * 4 * ps_native;
		int rows_per_slab = (int)(slab_budget / (row_bytes ? row_bytes : 1));
		if (rows_per_slab < 1)
			rows_per_slab = 1;
		for (int r0 = 0; r0 < hib && rc == DETEXB200_OK; r0 += rows_per_slab) {
			const int rows = (hib - r0) < rows_per_slab ? (hib - r0) : rows_per_slab;
			int slab_h = height - r0 * 4;
			if (slab_h > rows * 4)
				slab_h = rows * 4;
			if (slab_h < 0)
				slab_h = 0;
			rc = enqueue_decode(tf, (const uint8_t *)d_blocks + (size_t)r0 * wib * bb, width, slab_h, wib, rows,
				(uint8_t *)d_pixels + (size_t)r0 * 4 * width * ps, pf, layout, mode_mask, flags, d_invalid, tmp, map, s, hdr);
		}
	}
	else {
		const uint64_t n_blocks = (uint64_t)wib * hib;
		uint64_t per_slab = (slab_budget / (16u * ps_native)) & ~1023ull;
		if (per_slab < 1024)
			per_slab = 1024;
		for (uint64_t b0 = 0; b0 < n_blocks && rc == DETEXB200_OK; b0 += per_slab) {
			const uint64_t count = (n_blocks - b0) < per_slab ? (n_blocks - b0) : per_slab;
			rc = enqueue_decode(tf, (const uint8_t *)d_blocks + b0 * bb, (int)count * 4, 4, (int)count, 1,
				(uint8_t *)d_pixels + b0 * 16 * ps, pf, layout, mode_mask, flags, d_invalid, tmp, map, s, hdr);
		}
	}
	cudaFreeAsync(tmp, s);
	return rc;
}

int detexb200DecodeDevice(uint32_t tf, const void *d_blocks, int width, int height, int wib, int hib,
void *d_pixels, uint32_t pf, int layout, uint32_t mode_mask, uint32_t flags, uint32_t *d_invalid, void *stream) {
	return decode_device("detexb200DecodeDevice", tf, d_blocks, width, height, wib, hib, d_pixels, pf, layout, mode_mask,
		flags, d_invalid, stream, false);
}

int detexb200DecodeHDRDevice(uint32_t tf, const void *d_blocks, int width, int height, int wib, int hib,
void *d_pixels, uint32_t pf, int layout, float gamma, float range_min, float range_max, uint32_t *d_invalid,
void *stream) {
	const uint32_t native = tf & 0xFFFF;
	// the formats that have an _HDR twin in the conversion table: FLOAT_RGBX16 (BC6H unsigned)
	if (!is_compressed_format(tf) || native != 0x2721u) {
		set_error("detexb200DecodeHDRDevice: texture format 0x%08X does not decode to an HDR pixel format", tf);
		return DETEXB200_EINVAL;
	}
	float g0, lo0, hi0;
	get_hdr_parameters(&g0, &lo0, &hi0);
	set_hdr_parameters(gamma, range_min, range_max);
	const int rc = decode_device("detexb200DecodeHDRDevice", tf, d_blocks, width, height, wib, hib, d_pixels, pf, layout,
		0xFFFFFFFFu, 0u, d_invalid, stream, true);
	set_hdr_parameters(g0, lo0, hi0);
	return rc;
}

int detexb200ConversionPlan(uint32_t spf, uint32_t tpf, uint32_t formats_out[5]) {
	int steps[4];
	const int n = convert_plan(spf, tpf, steps);
	if (n < 0)
		return -1;
	formats_out[0] = spf;
	for (int i = 0; i < n; i++) {
		uint32_t a, b;
		conversion_table_entry(steps[i], &a, &b);
		formats_out[i + 1] = b;
	}
	return n;
}

void detexSetHDRParameters(float gamma, float range_min, float range_max) {
	set_hdr_parameters(gamma, range_min, range_max);
}

int detexb200ConvertPixelsDevice(const void *d_src, uint64_t n, uint32_t spf, void *d_dst, uint32_t tpf, void *stream) {
	if (n && (!d_src || !d_dst)) {
		set_error("detexb200ConvertPixelsDevice: null buffer");
		return DETEXB200_EINVAL;
	}
	if (!convert_supported(spf, tpf)) {
		set_error("detexConvertPixels: Unable to find conversion path (0x%04X -> 0x%04X on device)", spf, tpf);
		return DETEXB200_EUNSUPPORTED;
	}
	cudaError_t e = launch_convert(d_src, n, spf, d_dst, tpf, (cudaStream_t)stream);
	return e == cudaSuccess ? DETEXB200_OK : cuda_fail("convert kernel launch", e);
}

int detexb200GenerateBlocksDevice(void *d_blocks, uint64_t n_blocks, int bb, uint64_t seed, uint64_t first_block,
int force_modes, void *stream) {
	cudaError_t e = launch_generate(d_blocks, n_blocks, bb, seed, first_block, force_modes, (cudaStream_t)stream);
	if (e == cudaErrorInvalidValue) {
		set_error("detexb200GenerateBlocksDevice: invalid argument (sizes must be multiples of 16 bytes; "
			"mode forcing needs 16-byte blocks)");
		return DETEXB200_EINVAL;
	}
	return e == cudaSuccess ? DETEXB200_OK : cuda_fail("generator kernel launch", e);
}

int detexb200HashDevice(const void *d_buf, uint64_t n_bytes, uint64_t *h_hash_out, void *stream) {
	if (!h_hash_out || (n_bytes && !d_buf)) {
		set_error("detexb200HashDevice: null pointer");
		return DETEXB200_EINVAL;
	}
	cudaError_t e = device_hash(d_buf, n_bytes, h_hash_out, (cudaStream_t)stream);
	return e == cudaSuccess ? DETEXB200_OK : cuda_fail("device hash", e);
}

int detexb200DeviceCount(void) {
	int n = 0;
	return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

const char *detexb200Version(void) { return "detex_b200 0.1 (sm_100a; API of hglm/detex v0.1.2)"; }

uint64_t detexb200LaunchCount(void) { return launch_count(); }

// =============================================================================================
// include/detex.h  (reference API for the hot path)
// =============================================================================================
const char *detexGetErrorMessage(void) { return g_error.c_str(); }

#define DETEX_DEFINE_BLOCK_DECODER(NAME) \
	bool detexDecompressBlock##NAME(const uint8_t *bitstring, uint32_t mode_mask, uint32_t flags, \
	uint8_t *pixel_buffer) { \
		return decode_one_block(DETEX_TEXTURE_FORMAT_##NAME, bitstring, mode_mask, flags, pixel_buffer, \
			DETEX_TEXTURE_FORMAT_##NAME & 0xFFFF, false); \
	}
DETEX_DEFINE_BLOCK_DECODER(BC1) DETEX_DEFINE_BLOCK_DECODER(BC1A) DETEX_DEFINE_BLOCK_DECODER(BC2)
DETEX_DEFINE_BLOCK_DECODER(BC3) DETEX_DEFINE_BLOCK_DECODER(RGTC1) DETEX_DEFINE_BLOCK_DECODER(SIGNED_RGTC1)
DETEX_DEFINE_BLOCK_DECODER(RGTC2) DETEX_DEFINE_BLOCK_DECODER(SIGNED_RGTC2) DETEX_DEFINE_BLOCK_DECODER(BPTC_FLOAT)
DETEX_DEFINE_BLOCK_DECODER(BPTC_SIGNED_FLOAT) DETEX_DEFINE_BLOCK_DECODER(BPTC) DETEX_DEFINE_BLOCK_DECODER(ETC1)
DETEX_DEFINE_BLOCK_DECODER(ETC2) DETEX_DEFINE_BLOCK_DECODER(ETC2_PUNCHTHROUGH) DETEX_DEFINE_BLOCK_DECODER(ETC2_EAC)
DETEX_DEFINE_BLOCK_DECODER(EAC_R11) DETEX_DEFINE_BLOCK_DECODER(EAC_SIGNED_R11) DETEX_DEFINE_BLOCK_DECODER(EAC_RG11)
DETEX_DEFINE_BLOCK_DECODER(EAC_SIGNED_RG11)
#undef DETEX_DEFINE_BLOCK_DECODER

// texture.c:55-70
bool detexDecompressBlock(const uint8_t *bitstring, uint32_t texture_format, uint32_t mode_mask, uint32_t flags,
uint8_t *pixel_buffer, uint32_t pixel_format) {
	if (!is_compressed_format(texture_format)) {
		// The reference indexes past its function table here (format index 0 or 20, detex.h:610).
		set_error("detexDecompressBlock: no decompress function for format 0x%08X", texture_format);
		return false;
	}
	return decode_one_block(texture_format, bitstring, mode_mask, flags, pixel_buffer, pixel_format, true);
}

// texture.c:77-98
bool detexDecompressTextureTiled(const detexTexture *texture, uint8_t *pixel_buffer, uint32_t pixel_format) {
	if (!detexFormatIsCompressed(texture->format)) {
		set_error("detexDecompressTextureTiled: Cannot handle uncompressed texture format");
		return false;
	}
	return decode_texture_host(texture, pixel_buffer, pixel_format, DETEXB200_LAYOUT_TILED, "detexDecompressTextureTiled");
}

// texture.c:105-145
bool detexDecompressTextureLinear(const detexTexture *texture, uint8_t *pixel_buffer, uint32_t pixel_format) {
	if (!detexFormatIsCompressed(texture->format))
		return detexConvertPixels(texture->data, (uint32_t)(texture->width * texture->height),
			detexGetPixelFormat(texture->format), pixel_buffer, pixel_format);
	return decode_texture_host(texture, pixel_buffer, pixel_format, DETEXB200_LAYOUT_LINEAR, "detexDecompressTextureLinear");
}

// convert.c:1082-1166
bool detexConvertPixels(uint8_t *src, uint32_t n, uint32_t spf, uint8_t *dst, uint32_t tpf) {
	if (spf == tpf) {				// convert.c:1087-1092
		memcpy(dst, src, (size_t)n * pixel_size(spf));
		return true;
	}
	if (!convert_supported(spf, tpf)) {
		set_error("detexConvertPixels: Unable to find conversion path");	// convert.c:1096
		return false;
	}
	Context &c = g_ctx;
	if (!c.init())
		return false;
	const size_t in_bytes = (size_t)n * pixel_size(spf), out_bytes = (size_t)n * pixel_size(tpf);
	if (!c.reserve(&c.d_tmp, &c.cap_tmp, in_bytes + 16) || !c.reserve(&c.d_out, &c.cap_out, out_bytes + 16))
		return false;
	cudaStream_t s = c.streams[0];
	cudaError_t e;
	if ((e = cudaMemcpyAsync(c.d_tmp, src, in_bytes, cudaMemcpyHostToDevice, s)) != cudaSuccess ||
	    (e = launch_convert(c.d_tmp, n, spf, c.d_out, tpf, s)) != cudaSuccess ||
	    (e = cudaMemcpyAsync(dst, c.d_out, out_bytes, cudaMemcpyDeviceToHost, s)) != cudaSuccess ||
	    (e = cudaStreamSynchronize(s)) != cudaSuccess) {
		cuda_fail("detexConvertPixels", e);
		return false;
	}
	return true;
}

// convert.c:1168-1171 (+ :1100-1113): only conversions that keep the pixel size.
bool detexConvertPixelsInPlace(uint8_t *buf, uint32_t n, uint32_t spf, uint32_t tpf) {
	if (spf == tpf)
		return true;
	int steps[4];
	const int n_steps = convert_plan(spf, tpf, steps);
	if (n_steps < 0) {
		set_error("detexConvertPixels: Unable to find conversion path");	// convert.c:1096
		return false;
	}
	for (int i = 0; i < n_steps; i++) {		// every step of the chain has to keep the pixel size
		uint32_t a, b;
		conversion_table_entry(steps[i], &a, &b);
		if (pixel_size(a) != pixel_size(b)) {
			set_error("Unable to find in-place conversion path");		// convert.c:1112
			return false;
		}
	}
	return detexConvertPixels(buf, n, spf, buf, tpf);
}

}  // extern "C"
