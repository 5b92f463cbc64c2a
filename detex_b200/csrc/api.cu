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
#include <atomic>
#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <thread>
#include <string>
#include <vector>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

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

// ---- lanes: one stream + its grow-only buffers ----------------------------------------------
// A lane is the unit of the host-pointer pipeline: a slab of the texture goes host -> d_in -> kernel -> d_out
// -> host on the lane's stream; work on one lane is ordered by the stream, so its buffers are reused slab after
// slab without further synchronisation.  The calling thread owns N_STREAMS lanes on its current device (pinned
// caller buffers: the DMA engines read and write them directly); every thread of the copy pool owns one lane per
// device it has worked on (pageable caller buffers and multi-device calls, below).
constexpr int N_STREAMS = 4;
constexpr size_t SLAB_BYTES = 16u << 20;	// output bytes per slab, pinned single-device path
constexpr size_t POOL_SLAB_BYTES = 4u << 20;	// ... copy-pool path (more, smaller jobs: smoother tail)
constexpr size_t POOL_MIN_BYTES = 4u << 20;	// smaller textures are not worth a thread hop

std::atomic<bool> g_exiting{false};		// set by an atexit handler: no CUDA calls from destructors after it

struct Lane {
	int device = -1;
	cudaStream_t stream = nullptr;
	cudaEvent_t done = nullptr;		// blocking-sync event: pool threads sleep instead of spinning
	void *d_in = nullptr, *d_out = nullptr, *d_tmp = nullptr;
	size_t cap_in = 0, cap_out = 0, cap_tmp = 0;
	void *h_in = nullptr, *h_out = nullptr;	// pinned bounce buffers (pageable caller memory)
	size_t cap_hin = 0, cap_hout = 0;
	uint32_t *d_flag = nullptr;
	uint32_t *h_flag = nullptr;		// pinned

	bool init() {				// on the calling thread's current device
		if (stream)
			return true;
		cudaError_t e;
		cudaGetDevice(&device);
		if ((e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking)) != cudaSuccess ||
		    (e = cudaEventCreateWithFlags(&done, cudaEventBlockingSync | cudaEventDisableTiming)) != cudaSuccess ||
		    (e = cudaMalloc((void **)&d_flag, 4)) != cudaSuccess ||
		    (e = cudaMallocHost((void **)&h_flag, 4)) != cudaSuccess) {
			cuda_fail("lane allocation", e);
			release();
			return false;
		}
		return true;
	}
	bool reserve(void **p, size_t *cap, size_t need) {
		if (need <= *cap)
			return true;
		if (*p)
			cudaFree(*p);		// (synchronises with whatever still uses the old buffer)
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
	bool reserve_host(void **p, size_t *cap, size_t need) {
		if (need <= *cap)
			return true;
		if (*p) {
			cudaStreamSynchronize(stream);
			cudaFreeHost(*p);
		}
		*p = nullptr;
		*cap = 0;
		cudaError_t e = cudaMallocHost(p, need + 4096);
		if (e != cudaSuccess) {
			cuda_fail("cudaMallocHost", e);
			return false;
		}
		*cap = need + 4096;
		return true;
	}
	void release() {
		if (g_exiting.load())
			return;			// the runtime may be gone: leak rather than call into it
		if (stream)
			cudaStreamSynchronize(stream);
		cudaFree(d_in); cudaFree(d_out); cudaFree(d_tmp); cudaFree(d_flag);
		cudaFreeHost(h_in); cudaFreeHost(h_out); cudaFreeHost(h_flag);
		if (done)
			cudaEventDestroy(done);
		if (stream)
			cudaStreamDestroy(stream);
		(void)cudaGetLastError();	// (cudaErrorCudartUnloading and friends are not the caller's problem)
		*this = Lane();
	}
};

// ---- per-thread device context ----------------------------------------------------------------
struct Context {
	bool ready = false, failed = false;
	int device = -1;			// the context (lanes) belongs to this device
	Lane lanes[N_STREAMS];
	std::map<int, std::vector<Lane>> rings;	// pool path: per device, the lanes whose slabs are in flight
	void *h_small = nullptr;		// pinned scratch for single-block calls (512 bytes)

	bool init() {
		if (ready) {
			int cur = -1;
			if (cudaGetDevice(&cur) == cudaSuccess && cur == device)
				return true;
			set_error("detex_b200: calling thread switched CUDA device (%d -> %d); the host-pointer API keeps "
				"one device per thread (detexb200ReleaseThreadResources() lets the thread start over)", device, cur);
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
		static std::once_flag exit_hook;
		std::call_once(exit_hook, [] { atexit([] { g_exiting.store(true); }); });
		cudaGetDevice(&device);
		for (int i = 0; i < N_STREAMS; i++)
			if (!lanes[i].init()) {
				failed = true;
				return false;
			}
		if ((e = cudaMallocHost(&h_small, 512)) != cudaSuccess) {
			cuda_fail("context allocation", e);
			failed = true;
			return false;
		}
		ready = true;
		return true;
	}
	// Streams, device buffers and pinned buffers of this thread go back to the driver: at thread exit (the
	// destructor) or on request (detexb200ReleaseThreadResources).  Nothing is touched once process exit
	// has begun -- the order in which the CUDA runtime and this library are torn down is not ours to know.
	void release() {
		if (!ready || g_exiting.load())
			return;
		int cur = -1;
		const bool switched = cudaGetDevice(&cur) == cudaSuccess && cur != device && cudaSetDevice(device) == cudaSuccess;
		for (int i = 0; i < N_STREAMS; i++)
			lanes[i].release();
		for (auto &r : rings) {
			if (cudaSetDevice(r.first) == cudaSuccess)
				for (Lane &L : r.second)
					L.release();
		}
		rings.clear();
		cudaSetDevice(device);
		cudaFreeHost(h_small);
		h_small = nullptr;
		if (switched)
			cudaSetDevice(cur);
		(void)cudaGetLastError();
		ready = false;
	}
	~Context() { release(); }
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

// ---- host-pointer texture decode ----------------------------------------------------------------
// The texture is cut into slabs (whole block rows for LINEAR, runs of consecutive blocks for TILED); every slab
// is an independent H2D -> kernel -> D2H chain on one lane.
struct Slab {
	size_t in_off, in_len;		// bytes of the caller's block array
	size_t out_off, out_len;	// bytes of the caller's pixel buffer
	int width, height, wib, hib;	// geometry handed to the kernels
};

struct TexCall {
	uint32_t tf, pf;
	int layout;
	const uint8_t *in;
	uint8_t *out;
	int bb, ps, ps_native;
	bool needs_tmp;			// decode + separate conversion pass
	float gamma, range_min, range_max;	// HDR parameters of the calling thread (hdr.c:26-38)
};

std::vector<Slab> plan_slabs(const detexTexture *t, int layout, int bb, int ps, size_t slab_bytes) {
	std::vector<Slab> slabs;
	const int wib = t->width_in_blocks, hib = t->height_in_blocks;
	if (layout == DETEXB200_LAYOUT_LINEAR) {
		const size_t out_row_bytes = (size_t)t->width * 4 * ps;
		int rows_per_slab = (int)(slab_bytes / (out_row_bytes ? out_row_bytes : 1));
		// (a slab never holds more than 2^22 blocks, whatever the image width says)
		const int row_cap = (int)(((size_t)1 << 22) / (size_t)(wib > 0 ? wib : 1));
		if (rows_per_slab > row_cap)
			rows_per_slab = row_cap;
		if (rows_per_slab < 1)
			rows_per_slab = 1;
		for (int r0 = 0; r0 < hib; r0 += rows_per_slab) {
			const int rows = (hib - r0) < rows_per_slab ? (hib - r0) : rows_per_slab;
			int slab_h = t->height - r0 * 4;
			if (slab_h > rows * 4)
				slab_h = rows * 4;
			if (slab_h < 0)
				slab_h = 0;
			Slab s;
			s.in_off = (size_t)r0 * wib * bb;
			s.in_len = (size_t)rows * wib * bb;
			s.out_off = (size_t)r0 * 4 * t->width * ps;
			s.out_len = (size_t)slab_h * t->width * ps;
			s.width = t->width;
			s.height = slab_h;
			s.wib = wib;
			s.hib = rows;
			slabs.push_back(s);
		}
	}
	else {
		const uint64_t n_blocks = (uint64_t)wib * hib;
		// a multiple of 1024 blocks: slabs stay 16-byte aligned and tile-aligned for the sorted kernels
		uint64_t per_slab = (slab_bytes / (16u * ps)) & ~1023ull;
		if (per_slab < 1024)
			per_slab = 1024;
		for (uint64_t b0 = 0; b0 < n_blocks; b0 += per_slab) {
			const uint64_t count = (n_blocks - b0) < per_slab ? (n_blocks - b0) : per_slab;
			Slab s;
			s.in_off = (size_t)b0 * bb;
			s.in_len = (size_t)count * bb;
			s.out_off = (size_t)b0 * 16 * ps;
			s.out_len = (size_t)count * 16 * ps;
			s.width = (int)count * 4;
			s.height = 4;
			s.wib = (int)count;
			s.hib = 1;
			slabs.push_back(s);
		}
	}
	return slabs;
}

// scratch of the decode + conversion pass of one slab: native-format pixels, then one bit per block
inline size_t tmp_pixel_bytes(const TexCall &c, const Slab &s) {
	return (((size_t)s.wib * s.hib * 16 * c.ps_native) + 255) & ~(size_t)255;
}
inline size_t tmp_bytes(const TexCall &c, const Slab &s) {
	return tmp_pixel_bytes(c, s) + (((size_t)s.wib * s.hib + 31) / 32 + 64) * 4;
}

bool reserve_lane(Lane &L, const TexCall &c, size_t max_in, size_t max_out, size_t max_tmp, bool bounce_in, bool bounce_out) {
	return L.reserve(&L.d_in, &L.cap_in, max_in + 16) && L.reserve(&L.d_out, &L.cap_out, max_out + 32) &&
		(!c.needs_tmp || L.reserve(&L.d_tmp, &L.cap_tmp, max_tmp)) &&
		(!bounce_in || L.reserve_host(&L.h_in, &L.cap_hin, max_in)) &&
		(!bounce_out || L.reserve_host(&L.h_out, &L.cap_hout, max_out));
}

// H2D -> decode (+ conversion) -> D2H of one slab, enqueued on the lane's stream.  The lane's buffers have been
// reserved by the caller.
int enqueue_slab(Lane &L, const TexCall &c, const Slab &s, const void *h_src, void *h_dst) {
	cudaError_t e;
	if (s.in_len && (e = cudaMemcpyAsync(L.d_in, h_src, s.in_len, cudaMemcpyHostToDevice, L.stream)) != cudaSuccess)
		return cuda_fail("host to device copy", e);
	uint32_t *map = c.needs_tmp ? (uint32_t *)((uint8_t *)L.d_tmp + tmp_pixel_bytes(c, s)) : nullptr;
	const int rc = enqueue_decode(c.tf, L.d_in, s.width, s.height, s.wib, s.hib, L.d_out, c.pf, c.layout,
		DETEX_MODE_MASK_ALL, 0, L.d_flag, c.needs_tmp ? L.d_tmp : nullptr, map, L.stream);
	if (rc != DETEXB200_OK)
		return rc;
	if (s.out_len && (e = cudaMemcpyAsync(h_dst, L.d_out, s.out_len, cudaMemcpyDeviceToHost, L.stream)) != cudaSuccess)
		return cuda_fail("device to host copy", e);
	return DETEXB200_OK;
}

// ---- copy pool: pageable caller buffers and multi-device calls --------------------------------------
// cudaMemcpyAsync on pageable memory is staged by the driver on the calling thread, one copy at a time, and the
// pipeline above degenerates into a sequence.  Real callers of this API pass malloc'd buffers
// (detex-convert.c:303-310, ktx.c:142), so that case gets its own machinery: a process-wide pool of threads, each
// owning a lane with pinned bounce buffers.  A pool thread takes one slab at a time: memcpy blocks -> pinned,
// H2D, kernel, D2H into pinned, wait (blocking event, the thread sleeps), memcpy pinned -> caller's pixels.  With
// a handful of threads the host copies of some slabs overlap the DMA transfers of others and the PCIe link stays
// busy.  The same pool spreads ONE call over several GPUs (detexb200SetDevices / DETEXB200_DEVICES): slab k goes
// to device k mod G, every device moving its slabs over its own PCIe link.
// Copy a pinned bounce buffer into the caller's (pageable) pixel buffer without reading the destination lines
// first: plain stores cost a read-for-ownership per line on top of the write, and this copy runs beside the DMA
// engine's own writes into host memory -- the host memory system, not the PCIe link, is what a pageable call
// is bound by.  SSE2 streaming stores (baseline x86-64); head and tail through memcpy.
void stream_copy(void *dst, const void *src, size_t n) {
#if defined(__x86_64__)
	uint8_t *d = (uint8_t *)dst;
	const uint8_t *s = (const uint8_t *)src;
	if (n < 4096) {
		memcpy(d, s, n);
		return;
	}
	const size_t head = (size_t)((0 - (uintptr_t)d) & 63u);
	memcpy(d, s, head);
	d += head; s += head; n -= head;
	const size_t lines = n / 64;
	for (size_t i = 0; i < lines; i++) {
		const __m128i a = _mm_loadu_si128((const __m128i *)(s + 64 * i)), b = _mm_loadu_si128((const __m128i *)(s + 64 * i + 16));
		const __m128i c = _mm_loadu_si128((const __m128i *)(s + 64 * i + 32)), e = _mm_loadu_si128((const __m128i *)(s + 64 * i + 48));
		_mm_stream_si128((__m128i *)(d + 64 * i), a);
		_mm_stream_si128((__m128i *)(d + 64 * i + 16), b);
		_mm_stream_si128((__m128i *)(d + 64 * i + 32), c);
		_mm_stream_si128((__m128i *)(d + 64 * i + 48), e);
	}
	_mm_sfence();
	memcpy(d + 64 * lines, s + 64 * lines, n - 64 * lines);
#else
	memcpy(dst, src, n);
#endif
}

// tuning knobs of the pool path (environment, read once; see device_set())
size_t g_pool_slab_bytes = POOL_SLAB_BYTES;	// DETEXB200_POOL_SLAB_KB
bool g_pool_stream_copy = true;		// DETEXB200_POOL_COPY=memcpy selects plain memcpy out of the bounce buffers
int g_pool_ring = 48;			// DETEXB200_POOL_RING: lanes (slabs in flight) per device and calling thread.  Measured on a
					// 16-core box (BC1 8192^2, pageable): 8 -> 24 -> 48 lanes = 12.5 -> 28.6 -> 34.7 GB/s of pixels

struct CallState {
	std::mutex m;
	std::condition_variable cv;
	int pending = 0;
	int rc = DETEXB200_OK;
	bool invalid = false;
	std::string error;
};
// A slab goes through the pool twice: SUBMIT (copy the blocks into the lane's pinned buffer, enqueue H2D -> kernel ->
// D2H into the lane's pinned buffer, record the lane's event) and FINISH (wait for the event, copy the pixels into the
// caller's buffer, hand the lane back).  No thread sits idle while its slab is on the link: as many slabs as the
// caller's ring has lanes are in flight, and the pool's threads spend their time copying.
struct Job {
	enum Kind { SUBMIT, FINISH } kind;
	const TexCall *call;
	const Slab *slab;
	Lane *lane;
	bool *lane_busy;
	int device;
	bool bounce_in, bounce_out;
	CallState *state;
};

class Pool {
public:
	void ensure(int n) {
		std::lock_guard<std::mutex> g(m_);
		while ((int)threads_ < n) {
			std::thread(&Pool::worker, this).detach();
			threads_++;
		}
	}
	void submit(const Job &j) {
		{
			std::lock_guard<std::mutex> g(m_);
			q_.push_back(j);
		}
		cv_.notify_one();
	}
private:
	std::mutex m_;
	std::condition_variable cv_;
	std::deque<Job> q_;
	size_t threads_ = 0;

	static int run_submit(const Job &j) {
		const TexCall &c = *j.call;
		const Slab &s = *j.slab;
		Lane &L = *j.lane;
		cudaError_t e;
		if ((e = cudaSetDevice(j.device)) != cudaSuccess)
			return cuda_fail("cudaSetDevice", e);
		set_hdr_parameters(c.gamma, c.range_min, c.range_max);
		if ((e = cudaMemsetAsync(L.d_flag, 0, 4, L.stream)) != cudaSuccess)
			return cuda_fail("cudaMemsetAsync", e);
		const void *src = c.in + s.in_off;
		if (j.bounce_in) {
			memcpy(L.h_in, src, s.in_len);
			src = L.h_in;
		}
		void *dst = j.bounce_out ? L.h_out : (void *)(c.out + s.out_off);
		const int rc = enqueue_slab(L, c, s, src, dst);
		if (rc != DETEXB200_OK)
			return rc;
		if ((e = cudaMemcpyAsync(L.h_flag, L.d_flag, 4, cudaMemcpyDeviceToHost, L.stream)) != cudaSuccess ||
		    (e = cudaEventRecord(L.done, L.stream)) != cudaSuccess)
			return cuda_fail("slab pipeline", e);
		return DETEXB200_OK;
	}
	static int run_finish(const Job &j, bool *invalid) {
		const TexCall &c = *j.call;
		const Slab &s = *j.slab;
		Lane &L = *j.lane;
		cudaError_t e;
		if ((e = cudaEventSynchronize(L.done)) != cudaSuccess)
			return cuda_fail("slab pipeline", e);
		if (j.bounce_out) {
			if (g_pool_stream_copy)
				stream_copy(c.out + s.out_off, L.h_out, s.out_len);
			else
				memcpy(c.out + s.out_off, L.h_out, s.out_len);
		}
		*invalid = *L.h_flag != 0;
		return DETEXB200_OK;
	}

	void worker() {
		for (;;) {
			Job j;
			{
				std::unique_lock<std::mutex> g(m_);
				cv_.wait(g, [&] { return !q_.empty(); });
				j = q_.front();
				q_.pop_front();
			}
			bool invalid = false;
			int rc;
			if (j.kind == Job::SUBMIT) {
				rc = run_submit(j);
				if (rc == DETEXB200_OK) {		// second trip: behind the jobs already queued
					Job f = j;
					f.kind = Job::FINISH;
					submit(f);
					continue;
				}
			}
			else
				rc = run_finish(j, &invalid);
			CallState &st = *j.state;
			std::lock_guard<std::mutex> g(st.m);
			if (rc != DETEXB200_OK && st.rc == DETEXB200_OK) {
				st.rc = rc;
				st.error = g_error;
			}
			st.invalid = st.invalid || invalid;
			*j.lane_busy = false;
			st.pending--;
			st.cv.notify_all();		// the caller waits for free lanes and for the last slab
		}
	}
};

Pool &pool() {
	static Pool *p = new Pool;	// never destroyed: its threads outlive static destruction
	return *p;
}

// ---- device set of the host-pointer texture calls ---------------------------------------------------------
std::mutex g_devices_mutex;
std::vector<int> g_devices;		// empty: the calling thread's current device
bool g_devices_from_env = false;
int g_host_threads = 0;			// 0: default

std::vector<int> device_set() {
	std::lock_guard<std::mutex> g(g_devices_mutex);
	if (!g_devices_from_env) {
		g_devices_from_env = true;
		if (const char *t = getenv("DETEXB200_HOST_THREADS"))
			g_host_threads = atoi(t);
		if (const char *t = getenv("DETEXB200_POOL_SLAB_KB"))
			if (atoi(t) >= 64)
				g_pool_slab_bytes = (size_t)atoi(t) << 10;
		if (const char *t = getenv("DETEXB200_POOL_COPY"))
			g_pool_stream_copy = t[0] != 'm';
		if (const char *t = getenv("DETEXB200_POOL_RING"))
			if (atoi(t) >= 2 && atoi(t) <= 256)
				g_pool_ring = atoi(t);
		const char *e = getenv("DETEXB200_DEVICES");
		if (e && *e && g_devices.empty()) {
			int n = 0;
			if (cudaGetDeviceCount(&n) != cudaSuccess)
				n = 0;
			if (!strcmp(e, "all"))
				for (int i = 0; i < n; i++)
					g_devices.push_back(i);
			else if (strchr(e, ',') == nullptr && atoi(e) > 0 && e[strspn(e, "0123456789")] == 0 && atoi(e) <= n && e[0] != '0')
				for (int i = 0; i < atoi(e); i++)	// a count
					g_devices.push_back(i);
			else
				for (const char *p = e; *p;) {		// a list of ordinals
					char *end;
					const long d = strtol(p, &end, 10);
					if (end == p)
						break;
					bool seen = false;
					for (int v : g_devices)
						seen = seen || v == (int)d;
					if (d >= 0 && d < n && !seen)		// (every device once: a lane ring belongs to one device)
						g_devices.push_back((int)d);
					p = *end == ',' ? end + 1 : end;
				}
		}
	}
	return g_devices;
}

bool host_pointer_is_pinned(const void *p) {
	cudaPointerAttributes a;
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
		(void)cudaGetLastError();
		return false;
	}
	return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

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
	TexCall call;
	call.tf = tf;
	call.pf = pf;
	call.layout = layout;
	call.in = t->data;
	call.out = pixel_buffer;
	call.bb = block_bytes(tf);
	call.ps = pixel_size(pf);
	call.ps_native = pixel_size(native);
	const FusedConversion fuse = convert_fused(native, pf);
	call.needs_tmp = !fuse.ok;	// (a fused class on a misaligned target cannot happen: slabs start at cudaMalloc'd bases)
	get_hdr_parameters(&call.gamma, &call.range_min, &call.range_max);
	const size_t in_bytes = n_blocks * call.bb;
	const size_t out_bytes = layout == DETEXB200_LAYOUT_LINEAR ? (size_t)t->width * t->height * call.ps : n_blocks * 16 * call.ps;

	const std::vector<int> devices = device_set();
	const bool pinned_in = host_pointer_is_pinned(t->data), pinned_out = host_pointer_is_pinned(pixel_buffer);
	const bool use_pool = in_bytes + out_bytes >= POOL_MIN_BYTES && (!devices.empty() || !pinned_in || !pinned_out);
	const std::vector<Slab> slabs = plan_slabs(t, layout, call.bb, call.ps, use_pool ? g_pool_slab_bytes : SLAB_BYTES);
	size_t max_in = 0, max_out = 0, max_tmp = 0;
	for (const Slab &s : slabs) {
		max_in = s.in_len > max_in ? s.in_len : max_in;
		max_out = s.out_len > max_out ? s.out_len : max_out;
		const size_t tb = call.needs_tmp ? tmp_bytes(call, s) : 0;
		max_tmp = tb > max_tmp ? tb : max_tmp;
	}
	bool invalid = false;
	if (use_pool) {
		int cur = 0;
		cudaGetDevice(&cur);
		const std::vector<int> devs = devices.empty() ? std::vector<int>(1, cur) : devices;
		int want = g_host_threads;
		if (want <= 0) {
			// the pool's threads copy (blocks in, pixels out) and enqueue; one per hardware thread keeps the host
			// memory system busy, which is what a pageable call is bound by
			const int hw = (int)std::thread::hardware_concurrency();
			want = hw < 4 ? 4 : (hw > 32 ? 32 : hw);
		}
		if (want > 64)
			want = 64;
		pool().ensure(want);
		// the caller's ring of lanes on every device of the set: streams, device buffers, pinned bounce buffers
		const size_t per_dev = (slabs.size() + devs.size() - 1) / devs.size();
		const size_t ring = per_dev < (size_t)g_pool_ring ? per_dev : (size_t)g_pool_ring;
		std::vector<std::vector<char>> busy(devs.size(), std::vector<char>(ring, 0));
		bool ok = true;
		for (size_t d = 0; d < devs.size() && ok; d++) {
			if (cudaSetDevice(devs[d]) != cudaSuccess) {
				set_error("%s: cannot select CUDA device %d", who, devs[d]);
				ok = false;
				break;
			}
			std::vector<Lane> &r = c.rings[devs[d]];
			if (r.size() < ring)
				r.resize(ring);
			for (size_t i = 0; i < ring && ok; i++)
				ok = r[i].init() && reserve_lane(r[i], call, max_in, max_out, max_tmp, !pinned_in, !pinned_out);
		}
		cudaSetDevice(cur);
		if (!ok)
			return false;
		CallState st;
		{
			std::unique_lock<std::mutex> g(st.m);
			for (size_t k = 0; k < slabs.size(); k++) {
				const size_t d = k % devs.size(), i = (k / devs.size()) % ring;
				char *flag = &busy[d][i];
				st.cv.wait(g, [&] { return !*flag; });		// the lane's previous slab has been copied out
				if (st.rc != DETEXB200_OK)
					break;
				*flag = 1;
				st.pending++;
				Job j;
				j.kind = Job::SUBMIT;
				j.call = &call;
				j.slab = &slabs[k];
				j.lane = &c.rings[devs[d]][i];
				j.lane_busy = reinterpret_cast<bool *>(flag);
				j.device = devs[d];
				j.bounce_in = !pinned_in;
				j.bounce_out = !pinned_out;
				j.state = &st;
				pool().submit(j);
			}
			st.cv.wait(g, [&] { return st.pending == 0; });
		}
		if (st.rc != DETEXB200_OK) {
			// a failed slab may have left work behind on its lane: nothing of this call may still be running (and
			// writing into the caller's buffers) once it has returned
			for (size_t d = 0; d < devs.size(); d++)
				for (size_t i = 0; i < ring; i++)
					cudaStreamSynchronize(c.rings[devs[d]][i].stream);
			(void)cudaGetLastError();
			g_error = st.error;
			return false;
		}
		invalid = st.invalid;
	}
	else {
		cudaError_t e;
		const int used = (int)(slabs.size() < (size_t)N_STREAMS ? slabs.size() : (size_t)N_STREAMS);
		for (int i = 0; i < used; i++) {
			Lane &L = c.lanes[i];
			if (!reserve_lane(L, call, max_in, max_out, max_tmp, false, false))
				return false;
			if ((e = cudaMemsetAsync(L.d_flag, 0, 4, L.stream)) != cudaSuccess) {
				cuda_fail(who, e);
				return false;
			}
		}
		int rc = DETEXB200_OK;
		for (size_t k = 0; k < slabs.size() && rc == DETEXB200_OK; k++)
			rc = enqueue_slab(c.lanes[k % N_STREAMS], call, slabs[k], t->data + slabs[k].in_off, pixel_buffer + slabs[k].out_off);
		for (int i = 0; i < used; i++) {
			Lane &L = c.lanes[i];
			if (rc == DETEXB200_OK && (e = cudaMemcpyAsync(L.h_flag, L.d_flag, 4, cudaMemcpyDeviceToHost, L.stream)) != cudaSuccess)
				rc = cuda_fail(who, e);
			if ((e = cudaStreamSynchronize(L.stream)) != cudaSuccess && rc == DETEXB200_OK)
				rc = cuda_fail(who, e);
			if (rc == DETEXB200_OK && *L.h_flag)
				invalid = true;
		}
		if (rc != DETEXB200_OK)
			return false;
	}
	if (invalid) {
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
	Lane &L = c.lanes[0];
	if (!L.reserve(&L.d_in, &L.cap_in, 64) || !L.reserve(&L.d_out, &L.cap_out, 512) || !L.reserve(&L.d_tmp, &L.cap_tmp, 1024))
		return false;
	cudaStream_t s = L.stream;
	uint8_t *hs = (uint8_t *)c.h_small;
	memcpy(hs, bitstring, bb);
	cudaError_t e;
	if ((e = cudaMemsetAsync(L.d_flag, 0, 4, s)) != cudaSuccess ||
	    (e = cudaMemcpyAsync(L.d_in, hs, bb, cudaMemcpyHostToDevice, s)) != cudaSuccess) {
		cuda_fail("detexDecompressBlock", e);
		return false;
	}
	if (enqueue_decode(tf, L.d_in, 4, 4, 1, 1, L.d_out, pf, DETEXB200_LAYOUT_TILED, mode_mask, flags, L.d_flag, L.d_tmp,
			(uint32_t *)((uint8_t *)L.d_tmp + 512), s) != DETEXB200_OK)
		return false;
	if ((e = cudaMemcpyAsync(hs + 64, L.d_out, out_bytes, cudaMemcpyDeviceToHost, s)) != cudaSuccess ||
	    (e = cudaMemcpyAsync(L.h_flag, L.d_flag, 4, cudaMemcpyDeviceToHost, s)) != cudaSuccess ||
	    (e = cudaStreamSynchronize(s)) != cudaSuccess) {
		cuda_fail("detexDecompressBlock", e);
		return false;
	}
	memcpy(pixel_buffer, hs + 64, out_bytes);
	if (*L.h_flag) {
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
	// one bit per block of a slab.  A LINEAR slab holds rows * wib blocks whatever the image width says (width may be
	// much smaller than wib * 4), so the rows per slab are capped by the bitmap, not only by the pixel budget.
	const size_t map_blocks = pixel_region / (16 * ps_native);
	const size_t map_region = (map_blocks / 32 + 64) * 4;
	cudaError_t e = cudaMallocAsync(&tmp, pixel_region + map_region, s);
	if (e != cudaSuccess)
		return cuda_fail("cudaMallocAsync", e);
	uint32_t *const map = (uint32_t *)((uint8_t *)tmp + pixel_region);
	int rc = DETEXB200_OK;
	if (layout == DETEXB200_LAYOUT_LINEAR) {
		const size_t row_bytes = (size_t)width * 4 * ps_native;
		size_t rows_budget = slab_budget / (row_bytes ? row_bytes : 1);
		const size_t rows_cap = map_blocks / (size_t)(wib > 0 ? wib : 1);	// >= 4: pixel_region holds four block rows
		if (rows_budget > rows_cap)
			rows_budget = rows_cap;
		int rows_per_slab = rows_budget > (size_t)hib ? hib : (int)rows_budget;
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

#ifndef DETEXB200_BUILD_ID
#define DETEXB200_BUILD_ID "unknown"
#endif
#ifdef DETEXB200_EXPERIMENTS
#define DETEXB200_EXPERIMENTS_TAG "; +experiments"
#else
#define DETEXB200_EXPERIMENTS_TAG ""
#endif
// "build <id>": hash of the sources this library was compiled from (detex_b200/build.py); profiles/ncu_traffic.json
// records it with every ncu figure so that bench.py can tell a stale figure from a current one.
const char *detexb200Version(void) { return "detex_b200 0.2 (sm_100a; API of hglm/detex v0.1.2; build " DETEXB200_BUILD_ID DETEXB200_EXPERIMENTS_TAG ")"; }

int detexb200SetDevices(int n, const int *device_ids) {
	int count = 0;
	if (n < 0 || (n > 0 && device_ids == nullptr) || cudaGetDeviceCount(&count) != cudaSuccess) {
		set_error("detexb200SetDevices: invalid argument or no CUDA device");
		return DETEXB200_EINVAL;
	}
	std::vector<int> v;
	for (int i = 0; i < n; i++) {
		if (device_ids[i] < 0 || device_ids[i] >= count) {
			set_error("detexb200SetDevices: device ordinal %d out of range (%d devices)", device_ids[i], count);
			return DETEXB200_EINVAL;
		}
		for (int seen : v)
			if (seen == device_ids[i]) {
				set_error("detexb200SetDevices: device ordinal %d listed twice", device_ids[i]);
				return DETEXB200_EINVAL;
			}
		v.push_back(device_ids[i]);
	}
	device_set();			// (settle the environment first: an explicit call wins over it)
	std::lock_guard<std::mutex> g(g_devices_mutex);
	g_devices = v;
	return DETEXB200_OK;
}

int detexb200SetHostThreads(int n) {
	if (n < 0 || n > 64) {
		set_error("detexb200SetHostThreads: thread count %d out of range (0 = default, at most 64)", n);
		return DETEXB200_EINVAL;
	}
	device_set();
	std::lock_guard<std::mutex> g(g_devices_mutex);
	g_host_threads = n;
	return DETEXB200_OK;
}

void detexb200ReleaseThreadResources(void) { g_ctx.release(); }

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
	Lane &L = c.lanes[0];
	if (!L.reserve(&L.d_tmp, &L.cap_tmp, in_bytes + 16) || !L.reserve(&L.d_out, &L.cap_out, out_bytes + 16))
		return false;
	cudaStream_t s = L.stream;
	cudaError_t e;
	if ((e = cudaMemcpyAsync(L.d_tmp, src, in_bytes, cudaMemcpyHostToDevice, s)) != cudaSuccess ||
	    (e = launch_convert(L.d_tmp, n, spf, L.d_out, tpf, s)) != cudaSuccess ||
	    (e = cudaMemcpyAsync(dst, L.d_out, out_bytes, cudaMemcpyDeviceToHost, s)) != cudaSuccess ||
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
