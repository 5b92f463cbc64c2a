// Auxiliary device code around the decode path: pixel-format conversion for the
// formats callers of the hot path request (reference convert.c), the synthetic block
// generator used by the benchmark and a device-side hash for
// per-shard parity checks.
#include "common.cuh"
#include "launch.h"

namespace detexb200 {

// ---- pixel-format conversion ------------------------------------------------------------
// Pixel format constants (reference detex.h:83-379).
constexpr uint32_t PF_RGBA8 = 0x334, PF_BGRA8 = 0x33C, PF_RGBX8 = 0x320, PF_BGRX8 = 0x328,
	PF_RGB8 = 0x220, PF_FLOAT_RGBX16 = 0x2721, PF_FLOAT_BGRX16 = 0x2729, PF_FLOAT_RGB16 = 0x2521;

enum ConvertOp : int { OP_NONE = -1, OP_COPY32, OP_SWAP32, OP_SWAP64, OP_PACK24, OP_PACK48 };

static ConvertOp convert_op(uint32_t s, uint32_t d) {
	const bool s_rgb32 = s == PF_RGBA8 || s == PF_RGBX8, s_bgr32 = s == PF_BGRA8 || s == PF_BGRX8;
	// convert.c:765-774: X<->A of the same order is a no-op, R<->B swap is one in-place
	// step (ConvertPixel32RGBA8ToPixel32BGRA8, :37-52).
	if ((s == PF_RGBX8 && d == PF_RGBA8) || (s == PF_RGBA8 && d == PF_RGBX8) ||
	    (s == PF_BGRX8 && d == PF_BGRA8) || (s == PF_BGRA8 && d == PF_BGRX8))
		return OP_COPY32;
	// swap, possibly chained with the X<->A no-op (two-step paths of the reference's search)
	if ((s_rgb32 && (d == PF_BGRA8 || d == PF_BGRX8)) || (s_bgr32 && (d == PF_RGBA8 || d == PF_RGBX8)))
		return OP_SWAP32;
	if ((s == PF_FLOAT_RGBX16 && d == PF_FLOAT_BGRX16) || (s == PF_FLOAT_BGRX16 && d == PF_FLOAT_RGBX16))
		return OP_SWAP64;		// ConvertPixel64RGBX16ToPixel64BGRX16, convert.c:54-70
	if (s_rgb32 && d == PF_RGB8)
		return OP_PACK24;		// ConvertPixel32RGBX8ToPixel24RGB8, convert.c:671-690
	if (s == PF_FLOAT_RGBX16 && d == PF_FLOAT_RGB16)
		return OP_PACK48;		// ConvertPixel64RGBX16ToPixel48RGB16, convert.c:700-716
	return OP_NONE;
}

int convert_fusable(uint32_t s, uint32_t d) {
	switch (convert_op(s, d)) {
	case OP_COPY32: return 1;
	case OP_SWAP32: case OP_SWAP64: return 2;
	default: return 0;
	}
}

bool convert_supported(uint32_t s, uint32_t d) {
	return s == d || convert_op(s, d) != OP_NONE;
}

// Four 32-bit pixels per thread.
template <int OP>
__global__ void __launch_bounds__(256) convert32_kernel(const uint4 *__restrict__ src, void *dst, uint64_t n_quads) {
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_quads; i += (uint64_t)gridDim.x * blockDim.x) {
		uint4 v = src[i];
		if (OP == OP_SWAP32) {
			v.x = prmt(v.x, 0, 0x3012);	v.y = prmt(v.y, 0, 0x3012);
			v.z = prmt(v.z, 0, 0x3012);	v.w = prmt(v.w, 0, 0x3012);
		}
		if (OP == OP_PACK24) {
			// R0G0B0R1 | G1B1R2G2 | B2R3G3B3
			uint32_t *o = (uint32_t *)dst + 3 * i;
			o[0] = prmt(v.x, v.y, 0x4210);
			o[1] = prmt(v.y, v.z, 0x5421);
			o[2] = prmt(v.z, v.w, 0x6542);
		}
		else
			((uint4 *)dst)[i] = v;
	}
}

// Two 64-bit pixels per thread.
template <int OP>
__global__ void __launch_bounds__(256) convert64_kernel(const uint4 *__restrict__ src, void *dst, uint64_t n_pairs) {
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += (uint64_t)gridDim.x * blockDim.x) {
		uint4 v = src[i];
		if (OP == OP_SWAP64) {
			// R G B X (16 bit each) -> B G R X
			const uint32_t a0 = prmt(v.x, v.y, 0x3254), a1 = prmt(v.x, v.y, 0x7610);
			const uint32_t b0 = prmt(v.z, v.w, 0x3254), b1 = prmt(v.z, v.w, 0x7610);
			((uint4 *)dst)[i] = make_uint4(a0, a1, b0, b1);
		}
		else {
			// OP_PACK48: R0G0 B0R1 G1B1
			uint32_t *o = (uint32_t *)dst + 3 * i;
			o[0] = v.x;
			o[1] = prmt(v.y, v.z, 0x5410);
			o[2] = prmt(v.z, v.w, 0x5432);
		}
	}
}

// Scalar tail / unaligned fallback: one pixel per thread, byte granular.
__global__ void __launch_bounds__(256) convert_scalar_kernel(const uint8_t *src, uint8_t *dst, uint64_t first,
uint64_t n, int op) {
	for (uint64_t i = first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		if (op == OP_COPY32 || op == OP_SWAP32) {
			const uint8_t r = src[4 * i], g = src[4 * i + 1], b = src[4 * i + 2], a = src[4 * i + 3];
			const bool sw = op == OP_SWAP32;
			dst[4 * i] = sw ? b : r;	dst[4 * i + 1] = g;	dst[4 * i + 2] = sw ? r : b;	dst[4 * i + 3] = a;
		}
		else if (op == OP_PACK24) {
			const uint8_t r = src[4 * i], g = src[4 * i + 1], b = src[4 * i + 2];
			dst[3 * i] = r;	dst[3 * i + 1] = g;	dst[3 * i + 2] = b;
		}
		else {
			uint16_t c[4];
			for (int k = 0; k < 4; k++)
				c[k] = (uint16_t)(src[8 * i + 2 * k] | (src[8 * i + 2 * k + 1] << 8));
			if (op == OP_SWAP64) {
				const uint16_t t = c[0];
				c[0] = c[2];
				c[2] = t;
			}
			const int nc = op == OP_SWAP64 ? 4 : 3;
			for (int k = 0; k < nc; k++) {
				dst[2 * nc * i + 2 * k] = (uint8_t)(c[k] & 0xFF);
				dst[2 * nc * i + 2 * k + 1] = (uint8_t)(c[k] >> 8);
			}
		}
	}
}

static unsigned grid_for(uint64_t items) {
	const uint64_t want = (items + 255) / 256;
	return (unsigned)(want < 148ull * 64 ? (want ? want : 1) : 148ull * 64);
}

cudaError_t launch_convert(const void *src, uint64_t n, uint32_t s, void *dst, uint32_t d, cudaStream_t stream) {
	if (n == 0)
		return cudaSuccess;
	if (s == d) {
		const size_t bytes = (size_t)n * (1 + ((s & 0xF00) >> 8));
		if (src == dst)
			return cudaSuccess;
		return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, stream);
	}
	const ConvertOp op = convert_op(s, d);
	if (op == OP_NONE)
		return cudaErrorNotSupported;
	const bool aligned = (((uintptr_t)src | (uintptr_t)dst) & 15u) == 0;
	const bool is64 = op == OP_SWAP64 || op == OP_PACK48;
	const uint64_t per = is64 ? 2 : 4;
	const uint64_t groups = aligned ? n / per : 0;
	if (groups) {
		const unsigned g = grid_for(groups);
		switch (op) {
		case OP_COPY32: convert32_kernel<OP_COPY32><<<g, 256, 0, stream>>>((const uint4 *)src, dst, groups); break;
		case OP_SWAP32: convert32_kernel<OP_SWAP32><<<g, 256, 0, stream>>>((const uint4 *)src, dst, groups); break;
		case OP_PACK24: convert32_kernel<OP_PACK24><<<g, 256, 0, stream>>>((const uint4 *)src, dst, groups); break;
		case OP_SWAP64: convert64_kernel<OP_SWAP64><<<g, 256, 0, stream>>>((const uint4 *)src, dst, groups); break;
		default: convert64_kernel<OP_PACK48><<<g, 256, 0, stream>>>((const uint4 *)src, dst, groups); break;
		}
		count_launch();
	}
	if (groups * per < n) {
		convert_scalar_kernel<<<grid_for(n - groups * per), 256, 0, stream>>>((const uint8_t *)src, (uint8_t *)dst,
			groups * per, n, (int)op);
		count_launch();
	}
	return cudaGetLastError();
}

// ---- synthetic block generator ------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64_word(uint64_t seed, uint64_t j) {
	uint64_t z = seed + (j + 1ull) * 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

// One 16-byte unit (one 128-bit block or two 64-bit blocks) per thread.
__global__ void __launch_bounds__(256) generate_kernel(uint4 *out, uint64_t n_units, uint64_t seed, uint64_t first_unit,
int force_modes) {
	for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < n_units; u += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t g = first_unit + u;
		uint64_t w0 = splitmix64_word(seed, 2 * g), w1 = splitmix64_word(seed, 2 * g + 1);
		if (force_modes == 8) {			// BC7: unary mode prefix, block g gets mode g % 8
			const uint32_t m = (uint32_t)(g % 8ull);
			const uint64_t b0 = ((w0 & 0xFFull) & ((0xFFull << (m + 1)) & 0xFFull)) | (1ull << m);
			w0 = (w0 & ~0xFFull) | b0;
		}
		else if (force_modes == 14) {		// BC6H: 2/5-bit codes (decompress-bptc-float.c:660-662)
			const uint32_t codes[14] = { 0, 1, 2, 6, 10, 14, 18, 22, 26, 30, 3, 7, 11, 15 };
			const uint32_t m = (uint32_t)(g % 14ull);
			const uint64_t keep = m < 2 ? 0xFCull : 0xE0ull;
			w0 = (w0 & ~0xFFull) | (w0 & keep) | codes[m];
		}
		out[u] = make_uint4((uint32_t)w0, (uint32_t)(w0 >> 32), (uint32_t)w1, (uint32_t)(w1 >> 32));
	}
}

cudaError_t launch_generate(void *blocks, uint64_t n_blocks, int block_bytes, uint64_t seed, uint64_t first_block,
int force_modes, cudaStream_t stream) {
	const uint64_t bytes = n_blocks * (uint64_t)block_bytes, first_byte = first_block * (uint64_t)block_bytes;
	if ((bytes & 15u) || (first_byte & 15u) || (block_bytes != 8 && block_bytes != 16))
		return cudaErrorInvalidValue;
	if (force_modes && block_bytes != 16)
		return cudaErrorInvalidValue;
	if (bytes == 0)
		return cudaSuccess;
	generate_kernel<<<grid_for(bytes / 16), 256, 0, stream>>>((uint4 *)blocks, bytes / 16, seed, first_byte / 16, force_modes);
	count_launch();
	return cudaGetLastError();
}

// ---- device-side hash ----------------------------------------------------------------------
// FNV-1a-64 is byte-serial, so a flat hash of gigabytes cannot be computed in parallel.
// The device hash is therefore a hash of hashes: split the buffer into 4096-byte segments,
// FNV-1a-64 every segment (one thread each), treat the little-endian array of segment
// hashes as the next buffer, and repeat until at most 4096 bytes remain; the final level is
// hashed directly.  oracle/harness.py (tree_hash) is the CPU statement of the same function.
constexpr uint64_t SEG = 4096;

__global__ void __launch_bounds__(128) segment_hash_kernel(const uint8_t *buf, uint64_t n_bytes, uint64_t *out, uint64_t n_seg) {
	for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_seg; s += (uint64_t)gridDim.x * blockDim.x) {
		const uint8_t *p = buf + s * SEG;
		const uint64_t len = min(SEG, n_bytes - s * SEG);
		uint64_t h = 0xcbf29ce484222325ull;
		uint64_t i = 0;
		if ((((uintptr_t)p) & 15u) == 0)
			for (; i + 16 <= len; i += 16) {
				const uint4 v = *(const uint4 *)(p + i);
				const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
				for (int k = 0; k < 4; k++)
#pragma unroll
					for (int b = 0; b < 4; b++) {
						h ^= (w[k] >> (8 * b)) & 0xFFu;
						h *= 0x100000001b3ull;
					}
			}
		for (; i < len; i++) {
			h ^= p[i];
			h *= 0x100000001b3ull;
		}
		out[s] = h;
	}
}

cudaError_t device_hash(const void *buf, uint64_t n_bytes, uint64_t *hash_out, cudaStream_t stream) {
	const uint8_t *cur = (const uint8_t *)buf;
	uint64_t len = n_bytes;
	uint64_t *levels[2] = { nullptr, nullptr };
	cudaError_t err = cudaSuccess;
	int which = 0;
	// always reduce at least once so the final level lives in a buffer we own
	do {
		const uint64_t n_seg = len ? (len + SEG - 1) / SEG : 1;
		uint64_t *next = nullptr;
		if ((err = cudaMallocAsync((void **)&next, n_seg * 8, stream)) != cudaSuccess)
			break;
		segment_hash_kernel<<<grid_for(n_seg) * 2, 128, 0, stream>>>(cur, len, next, n_seg);
		count_launch();
		if (levels[which])
			cudaFreeAsync(levels[which], stream);
		levels[which] = next;
		cur = (const uint8_t *)next;
		len = n_seg * 8;
		which ^= 1;
	} while (len > 8);
	if (err == cudaSuccess)
		err = cudaMemcpyAsync(hash_out, cur, 8, cudaMemcpyDeviceToHost, stream);
	if (err == cudaSuccess)
		err = cudaStreamSynchronize(stream);
	for (int k = 0; k < 2; k++)
		if (levels[k])
			cudaFreeAsync(levels[k], stream);
	return err;
}

}  // namespace detexb200
