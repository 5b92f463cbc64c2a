// Auxiliary device code around the decode path: the synthetic block generator used by the
// benchmark and a device-side hash for per-shard parity checks.
// (Pixel-format conversion lives in convert.cu.)
#include "common.cuh"
#include "launch.h"

namespace detexb200 {

static unsigned grid_for(uint64_t items) {
	const uint64_t want = (items + 255) / 256;
	return (unsigned)(want < 148ull * 64 ? (want ? want : 1) : 148ull * 64);
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
		if (force_modes == 8 || (force_modes >= 100 && force_modes < 108)) {
			// BC7: unary mode prefix, block g gets mode g % 8 (100 + m: every block mode m, profiling aid)
			const uint32_t m = force_modes == 8 ? (uint32_t)(g % 8ull) : (uint32_t)(force_modes - 100);
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
