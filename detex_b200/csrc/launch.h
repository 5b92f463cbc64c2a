// Internal interface between the C-ABI layer (api.cu) and the kernels (kernels.cu, aux.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace detexb200 {

constexpr int LAYOUT_TILED = 0, LAYOUT_LINEAR = 1;

struct DecodeArgs {
	const void *blocks;	// device, wib*hib blocks
	void *pixels;		// device
	int width, height;	// pixels (LINEAR clipping)
	int wib, hib;		// blocks
	int layout;
	uint32_t mode_mask, flags;
	uint32_t post_op;	// bit 0: R<->B swap, bit 1: signed<->unsigned 16-bit components, fused into the store
	int post_class;		// 0 native tile size, 1 drop the fourth component, 2 16-bit -> 8-bit components,
				// 3..11 BC6H half -> integer pixels (tone.cuh): 3 + 3 * mode + kind, mode 0 plain,
				// 1 HDR range, 2 HDR range + gamma; kind 0 RGBX8, 1 RGB8, 2 RGBX16
	// tone classes only.  Mode 1: tone_min / tone_max = the HDR range (hdr.c:119-141).  Mode 2: 1 / gamma and
	// the gamma-corrected range as the reference computes it on the host (hdr.c:149-158).
	// tone_zero_target: what an invalid block leaves behind (1: zeros in the target format,
	// texture.c:90-93; 0: the conversion of zero pixels)
	float tone_min, tone_max, tone_inv_gamma, tone_gamma_lo, tone_gamma_factor;
	uint32_t tone_zero_target;
	uint32_t *invalid;	// device flag word or nullptr
	uint32_t *invalid_map;	// device bitmap, one bit per block, set for invalid blocks; or nullptr
};

// Enqueue the decode of one texture in its native pixel format.  format_index = texture_format >> 24.
cudaError_t launch_decode(int format_index, const DecodeArgs &args, cudaStream_t stream);

// Pixel-format conversion on device (convert.cu).  convert_plan restates the reference's search
// (convert.c:887-1045): number of steps (0 = identical formats, -1 = no path) and the indices of
// the table entries used; launch_convert returns cudaErrorNotSupported when there is no path.
int convert_plan(uint32_t source_pixel_format, uint32_t target_pixel_format, int steps[4]);
int conversion_table_size();
void conversion_table_entry(int index, uint32_t *source_pixel_format, uint32_t *target_pixel_format);
void set_hdr_parameters(float gamma, float range_min, float range_max);	// thread-local, hdr.c:26-38
void get_hdr_parameters(float *gamma, float *range_min, float *range_max);
bool convert_supported(uint32_t source_pixel_format, uint32_t target_pixel_format);
// A conversion the decode kernels can do while storing: post_op / post_class as in DecodeArgs.
// tone_hdr: the tone classes take the HDR range parameters of the calling thread.
struct FusedConversion { bool ok; uint32_t post_op; int post_class; bool tone_hdr; };
FusedConversion convert_fused(uint32_t source_pixel_format, uint32_t target_pixel_format);
// Clear the target pixels of the blocks marked in invalid_map (after a separate conversion pass).
cudaError_t launch_zero_invalid(const uint32_t *invalid_map, void *pixels, int width, int height, int wib, int hib,
	int layout, int pixel_bytes, cudaStream_t stream);
cudaError_t launch_convert(const void *src, uint64_t nu_pixels, uint32_t source_pixel_format,
	void *dst, uint32_t target_pixel_format, cudaStream_t stream);

// Synthetic block generator and device-side hash (aux.cu).
cudaError_t launch_generate(void *blocks, uint64_t n_blocks, int block_bytes, uint64_t seed,
	uint64_t first_block, int force_modes, cudaStream_t stream);
cudaError_t device_hash(const void *buf, uint64_t n_bytes, uint64_t *hash_out, cudaStream_t stream);

// Thread-local error text behind detexGetErrorMessage (api.cu).
void set_error_text(const char *text);

uint64_t launch_count();
void count_launch();
int sm_count();		// multiprocessors of the current device (148 on B200)

}  // namespace detexb200
