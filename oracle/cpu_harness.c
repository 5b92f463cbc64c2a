/*
 * TEST / BENCH INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Multi-threaded driver around a CPU texture decoder with the signature of
 * detexDecompressTextureLinear / detexDecompressTextureTiled
 * (reference detex.h:747-765, texture.c:77-145).  The reference library is
 * single-threaded but re-entrant on this path, so the
 * CPU baseline of record row-shards the texture: every
 * thread decodes a contiguous range of block rows through its own
 * detexTexture view.  The decoder is passed in as a function pointer so the
 * same harness times either the compiled reference (oracle/_ref) or the
 * restatement (liboracle.so).
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdlib.h>
#include <time.h>

/* Same layout as the reference's detexTexture (detex.h:729-736). */
typedef struct {
	uint32_t format;
	uint8_t *data;
	int width, height, width_in_blocks, height_in_blocks;
} tex_view_t;

typedef bool (*texture_decode_fn)(const tex_view_t *, uint8_t *, uint32_t);

typedef struct {
	texture_decode_fn fn;
	tex_view_t view;
	uint8_t *out;
	uint32_t pixel_format;
	bool ok;
} job_t;

static void *run_job(void *p) {
	job_t *j = (job_t *)p;
	j->ok = j->fn(&j->view, j->out, j->pixel_format);
	return NULL;
}

static double now_s(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/*
 * Decode a texture with `nthreads` threads.  linear != 0: row-major image
 * output (texture.c:105-145), else flat tile output (texture.c:77-98).
 * Returns 1 when every block was valid, 0 otherwise; *seconds gets the wall time
 * of the decode itself (thread start to last join).
 */
int cpu_harness_decode(texture_decode_fn fn, uint32_t format, uint8_t *data, int width,
int height, int wib, int hib, uint8_t *out, uint32_t pixel_format, int linear,
int nthreads, double *seconds) {
	int block_bytes = 8 + ((format & 0x00800000u) >> 20);		/* detex.h:918-920 */
	int ps = 1 + ((pixel_format & 0xF00) >> 8);			/* detex.h:879-881 */
	if (nthreads < 1)
		nthreads = 1;
	if (nthreads > hib)
		nthreads = hib > 0 ? hib : 1;
	job_t *jobs = (job_t *)calloc(nthreads, sizeof(job_t));
	pthread_t *tids = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
	for (int t = 0; t < nthreads; t++) {
		int r0 = (int)((int64_t)hib * t / nthreads);
		int r1 = (int)((int64_t)hib * (t + 1) / nthreads);
		job_t *j = &jobs[t];
		j->fn = fn;
		j->pixel_format = pixel_format;
		j->view.format = format;
		j->view.data = data + (size_t)r0 * wib * block_bytes;
		j->view.width = width;
		j->view.width_in_blocks = wib;
		j->view.height_in_blocks = r1 - r0;
		int h = height - r0 * 4;
		if (h > (r1 - r0) * 4)
			h = (r1 - r0) * 4;
		j->view.height = h;
		if (linear)
			j->out = out + (size_t)r0 * 4 * width * ps;
		else
			j->out = out + (size_t)r0 * wib * 16 * ps;
	}
	double t0 = now_s();
	if (nthreads == 1)
		run_job(&jobs[0]);
	else {
		for (int t = 0; t < nthreads; t++)
			pthread_create(&tids[t], NULL, run_job, &jobs[t]);
		for (int t = 0; t < nthreads; t++)
			pthread_join(tids[t], NULL);
	}
	double t1 = now_s();
	if (seconds)
		*seconds = t1 - t0;
	int ok = 1;
	for (int t = 0; t < nthreads; t++)
		ok &= jobs[t].ok ? 1 : 0;
	free(jobs);
	free(tids);
	return ok;
}

/* FNV-1a-64 over a buffer (offset 0xcbf29ce484222325, prime 0x100000001b3): the
 * hash the SURVEY.md 8c golden tables are quoted in. */
uint64_t cpu_harness_fnv1a64(const uint8_t *p, size_t n) {
	uint64_t h = 0xcbf29ce484222325ull;
	for (size_t i = 0; i < n; i++) {
		h ^= p[i];
		h *= 0x100000001b3ull;
	}
	return h;
}

/* Counter-based splitmix64 block generator of SURVEY.md 8c: 64-bit word j of the
 * array = mix(seed + (j+1)*0x9E3779B97F4A7C15), little-endian.  first_word lets a
 * shard generate its own range. */
void cpu_harness_splitmix64(uint64_t seed, uint64_t first_word, uint64_t n_words, uint64_t *out) {
	for (uint64_t j = 0; j < n_words; j++) {
		uint64_t z = seed + (first_word + j + 1) * 0x9E3779B97F4A7C15ull;
		z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
		z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
		out[j] = z ^ (z >> 31);
	}
}
