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

static void run_job(job_t *j) {
	j->ok = j->fn(&j->view, j->out, j->pixel_format);
}

static double now_s(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/*
 * Persistent worker threads (BASELINE.md section 4 asks for warm buffers and steady-state numbers: creating
 * `nthreads` threads inside every timed call showed up as a 1.5x disagreement between two CPU numbers of the same
 * box in round 1).  Workers are created on first use and parked on a condition variable; one decode call hands
 * every worker one job (worker t takes job t of the current generation) and waits for the count to drain.
 * One call at a time (the Python callers are single-threaded); a mutex serialises callers anyway.
 */
#define MAX_WORKERS 512
static pthread_mutex_t pool_mutex = PTHREAD_MUTEX_INITIALIZER, call_mutex = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t pool_wake = PTHREAD_COND_INITIALIZER, pool_done = PTHREAD_COND_INITIALIZER;
static int pool_threads = 0, pool_pending = 0, pool_jobs = 0;
static unsigned long pool_generation = 0;
static job_t *pool_job_list = NULL;

static void *pool_worker(void *arg) {
	int me = (int)(intptr_t)arg;
	unsigned long seen = 0;
	pthread_mutex_lock(&pool_mutex);
	for (;;) {
		while (pool_generation == seen || me >= pool_jobs) {
			if (pool_generation != seen)
				seen = pool_generation;		/* this round has no job for me */
			pthread_cond_wait(&pool_wake, &pool_mutex);
		}
		seen = pool_generation;
		job_t *j = &pool_job_list[me];
		pthread_mutex_unlock(&pool_mutex);
		run_job(j);
		pthread_mutex_lock(&pool_mutex);
		if (--pool_pending == 0)
			pthread_cond_signal(&pool_done);
	}
	return NULL;
}

static void pool_run(job_t *jobs, int n) {
	pthread_mutex_lock(&pool_mutex);
	while (pool_threads < n && pool_threads < MAX_WORKERS) {
		pthread_t tid;
		pthread_attr_t attr;
		pthread_attr_init(&attr);
		pthread_attr_setdetachstate(&attr, PTHREAD_CREATE_DETACHED);
		if (pthread_create(&tid, &attr, pool_worker, (void *)(intptr_t)pool_threads) != 0)
			break;
		pool_threads++;
	}
	int n_pool = n < pool_threads ? n : pool_threads;
	pool_job_list = jobs;
	pool_jobs = n_pool;
	pool_pending = n_pool;
	pool_generation++;
	pthread_cond_broadcast(&pool_wake);
	pthread_mutex_unlock(&pool_mutex);
	for (int t = n_pool; t < n; t++)		/* (thread creation failed: the caller does the rest) */
		run_job(&jobs[t]);
	pthread_mutex_lock(&pool_mutex);
	while (pool_pending > 0)
		pthread_cond_wait(&pool_done, &pool_mutex);
	pool_jobs = 0;
	pthread_mutex_unlock(&pool_mutex);
}

/*
 * Decode a texture with `nthreads` threads.  linear != 0: row-major image
 * output (texture.c:105-145), else flat tile output (texture.c:77-98).
 * Returns 1 when every block was valid, 0 otherwise; *seconds gets the wall time
 * of the decode itself (jobs handed to the parked workers until the last one is done).
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
	if (nthreads > MAX_WORKERS)
		nthreads = MAX_WORKERS;
	job_t *jobs = (job_t *)calloc(nthreads, sizeof(job_t));
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
	pthread_mutex_lock(&call_mutex);
	double t0 = now_s();
	if (nthreads == 1)
		run_job(&jobs[0]);
	else
		pool_run(jobs, nthreads);
	double t1 = now_s();
	pthread_mutex_unlock(&call_mutex);
	if (seconds)
		*seconds = t1 - t0;
	int ok = 1;
	for (int t = 0; t < nthreads; t++)
		ok &= jobs[t].ok ? 1 : 0;
	free(jobs);
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

/*
 * CPU statement of detexb200HashDevice (detex_b200/csrc/aux.cu): FNV-1a-64 of every 4096-byte segment, the
 * little-endian array of segment hashes taken as the next buffer, until one hash is left.  The first level is
 * spread over `nthreads` threads (it is all the work); harness.py:tree_hash is the same thing in numpy and
 * tests/test_oracle.py keeps the two equal.
 */
typedef struct {
	const uint8_t *p;
	size_t n, seg0, seg1;
	uint64_t *out;
} hash_job_t;

static void *run_hash_job(void *q) {
	hash_job_t *j = (hash_job_t *)q;
	for (size_t s = j->seg0; s < j->seg1; s++) {
		size_t a = s * 4096, b = a + 4096 < j->n ? a + 4096 : j->n;
		j->out[s] = cpu_harness_fnv1a64(j->p + a, b - a);
	}
	return NULL;
}

uint64_t cpu_harness_tree_hash(const uint8_t *p, size_t n, int nthreads) {
	uint64_t *level = NULL;
	for (;;) {
		size_t n_seg = (n + 4095) / 4096;
		if (n_seg <= 1) {
			uint64_t h = cpu_harness_fnv1a64(p, n);
			free(level);
			return h;
		}
		uint64_t *next = (uint64_t *)malloc(n_seg * sizeof(uint64_t));
		int t_count = nthreads < 1 ? 1 : nthreads;
		if ((size_t)t_count > n_seg / 64 + 1)
			t_count = (int)(n_seg / 64 + 1);
		hash_job_t *jobs = (hash_job_t *)calloc(t_count, sizeof(hash_job_t));
		pthread_t *tids = (pthread_t *)calloc(t_count, sizeof(pthread_t));
		for (int t = 0; t < t_count; t++) {
			jobs[t].p = p;
			jobs[t].n = n;
			jobs[t].seg0 = n_seg * t / t_count;
			jobs[t].seg1 = n_seg * (t + 1) / t_count;
			jobs[t].out = next;
			if (t_count > 1)
				pthread_create(&tids[t], NULL, run_hash_job, &jobs[t]);
			else
				run_hash_job(&jobs[t]);
		}
		for (int t = 0; t_count > 1 && t < t_count; t++)
			pthread_join(tids[t], NULL);
		free(jobs);
		free(tids);
		free(level);
		level = next;
		p = (const uint8_t *)level;		/* (x86-64: little-endian, as the device stores them) */
		n = n_seg * sizeof(uint64_t);
	}
}
