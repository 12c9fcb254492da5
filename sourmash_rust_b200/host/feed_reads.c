/* feed_reads.c -- a minimal C HOST PROGRAM LOOP over the unmodified reference ABI: what a C (or Rust, through
 * `extern "C"`) caller of include/sourmash.h does when it sketches a read set the way the reference is used for
 * BASELINE config 2 -- one kmerminhash_add_sequence call per read and per k-size (src/ffi.rs:55-70).
 *
 * It exists so that bench.py and tests/manual/percall.py can time that calling pattern without the ~1 us per call
 * a ctypes round trip costs.  It links against nothing: the entry point is passed in as a function pointer, so
 * the same loop drives libsourmash.so (this build) or any other implementation of the header.
 */
#include <pthread.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdlib.h>
#include <time.h>

typedef void (*add_sequence_fn)(void *mh, const char *sequence, bool force);

static uint64_t now_ns(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}

/* for every read i (a NUL-terminated string at reads + i * stride) and every sketch j: add_sequence(mhs[j], read_i, force).
 * Returns the wall time of the loop in nanoseconds. */
uint64_t feed_reads(add_sequence_fn add_sequence, void *const *mhs, int n_mhs, const char *reads, uint64_t n_reads,
                    uint64_t stride, bool force) {
    const uint64_t t0 = now_ns();
    for (uint64_t i = 0; i < n_reads; i++) {
        const char *r = reads + i * stride;
        for (int j = 0; j < n_mhs; j++) add_sequence(mhs[j], r, force);
    }
    return now_ns() - t0;
}

typedef uintptr_t (*size_fn)(void *mh);

struct feed_job {
    add_sequence_fn fn;
    size_fn size;             /* nullable: kmerminhash_get_mins_size, called on every sketch to flush deferred work */
    void *const *mhs;
    int n_mhs;
    const char *reads;
    uint64_t n_reads, n_warm, stride;
    bool force;
    pthread_barrier_t *start;
    uint64_t t_begin, t_end;
};
static void *feed_thread(void *p) {
    struct feed_job *j = (struct feed_job *)p;
    /* untimed: the thread's first calls set up its CUDA stream and scratch (one context per host thread) */
    if (j->n_warm) {
        feed_reads(j->fn, j->mhs, j->n_mhs, j->reads, j->n_warm, j->stride, j->force);
        if (j->size) for (int k = 0; k < j->n_mhs; k++) j->size(j->mhs[k]);
    }
    pthread_barrier_wait(j->start);
    j->t_begin = now_ns();
    feed_reads(j->fn, j->mhs, j->n_mhs, j->reads + j->n_warm * j->stride, j->n_reads - j->n_warm, j->stride, j->force);
    if (j->size) for (int k = 0; k < j->n_mhs; k++) j->size(j->mhs[k]);
    j->t_end = now_ns();
    return NULL;
}
/* n_threads host threads, thread t feeding its own sketches mhs[t * n_mhs .. (t + 1) * n_mhs) with its own
 * contiguous share of the reads (distinct handles are independent: SURVEY 8(b) Threading).  Each thread first feeds
 * `warm_reads` of its share untimed; then all start together.  The timed region ends when the last thread has fed
 * its share AND (when `size` is given) read the size of each of its sketches, which makes the library finish all
 * deferred work.  Returns that region's wall time in nanoseconds. */
uint64_t feed_reads_mt(add_sequence_fn add_sequence, size_fn size, void *const *mhs, int n_mhs, int n_threads, const char *reads,
                       uint64_t n_reads, uint64_t stride, bool force, uint64_t warm_reads) {
    if (n_threads < 1) n_threads = 1;
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof *th);
    struct feed_job *jobs = (struct feed_job *)calloc((size_t)n_threads, sizeof *jobs);
    pthread_barrier_t start;
    pthread_barrier_init(&start, NULL, (unsigned)n_threads);
    const uint64_t per = (n_reads + (uint64_t)n_threads - 1) / (uint64_t)n_threads;
    for (int t = 0; t < n_threads; t++) {
        const uint64_t lo = per * (uint64_t)t < n_reads ? per * (uint64_t)t : n_reads;
        const uint64_t hi = lo + per < n_reads ? lo + per : n_reads;
        jobs[t].fn = add_sequence; jobs[t].size = size; jobs[t].mhs = mhs + (size_t)t * (size_t)n_mhs; jobs[t].n_mhs = n_mhs;
        jobs[t].reads = reads + lo * stride; jobs[t].n_reads = hi - lo; jobs[t].stride = stride; jobs[t].force = force;
        jobs[t].n_warm = warm_reads < hi - lo ? warm_reads : hi - lo;
        jobs[t].start = &start;
        pthread_create(&th[t], NULL, feed_thread, &jobs[t]);
    }
    uint64_t t0 = ~0ull, t1 = 0;
    for (int t = 0; t < n_threads; t++) {
        pthread_join(th[t], NULL);
        if (jobs[t].t_begin < t0) t0 = jobs[t].t_begin;
        if (jobs[t].t_end > t1) t1 = jobs[t].t_end;
    }
    pthread_barrier_destroy(&start);
    free(th);
    free(jobs);
    return t1 - t0;
}
