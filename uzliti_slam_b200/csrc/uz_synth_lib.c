/* The benchmark's synthetic map generator (include/uz_synth.h) as a small shared library, so that bench.py fills its
 * 10 000-keyframe map in seconds; synth_splitmix.py holds the numpy mirror that produces the same bytes. */
#include "uz_synth.h"
#include <pthread.h>

typedef struct { const uz_synth_cfg* c; uint8_t* desc; double* pos; uint8_t* valid; int32_t lo, hi; int rc; } job_t;

static void* run(void* a) {
    job_t* j = (job_t*)a;
    for (int32_t i = j->lo; i < j->hi; ++i) {
        const size_t at = (size_t)i * (size_t)j->c->n_features;
        if (uz_synth_keyframe(j->c, i, j->desc + at * (size_t)j->c->desc_bytes, j->pos + at * 3, j->valid + at) != 0) { j->rc = -1; break; }
    }
    return 0;
}

/* keyframes on `threads` host threads (keyframes are independent streams), then the candidate pairs; returns the pair count or -1 */
long long uz_synth_map_mt(const uz_synth_cfg* c, uint8_t* desc, double* pos, uint8_t* valid, int32_t* pairs, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    pthread_t th[64];
    job_t jobs[64];
    for (int t = 0; t < threads; ++t) {
        jobs[t].c = c; jobs[t].desc = desc; jobs[t].pos = pos; jobs[t].valid = valid; jobs[t].rc = 0;
        jobs[t].lo = (int32_t)((long long)c->n_keyframes * t / threads);
        jobs[t].hi = (int32_t)((long long)c->n_keyframes * (t + 1) / threads);
        if (pthread_create(&th[t], 0, run, &jobs[t]) != 0) { run(&jobs[t]); th[t] = 0; }
    }
    int rc = 0;
    for (int t = 0; t < threads; ++t) { if (th[t]) pthread_join(th[t], 0); rc |= jobs[t].rc; }
    if (rc) return -1;
    long long n_pairs = 0;
    for (int32_t i = 0; i < c->n_keyframes; ++i) {
        const int32_t k = uz_synth_candidates(c, i, pairs + 2 * n_pairs);
        if (k < 0) return -1;
        n_pairs += k;
    }
    return n_pairs;
}

void uz_synth_pose_c(const uz_synth_cfg* c, int32_t i, double* T12) { uz_synth_pose(c, i, T12); }
unsigned long long uz_synth_checksum_c(const void* p, size_t n) { return uz_synth_checksum(p, n); }
