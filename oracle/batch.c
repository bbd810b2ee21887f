/*
 * batch.c -- thread-pooled drivers over the CPU oracles (TEST INFRASTRUCTURE ONLY).
 * Frames / KTX2 segments are independent units, so the CPU baseline hands them to a pool of
 * worker threads, mirroring the reference's worker pools (src/lib/DRACOLoader.js:24,312-364:
 * <=4 Draco workers; src/lib/WorkerPool.js:7: <=4 Basis workers) but with as many threads as
 * the caller asks for.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

static uint64_t fnv(uint64_t h, const void *p, size_t n) {
    const uint8_t *b = (const uint8_t *)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

typedef struct {
    const uint8_t *const *data; const size_t *len; int n; int kind;
    int next; pthread_mutex_t mu;
    int ok; uint64_t checksum, a, b; int want_sum;
} job_t;

static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    for (;;) {
        pthread_mutex_lock(&j->mu); int i = j->next++; pthread_mutex_unlock(&j->mu);
        if (i >= j->n) break;
        uint64_t cs = 0, a = 0, b = 0; int ok = 0;
        if (j->kind == 0) {
            uvo_draco_mesh m;
            if (uvo_draco_decode(j->data[i], j->len[i], &m) == 0) {
                ok = 1; a = m.num_points; b = m.num_faces;
                if (j->want_sum) {
                    cs = fnv(1469598103934665603ull, m.index, (size_t)m.num_faces * 12);
                    if (m.position) cs = fnv(cs, m.position, (size_t)m.num_points * 12);
                    if (m.normal) cs = fnv(cs, m.normal, (size_t)m.num_points * 12);
                    if (m.uv) cs = fnv(cs, m.uv, (size_t)m.num_points * 8);
                }
            }
            uvo_draco_free(&m);
        } else {
            uvo_ktx2_image t;
            if (uvo_ktx2_decode(j->data[i], j->len[i], &t) == 0) {
                ok = 1; a = (uint64_t)t.width * t.height * t.layers;
                if (j->want_sum) cs = fnv(1469598103934665603ull, t.rgba, t.rgba_bytes);
            }
            uvo_ktx2_free(&t);
        }
        pthread_mutex_lock(&j->mu);
        j->ok += ok; j->a += a; j->b += b; j->checksum += cs * (uint64_t)(2 * i + 1);
        pthread_mutex_unlock(&j->mu);
    }
    return NULL;
}

static int run(job_t *j, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 1024) threads = 1024;
    pthread_mutex_init(&j->mu, NULL);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, worker, j);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th); pthread_mutex_destroy(&j->mu);
    return j->ok;
}

int uvo_draco_decode_batch(const uint8_t *const *data, const size_t *len, int n, int threads, uint64_t *checksum,
                           uint64_t *total_points, uint64_t *total_faces) {
    job_t j; memset(&j, 0, sizeof j); j.data = data; j.len = len; j.n = n; j.kind = 0; j.want_sum = checksum != NULL;
    int ok = run(&j, threads);
    if (checksum) *checksum = j.checksum;
    if (total_points) *total_points = j.a;
    if (total_faces) *total_faces = j.b;
    return ok;
}

int uvo_ktx2_decode_batch(const uint8_t *const *data, const size_t *len, int n, int threads, uint64_t *checksum,
                          uint64_t *total_texels) {
    job_t j; memset(&j, 0, sizeof j); j.data = data; j.len = len; j.n = n; j.kind = 1; j.want_sum = checksum != NULL;
    int ok = run(&j, threads);
    if (checksum) *checksum = j.checksum;
    if (total_texels) *total_texels = j.a;
    return ok;
}
