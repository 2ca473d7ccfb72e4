/* coalescer_load.c — T client threads (standing in for Postgres backends connected to a sidecar) each issue one
 * blocking single-query request after another through dann_coalescer_search; reports the aggregate query rate and how
 * the requests were batched.  This is the measurement SURVEY.md §8f row 4 calls for ("batching across backends is
 * where the QPS comes from"); plain C99 + pthreads over the C ABI.
 *
 *   gcc -std=c99 -O2 -Wall -Wextra -Iinclude -Iharness harness/coalescer_load.c -Lpgvectorscale_b200 \
 *       -l:libdiskann_b200.so -Wl,-rpath,$PWD/pgvectorscale_b200 -lpthread -o /tmp/coalescer_load
 *   /tmp/coalescer_load snap.raw queries.f32 <threads> <queries per thread> <L> <rescore> <k> <max_batch> <max_wait_us>
 *
 * queries.f32 = raw little-endian f32 rows of `dim` values.  Prints one JSON line. */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "snapshot_raw.h"

typedef struct {
    dann_coalescer *co;
    const float *queries;
    size_t nq_total, first;
    int per_thread, dim, k, L, rescore;
    uint64_t *tids; /* [per_thread][k] of this thread */
    unsigned long rows;
    int rc;
} client;

static void *client_main(void *arg) {
    client *c = (client *)arg;
    for (int i = 0; i < c->per_thread; i++) {
        const float *q = c->queries + ((c->first + (size_t)i) % c->nq_total) * (size_t)c->dim;
        uint32_t count = 0;
        int rc = dann_coalescer_search(c->co, q, NULL, -1, c->k, c->L, c->rescore, c->tids + (size_t)i * c->k, NULL, &count, NULL);
        if (rc != DANN_OK) {
            c->rc = rc;
            return NULL;
        }
        c->rows += count;
    }
    return NULL;
}

int main(int argc, char **argv) {
    if (argc < 10) {
        fprintf(stderr, "usage: %s snap.raw queries.f32 threads per_thread L rescore k max_batch max_wait_us\n", argv[0]);
        return 2;
    }
    const int T = atoi(argv[3]), per = atoi(argv[4]), L = atoi(argv[5]), rescore = atoi(argv[6]), k = atoi(argv[7]);
    dann_snapshot_desc s;
    const float *iv = NULL;
    void *buf = dann_snapshot_raw_read(argv[1], &s, &iv);
    if (!buf) {
        fprintf(stderr, "cannot read %s\n", argv[1]);
        return 1;
    }
    FILE *f = fopen(argv[2], "rb");
    if (!f) return 1;
    fseek(f, 0, SEEK_END);
    const long qbytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    float *queries = (float *)malloc((size_t)qbytes);
    if (!queries || fread(queries, 1, (size_t)qbytes, f) != (size_t)qbytes) return 1;
    fclose(f);
    const size_t nq = (size_t)qbytes / 4 / s.dim;
    dann_index *ix = NULL;
    int rc = iv ? dann_index_load_plain(&s, iv, 0, &ix) : dann_index_load(&s, 0, &ix);
    free(buf);
    if (rc == DANN_ERR_NO_DEVICE) {
        fprintf(stderr, "coalescer_load: %s\n", dann_last_error());
        return 3;
    }
    if (rc != DANN_OK) {
        fprintf(stderr, "load failed (%d): %s\n", rc, dann_last_error());
        return 1;
    }
    dann_coalescer *co = NULL;
    if (dann_coalescer_create(ix, atoi(argv[8]), atoi(argv[9]), &co) != DANN_OK) {
        fprintf(stderr, "%s\n", dann_last_error());
        return 1;
    }
    client *cl = (client *)calloc((size_t)T, sizeof *cl);
    pthread_t *th = (pthread_t *)calloc((size_t)T, sizeof *th);
    for (int t = 0; t < T; t++) {
        cl[t].co = co;
        cl[t].queries = queries;
        cl[t].nq_total = nq;
        cl[t].first = (size_t)t * (size_t)per;
        cl[t].per_thread = per;
        cl[t].dim = (int)s.dim;
        cl[t].k = k;
        cl[t].L = L;
        cl[t].rescore = rescore;
        cl[t].tids = (uint64_t *)malloc((size_t)per * (size_t)k * 8);
    }
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < T; t++) pthread_create(&th[t], NULL, client_main, &cl[t]);
    for (int t = 0; t < T; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    unsigned long rows = 0;
    uint64_t checksum = 0;
    for (int t = 0; t < T; t++) {
        if (cl[t].rc) {
            fprintf(stderr, "client %d failed (%d)\n", t, cl[t].rc);
            return 1;
        }
        rows += cl[t].rows;
        for (size_t i = 0; i < (size_t)per * (size_t)k; i++) checksum = checksum * 1099511628211ull + cl[t].tids[i];
    }
    uint64_t batches = 0, qs = 0, largest = 0;
    dann_coalescer_stats(co, &batches, &qs, &largest);
    const double sec = (double)(t1.tv_sec - t0.tv_sec) + (double)(t1.tv_nsec - t0.tv_nsec) * 1e-9;
    printf("{\"clients\": %d, \"queries\": %llu, \"seconds\": %.6f, \"queries_per_s\": %.1f, \"batches\": %llu, "
           "\"mean_batch\": %.2f, \"largest_batch\": %llu, \"rows\": %lu, \"tid_checksum\": %llu}\n",
           T, (unsigned long long)qs, sec, (double)qs / sec, (unsigned long long)batches,
           batches ? (double)qs / (double)batches : 0.0, (unsigned long long)largest, rows, (unsigned long long)checksum);
    dann_coalescer_destroy(co);
    dann_index_free(ix);
    return 0;
}
