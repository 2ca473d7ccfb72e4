/* executor_harness.c — plays the Postgres executor's role against the C ABI, in plain C99:
 * dann_index_load -> dann_scan_begin -> dann_scan_rescan -> dann_scan_gettuple x k -> dann_scan_end.
 * This is the call sequence the Rust/pgrx shim of INTEGRATION.md performs from ambeginscan / amrescan /
 * amgettuple / amendscan; it exists to show that include/diskann_b200.h is a self-contained C header and
 * that the library links without any C++/CUDA/torch dependency leaking into its interface.
 *
 *   gcc -std=c99 -Wall -Wextra -Werror -Iinclude harness/executor_harness.c \
 *       -Lpgvectorscale_b200 -ldiskann_b200 -Wl,-rpath,$PWD/pgvectorscale_b200 -o /tmp/executor_harness
 *
 * Exit code: 0 = a scan ran and returned rows, 3 = no CUDA device (DANN_ERR_NO_DEVICE reported cleanly),
 * 1 = any other failure. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "diskann_b200.h"

enum { N = 64, DIM = 8, R = 4 };

int main(void) {
    static float vectors[N * DIM], mean[DIM], m2[DIM];
    static uint64_t codes[N], tids[N];
    static uint32_t nbrs[N * R];
    for (int i = 0; i < N; i++) {
        for (int d = 0; d < DIM; d++) vectors[i * DIM + d] = (float)((i * 7 + d * 3) % 11) - 5.0f;
        codes[i] = (uint64_t)i * 0x9E3779B97F4A7C15ull & 0xFFFFull; /* 8 dims x 2 bits = 16 bits */
        tids[i] = ((uint64_t)(i / 2) << 16) | (uint64_t)(i % 2 + 1);
        for (int j = 0; j < R; j++) nbrs[i * R + j] = (uint32_t)((i + 1 + j * 5) % N);
    }
    for (int d = 0; d < DIM; d++) {
        mean[d] = 0.0f;
        m2[d] = (float)N;
    }
    dann_snapshot_desc s;
    memset(&s, 0, sizeof s);
    s.n = N;
    s.dim = DIM;
    s.dim_index = DIM;
    s.bits = 2;
    s.words = 1;
    s.R = R;
    s.distance_type = DANN_L2;
    s.count = N;
    s.mean = mean;
    s.m2 = m2;
    s.codes = codes;
    s.nbrs = nbrs;
    s.heap_tid = tids;
    s.vectors = vectors;
    s.start_default = 0;

    dann_index *ix = NULL;
    int rc = dann_index_load(&s, 0, &ix);
    if (rc == DANN_ERR_NO_DEVICE) {
        fprintf(stderr, "harness: %s\n", dann_last_error());
        return 3;
    }
    if (rc != DANN_OK) {
        fprintf(stderr, "harness: load failed (%d): %s\n", rc, dann_last_error());
        return 1;
    }
    dann_scan *sc = NULL;
    float query[DIM] = {1, 0, -1, 2, 0, 0, 3, -2};
    if (dann_scan_begin(ix, &sc) != DANN_OK || dann_scan_rescan(sc, query, NULL, -1, 100, 50) != DANN_OK) {
        fprintf(stderr, "harness: %s\n", dann_last_error());
        return 1;
    }
    int rows = 0;
    for (;;) {
        uint32_t block, node;
        uint16_t offset;
        float dist;
        rc = dann_scan_gettuple(sc, &block, &offset, &node, &dist);
        if (rc <= 0) break;
        if (rows < 5) printf("row %d: tid (%u,%u) node %u dist %g\n", rows, block, (unsigned)offset, node, dist);
        rows++;
    }
    dann_query_stats st;
    dann_scan_stats(sc, &st);
    printf("%d rows; visits=%u d_quantized=%u d_full=%u\n", rows, st.visits, st.d_quantized, st.d_full);
    dann_scan_end(sc);
    dann_index_free(ix);
    return rc < 0 || rows == 0 ? 1 : 0;
}
