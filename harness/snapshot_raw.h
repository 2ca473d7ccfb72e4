/* snapshot_raw.h — reads the flat snapshot file written by Snapshot.save_raw (pgvectorscale_b200/snapshot.py) into a
 * dann_snapshot_desc.  Plain C99, for hosts that are not Python (the harnesses here; a sidecar).  The arrays stay in
 * one malloc'ed buffer that the caller frees after dann_index_load has copied them to the GPU. */
#ifndef DANN_SNAPSHOT_RAW_H
#define DANN_SNAPSHOT_RAW_H
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "diskann_b200.h"

/* returns the buffer to free(), or NULL on error; *index_vectors is set for plain-layout snapshots */
static void *dann_snapshot_raw_read(const char *path, dann_snapshot_desc *s, const float **index_vectors) {
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    unsigned char *buf = (unsigned char *)malloc((size_t)size + 64);
    if (!buf || fread(buf, 1, (size_t)size, f) != (size_t)size) {
        fclose(f);
        free(buf);
        return NULL;
    }
    fclose(f);
    if (size < 8 + 16 * 8 || memcmp(buf, "DANNSNP1", 8) != 0) {
        free(buf);
        return NULL;
    }
    uint64_t h[16];
    memcpy(h, buf + 8, sizeof h);
    memset(s, 0, sizeof *s);
    s->n = (uint32_t)h[0];
    s->dim = (uint32_t)h[1];
    s->dim_index = (uint32_t)h[2];
    s->bits = (uint32_t)h[3];
    s->words = (uint32_t)h[4];
    s->R = (uint32_t)h[5];
    s->distance_type = (int32_t)(uint32_t)h[6];
    s->has_labels = (int32_t)h[7];
    s->count = h[8];
    s->start_default = (uint32_t)h[9];
    s->n_start_labels = (uint32_t)h[10];
    const void *arr[11];
    uint64_t alen[11];
    size_t off = 8 + 16 * 8;
    for (int i = 0; i < 11; i++) {
        uint64_t len;
        if ((size_t)size - off < 8) { /* off <= size always holds here */
            free(buf);
            return NULL;
        }
        memcpy(&len, buf + off, 8);
        off += 8;
        if (len > (uint64_t)((size_t)size - off)) { /* untrusted length: compare without adding to it */
            free(buf);
            return NULL;
        }
        arr[i] = len ? (const void *)(buf + off) : NULL;
        alen[i] = len;
        uint64_t adv = len + (64 - len % 64) % 64;
        if (adv > (uint64_t)((size_t)size - off)) adv = (uint64_t)((size_t)size - off); /* padding of the last array may be cut */
        off += (size_t)adv;
    }
    /* every array's byte length must match the header geometry (64-bit products: n, dim, words, R are 32-bit) - a
     * truncated or inconsistent file must not hand dann_index_load pointers it would read past */
    {
        const uint64_t n = s->n, plain = h[11];
        const uint64_t nlab = (s->has_labels && alen[8] >= (n + 1) * 4) ? ((const uint32_t *)arr[8])[n] : 0;
        const uint64_t want[11] = {
            (uint64_t)s->dim_index * 4,                 /* mean */
            (uint64_t)s->dim_index * 4,                 /* m2 (may be absent for 1-bit codes) */
            n * s->words * 8,                           /* codes */
            n * s->R * 4,                               /* nbrs */
            n * 8,                                      /* heap_tid */
            n * s->dim * 4,                             /* vectors (optional: dann_index_set_vectors) */
            (uint64_t)s->n_start_labels * 2,            /* start_labels */
            (uint64_t)s->n_start_labels * 4,            /* start_label_nodes */
            s->has_labels ? (n + 1) * 4 : 0,            /* label_off */
            nlab * 2,                                   /* labels */
            plain ? n * s->dim_index * 4 : 0,           /* index_vectors (plain layout only) */
        };
        const int optional[11] = {plain ? 1 : 0, 1, plain ? 1 : 0, 0, 0, 1, 0, 0, 0, 0, 0};
        for (int i = 0; i < 11; i++) {
            if (alen[i] == want[i]) continue;
            if (alen[i] == 0 && optional[i]) continue;
            free(buf);
            return NULL;
        }
        if (plain > 1 || (plain && !arr[10] && n)) { /* storage_type must agree with the presence of index_vectors */
            free(buf);
            return NULL;
        }
    }
    s->mean = (const float *)arr[0];
    s->m2 = (const float *)arr[1];
    s->codes = (const uint64_t *)arr[2];
    s->nbrs = (const uint32_t *)arr[3];
    s->heap_tid = (const uint64_t *)arr[4];
    s->vectors = (const float *)arr[5];
    s->start_labels = (const int16_t *)arr[6];
    s->start_label_nodes = (const uint32_t *)arr[7];
    s->label_off = (const uint32_t *)arr[8];
    s->labels = (const int16_t *)arr[9];
    if (index_vectors) *index_vectors = h[11] ? (const float *)arr[10] : NULL;
    return buf;
}
#endif
