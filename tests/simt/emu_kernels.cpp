// emu_kernels.cpp — the non-graph kernels of the scan path (pgvectorscale_b200/csrc/dann_kernels.cuh: query
// preparation, row normalisation, rerank window) under the CPU SIMT emulator, so that a whole dann_search_batch call
// can be replayed on the CPU: emu_prepare -> emu_search (emu_search.cpp) -> emu_rerank.  TEST INFRASTRUCTURE ONLY.
#include <cuda_runtime.h> /* tests/simt/shim */

#include <vector>

#include "dann_kernels.cuh"

static IndexView view_of(const dann_snapshot_desc *s) {
    IndexView v{};
    v.n = s->n;
    v.dim = s->dim;
    v.dim_index = s->dim_index;
    v.bits = s->bits;
    v.words = s->words;
    v.cw = (s->words + 1u) & ~1u;
    v.R = s->R;
    v.Rp = (s->R + 7u) & ~7u;
    v.distance_type = s->distance_type;
    v.has_labels = 0;
    v.count = s->count;
    v.mean = s->mean;
    v.m2 = s->m2;
    v.tids = s->heap_tid;
    return v;
}

/* amrescan's vector preparation for B queries: q_full [B][dim], q_codes [B][words] */
extern "C" int emu_prepare(const dann_snapshot_desc *s, const float *queries, int B, float *q_full, uint64_t *q_codes) {
    IndexView v = view_of(s);
    std::vector<uint64_t> codes((size_t)B * v.cw);
    simt::launch((unsigned)B, 128, [&] { dann_prepare_kernel(v, queries, B, q_full, codes.data()); });
    for (int b = 0; b < B; b++)
        for (uint32_t w = 0; w < s->words; w++) q_codes[(size_t)b * s->words + w] = codes[(size_t)b * v.cw + w];
    return 0;
}

extern "C" int emu_prepare_plain(uint32_t dim, uint32_t dim_index, int cosine, const float *queries, int B, float *q_full,
                                 float *q_index) {
    simt::launch((unsigned)B, 128, [&] { dann_prepare_plain_kernel(dim, dim_index, cosine, queries, q_full, q_index); });
    return 0;
}

/* dann_index_load's normalisation of the heap vectors (cosine) followed by the rerank kernel over the streams the
 * search produced.  out_* as in dann_search_batch; stats[b].d_full is filled, the other counters are left alone. */
extern "C" int emu_rerank(const dann_snapshot_desc *s, const float *q_full, const uint32_t *stream,
                          const uint32_t *stream_len, int B, uint32_t c_target, uint32_t k, uint32_t rescore,
                          uint64_t *out_tid, float *out_dist, uint32_t *out_node, uint32_t *out_count,
                          dann_query_stats *stats) {
    IndexView v = view_of(s);
    std::vector<float4> store(((size_t)s->n * s->dim + 3) / 4 + 1);
    float *vec = reinterpret_cast<float *>(store.data());
    memcpy(vec, s->vectors, (size_t)s->n * s->dim * sizeof(float));
    if (s->distance_type == DANN_COSINE && s->n)
        simt::launch(4, 256, [&] { dann_normalize_rows_kernel(vec, s->n, s->dim); });
    v.vectors = vec;
    std::vector<float4> qstore(((size_t)B * s->dim + 3) / 4 + 1);
    memcpy(qstore.data(), q_full, (size_t)B * s->dim * sizeof(float));
    RerankArgs r{};
    r.ix = v;
    r.q_full = reinterpret_cast<const float *>(qstore.data());
    r.stream = stream;
    r.stream_len = stream_len;
    r.c_target = c_target;
    r.k = k;
    r.rescore = rescore;
    r.out_tid = out_tid;
    r.out_dist = out_dist;
    r.out_node = out_node;
    r.out_count = out_count;
    r.stats = stats;
    simt::launch((unsigned)B, 128, [&] { dann_rerank_kernel(r); });
    return 0;
}

extern "C" int emu_plain_stats(dann_query_stats *stats, int B) {
    simt::launch((unsigned)(B + 127) / 128, 128, [&] { dann_plain_stats_kernel(stats, B); });
    return 0;
}
