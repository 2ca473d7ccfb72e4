// dann_pgreader.h — host-side reader of a pgvectorscale `diskann` index RELATION FILE and of the table's heap / TOAST
// files: Postgres pages -> the flat arrays of dann_snapshot_desc.  Pure host C++ (no CUDA); compiled into the C-ABI library and
// exported as dann_pg_* (include/diskann_b200.h).  Reference paths are relative to /root/reference/pgvectorscale/src/.
//
// What is read, and what pins its layout:
//   * the page itself - PageHeaderData, line pointers, MAXALIGNed items, special area: Postgres' documented on-disk
//     format (storage/bufpage.h, storage/itemid.h), the same for every supported server version (layout version 4);
//   * the special area the extension puts on every page it owns: TsvPageOpaqueData, #[repr(C)] { page_type: u8,
//     _reserved: u8, page_id: u16 = 0xAE24 }   util/page.rs:23-71 (PageInit with a 4-byte special -> the last 8 bytes);
//   * chained items (meta page, SbqMeans): every chunk starts with ArchivedChainItemHeader = ArchivedItemPointer
//     { block_number: u32, offset: u16 } (8 bytes), next == (InvalidBlockNumber, 0) ends the chain   util/chain.rs:26-33,
//     125-183; the meta header (block 0, item 1) is { magic_number: u32 = 768756476, version: u32 }  meta_page.rs:22-28,
//     181-189,386-419;
//   * node items (PageType::SbqNode pages, one rkyv archive per line pointer, written by Tape::write, util/tape.rs:50-72):
//     ClassicSbqNode { heap_item_pointer, bq_vector: Vec<u64>, neighbor_index_pointers: Vec<ItemPointer>,
//     _neighbor_vectors: Vec<Vec<u64>> } / LabeledSbqNode { .., labels: LabelSet{Vec<u16>} }   sbq/node.rs:26-42.
//
// The rkyv limitation.  rkyv 0.7 (Cargo.toml:32, default size_32 / native-endian) puts the archived root struct at the END
// of the item; an ArchivedVec is { RelPtr: i32 offset from the field's own address, len: u32 }, an ArchivedItemPointer
// { u32, u16, 2 bytes padding } - those are fixed by rkyv.  What is NOT fixed is the ORDER of the fields inside an
// archived struct: the derive emits repr(Rust) types (no #[archive_attr(repr(C))] in the reference), so rustc may
// reorder them, and no toolchain or index file exists in this sandbox to pin what it does.  For the node structs this
// reader therefore does not assume an order: all their fields are 8 bytes wide, so it tries every assignment of the
// root's 8-byte cells to (heap pointer, code vector, neighbour vector, fourth vector) on a sample of items and keeps
// the one under which every sampled item satisfies the format's invariants (vector lengths equal to the meta page's
// num_neighbors / code words, relative pointers landing inside the item with the element alignment, neighbour
// pointers naming live node items or being invalid); declaration order wins a tie; no fit -> DANN_ERR_FORMAT, never
// a guess.  The MetaPage body (14 fields of five different widths, four of them indistinguishable u32s) cannot be
// recovered that way: the caller passes its scalars (the Rust host reads them with MetaPage's own getters,
// meta_page.rs:212-282, exactly as it fills dann_snapshot_desc today) and this reader cross-checks them against the
// node items.  The same limitation is why parity here is "self-consistent" only: tests/pgpages.py writes relation
// images by the rules above, nothing in the sandbox can produce one with the reference itself.
//
// Invalidation rule (when an HBM snapshot made from this file is stale).  The extension changes the relation only
// through WritablePage::commit -> GenericXLogFinish (util/page.rs:224-231), which stamps pd_lsn; node items are never
// moved or removed (vacuum rewrites heap pointers / neighbour lists in place, inserts append items and rewrite
// neighbour lists), so   fingerprint = hash over every page of (block, pd_lsn, pd_lower, pd_upper, pd_checksum)
// changes iff some page changed, and nblocks / max_lsn order two fingerprints in time.  A snapshot is valid for
// exactly the fingerprint it was extracted under: the loader records it (dann_pg_snapshot.fingerprint), the host compares
// dann_pg_relation_stat() again before reusing a cached index handle - in a backend additionally on relcache
// invalidation and after its own aminsert / ambulkdelete - and reloads on any difference.
// A relation FILE shows only what has been written back: an external reader (the sidecar) needs a CHECKPOINT or a
// clean shutdown first; inside a backend the exporter of INTEGRATION.md §4b(ii) walks the buffer manager with the
// reference's own accessors instead.
//
// The table side (the vector column a rerank reads) is at the end of this file: heap tuples by TID, varlena headers,
// TOAST chunks - Postgres' own documented formats and pgvector's value layout, nothing from rkyv.
#pragma once
#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/diskann_b200.h"

namespace dannpg {

constexpr uint32_t BLCKSZ = 8192;
constexpr uint32_t PAGE_HEADER = 24;          /* SizeOfPageHeaderData */
constexpr uint16_t TSV_PAGE_ID = 0xAE24;      /* util/page.rs:23 */
constexpr uint32_t TSV_MAGIC = 768756476u;    /* meta_page.rs:22 */
constexpr uint32_t INVALID_BLOCK = 0xFFFFFFFFu;
constexpr uint32_t RELSEG_BLOCKS = 131072;    /* 1 GB segments of 8 KB pages */
enum { PT_META_V1 = 0, PT_NODE = 1, PT_PQ_DEF = 2, PT_PQ_VEC = 3, PT_SBQ_MEANS_V1 = 4, PT_SBQ_NODE = 5, PT_META_V2 = 6,
       PT_SBQ_MEANS = 7, PT_META = 8, PT_COUNT = 9 }; /* util/page.rs:28-38 */

static inline uint16_t rd16(const unsigned char *p) { uint16_t v; memcpy(&v, p, 2); return v; }
static inline uint32_t rd32(const unsigned char *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t rd64(const unsigned char *p) { uint64_t v; memcpy(&v, p, 8); return v; }

struct Segment {
    const unsigned char *base = nullptr;
    size_t bytes = 0;
};

struct Relation {
    std::vector<Segment> segs;
    uint32_t nblocks = 0;
    uint32_t seg_blocks = RELSEG_BLOCKS; /* RELSEG_SIZE; DANN_PG_RELSEG_BLOCKS overrides it (test hook, non-default builds) */
    std::string err;

    ~Relation() {
        for (auto &s : segs)
            if (s.base) munmap(const_cast<unsigned char *>(s.base), s.bytes);
    }
    const unsigned char *page(uint32_t block) const {
        if (block >= nblocks) return nullptr;
        const Segment &s = segs[block / seg_blocks];
        return s.base + (size_t)(block % seg_blocks) * BLCKSZ;
    }
};

/* one decoded page header; ok == false: not a page this reader may touch (reason in why) */
struct PageView {
    const unsigned char *p = nullptr;
    bool is_new = false; /* all-zero page (PageIsNew: pd_upper == 0), e.g. after an extension that was never initialised */
    bool ok = false;
    uint64_t lsn = 0;
    uint16_t checksum = 0, lower = 0, upper = 0, special = 0;
    uint8_t page_type = 0xFF;
    uint32_t nitems = 0;
    const char *why = "";
};

static inline PageView view_page(const unsigned char *p) {
    PageView v;
    v.p = p;
    v.lsn = ((uint64_t)rd32(p) << 32) | rd32(p + 4); /* PageXLogRecPtr { xlogid, xrecoff } */
    v.checksum = rd16(p + 8);
    v.lower = rd16(p + 12);
    v.upper = rd16(p + 14);
    v.special = rd16(p + 16);
    const uint16_t psv = rd16(p + 18);
    if (v.upper == 0) {
        v.is_new = true;
        v.why = "new (all-zero) page";
        return v;
    }
    if ((psv & 0xFF00u) != BLCKSZ || (psv & 0x00FFu) != 4) {
        v.why = "pd_pagesize_version is not 8192 | layout version 4";
        return v;
    }
    /* PageInit(page, BLCKSZ, sizeof(TsvPageOpaqueData) = 4) -> MAXALIGN(4) = 8 bytes of special space */
    if (v.special != BLCKSZ - 8 || v.lower < PAGE_HEADER || v.lower > v.upper || v.upper > v.special || ((v.lower - PAGE_HEADER) & 3)) {
        v.why = "pd_lower / pd_upper / pd_special are not those of a page initialised by the extension";
        return v;
    }
    if (rd16(p + v.special + 2) != TSV_PAGE_ID) {
        v.why = "special area does not carry the extension's page id 0xAE24";
        return v;
    }
    v.page_type = p[v.special];
    if (v.page_type >= PT_COUNT) {
        v.why = "unknown PageType";
        return v;
    }
    v.nitems = (v.lower - PAGE_HEADER) / 4;
    v.ok = true;
    return v;
}

/* line pointer `off` (1-based OffsetNumber) of a valid page -> item bytes; false unless LP_NORMAL and inside the page */
static inline bool page_item(const PageView &v, uint32_t off, const unsigned char **item, uint32_t *len) {
    if (off == 0 || off > v.nitems) return false;
    const uint32_t lp = rd32(v.p + PAGE_HEADER + 4 * (off - 1)); /* ItemIdData: lp_off:15, lp_flags:2, lp_len:15 */
    const uint32_t lp_off = lp & 0x7FFFu, lp_flags = (lp >> 15) & 3u, lp_len = lp >> 17;
    if (lp_flags != 1u /* LP_NORMAL */ || lp_len == 0) return false;
    if (lp_off < v.upper || lp_off + lp_len > v.special || (lp_off & 7u)) return false;
    *item = v.p + lp_off;
    *len = lp_len;
    return true;
}

static inline uint64_t mix64(uint64_t h, uint64_t x) { /* splitmix-style accumulate, order-sensitive */
    h ^= x + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xBF58476D1CE4E5B9ull;
    return h ^ (h >> 31);
}

/* util/chain.rs:125-183 ChainItemIterator: follow `next` from (block, offset), appending each chunk's payload */
static inline int read_chain(const Relation &r, uint32_t block, uint16_t offset, int expect_type, std::vector<unsigned char> &out,
                             std::string &err) {
    out.clear();
    uint32_t hops = 0;
    while (block != INVALID_BLOCK) {
        if (++hops > r.nblocks + 1u) {
            err = "chain does not terminate";
            return DANN_ERR_FORMAT;
        }
        const unsigned char *pg = r.page(block);
        if (!pg) {
            err = "chain points past the end of the relation";
            return DANN_ERR_FORMAT;
        }
        const PageView v = view_page(pg);
        if (!v.ok) {
            err = std::string("chain page: ") + v.why;
            return DANN_ERR_FORMAT;
        }
        if (expect_type >= 0 && v.page_type != expect_type) { /* assert!(page.get_type() == self.page_type) */
            err = "chain page has another PageType";
            return DANN_ERR_FORMAT;
        }
        const unsigned char *it;
        uint32_t len;
        if (!page_item(v, offset, &it, &len) || len <= 8) { /* assert!(slice.len() > CHAIN_ITEM_HEADER_SIZE) */
            err = "chain item missing or shorter than its header";
            return DANN_ERR_FORMAT;
        }
        out.insert(out.end(), it + 8, it + len);
        block = rd32(it);
        offset = rd16(it + 4);
    }
    return DANN_OK;
}

/* ---- rkyv node items ------------------------------------------------------------------------------------------- */
struct NodeLayout {
    int cell_heap = 0, cell_code = 1, cell_nbrs = 2, cell_fourth = 3; /* which 8-byte cell of the 32-byte root holds what */
};

struct NodeFields {
    uint32_t heap_block;
    uint16_t heap_offset;
    const unsigned char *code;  /* words * 8 bytes */
    const unsigned char *nbrs;  /* R * 8 bytes: ArchivedItemPointer each */
    const unsigned char *fourth; /* labeled: u16 labels; classic: ArchivedVec<u64> cells (must be empty) */
    uint32_t n_fourth;
};

/* ArchivedVec cell at `cell` inside item[0..len): { i32 rel, u32 n }, target = cell address + rel */
static inline bool vec_cell(const unsigned char *item, uint32_t len, uint32_t cell_off, uint32_t elem_size, uint32_t elem_align,
                            const unsigned char **data, uint32_t *n) {
    const int32_t rel = (int32_t)rd32(item + cell_off);
    const uint32_t cnt = rd32(item + cell_off + 4);
    const int64_t tgt = (int64_t)cell_off + rel;
    if (cnt == 0) { /* an empty vector's pointer is not dereferenced; rkyv leaves it pointing at the write position */
        *data = item;
        *n = 0;
        return tgt >= 0 && tgt <= (int64_t)len;
    }
    if (tgt < 0 || (uint64_t)tgt + (uint64_t)cnt * elem_size > len || ((uint64_t)tgt % elem_align)) return false;
    *data = item + tgt;
    *n = cnt;
    return true;
}

/* what the four 8-byte cells of a node root are, per storage layout:
 *   SBQ   (sbq/node.rs:26-42):   heap pointer | bq_vector: Vec<u64> [words] | neighbours [R] | Vec<Vec<u64>> (empty) or labels: Vec<u16>
 *   plain (plain/node.rs:15-22): heap pointer | vector: Vec<f32> [dim_index] | neighbours [R] | pq_vector: Vec<u8> (empty)
 * (the struct's declaration order differs - plain declares vector, pq_vector, neighbours, heap pointer - see decl_order) */
struct NodeShape {
    uint32_t main_len, main_elem, main_align; /* the code / vector cell */
    uint32_t R;
    uint32_t fourth_elem, fourth_align;
    bool fourth_must_be_empty;
};

static inline bool parse_node(const unsigned char *item, uint32_t len, const NodeLayout &lay, const NodeShape &sh, NodeFields *f) {
    if (len < 32 || (len & 3)) return false;
    const uint32_t root = len - 32; /* archived root = the last size_of::<ArchivedNode>() bytes */
    const unsigned char *hp = item + root + 8 * lay.cell_heap;
    f->heap_block = rd32(hp);
    f->heap_offset = rd16(hp + 4);
    uint32_t n;
    if (!vec_cell(item, len, root + 8 * lay.cell_code, sh.main_elem, sh.main_align, &f->code, &n) || n != sh.main_len) return false;
    if (!vec_cell(item, len, root + 8 * lay.cell_nbrs, 8, 4, &f->nbrs, &n) || n != sh.R) return false;
    if (!vec_cell(item, len, root + 8 * lay.cell_fourth, sh.fourth_elem, sh.fourth_align, &f->fourth, &f->n_fourth)) return false;
    if (sh.fourth_must_be_empty && f->n_fourth != 0) return false; /* _neighbor_vectors / pq_vector: "no longer used", always empty */
    return true;
}

/* ---- whole-relation passes ------------------------------------------------------------------------------------------ */
static inline int open_relation(const char *path, Relation *r) {
    if (const char *e = getenv("DANN_PG_RELSEG_BLOCKS")) {
        const unsigned long v = strtoul(e, nullptr, 10);
        if (v >= 1 && v <= RELSEG_BLOCKS) r->seg_blocks = (uint32_t)v;
    }
    const size_t seg_bytes = (size_t)r->seg_blocks * BLCKSZ;
    for (uint32_t seg = 0;; seg++) { /* relfilenode, relfilenode.1, relfilenode.2, ...: 1 GB each but the last */
        std::string p = path;
        if (seg) p += "." + std::to_string(seg);
        const int fd = open(p.c_str(), O_RDONLY);
        if (fd < 0) {
            if (seg == 0) {
                r->err = "cannot open " + p + ": " + strerror(errno);
                return DANN_ERR_INVALID_ARG;
            }
            break;
        }
        struct stat st;
        if (fstat(fd, &st) != 0 || st.st_size % BLCKSZ != 0 || (uint64_t)st.st_size > seg_bytes) {
            close(fd);
            r->err = p + ": size is not a whole number of 8 KB pages (or exceeds one segment)";
            return DANN_ERR_FORMAT;
        }
        Segment s;
        s.bytes = (size_t)st.st_size;
        if (s.bytes) {
            void *m = mmap(nullptr, s.bytes, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m == MAP_FAILED) {
                close(fd);
                r->err = "mmap failed for " + p;
                return DANN_ERR_OOM;
            }
            s.base = static_cast<const unsigned char *>(m);
        }
        close(fd);
        if (!r->segs.empty() && r->segs.back().bytes != seg_bytes) {
            r->err = "segment before " + p + " is not full";
            return DANN_ERR_FORMAT;
        }
        r->segs.push_back(s);
        if ((uint64_t)r->nblocks + s.bytes / BLCKSZ > 0xFFFFFFFEull) {
            r->err = "relation too large";
            return DANN_ERR_FORMAT;
        }
        r->nblocks += (uint32_t)(s.bytes / BLCKSZ);
        if (s.bytes != seg_bytes) break;
    }
    return DANN_OK;
}

static inline void stat_relation(const Relation &r, dann_pg_relation_info *o) {
    memset(o, 0, sizeof *o);
    o->nblocks = r.nblocks;
    uint64_t fp = 0x243F6A8885A308D3ull ^ r.nblocks;
    for (uint32_t b = 0; b < r.nblocks; b++) {
        const PageView v = view_page(r.page(b));
        fp = mix64(fp, ((uint64_t)b << 32) ^ v.lsn);
        fp = mix64(fp, ((uint64_t)v.checksum << 32) | ((uint64_t)v.lower << 16) | v.upper);
        o->max_lsn = std::max(o->max_lsn, v.lsn);
        if (v.is_new) {
            o->new_pages++;
            continue;
        }
        if (!v.ok) {
            o->foreign_pages++;
            continue;
        }
        o->pages_by_type[v.page_type]++;
        if (v.page_type == PT_SBQ_NODE) {
            const unsigned char *it;
            uint32_t len;
            for (uint32_t off = 1; off <= v.nitems; off++)
                if (page_item(v, off, &it, &len)) o->node_items++;
        }
    }
    o->fingerprint = fp;
    /* MetaPageHeader: block 0, item 1, behind the 8-byte chain header (meta_page.rs:26-28,368-384) */
    if (r.nblocks) {
        const PageView v = view_page(r.page(0));
        const unsigned char *it;
        uint32_t len;
        if (v.ok && v.page_type == PT_META && page_item(v, 1, &it, &len) && len == 16) {
            const uint32_t a = rd32(it + 8), b = rd32(it + 12);
            /* two u32 fields: the magic number identifies itself whichever way rustc ordered them */
            if (a == TSV_MAGIC) o->meta_magic = a, o->meta_version = b;
            else if (b == TSV_MAGIC) o->meta_magic = b, o->meta_version = a;
        }
    }
}

struct SbqOut { /* owns everything dann_pg_snapshot points into */
    dann_pg_snapshot pub;
    std::vector<float> index_vectors; /* plain layout */
    std::vector<uint64_t> codes, heap_tid, index_tid;
    std::vector<uint32_t> nbrs, label_off, start_label_nodes;
    std::vector<int16_t> labels, start_labels;
    std::vector<float> mean, m2;
};

static inline uint64_t ip_key(uint32_t block, uint16_t off) { return ((uint64_t)block << 16) | off; }

/* SbqMeans { count: u64, means: Vec<f32>, m2: Vec<f32> } (sbq/mod.rs:79-86): 24-byte root at the end of the chain's bytes */
static inline int parse_means(const std::vector<unsigned char> &buf, uint32_t dim_index, SbqOut *o, std::string &err) {
    const uint32_t len = (uint32_t)buf.size();
    if (len < 24 || (len & 3)) {
        err = "SbqMeans archive too short";
        return DANN_ERR_FORMAT;
    }
    const unsigned char *b = buf.data();
    const uint32_t root = len - 24;
    int found = -1;
    const unsigned char *d0 = nullptr, *d1 = nullptr;
    uint32_t n0 = 0, n1 = 0;
    for (int cc = 0; cc < 3; cc++) { /* which cell is `count`; the other two are the vectors */
        const int va = cc == 0 ? 1 : 0, vb = cc == 2 ? 1 : 2;
        const unsigned char *da, *db;
        uint32_t na, nb;
        if (!vec_cell(b, len, root + 8 * va, 4, 4, &da, &na) || !vec_cell(b, len, root + 8 * vb, 4, 4, &db, &nb)) continue;
        if (na != dim_index || (nb != dim_index && nb != 0)) continue;
        if (found >= 0 && cc != 0) continue; /* declaration order (count first) wins a tie */
        found = cc;
        d0 = da, n0 = na, d1 = db, n1 = nb;
        if (cc == 0) break;
    }
    if (found < 0) {
        err = "SbqMeans archive: no field order satisfies means.len == m2.len == num_dimensions_to_index";
        return DANN_ERR_FORMAT;
    }
    if (n1 && d1 < d0) std::swap(d0, d1); /* `means` is serialised before `m2`: the lower address */
    o->pub.snap.count = rd64(b + root + 8 * found);
    o->mean.resize(dim_index);
    memcpy(o->mean.data(), d0, (size_t)dim_index * 4);
    if (n1) {
        o->m2.resize(dim_index);
        memcpy(o->m2.data(), d1, (size_t)dim_index * 4);
    }
    return DANN_OK;
}

static inline int extract_nodes(const Relation &r, const dann_pg_meta *m, bool plain, SbqOut *o, std::string &err) {
    const uint32_t R = m->num_neighbors, bits = plain ? 1u : m->bq_bits, dimx = m->num_dimensions_to_index;
    if (!R || !bits || !dimx || m->num_dimensions < dimx || (uint64_t)dimx * bits > (1u << 24)) {
        err = "meta scalars: num_neighbors, bq_bits and num_dimensions_to_index (<= num_dimensions) must be positive";
        return DANN_ERR_INVALID_ARG;
    }
    const uint32_t words = (uint32_t)(((uint64_t)dimx * bits + 63) / 64); /* sbq/quantize.rs:38-46 */
    const bool labeled = m->has_labels != 0;
    if (plain && labeled) { /* build.rs:264-290: label filtering needs the memory_optimized layout */
        err = "a plain-storage index has no labels";
        return DANN_ERR_INVALID_ARG;
    }
    const int node_page = plain ? PT_NODE : PT_SBQ_NODE; /* plain/storage.rs:124-126, sbq/storage.rs page_type() */
    NodeShape sh;
    sh.main_len = plain ? dimx : words;
    sh.main_elem = plain ? 4u : 8u;
    sh.main_align = plain ? 4u : 8u;
    sh.R = R;
    sh.fourth_elem = plain ? 1u : (labeled ? 2u : 8u);
    sh.fourth_align = plain ? 1u : (labeled ? 2u : 4u);
    sh.fourth_must_be_empty = plain || !labeled;
    /* the struct's declaration order, as cells (heap, main, neighbours, fourth): PlainNode declares vector, pq_vector,
     * neighbor_index_pointers, heap_item_pointer */
    const int decl[4] = {plain ? 3 : 0, plain ? 0 : 1, 2, plain ? 1 : 3};
    /* pass 1: every live item of every SbqNode page, in (block, offset) order = dense node ids */
    std::vector<std::pair<const unsigned char *, uint32_t>> items;
    for (uint32_t b = 0; b < r.nblocks; b++) {
        const PageView v = view_page(r.page(b));
        if (v.is_new) continue;
        if (!v.ok) {
            err = "block " + std::to_string(b) + ": " + v.why;
            return DANN_ERR_FORMAT;
        }
        if (v.page_type != node_page) continue;
        for (uint32_t off = 1; off <= v.nitems; off++) {
            const unsigned char *it;
            uint32_t len;
            if (!page_item(v, off, &it, &len)) continue; /* unused / dead line pointers are not nodes */
            o->index_tid.push_back(ip_key(b, (uint16_t)off));
            items.emplace_back(it, len);
        }
    }
    const uint64_t n64 = items.size();
    if (n64 >= 0xFFFFFFFFull) {
        err = "more than 2^32 - 2 nodes";
        return DANN_ERR_CAPACITY;
    }
    const uint32_t n = (uint32_t)n64;
    /* IndexPointer -> dense id: the items of one block are consecutive in index_tid (at most a few hundred), so a
     * per-block start index turns every lookup into a search inside one page's worth of entries */
    std::vector<uint32_t> block_start((size_t)r.nblocks + 1, n);
    for (uint32_t i = n; i-- > 0;) block_start[(size_t)(o->index_tid[i] >> 16)] = i;
    for (uint32_t b = r.nblocks; b-- > 0;)
        if (block_start[b] > block_start[b + 1]) block_start[b] = block_start[b + 1]; /* blocks without node items */
    auto dense_of = [&](uint32_t block, uint16_t off, uint32_t *id) {
        if (block >= r.nblocks) return false;
        const uint64_t k = ip_key(block, off);
        auto lo = o->index_tid.begin() + block_start[block], hi = o->index_tid.begin() + block_start[block + 1];
        auto it = std::lower_bound(lo, hi, k);
        if (it == hi || *it != k) return false;
        *id = (uint32_t)(it - o->index_tid.begin());
        return true;
    };
    /* the root's field order: every assignment of its four 8-byte cells, judged on a sample spread over the relation */
    NodeLayout lay;
    if (n) {
        const uint32_t sample = std::min<uint32_t>(n, 256);
        int perm[4] = {0, 1, 2, 3}, best[4] = {-1, -1, -1, -1};
        int nfit = 0;
        do {
            NodeLayout c;
            c.cell_heap = perm[0], c.cell_code = perm[1], c.cell_nbrs = perm[2], c.cell_fourth = perm[3];
            bool fit = true;
            for (uint32_t s = 0; s < sample && fit; s++) {
                const auto &it = items[(uint64_t)s * n / sample];
                NodeFields f;
                if (!parse_node(it.first, it.second, c, sh, &f)) {
                    fit = false;
                    break;
                }
                bool ended = false;
                for (uint32_t j = 0; j < R && fit; j++) { /* ArchivedItemPointer: zero padding, invalid or a live node */
                    const uint32_t nb = rd32(f.nbrs + 8 * j);
                    const uint16_t no = rd16(f.nbrs + 8 * j + 4);
                    uint32_t id;
                    if (rd16(f.nbrs + 8 * j + 6) != 0) fit = false;
                    else if (nb == INVALID_BLOCK) ended = true;
                    else if (!ended && !dense_of(nb, no, &id)) fit = false;
                }
                if (rd16(it.first + it.second - 32 + 8 * c.cell_heap + 6) != 0) fit = false; /* the heap pointer's padding */
            }
            if (fit) {
                const bool is_decl = perm[0] == decl[0] && perm[1] == decl[1] && perm[2] == decl[2] && perm[3] == decl[3];
                if (nfit == 0 || is_decl) memcpy(best, perm, sizeof best);
                nfit++;
                if (is_decl) break;
            }
        } while (std::next_permutation(perm, perm + 4));
        if (nfit == 0) {
            err = std::string("node items: no order of the archived root's fields satisfies ") + (plain ? "vector.len == " : "bq_vector.len == ") +
                  std::to_string(sh.main_len) +
                  ", neighbor_index_pointers.len == " + std::to_string(R) + " with pointers to live node items (wrong meta scalars, "
                  "another storage layout, or an rkyv layout this reader does not know)";
            return DANN_ERR_FORMAT;
        }
        lay.cell_heap = best[0], lay.cell_code = best[1], lay.cell_nbrs = best[2], lay.cell_fourth = best[3];
    }
    o->pub.layout[0] = lay.cell_heap, o->pub.layout[1] = lay.cell_code, o->pub.layout[2] = lay.cell_nbrs, o->pub.layout[3] = lay.cell_fourth;
    /* pass 2: every item into the flat arrays */
    if (plain) o->index_vectors.resize((size_t)n * dimx);
    else o->codes.resize((size_t)n * words);
    o->nbrs.assign((size_t)n * R, DANN_INVALID_NODE);
    o->heap_tid.resize(n);
    if (labeled) o->label_off.assign((size_t)n + 1, 0);
    for (uint32_t i = 0; i < n; i++) {
        NodeFields f;
        if (!parse_node(items[i].first, items[i].second, lay, sh, &f)) {
            err = "node item (" + std::to_string(o->index_tid[i] >> 16) + "," + std::to_string(o->index_tid[i] & 0xFFFF) +
                  ") does not parse under the layout the sample fixed";
            return DANN_ERR_FORMAT;
        }
        if (plain) memcpy(&o->index_vectors[(size_t)i * dimx], f.code, (size_t)dimx * 4);
        else memcpy(&o->codes[(size_t)i * words], f.code, (size_t)words * 8);
        /* heap pointer: offset 0 (InvalidOffsetNumber) = deleted tuple, kept as such (scan.rs:231-234) */
        o->heap_tid[i] = ip_key(f.heap_block, f.heap_offset);
        for (uint32_t j = 0; j < R; j++) { /* iter_neighbors: up to the first InvalidBlockNumber (sbq/node.rs:261-285) */
            const uint32_t nb = rd32(f.nbrs + 8 * j);
            if (nb == INVALID_BLOCK) break;
            uint32_t id;
            if (!dense_of(nb, rd16(f.nbrs + 8 * j + 4), &id)) {
                err = "node item (" + std::to_string(o->index_tid[i] >> 16) + "," + std::to_string(o->index_tid[i] & 0xFFFF) +
                      "): neighbour " + std::to_string(j) + " is not a live node item";
                return DANN_ERR_FORMAT;
            }
            o->nbrs[(size_t)i * R + j] = id;
        }
        if (labeled) {
            std::vector<int16_t> ls(f.n_fourth);
            for (uint32_t j = 0; j < f.n_fourth; j++) ls[j] = (int16_t)rd16(f.fourth + 2 * j);
            std::sort(ls.begin(), ls.end()); /* LabelSet is sorted + deduplicated at construction (labels/mod.rs:30-37) */
            ls.erase(std::unique(ls.begin(), ls.end()), ls.end());
            o->labels.insert(o->labels.end(), ls.begin(), ls.end());
            if (o->labels.size() > 0xFFFFFFFFull) {
                err = "label CSR exceeds 2^32 entries";
                return DANN_ERR_CAPACITY;
            }
            o->label_off[i + 1] = (uint32_t)o->labels.size();
        }
    }
    /* start nodes (graph/start_nodes.rs:16-48): IndexPointers -> dense ids */
    dann_snapshot_desc &sn = o->pub.snap;
    sn.start_default = DANN_INVALID_NODE;
    if (m->start_block != INVALID_BLOCK && !dense_of(m->start_block, m->start_offset, &sn.start_default)) {
        err = "start_nodes.default_node is not a live node item";
        return DANN_ERR_FORMAT;
    }
    if (m->n_start_labels) {
        if (!m->start_labels || !m->start_label_block || !m->start_label_offset) {
            err = "n_start_labels > 0 without the three arrays";
            return DANN_ERR_INVALID_ARG;
        }
        std::vector<std::pair<int16_t, uint32_t>> sl(m->n_start_labels);
        for (uint32_t i = 0; i < m->n_start_labels; i++) {
            sl[i].first = m->start_labels[i];
            if (!dense_of(m->start_label_block[i], m->start_label_offset[i], &sl[i].second)) {
                err = "a labeled start node is not a live node item";
                return DANN_ERR_FORMAT;
            }
        }
        std::sort(sl.begin(), sl.end()); /* BTreeMap order */
        for (auto &e : sl) {
            o->start_labels.push_back(e.first);
            o->start_label_nodes.push_back(e.second);
        }
    }
    /* quantizer (SbqQuantizer::new: use_mean unless 1 bit with the default options - then there is no SbqMeans item) */
    if (plain) {
        /* no quantizer */
    } else if (m->means_block != INVALID_BLOCK) {
        std::vector<unsigned char> buf;
        int rc = read_chain(r, m->means_block, m->means_offset, PT_SBQ_MEANS, buf, err);
        if (rc != DANN_OK) return rc;
        rc = parse_means(buf, dimx, o, err);
        if (rc != DANN_OK) return rc;
    } else {
        o->mean.assign(dimx, 0.0f);
    }
    sn.n = n;
    sn.dim = m->num_dimensions;
    sn.dim_index = dimx;
    sn.bits = plain ? 0u : bits;
    sn.words = plain ? 0u : words;
    sn.R = R;
    sn.distance_type = m->distance_type;
    sn.has_labels = labeled ? 1 : 0;
    sn.mean = o->mean.empty() ? nullptr : o->mean.data();
    sn.m2 = o->m2.empty() ? nullptr : o->m2.data();
    sn.codes = plain ? nullptr : o->codes.data();
    o->pub.index_vectors = plain ? o->index_vectors.data() : nullptr;
    sn.nbrs = o->nbrs.data();
    sn.heap_tid = o->heap_tid.data();
    sn.vectors = nullptr; /* heap rows live in the TABLE, not in the index: the caller fetches heap_tid[i] in order */
    sn.n_start_labels = (uint32_t)o->start_labels.size();
    sn.start_labels = o->start_labels.empty() ? nullptr : o->start_labels.data();
    sn.start_label_nodes = o->start_label_nodes.empty() ? nullptr : o->start_label_nodes.data();
    sn.label_off = labeled ? o->label_off.data() : nullptr;
    sn.labels = labeled ? (o->labels.empty() ? reinterpret_cast<const int16_t *>(o->label_off.data()) : o->labels.data()) : nullptr;
    o->pub.index_tid = o->index_tid.data();
    dann_pg_relation_info info;
    stat_relation(r, &info);
    o->pub.fingerprint = info.fingerprint;
    return DANN_OK;
}

/* ---- heap rows: the vector column of the TABLE the index was built on ------------------------------------------------
 * The rerank reads heap tuples by TID (sbq/storage.rs:304-328 -> TableSlot::from_index_heap_pointer, util/table_slot.rs:
 * 13-53) and detoasts the pgvector value (pg_vector.rs:125-199).  An external loader does the same from the heap's and
 * its TOAST table's relation files.  Everything here is Postgres' own documented on-disk format (htup_details.h,
 * postgres.h varlena headers, detoast.h varatt_external, heaptoast.h TOAST_MAX_CHUNK_SIZE) plus pgvector's value layout
 * (vector.h: int32 vl_len_, int16 dim, int16 unused, float x[dim]) - no rkyv involved.  Not handled, and refused
 * rather than guessed: compressed values (pgvector declares the type STORAGE external, so a vector is stored out of
 * line uncompressed or inline), attributes in front of the vector column whose type the caller did not describe.
 * No visibility test: the index names the tuples (HOT redirects are followed); the executor re-checks every row. */
struct HeapPage {
    const unsigned char *p = nullptr;
    bool ok = false;
    uint32_t nitems = 0;
    uint16_t lower = 0, upper = 0, special = 0;
};

static inline HeapPage view_heap_page(const unsigned char *p) {
    HeapPage v;
    v.p = p;
    v.lower = rd16(p + 12);
    v.upper = rd16(p + 14);
    v.special = rd16(p + 16);
    const uint16_t psv = rd16(p + 18);
    if (v.upper == 0) return v; /* new page */
    if ((psv & 0xFF00u) != BLCKSZ || (psv & 0x00FFu) != 4) return v;
    if (v.special > BLCKSZ || v.lower < PAGE_HEADER || v.lower > v.upper || v.upper > v.special || ((v.lower - PAGE_HEADER) & 3)) return v;
    v.nitems = (v.lower - PAGE_HEADER) / 4;
    v.ok = true;
    return v;
}

/* line pointer -> tuple bytes, following one LP_REDIRECT (a HOT chain's root) */
static inline bool heap_item(const HeapPage &v, uint32_t off, const unsigned char **item, uint32_t *len) {
    for (int hop = 0; hop < 2; hop++) {
        if (off == 0 || off > v.nitems) return false;
        const uint32_t lp = rd32(v.p + PAGE_HEADER + 4 * (off - 1));
        const uint32_t lp_off = lp & 0x7FFFu, lp_flags = (lp >> 15) & 3u, lp_len = lp >> 17;
        if (lp_flags == 2u) { /* LP_REDIRECT: lp_off holds the offset number of the live tuple */
            off = lp_off;
            continue;
        }
        if (lp_flags != 1u || lp_len < 23) return false;
        if (lp_off < v.upper || lp_off + lp_len > v.special) return false;
        *item = v.p + lp_off;
        *len = lp_len;
        return true;
    }
    return false;
}

struct Varlena { /* one decoded varlena header */
    const unsigned char *data = nullptr; /* payload (inline) */
    uint32_t len = 0;                    /* payload bytes (inline) */
    uint32_t total = 0;                  /* bytes the value occupies in the tuple, header included */
    bool external = false, compressed = false;
    uint32_t ext_rawsize = 0, ext_size = 0, ext_valueid = 0, ext_toastrelid = 0;
};

static inline bool decode_varlena(const unsigned char *p, const unsigned char *end, Varlena *v) {
    if (p >= end) return false;
    const unsigned char b = p[0];
    if (b == 0x01) { /* VARATT_IS_1B_E: TOAST pointer, 1-byte tag then the struct (unaligned) */
        if (p + 2 > end || p[1] != 18 /* VARTAG_ONDISK */ || p + 2 + 16 > end) return false;
        v->external = true;
        v->ext_rawsize = rd32(p + 2);
        const uint32_t extinfo = rd32(p + 6);
        v->ext_size = extinfo & 0x3FFFFFFFu;
        v->compressed = v->ext_size < v->ext_rawsize - 4u; /* VARATT_EXTERNAL_IS_COMPRESSED */
        v->ext_valueid = rd32(p + 10);
        v->ext_toastrelid = rd32(p + 14);
        v->total = 18;
        return true;
    }
    if (b & 0x01) { /* VARATT_IS_1B: short header, length includes the header byte */
        const uint32_t tot = b >> 1;
        if (tot < 1 || p + tot > end) return false;
        v->data = p + 1;
        v->len = tot - 1;
        v->total = tot;
        return true;
    }
    if (p + 4 > end) return false;
    const uint32_t h = rd32(p);
    const uint32_t tot = h >> 2;
    if (tot < 4 || p + tot > end) return false;
    v->compressed = (b & 0x03) == 0x02; /* VARATT_IS_4B_C */
    v->data = p + 4;
    v->len = tot - 4;
    v->total = tot;
    return true;
}

static inline uint32_t align_up(uint32_t off, char a) {
    const uint32_t n = a == 'd' ? 8u : a == 'i' ? 4u : a == 's' ? 2u : 1u;
    return (off + n - 1) & ~(n - 1);
}

/* walks a heap tuple to attribute `target` (0-based), given attlen / attalign of attributes 0..target (heap_deform_tuple's
 * rules: NULLs take no space; a varlena is aligned only if the next byte is a pad byte, att_align_pointer) */
static inline int tuple_attribute(const unsigned char *tup, uint32_t len, const int16_t *attlen, const char *attalign, uint32_t target,
                                  const unsigned char **att, const unsigned char **end, std::string &err) {
    const uint16_t infomask2 = rd16(tup + 18), infomask = rd16(tup + 20);
    const uint32_t natts = infomask2 & 0x07FFu, hoff = tup[22];
    if (hoff < 23 || hoff > len || (hoff & 7)) {
        err = "heap tuple: bad t_hoff";
        return DANN_ERR_FORMAT;
    }
    if (target >= natts) { /* added after the row was written: reads as NULL / missing */
        *att = nullptr;
        return DANN_OK;
    }
    const bool hasnull = infomask & 0x0001u;
    const unsigned char *bits = tup + 23;
    if (hasnull && 23 + (natts + 7) / 8 > hoff) {
        err = "heap tuple: null bitmap does not fit t_hoff";
        return DANN_ERR_FORMAT;
    }
    uint32_t off = hoff;
    for (uint32_t a = 0; a <= target; a++) {
        const bool isnull = hasnull && !((bits[a >> 3] >> (a & 7)) & 1u);
        if (isnull) {
            if (a == target) {
                *att = nullptr;
                return DANN_OK;
            }
            continue;
        }
        if (attlen[a] == -1) {
            if (off >= len) break;
            if (tup[off] == 0) off = align_up(off, attalign[a]); /* pad bytes: a 4-byte header follows, aligned */
            if (a == target) {
                *att = tup + off;
                *end = tup + len;
                return DANN_OK;
            }
            Varlena v;
            if (!decode_varlena(tup + off, tup + len, &v)) break;
            off += v.total;
        } else if (attlen[a] > 0) {
            off = align_up(off, attalign[a]);
            if (a == target) {
                *att = tup + off;
                *end = tup + len;
                return off + (uint32_t)attlen[a] <= len ? DANN_OK : DANN_ERR_FORMAT;
            }
            off += (uint32_t)attlen[a];
        } else {
            err = "heap tuple: an attribute in front of the vector column has a type this reader was not told how to skip (cstring)";
            return DANN_ERR_FORMAT;
        }
        if (off > len) break;
    }
    err = "heap tuple: attribute walk ran past the tuple";
    return DANN_ERR_FORMAT;
}

struct ToastChunk {
    uint32_t valueid, seq;
    const unsigned char *data;
    uint32_t len;
    bool operator<(const ToastChunk &o) const { return valueid != o.valueid ? valueid < o.valueid : seq < o.seq; }
};

/* every chunk of a TOAST relation: (chunk_id oid, chunk_seq int4, chunk_data bytea), heaptoast.h */
static inline int index_toast(const Relation &t, std::vector<ToastChunk> &out, std::string &err) {
    static const int16_t tl[3] = {4, 4, -1};
    static const char ta[3] = {'i', 'i', 'i'};
    for (uint32_t b = 0; b < t.nblocks; b++) {
        const HeapPage v = view_heap_page(t.page(b));
        if (!v.ok) continue;
        for (uint32_t off = 1; off <= v.nitems; off++) {
            const uint32_t lp = rd32(v.p + PAGE_HEADER + 4 * (off - 1));
            if (((lp >> 15) & 3u) != 1u) continue; /* only LP_NORMAL: nothing points into a TOAST table by TID */
            const unsigned char *tup, *att, *end;
            uint32_t len;
            if (!heap_item(v, off, &tup, &len)) continue;
            ToastChunk c;
            int rc = tuple_attribute(tup, len, tl, ta, 0, &att, &end, err);
            if (rc != DANN_OK || !att) return rc != DANN_OK ? rc : DANN_ERR_FORMAT;
            c.valueid = rd32(att);
            rc = tuple_attribute(tup, len, tl, ta, 1, &att, &end, err);
            if (rc != DANN_OK || !att) return rc != DANN_OK ? rc : DANN_ERR_FORMAT;
            c.seq = rd32(att);
            rc = tuple_attribute(tup, len, tl, ta, 2, &att, &end, err);
            if (rc != DANN_OK || !att) return rc != DANN_OK ? rc : DANN_ERR_FORMAT;
            Varlena d;
            if (!decode_varlena(att, end, &d) || d.external || d.compressed) {
                err = "TOAST chunk data is not a plain inline bytea";
                return DANN_ERR_FORMAT;
            }
            c.data = d.data;
            c.len = d.len;
            out.push_back(c);
        }
    }
    std::sort(out.begin(), out.end());
    return DANN_OK;
}

/* vectors of the rows heap_tid[0..n) -> out[n * dim]; rows that are gone (offset 0, dead line pointer, NULL value)
 * are zero-filled and counted in *missing */
static inline int fetch_vectors(const Relation &heap, const Relation *toast, const dann_pg_heap_layout *lay, const uint64_t *heap_tid,
                                uint32_t n, float *out, uint32_t *missing, std::string &err) {
    std::vector<ToastChunk> chunks;
    bool have_toast_index = false;
    const uint32_t dim = lay->dim, target = lay->natts_before;
    std::vector<int16_t> al(lay->attlen, lay->attlen + target);
    std::vector<char> aa(lay->attalign, lay->attalign + target);
    al.push_back(-1); /* the vector column itself: a varlena; CREATE TYPE vector names no ALIGNMENT, i.e. int4 */
    aa.push_back(lay->vector_align ? lay->vector_align : 'i');
    *missing = 0;
    std::vector<unsigned char> tmp;
    for (uint32_t i = 0; i < n; i++) {
        float *dst = out + (size_t)i * dim;
        const uint32_t block = (uint32_t)(heap_tid[i] >> 16), off = (uint32_t)(heap_tid[i] & 0xFFFFu);
        const unsigned char *pg = off ? heap.page(block) : nullptr;
        const unsigned char *tup = nullptr, *att = nullptr, *end = nullptr;
        uint32_t len = 0;
        HeapPage v;
        if (pg) v = view_heap_page(pg);
        if (!pg || !v.ok || !heap_item(v, off, &tup, &len)) {
            memset(dst, 0, (size_t)dim * 4);
            (*missing)++;
            continue;
        }
        int rc = tuple_attribute(tup, len, al.data(), aa.data(), target, &att, &end, err);
        if (rc != DANN_OK) return rc;
        if (!att) {
            memset(dst, 0, (size_t)dim * 4);
            (*missing)++;
            continue;
        }
        Varlena d;
        if (!decode_varlena(att, end, &d)) {
            err = "heap tuple (" + std::to_string(block) + "," + std::to_string(off) + "): the vector column is not a varlena";
            return DANN_ERR_FORMAT;
        }
        const unsigned char *val = d.data;
        uint32_t vlen = d.len;
        if (d.external) {
            if (d.compressed) {
                err = "a vector value is stored compressed (the column's STORAGE is not external / plain)";
                return DANN_ERR_FORMAT;
            }
            if (!toast) {
                err = "a vector value is stored out of line and no TOAST relation was given";
                return DANN_ERR_INVALID_ARG;
            }
            if (!have_toast_index) {
                rc = index_toast(*toast, chunks, err);
                if (rc != DANN_OK) return rc;
                have_toast_index = true;
            }
            ToastChunk key{d.ext_valueid, 0, nullptr, 0};
            auto it = std::lower_bound(chunks.begin(), chunks.end(), key);
            tmp.clear();
            uint32_t want_seq = 0;
            for (; it != chunks.end() && it->valueid == d.ext_valueid; ++it) {
                if (it->seq != want_seq) break; /* a missing or repeated chunk */
                tmp.insert(tmp.end(), it->data, it->data + it->len);
                want_seq++;
            }
            if (tmp.size() != d.ext_size) {
                err = "TOAST value " + std::to_string(d.ext_valueid) + ": chunks add up to " + std::to_string(tmp.size()) + " bytes, the pointer says " +
                      std::to_string(d.ext_size);
                return DANN_ERR_FORMAT;
            }
            val = tmp.data();
            vlen = (uint32_t)tmp.size();
        } else if (d.compressed) {
            err = "a vector value is stored compressed inline";
            return DANN_ERR_FORMAT;
        }
        /* pgvector Vector behind the varlena header: int16 dim, int16 unused, float x[dim] */
        if (vlen != 4u + dim * 4u || rd16(val) != dim) {
            err = "heap tuple (" + std::to_string(block) + "," + std::to_string(off) + "): not a vector(" + std::to_string(dim) + ") value";
            return DANN_ERR_FORMAT;
        }
        memcpy(dst, val + 4, (size_t)dim * 4);
    }
    return DANN_OK;
}

} // namespace dannpg
