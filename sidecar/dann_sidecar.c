/* dann_sidecar.c — the process that owns the GPU on a Postgres host.
 *
 * Postgres backends are separate processes and cannot share a CUDA context; each would otherwise hold its own copy of
 * the index in HBM and run one lonely scan at a time.  The sidecar loads the index once (DANNSNP1 file written by the
 * Rust-side exporter), listens on a Unix-domain socket, and gives every connection (= one backend)
 * a thread that forwards its scans to the library's coalescer: whatever the backends ask for within one window runs as
 * one batch on the GPU.  Plain C99 + pthreads over include/diskann_b200.h.
 *
 *   dann_sidecar <snapshot.raw> <socket path> [max_batch=1024] [max_wait_us=200]
 *   dann_sidecar --relation <index file> <heap file> <toast file | -> <socket path> key=value...
 *       cold start from a checkpointed data directory (dann_pg_*): the index relation's pages, the
 *       table's vector column.  Keys = what MetaPage's getters and pg_attribute say: dim= dim_index= bits= R= distance=
 *       (0 cosine, 1 l2, 2 ip) start=<block>:<offset> means=<block>:<offset> [atts=8d,-1i (attlen+attalign of the columns
 *       in front of the vector column)] [max_batch=] [max_wait_us=].  SIGHUP re-reads the index relation's page headers
 *       and exits with status 5 when its fingerprint has moved (the supervisor restarts it: that is the reload).
 *
 * Wire protocol (little-endian, one request -> one reply, any number per connection):
 *   request : u32 magic 'DANQ', i32 k, i32 search_list_size, i32 rescore, i32 nlabels (-1 = no scan key),
 *             f32 query[dim], i16 labels[max(nlabels,0)]
 *   reply   : i32 status (dann_status), u32 count, then if status == 0: u64 tid[k], f32 dist[k],
 *             u32 stats[6] (visits, d_quantized, candidates, d_full, stream_len, status); else u32 len + message.
 * The first message on a connection is the server's hello: u32 magic 'DANH', u32 dim, u32 n.
 * A backend's amrescan sends one request with k = the rows it expects to need (a LIMIT hint or a chunk size) and
 * amgettuple serves rows from the reply; running past k re-requests with a larger k (scans are deterministic, so the
 * first k rows repeat). */
#define _POSIX_C_SOURCE 200809L
#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/select.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <time.h>
#include <sys/un.h>
#include <unistd.h>

#include "../harness/snapshot_raw.h"

static dann_coalescer *g_co;
static uint32_t g_dim, g_n;
static volatile sig_atomic_t g_stop;
/* open connections, so that shutdown can wake their threads and wait for them before the coalescer goes away */
enum { MAX_CONN = 4096 };
static pthread_mutex_t g_conn_mu = PTHREAD_MUTEX_INITIALIZER;
static int g_conn_fd[MAX_CONN];
static int g_conn_n;

static int conn_add(int fd) {
    int ok = 0;
    pthread_mutex_lock(&g_conn_mu);
    if (g_conn_n < MAX_CONN) {
        g_conn_fd[g_conn_n++] = fd;
        ok = 1;
    }
    pthread_mutex_unlock(&g_conn_mu);
    return ok;
}
static void conn_remove(int fd) {
    pthread_mutex_lock(&g_conn_mu);
    for (int i = 0; i < g_conn_n; i++)
        if (g_conn_fd[i] == fd) {
            g_conn_fd[i] = g_conn_fd[--g_conn_n];
            break;
        }
    pthread_mutex_unlock(&g_conn_mu);
}

static int read_full(int fd, void *buf, size_t n) {
    unsigned char *p = (unsigned char *)buf;
    while (n) {
        ssize_t r = read(fd, p, n);
        if (r == 0) return -1;
        if (r < 0) {
            if (errno == EINTR) continue;
            return -1;
        }
        p += r;
        n -= (size_t)r;
    }
    return 0;
}
static int write_full(int fd, const void *buf, size_t n) {
    const unsigned char *p = (const unsigned char *)buf;
    while (n) {
        ssize_t r = write(fd, p, n);
        if (r < 0) {
            if (errno == EINTR) continue;
            return -1;
        }
        p += r;
        n -= (size_t)r;
    }
    return 0;
}

/* SIGTERM / SIGINT / SIGHUP are for the main thread alone (it sits in accept() and acts on them): every other thread -
 * the library's dispatcher and CUDA's own, created while main() still has them blocked, and the connection threads -
 * keeps them blocked, so the kernel cannot hand one to a thread that would swallow it. */
static void block_control_signals(int how) {
    sigset_t set;
    sigemptyset(&set);
    sigaddset(&set, SIGTERM);
    sigaddset(&set, SIGINT);
    sigaddset(&set, SIGHUP);
    pthread_sigmask(how, &set, NULL);
}

static void *serve(void *arg) {
    const int fd = (int)(intptr_t)arg;
    block_control_signals(SIG_BLOCK);
    const uint32_t hello[3] = {0x484E4144u /* 'DANH' */, g_dim, g_n};
    float *query = (float *)malloc((size_t)g_dim * sizeof(float));
    if (!query || write_full(fd, hello, sizeof hello) != 0) goto out;
    for (;;) {
        int32_t h[5];
        if (read_full(fd, h, sizeof h) != 0) break;
        if ((uint32_t)h[0] != 0x514E4144u /* 'DANQ' */ || h[1] < 1 || h[1] > 65536 || h[4] < -1 || h[4] > 32767) {
            /* the stream cannot be resynchronised after a bad header: say why, then drop the connection */
            static const char msg[] = "dann_sidecar: malformed request header (magic, k in 1..65536, nlabels in -1..32767)";
            const int32_t head[2] = {DANN_ERR_INVALID_ARG, 0};
            const uint32_t len = (uint32_t)(sizeof msg - 1);
            if (write_full(fd, head, sizeof head) == 0 && write_full(fd, &len, 4) == 0) write_full(fd, msg, len);
            break;
        }
        const int k = h[1], nlabels = h[4];
        int16_t labels[64];
        int16_t *lab = labels, *big = NULL;
        if (read_full(fd, query, (size_t)g_dim * sizeof(float)) != 0) break;
        if (nlabels > 64) lab = big = (int16_t *)malloc((size_t)nlabels * 2);
        if (nlabels > 0 && (!lab || read_full(fd, lab, (size_t)nlabels * 2) != 0)) {
            free(big);
            break;
        }
        uint64_t *tid = (uint64_t *)malloc((size_t)k * 8);
        float *dist = (float *)malloc((size_t)k * 4);
        uint32_t count = 0;
        dann_query_stats st;
        memset(&st, 0, sizeof st);
        int32_t rc = tid && dist ? dann_coalescer_search(g_co, query, lab, nlabels, k, h[2], h[3], tid, dist, &count, &st)
                                 : DANN_ERR_OOM;
        int bad = 0;
        int32_t head[2] = {rc, (int32_t)count};
        bad |= write_full(fd, head, sizeof head);
        if (rc == DANN_OK) {
            bad |= write_full(fd, tid, (size_t)k * 8);
            bad |= write_full(fd, dist, (size_t)k * 4);
            bad |= write_full(fd, &st, sizeof st);
        } else {
            const char *msg = dann_last_error();
            uint32_t len = (uint32_t)strlen(msg);
            bad |= write_full(fd, &len, 4);
            bad |= write_full(fd, msg, len);
        }
        free(tid);
        free(dist);
        free(big);
        if (bad) break;
    }
out:
    free(query);
    conn_remove(fd);
    close(fd);
    return NULL;
}

static void on_term(int sig) {
    (void)sig;
    g_stop = 1;
}

static volatile sig_atomic_t g_hup;
static void on_hup(int sig) {
    (void)sig;
    g_hup = 1;
}

static const char *kv(int argc, char **argv, int from, const char *key) {
    const size_t n = strlen(key);
    for (int i = from; i < argc; i++)
        if (strncmp(argv[i], key, n) == 0 && argv[i][n] == '=') return argv[i] + n + 1;
    return NULL;
}

/* --relation: index pages + the table's vector column -> a loaded index; *fingerprint = the relation state it is valid for */
static int load_from_relation(int argc, char **argv, dann_index **ix, dann_pg_relation **index_rel, uint64_t *fingerprint) {
    const char *need[] = {"dim", "R", "start"};
    for (int i = 0; i < 3; i++)
        if (!kv(argc, argv, 6, need[i])) {
            fprintf(stderr, "dann_sidecar: --relation needs %s=\n", need[i]);
            return 2;
        }
    dann_pg_meta m;
    memset(&m, 0, sizeof m);
    m.num_dimensions = (uint32_t)atoi(kv(argc, argv, 6, "dim"));
    m.num_dimensions_to_index = kv(argc, argv, 6, "dim_index") ? (uint32_t)atoi(kv(argc, argv, 6, "dim_index")) : m.num_dimensions;
    m.bq_bits = kv(argc, argv, 6, "bits") ? (uint32_t)atoi(kv(argc, argv, 6, "bits")) : (m.num_dimensions_to_index < 900 ? 2u : 1u);
    m.num_neighbors = (uint32_t)atoi(kv(argc, argv, 6, "R"));
    m.distance_type = kv(argc, argv, 6, "distance") ? atoi(kv(argc, argv, 6, "distance")) : DANN_COSINE;
    unsigned sb = 0, so = 0, mb = DANN_INVALID_NODE, mo = 0;
    if (sscanf(kv(argc, argv, 6, "start"), "%u:%u", &sb, &so) != 2) return 2;
    if (kv(argc, argv, 6, "means") && sscanf(kv(argc, argv, 6, "means"), "%u:%u", &mb, &mo) != 2) return 2;
    m.start_block = sb;
    m.start_offset = (uint16_t)so;
    m.means_block = mb;
    m.means_offset = (uint16_t)mo;
    int16_t attlen[32];
    char attalign[32];
    uint32_t natts = 0;
    const char *atts = kv(argc, argv, 6, "atts");
    while (atts && *atts && natts < 32) { /* "8d,-1i" */
        char *end;
        const long l = strtol(atts, &end, 10);
        if (end == atts || !*end) return 2;
        attlen[natts] = (int16_t)l;
        attalign[natts++] = *end;
        atts = end + 1;
        if (*atts == ',') atts++;
    }
    dann_pg_relation *heap = NULL, *toast = NULL;
    dann_pg_snapshot *snap = NULL;
    float *vectors = NULL;
    int rc = dann_pg_relation_open(argv[2], index_rel);
    if (rc == DANN_OK) rc = dann_pg_relation_open(argv[3], &heap);
    if (rc == DANN_OK && strcmp(argv[4], "-") != 0) rc = dann_pg_relation_open(argv[4], &toast);
    if (rc == DANN_OK) rc = dann_pg_extract_sbq(*index_rel, &m, &snap);
    if (rc == DANN_OK) {
        const dann_pg_heap_layout lay = {natts, attlen, attalign, m.num_dimensions, 0};
        uint32_t missing = 0;
        vectors = (float *)malloc((size_t)snap->snap.n * m.num_dimensions * sizeof(float) + 4);
        rc = vectors ? dann_pg_heap_fetch_vectors(heap, toast, &lay, snap->snap.heap_tid, snap->snap.n, vectors, &missing) : DANN_ERR_OOM;
        if (rc == DANN_OK) {
            dann_snapshot_desc d = snap->snap;
            d.vectors = vectors;
            rc = dann_index_load(&d, 0, ix);
            g_dim = d.dim;
            g_n = d.n;
            *fingerprint = snap->fingerprint;
            fprintf(stderr, "dann_sidecar: %u nodes from %s (%u heap rows gone), relation fingerprint %016llx\n", d.n, argv[2], missing,
                    (unsigned long long)snap->fingerprint);
        }
    }
    if (rc != DANN_OK) fprintf(stderr, "dann_sidecar: %s\n", dann_last_error());
    free(vectors);
    dann_pg_snapshot_free(snap);
    dann_pg_relation_close(heap);
    dann_pg_relation_close(toast);
    return rc == DANN_OK ? 0 : (rc == DANN_ERR_NO_DEVICE ? 3 : 1);
}

int main(int argc, char **argv) {
    const int from_relation = argc > 1 && strcmp(argv[1], "--relation") == 0;
    if (argc < 3 || (from_relation && argc < 6)) {
        fprintf(stderr, "usage: %s snapshot.raw socket_path [max_batch] [max_wait_us]\n"
                        "       %s --relation index_file heap_file toast_file|- socket_path dim= R= start=B:O [means=B:O bits= distance= atts=]\n",
                argv[0], argv[0]);
        return 2;
    }
    block_control_signals(SIG_BLOCK); /* until the accept loop: threads created by the load inherit the mask */
    dann_index *ix = NULL;
    dann_pg_relation *index_rel = NULL;
    uint64_t fingerprint = 0;
    int max_batch = 1024, max_wait = 200;
    const char *index_path = from_relation ? argv[2] : NULL;
    if (from_relation) {
        const int lrc = load_from_relation(argc, argv, &ix, &index_rel, &fingerprint);
        if (lrc) return lrc;
        if (kv(argc, argv, 6, "max_batch")) max_batch = atoi(kv(argc, argv, 6, "max_batch"));
        if (kv(argc, argv, 6, "max_wait_us")) max_wait = atoi(kv(argc, argv, 6, "max_wait_us"));
        argv[2] = argv[5]; /* the socket path, where the code below expects it */
    } else {
        dann_snapshot_desc s;
        const float *iv = NULL;
        void *buf = dann_snapshot_raw_read(argv[1], &s, &iv);
        if (!buf) {
            fprintf(stderr, "dann_sidecar: cannot read %s\n", argv[1]);
            return 1;
        }
        int rc = iv ? dann_index_load_plain(&s, iv, 0, &ix) : dann_index_load(&s, 0, &ix);
        g_dim = s.dim;
        g_n = s.n;
        free(buf);
        if (rc != DANN_OK) {
            fprintf(stderr, "dann_sidecar: %s\n", dann_last_error());
            return rc == DANN_ERR_NO_DEVICE ? 3 : 1;
        }
        if (argc > 3) max_batch = atoi(argv[3]);
        if (argc > 4) max_wait = atoi(argv[4]);
    }
    if (dann_coalescer_create(ix, max_batch, max_wait, &g_co) != DANN_OK) {
        fprintf(stderr, "dann_sidecar: %s\n", dann_last_error());
        return 1;
    }
    struct sigaction sa;
    memset(&sa, 0, sizeof sa);
    sa.sa_handler = on_term;
    sigaction(SIGTERM, &sa, NULL);
    sigaction(SIGINT, &sa, NULL);
    sa.sa_handler = on_hup;
    sigaction(SIGHUP, &sa, NULL);
    signal(SIGPIPE, SIG_IGN);
    int ls = socket(AF_UNIX, SOCK_STREAM, 0);
    struct sockaddr_un addr;
    memset(&addr, 0, sizeof addr);
    addr.sun_family = AF_UNIX;
    strncpy(addr.sun_path, argv[2], sizeof addr.sun_path - 1);
    unlink(argv[2]);
    umask(0077); /* the socket belongs to the account the sidecar runs under (postgres): no other local user may connect */
    if (ls < 0 || bind(ls, (struct sockaddr *)&addr, sizeof addr) != 0 || chmod(argv[2], 0600) != 0 || listen(ls, 512) != 0) {
        perror("dann_sidecar: socket");
        return 1;
    }
    fprintf(stderr, "dann_sidecar: %u nodes x %u dims in HBM (%.2f GB), listening on %s\n", g_n, g_dim,
            (double)dann_index_hbm_bytes(ix) / 1e9, argv[2]);
    int stale = 0;
    /* the control signals stay blocked except inside pselect(): none can slip in between the flag test and the wait */
    sigset_t wait_mask;
    pthread_sigmask(SIG_BLOCK, NULL, &wait_mask);
    sigdelset(&wait_mask, SIGTERM);
    sigdelset(&wait_mask, SIGINT);
    sigdelset(&wait_mask, SIGHUP);
    fcntl(ls, F_SETFL, fcntl(ls, F_GETFL, 0) | O_NONBLOCK);
    while (!g_stop) {
        fd_set rf;
        FD_ZERO(&rf);
        FD_SET(ls, &rf);
        int fd = -1;
        if (pselect(ls + 1, &rf, NULL, NULL, NULL, &wait_mask) > 0) {
            fd = accept(ls, NULL, NULL);
            if (fd < 0 && (errno == EAGAIN || errno == EWOULDBLOCK || errno == ECONNABORTED)) errno = EINTR;
        } else {
            errno = EINTR; /* a signal (or a spurious wake-up): look at the flags */
        }
        if (fd < 0) {
            if (errno == EINTR) {
                if (g_hup && index_rel) { /* invalidation rule: is the snapshot still the relation's state? */
                    g_hup = 0;
                    dann_pg_relation *now = NULL;
                    dann_pg_relation_info info;
                    if (dann_pg_relation_open(index_path, &now) == DANN_OK &&
                        dann_pg_relation_stat(now, &info) == DANN_OK && info.fingerprint != fingerprint) {
                        fprintf(stderr, "dann_sidecar: the index relation changed (fingerprint %016llx -> %016llx): reload needed\n",
                                (unsigned long long)fingerprint, (unsigned long long)info.fingerprint);
                        stale = 1;
                        g_stop = 1;
                    }
                    dann_pg_relation_close(now);
                }
                continue;
            }
            break;
        }
        pthread_t th;
        if (!conn_add(fd)) {
            close(fd);
            continue;
        }
        if (pthread_create(&th, NULL, serve, (void *)(intptr_t)fd) == 0) pthread_detach(th);
        else {
            conn_remove(fd);
            close(fd);
        }
    }
    close(ls);
    unlink(argv[2]);
    /* wake every connection thread (its read() returns 0) and wait until the last one has left the coalescer: a request
     * already inside dann_coalescer_search completes (the dispatcher is still running), so this terminates.  The
     * coalescer and the index are destroyed only when no thread can touch them; if threads are still around after a
     * minute something is wedged and the process exits without tearing down under them. */
    pthread_mutex_lock(&g_conn_mu);
    for (int i = 0; i < g_conn_n; i++) shutdown(g_conn_fd[i], SHUT_RDWR);
    pthread_mutex_unlock(&g_conn_mu);
    int left = 0;
    for (int spin = 0; spin < 12000; spin++) {
        pthread_mutex_lock(&g_conn_mu);
        left = g_conn_n;
        pthread_mutex_unlock(&g_conn_mu);
        if (!left) break;
        struct timespec ts = {0, 5 * 1000 * 1000};
        nanosleep(&ts, NULL);
    }
    if (left) {
        fprintf(stderr, "dann_sidecar: %d connection threads still busy after 60 s, exiting without teardown\n", left);
        _exit(1);
    }
    uint64_t batches = 0, queries = 0, largest = 0;
    dann_coalescer_stats(g_co, &batches, &queries, &largest);
    fprintf(stderr, "dann_sidecar: %llu queries in %llu batches (largest %llu)\n", (unsigned long long)queries,
            (unsigned long long)batches, (unsigned long long)largest);
    dann_coalescer_destroy(g_co);
    dann_index_free(ix);
    dann_pg_relation_close(index_rel);
    return stale ? 5 : 0;
}
