// Loop-closure keyframe-descriptor database resident in HBM.
//
// Replaces the descriptor side of LoopClosureDetector in the reference
// (src/legacy/LoopClosure.cpp): addKeyFrame's FIFO store (:24-31) and findCandidates'
// per-keyframe knnMatch + ratio count (:86-95).  Layout: one [rows, 32] u8 array, keyframes are
// contiguous row segments in insertion order; eviction advances the head, and the live region
// is slid back to row 0 only when the tail would run past the capacity.
#include <cstdio>
#include <vector>

#include "match.cuh"

using namespace aslam;

struct aslam_loopdb {
    aslam_ctx *ctx = nullptr;
    uint8_t *d_rows = nullptr;
    int64_t capacity = 0;
    int max_keyframes = 0;
    int64_t head = 0, tail = 0;          // live rows are [head, tail)
    int64_t row_base = 0;                // global row of local row 0 of the live region
    struct Seg { uint64_t id; int64_t start; int32_t n; };
    std::vector<Seg> segs;               // oldest first; start is absolute in d_rows
    // query scratch
    uint8_t *d_q = nullptr; int q_cap = 0;
    KnnProblem *d_probs = nullptr; int prob_cap = 0;
    KnnProblem *h_probs = nullptr;       // pinned
    uint2 *d_partial = nullptr; size_t partial_cap = 0;
    int32_t *d_good = nullptr; int32_t *h_good = nullptr; int good_cap = 0;
    uint64_t *d_top2 = nullptr; int top2_cap = 0;
    unsigned int *d_arrival = nullptr;   // CTA arrival counter of the fused exchange kernel
    // flat-query fast path: the row-range table of the last query (rebuilt only when the query buffer, the
    // live region or the cut changes) and whether the merged result slots were left holding KEY_EMPTY
    const uint8_t *tab_q = nullptr; int tab_nq = 0, tab_splits = 0; int64_t tab_head = -1, tab_rows = -1;
    bool armed = false; int armed_stride = 0;
};

static int db_reserve(aslam_loopdb *db, int nq, int nprobs, size_t partial_elems)
{
    aslam_ctx *ctx = db->ctx;
    if (nq > db->q_cap) {
        if (db->d_q) { ASLAM_CUDA(cudaStreamSynchronize(ctx->stream)); ASLAM_CUDA(cudaFree(db->d_q)); db->d_q = nullptr; }
        if (db->d_top2) { ASLAM_CUDA(cudaFree(db->d_top2)); db->d_top2 = nullptr; }
        int cap = (int)align_up((size_t)nq, 1024);
        ASLAM_CUDA(cudaMalloc(&db->d_q, (size_t)cap * 32));
        ASLAM_CUDA(cudaMalloc(&db->d_top2, sizeof(uint64_t) * 2 * (size_t)cap));
        db->q_cap = cap;
    }
    if (nprobs > db->prob_cap) {
        ASLAM_CUDA(cudaStreamSynchronize(ctx->stream));
        if (db->d_probs) ASLAM_CUDA(cudaFree(db->d_probs));
        if (db->h_probs) ASLAM_CUDA(cudaFreeHost(db->h_probs));
        if (db->d_good) ASLAM_CUDA(cudaFree(db->d_good));
        if (db->h_good) ASLAM_CUDA(cudaFreeHost(db->h_good));
        db->d_probs = nullptr; db->h_probs = nullptr; db->d_good = nullptr; db->h_good = nullptr;
        int cap = (int)align_up((size_t)nprobs, 512);
        ASLAM_CUDA(cudaMalloc(&db->d_probs, sizeof(KnnProblem) * (size_t)cap));
        ASLAM_CUDA(cudaMallocHost(&db->h_probs, sizeof(KnnProblem) * (size_t)cap));
        ASLAM_CUDA(cudaMalloc(&db->d_good, sizeof(int32_t) * (size_t)cap));
        ASLAM_CUDA(cudaMallocHost(&db->h_good, sizeof(int32_t) * (size_t)cap));
        db->prob_cap = cap; db->good_cap = cap;
    }
    if (partial_elems > db->partial_cap) {
        ASLAM_CUDA(cudaStreamSynchronize(ctx->stream));
        if (db->d_partial) ASLAM_CUDA(cudaFree(db->d_partial));
        db->d_partial = nullptr;
        size_t cap = align_up(partial_elems + partial_elems / 4, 4096);
        ASLAM_CUDA(cudaMalloc(&db->d_partial, sizeof(uint2) * cap));
        db->partial_cap = cap;
    }
    return ASLAM_OK;
}

extern "C" {

int aslam_loopdb_create(aslam_ctx *ctx, int64_t capacity_rows, int max_keyframes, aslam_loopdb **out)
{
    ASLAM_REQUIRE(ctx && out, "NULL argument");
    ASLAM_REQUIRE(capacity_rows > 0 && max_keyframes > 0, "capacity and max_keyframes must be positive");
    ASLAM_CUDA(cudaSetDevice(ctx->device));
    aslam_loopdb *db = new aslam_loopdb();
    db->ctx = ctx; db->capacity = capacity_rows; db->max_keyframes = max_keyframes;
    cudaError_t e = cudaMalloc(&db->d_rows, (size_t)capacity_rows * 32);
    if (e != cudaSuccess) { delete db; return fail(ASLAM_E_NOMEM, "cudaMalloc of %lld rows -> %s", (long long)capacity_rows, cudaGetErrorString(e)); }
    *out = db;
    return ASLAM_OK;
}

void aslam_loopdb_destroy(aslam_loopdb *db)
{
    if (!db) return;
    cudaSetDevice(db->ctx->device);
    cudaStreamSynchronize(db->ctx->stream);
    cudaFree(db->d_rows); cudaFree(db->d_q); cudaFree(db->d_probs); cudaFree(db->d_partial);
    cudaFree(db->d_good); cudaFree(db->d_top2); cudaFree(db->d_arrival);
    if (db->h_probs) cudaFreeHost(db->h_probs);
    if (db->h_good) cudaFreeHost(db->h_good);
    delete db;
}

static int db_append(aslam_loopdb *db, const uint8_t *desc, int n, uint64_t id, cudaMemcpyKind kind)
{
    ASLAM_REQUIRE(db, "db is NULL");
    ASLAM_REQUIRE(n >= 0 && (n == 0 || desc), "bad descriptor block");
    ASLAM_REQUIRE(n <= KNN_MAX_ROWS, "keyframe has too many rows");
    aslam_ctx *ctx = db->ctx;
    ASLAM_CUDA(cudaSetDevice(ctx->device));
    // LoopClosure.cpp:24-31: push_back, then pop_front while size > cap.
    while ((int)db->segs.size() + 1 > db->max_keyframes) {
        db->head = db->segs.front().start + db->segs.front().n;
        db->row_base += db->segs.front().n;
        db->segs.erase(db->segs.begin());
    }
    if (db->segs.empty()) { db->row_base += db->tail - db->head; db->head = db->tail = 0; }
    if (db->tail + n > db->capacity) {
        const int64_t live = db->tail - db->head;
        if (live + n > db->capacity) return fail(ASLAM_E_NOMEM, "loop database full: %lld live rows + %d > capacity %lld",
                                                 (long long)live, n, (long long)db->capacity);
        // slide the live region to row 0 (regions may overlap: copy in stream order, chunk by chunk)
        const int64_t shift = db->head;
        int64_t done = 0;
        while (done < live) {
            int64_t chunk = live - done < shift ? live - done : shift;
            ASLAM_CUDA(cudaMemcpyAsync(db->d_rows + done * 32, db->d_rows + (shift + done) * 32, (size_t)chunk * 32,
                                       cudaMemcpyDeviceToDevice, ctx->stream));
            done += chunk;
        }
        for (auto &s : db->segs) s.start -= shift;
        db->head = 0; db->tail = live;
    }
    if (n) ASLAM_CUDA(cudaMemcpyAsync(db->d_rows + db->tail * 32, desc, (size_t)n * 32, kind, ctx->stream));
    db->segs.push_back({id, db->tail, n});
    db->tail += n;
    if (kind == cudaMemcpyHostToDevice) ASLAM_CUDA(cudaStreamSynchronize(ctx->stream));   // caller may reuse desc
    return ASLAM_OK;
}

int aslam_loopdb_add(aslam_loopdb *db, const uint8_t *desc, int n, uint64_t keyframe_id)
{
    return db_append(db, desc, n, keyframe_id, cudaMemcpyHostToDevice);
}

int aslam_loopdb_add_device(aslam_loopdb *db, const uint8_t *d_desc, int n, uint64_t keyframe_id)
{
    return db_append(db, d_desc, n, keyframe_id, cudaMemcpyDeviceToDevice);
}

int aslam_loopdb_clear(aslam_loopdb *db)
{
    ASLAM_REQUIRE(db, "db is NULL");
    db->segs.clear(); db->head = db->tail = 0; db->row_base = 0;
    return ASLAM_OK;
}

// Persistence (the keyframe store survives a restart / moves between shards): one file =
// header {magic, version, keyframes, rows, row_base}, the (id, rows) table, then the live rows.
static const uint64_t k_db_magic = 0x31424450554F4C41ull;      // "ALOUPDB1"

int aslam_loopdb_save(aslam_loopdb *db, const char *path)
{
    ASLAM_REQUIRE(db && path, "NULL argument");
    aslam_ctx *ctx = db->ctx;
    ASLAM_CUDA(cudaSetDevice(ctx->device));
    const int64_t rows = db->tail - db->head;
    std::vector<uint8_t> host((size_t)rows * 32);
    if (rows) ASLAM_CUDA(cudaMemcpyAsync(host.data(), db->d_rows + db->head * 32, host.size(), cudaMemcpyDeviceToHost, ctx->stream));
    ASLAM_CUDA(cudaStreamSynchronize(ctx->stream));
    FILE *f = fopen(path, "wb");
    if (!f) return fail(ASLAM_E_INVALID, "cannot open %s for writing", path);
    const uint64_t hdr[5] = {k_db_magic, 1, (uint64_t)db->segs.size(), (uint64_t)rows, (uint64_t)db->row_base};
    bool ok = fwrite(hdr, sizeof(hdr), 1, f) == 1;
    for (const auto &sg : db->segs) {
        const uint64_t e[2] = {sg.id, (uint64_t)sg.n};
        ok = ok && fwrite(e, sizeof(e), 1, f) == 1;
    }
    ok = ok && (host.empty() || fwrite(host.data(), 1, host.size(), f) == host.size());
    ok = (fclose(f) == 0) && ok;
    return ok ? ASLAM_OK : fail(ASLAM_E_INVALID, "short write to %s", path);
}

int aslam_loopdb_load(aslam_loopdb *db, const char *path)
{
    ASLAM_REQUIRE(db && path, "NULL argument");
    FILE *f = fopen(path, "rb");
    if (!f) return fail(ASLAM_E_INVALID, "cannot open %s", path);
    uint64_t hdr[5];
    if (fread(hdr, sizeof(hdr), 1, f) != 1 || hdr[0] != k_db_magic || hdr[1] != 1) { fclose(f); return fail(ASLAM_E_INVALID, "%s is not a loop database file", path); }
    const uint64_t nkf = hdr[2], rows = hdr[3];
    if ((int64_t)rows > db->capacity || (int64_t)nkf > db->max_keyframes) {
        fclose(f);
        return fail(ASLAM_E_NOMEM, "%s holds %llu keyframes / %llu rows; this database takes %d / %lld",
                    path, (unsigned long long)nkf, (unsigned long long)rows, db->max_keyframes, (long long)db->capacity);
    }
    std::vector<uint64_t> table(2 * (size_t)nkf);
    std::vector<uint8_t> host((size_t)rows * 32);
    bool ok = (nkf == 0 || fread(table.data(), 16, (size_t)nkf, f) == nkf) && (rows == 0 || fread(host.data(), 32, (size_t)rows, f) == rows);
    fclose(f);
    uint64_t sum = 0;
    for (uint64_t i = 0; ok && i < nkf; ++i) sum += table[2 * i + 1];
    if (!ok || sum != rows) return fail(ASLAM_E_INVALID, "%s is truncated or inconsistent", path);
    aslam_ctx *ctx = db->ctx;
    ASLAM_CUDA(cudaSetDevice(ctx->device));
    if (rows) ASLAM_CUDA(cudaMemcpyAsync(db->d_rows, host.data(), host.size(), cudaMemcpyHostToDevice, ctx->stream));
    ASLAM_CUDA(cudaStreamSynchronize(ctx->stream));
    db->segs.clear();
    int64_t at = 0;
    for (uint64_t i = 0; i < nkf; ++i) { db->segs.push_back({table[2 * i], at, (int32_t)table[2 * i + 1]}); at += (int64_t)table[2 * i + 1]; }
    db->head = 0; db->tail = (int64_t)rows; db->row_base = (int64_t)hdr[4];
    db->tab_splits = 0;
    return ASLAM_OK;
}

int aslam_loopdb_num_keyframes(const aslam_loopdb *db) { return db ? (int)db->segs.size() : 0; }
int64_t aslam_loopdb_num_rows(const aslam_loopdb *db) { return db ? db->tail - db->head : 0; }

int aslam_loopdb_set_row_base(aslam_loopdb *db, int64_t row_base)
{
    ASLAM_REQUIRE(db, "db is NULL");
    ASLAM_REQUIRE(row_base >= 0 && row_base < (1LL << 39), "row_base out of range");
    db->row_base = row_base;
    return ASLAM_OK;
}

int aslam_loopdb_keyframes(const aslam_loopdb *db, uint64_t *ids, int32_t *rows)
{
    ASLAM_REQUIRE(db, "db is NULL");
    for (size_t i = 0; i < db->segs.size(); ++i) {
        if (ids) ids[i] = db->segs[i].id;
        if (rows) rows[i] = db->segs[i].n;
    }
    return ASLAM_OK;
}

int aslam_loopdb_query_segmented(aslam_loopdb *db, const uint8_t *q, int nq, double ratio, int32_t *good)
{
    ASLAM_REQUIRE(db && (good || db->segs.empty()), "NULL argument");
    ASLAM_REQUIRE(nq >= 0, "negative nq");
    const int nseg = (int)db->segs.size();
    for (int s = 0; s < nseg; ++s) good[s] = 0;
    if (nq == 0 || nseg == 0) return ASLAM_OK;
    ASLAM_REQUIRE(q, "q is NULL");
    aslam_ctx *ctx = db->ctx;
    ASLAM_CUDA(cudaSetDevice(ctx->device));
    const int q_stride = (int)align_up((size_t)nq, 64);
    int rc = db_reserve(db, nq, nseg, (size_t)nseg * q_stride);
    if (rc) return rc;
    ASLAM_CUDA(cudaMemcpyAsync(db->d_q, q, (size_t)nq * 32, cudaMemcpyHostToDevice, ctx->stream));
    for (int s = 0; s < nseg; ++s) {
        KnnProblem &P = db->h_probs[s];
        P.q = db->d_q; P.t = db->d_rows + db->segs[s].start * 32;
        P.nq_dev = nullptr; P.nt_dev = nullptr; P.nq = nq; P.nt = db->segs[s].n; P.t_lo = 0; P.t_hi = db->segs[s].n;
    }
    db->tab_splits = 0; db->armed = false;      // the flat query's table and preset slots are overwritten here
    ASLAM_CUDA(cudaMemcpyAsync(db->d_probs, db->h_probs, sizeof(KnnProblem) * (size_t)nseg, cudaMemcpyHostToDevice, ctx->stream));
    rc = launch_knn2_partial(ctx, db->d_probs, nseg, nq, db->d_partial, q_stride);
    if (rc) return rc;
    RatioTest rt; rt.ratio = ratio; rt.as_double = 1; rt.disabled = ratio < 0;
    rc = launch_knn2_finish_counts(ctx, db->d_probs, db->d_partial, nseg, q_stride, rt, db->d_good);
    if (rc) return rc;
    ASLAM_CUDA(cudaMemcpyAsync(db->h_good, db->d_good, sizeof(int32_t) * (size_t)nseg, cudaMemcpyDeviceToHost, ctx->stream));
    ASLAM_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(good, db->h_good, sizeof(int32_t) * (size_t)nseg);
    return ASLAM_OK;
}

// Shared body of the flat query: scan this shard, then either write d_top2 or push to the peers.
static int db_query_global(aslam_loopdb *db, const uint8_t *d_q, int nq, uint64_t *d_top2,
                           uint64_t *const *d_peer_bufs, int nparts, int my_part, uint64_t round = 0)
{
    aslam_ctx *ctx = db->ctx;
    ASLAM_CUDA(cudaSetDevice(ctx->device));
    const int64_t rows = db->tail - db->head;
    if (rows > KNN_MAX_ROWS) return fail(ASLAM_E_INVALID, "shard holds %lld rows; at most %d per shard", (long long)rows, KNN_MAX_ROWS);
    const int q_stride = (int)align_up((size_t)nq, 64);
    int splits = 0;
    bool rearm = false;
    if (rows > 0) {
        splits = choose_splits(ctx, nq, rows, 1);
        const KnnProblem *probs_before = db->d_probs;
        const uint2 *partial_before = db->d_partial;
        int rc = db_reserve(db, 0, splits, (size_t)splits * q_stride);
        if (rc) return rc;
        if (db->d_probs != probs_before) db->tab_splits = 0;
        if (db->d_partial != partial_before) db->armed = false;
        if (db->tab_q != d_q || db->tab_nq != nq || db->tab_splits != splits || db->tab_head != db->head || db->tab_rows != rows) {
            rc = launch_fill_split_problems(ctx, db->d_probs, d_q, db->d_rows + db->head * 32, nullptr, nullptr, nq, (int)rows, splits);
            if (rc) return rc;
            db->tab_q = d_q; db->tab_nq = nq; db->tab_splits = splits; db->tab_head = db->head; db->tab_rows = rows;
        }
        const bool armed = db->armed && db->armed_stride == q_stride;
        db->armed = false;
        rc = launch_knn2_partial(ctx, db->d_probs, splits, nq, db->d_partial, q_stride, splits, armed);
        if (rc) return rc;
        rearm = splits > 1;                     // a single range writes its slot directly; nothing to preset
    } else {
        int rc = db_reserve(db, 0, 1, 1);       // an empty shard reports (empty, empty) per query
        if (rc) return rc;
    }
    const int nslots = rows > 0 ? 1 : 0;
    int rc;
    if (d_peer_bufs && round) {
        if (!db->d_arrival) {
            ASLAM_CUDA(cudaMalloc(&db->d_arrival, sizeof(unsigned int)));
            ASLAM_CUDA(cudaMemsetAsync(db->d_arrival, 0, sizeof(unsigned int), ctx->stream));
        }
        rc = launch_knn2_finish_exchange(ctx, db->d_partial, nslots, nq, q_stride, db->row_base, d_peer_bufs, nparts,
                                         my_part, round, db->d_arrival, d_top2, rearm);
    } else if (d_peer_bufs) {
        rc = launch_knn2_finish_global_push(ctx, db->d_partial, nslots, nq, q_stride, db->row_base, d_peer_bufs, nparts, my_part, rearm);
    } else {
        rc = launch_knn2_finish_global(ctx, db->d_partial, nslots, nq, q_stride, db->row_base, d_top2, rearm);
    }
    if (rc == ASLAM_OK && rearm) { db->armed = true; db->armed_stride = q_stride; }
    return rc;
}

int aslam_loopdb_query_global_device(aslam_loopdb *db, const uint8_t *d_q, int nq, uint64_t *d_top2)
{
    ASLAM_REQUIRE(db && (nq == 0 || (d_q && d_top2)), "NULL argument");
    ASLAM_REQUIRE(nq >= 0, "negative nq");
    if (nq == 0) return ASLAM_OK;
    return db_query_global(db, d_q, nq, d_top2, nullptr, 0, 0);
}

int aslam_loopdb_query_global_push_device(aslam_loopdb *db, const uint8_t *d_q, int nq,
                                          uint64_t *const *d_peer_bufs, int nparts, int my_part)
{
    ASLAM_REQUIRE(db && (nq == 0 || (d_q && d_peer_bufs)), "NULL argument");
    ASLAM_REQUIRE(nq >= 0 && nparts >= 1 && my_part >= 0 && my_part < nparts, "bad sizes");
    if (nq == 0) return ASLAM_OK;
    return db_query_global(db, d_q, nq, nullptr, d_peer_bufs, nparts, my_part);
}

int aslam_loopdb_query_global_exchange_device(aslam_loopdb *db, const uint8_t *d_q, int nq,
                                              uint64_t *const *d_peer_bufs, int nparts, int my_part,
                                              uint64_t round, uint64_t *d_out)
{
    ASLAM_REQUIRE(db && (nq == 0 || (d_q && d_peer_bufs && d_out)), "NULL argument");
    ASLAM_REQUIRE(nq >= 0 && nparts >= 1 && my_part >= 0 && my_part < nparts && round >= 1, "bad sizes or round");
    if (nq == 0) return ASLAM_OK;
    return db_query_global(db, d_q, nq, d_out, d_peer_bufs, nparts, my_part, round);
}

int aslam_loopdb_query_global(aslam_loopdb *db, const uint8_t *q, int nq, uint64_t *top2)
{
    ASLAM_REQUIRE(db && (nq == 0 || (q && top2)), "NULL argument");
    ASLAM_REQUIRE(nq >= 0, "negative nq");
    if (nq == 0) return ASLAM_OK;
    aslam_ctx *ctx = db->ctx;
    ASLAM_CUDA(cudaSetDevice(ctx->device));
    int rc = db_reserve(db, nq, 1, 1);
    if (rc) return rc;
    ASLAM_CUDA(cudaMemcpyAsync(db->d_q, q, (size_t)nq * 32, cudaMemcpyHostToDevice, ctx->stream));
    rc = aslam_loopdb_query_global_device(db, db->d_q, nq, db->d_top2);
    if (rc) return rc;
    ASLAM_CUDA(cudaMemcpyAsync(top2, db->d_top2, sizeof(uint64_t) * 2 * (size_t)nq, cudaMemcpyDeviceToHost, ctx->stream));
    ASLAM_CUDA(cudaStreamSynchronize(ctx->stream));
    return ASLAM_OK;
}

}  // extern "C"
