// Loop-closure keyframe-descriptor database resident in HBM.
//
// Replaces the descriptor side of LoopClosureDetector in the reference
// (src/legacy/LoopClosure.cpp): addKeyFrame's FIFO store (:24-31), findCandidates'
// per-keyframe knnMatch + ratio count (:86-95) and verifyGeometry's re-match of a candidate (:116-130).
// Layout: one [rows, 32] u8 array, keyframes are contiguous row segments in insertion order; eviction
// advances the head, the live region is slid back to row 0 only when the tail would run past the
// capacity, and the array grows (doubling) when the live rows themselves no longer fit -- like the
// reference's std::deque, the store never refuses a keyframe.
//
// Two sharded forms:
//  * one process per GPU: a database is one shard; peers' exchange buffers are opened through CUDA IPC
//    (aslam_loopdb_shard_export / aslam_loopdb_shard_init) or handed in as pointers (torch symmetric memory);
//  * one process, several GPUs: aslam_loopdb_shard_group owns one shard per device with peer access enabled.
#include <cstdio>
#include <deque>
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
    struct Seg { uint64_t id; int64_t start; int32_t n; int64_t gbase; };
    std::vector<Seg> segs;               // oldest first; start is absolute in d_rows
    bool scattered = false;              // keyframes carry their own global row base (shard group, round-robin placement)
    // device copy of the keyframe table: lstart[k] = first row of keyframe k relative to head (nseg + 1 entries), gbase[k]
    int *d_lstart = nullptr; long long *d_gbase = nullptr; int map_cap = 0; bool map_dirty = true;
    // query scratch
    uint8_t *d_q = nullptr; int q_cap = 0;
    KnnProblem *d_probs = nullptr; int prob_cap = 0;
    uint2 *d_partial = nullptr; size_t partial_cap = 0;
    int32_t *d_good = nullptr; int32_t *h_good = nullptr; int good_cap = 0;
    uint64_t *d_top2 = nullptr; int top2_cap = 0;
    // fused exchange: CTA arrival counters per query block, status word, %globaltimer stamps
    unsigned int *d_arrive = nullptr; unsigned int *d_status = nullptr; unsigned long long *d_stamps = nullptr;
    // flat-query fast path: the row-range table of the last query (rebuilt only when the query buffer, the
    // live region or the cut changes) and whether the merged result slots were left holding KEY_EMPTY
    const uint8_t *tab_q = nullptr; int tab_nq = 0, tab_splits = 0; int64_t tab_head = -1, tab_rows = -1;
    bool armed = false; int armed_stride = 0;
    // one-process-per-GPU sharding set up by aslam_loopdb_shard_export / _init
    uint64_t *x_buf = nullptr;           // this rank's exchange buffer (cudaMalloc)
    int x_nparts = 0, x_rank = 0, x_max_nq = 0;
    std::vector<void *> x_opened;        // peers' buffers opened through IPC (closed on destroy)
    unsigned long long **d_x_peers = nullptr;      // device table of the nparts buffer addresses
    uint64_t x_round = 0;
    // aslam_loopdb_set_pipelining: consecutive fused queries may overlap (programmatic dependent launch)
    bool pipelining = false;
    bool last_was_fused = false;         // the last operation this database put on the stream was a fused query kernel
};

static void db_invalidate_tables(aslam_loopdb *db)
{
    db->tab_splits = 0;
    db->map_dirty = true;
    db->last_was_fused = false;
}

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
        db->tab_splits = 0;
    }
    if (nprobs > db->prob_cap) {
        ASLAM_CUDA(cudaStreamSynchronize(ctx->stream));
        if (db->d_probs) ASLAM_CUDA(cudaFree(db->d_probs));
        if (db->d_good) ASLAM_CUDA(cudaFree(db->d_good));
        if (db->h_good) ASLAM_CUDA(cudaFreeHost(db->h_good));
        db->d_probs = nullptr; db->d_good = nullptr; db->h_good = nullptr;
        int cap = (int)align_up((size_t)nprobs, 512);
        ASLAM_CUDA(cudaMalloc(&db->d_probs, sizeof(KnnProblem) * (size_t)cap));
        ASLAM_CUDA(cudaMalloc(&db->d_good, sizeof(int32_t) * (size_t)cap));
        ASLAM_CUDA(cudaMallocHost(&db->h_good, sizeof(int32_t) * (size_t)cap));
        db->prob_cap = cap; db->good_cap = cap;
        db->tab_splits = 0;
    }
    if (partial_elems > db->partial_cap) {
        ASLAM_CUDA(cudaStreamSynchronize(ctx->stream));
        if (db->d_partial) ASLAM_CUDA(cudaFree(db->d_partial));
        db->d_partial = nullptr;
        size_t cap = align_up(partial_elems + partial_elems / 4, 4096);
        ASLAM_CUDA(cudaMalloc(&db->d_partial, sizeof(uint2) * cap));
        db->partial_cap = cap;
        db->armed = false;
    }
    return ASLAM_OK;
}

// Device copy of the keyframe table (first local row and global row base per keyframe), refreshed when the store changed.
// The source vectors are pageable: cudaMemcpyAsync stages them before it returns.
static int db_sync_map(aslam_loopdb *db)
{
    if (!db->map_dirty) return ASLAM_OK;
    aslam_ctx *ctx = db->ctx;
    const int nseg = (int)db->segs.size();
    if (nseg + 1 > db->map_cap) {
        ASLAM_CUDA(cudaStreamSynchronize(ctx->stream));
        if (db->d_lstart) ASLAM_CUDA(cudaFree(db->d_lstart));
        if (db->d_gbase) ASLAM_CUDA(cudaFree(db->d_gbase));
        db->d_lstart = nullptr; db->d_gbase = nullptr;
        const int cap = (int)align_up((size_t)nseg + 1, 512);
        ASLAM_CUDA(cudaMalloc(&db->d_lstart, sizeof(int) * (size_t)cap));
        ASLAM_CUDA(cudaMalloc(&db->d_gbase, sizeof(long long) * (size_t)cap));
        db->map_cap = cap;
    }
    std::vector<int> lstart((size_t)nseg + 1);
    std::vector<long long> gbase((size_t)nseg + 1, 0);
    for (int k = 0; k < nseg; ++k) {
        lstart[(size_t)k] = (int)(db->segs[(size_t)k].start - db->head);
        gbase[(size_t)k] = db->scattered ? db->segs[(size_t)k].gbase : db->row_base + lstart[(size_t)k];
    }
    lstart[(size_t)nseg] = (int)(db->tail - db->head);
    ASLAM_CUDA(cudaMemcpyAsync(db->d_lstart, lstart.data(), sizeof(int) * lstart.size(), cudaMemcpyHostToDevice, ctx->stream));
    ASLAM_CUDA(cudaMemcpyAsync(db->d_gbase, gbase.data(), sizeof(long long) * gbase.size(), cudaMemcpyHostToDevice, ctx->stream));
    db->map_dirty = false;
    return ASLAM_OK;
}

static RowMap db_row_map(const aslam_loopdb *db)
{
    RowMap m;
    m.row_base = db->row_base;
    if (db->scattered) { m.gbase = db->d_gbase; m.lstart = db->d_lstart; m.nseg = (int)db->segs.size(); }
    return m;
}

// `splits` problems per stored keyframe: the query set against row range j of that keyframe's rows (the scan merges the
// ranges of a keyframe into one slot per query, so short keyframe lists still fill the GPU).
__global__ void fill_segment_problems_kernel(KnnProblem *probs, const uint8_t *q, const uint8_t *rows_at_head,
                                             const int *__restrict__ lstart, int nseg, int nq, int splits)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nseg * splits) return;
    const int s = i / splits, j = i - s * splits;
    KnnProblem P;
    P.q = q; P.t = rows_at_head + (size_t)lstart[s] * 32;
    P.nq_dev = nullptr; P.nt_dev = nullptr; P.nq = nq; P.nt = lstart[s + 1] - lstart[s];
    const int per = (P.nt + splits - 1) / splits;
    P.t_lo = min(j * per, P.nt); P.t_hi = min((j + 1) * per, P.nt);
    probs[i] = P;
}

static void db_evict_oldest(aslam_loopdb *db)
{
    const aslam_loopdb::Seg &s = db->segs.front();
    db->head = s.start + s.n;
    db->row_base += s.n;
    db->segs.erase(db->segs.begin());
    if (db->segs.empty()) { db->row_base += db->tail - db->head; db->head = db->tail = 0; }
    db_invalidate_tables(db);
}

// Append one keyframe.  Nothing is changed unless the append can complete: the eviction the FIFO cap asks for
// (LoopClosure.cpp:24-31: push_back, then pop_front while size > cap) is planned first, the array is slid or grown if
// the rows do not fit behind the tail, and only then are head/segs touched.
static int db_append(aslam_loopdb *db, const uint8_t *desc, int n, uint64_t id, cudaMemcpyKind kind, int64_t gbase = -1)
{
    ASLAM_REQUIRE(db, "db is NULL");
    ASLAM_REQUIRE(n >= 0 && (n == 0 || desc), "bad descriptor block");
    ASLAM_REQUIRE(n <= KNN_MAX_ROWS, "keyframe has too many rows");
    aslam_ctx *ctx = db->ctx;
    ASLAM_CUDA(cudaSetDevice(ctx->device));
    size_t evict = 0;
    int64_t new_head = db->head, evicted_rows = 0;
    while (db->segs.size() - evict + 1 > (size_t)db->max_keyframes) {
        new_head = db->segs[evict].start + db->segs[evict].n;
        evicted_rows += db->segs[evict].n;
        ++evict;
    }
    const int64_t live = evict == db->segs.size() ? 0 : db->tail - new_head;
    if (live + n > db->capacity) {
        // grow: a new array of at least twice the rows, the surviving rows copied to its start
        int64_t cap = db->capacity * 2;
        if (cap < live + n) cap = live + n;
        uint8_t *bigger = nullptr;
        cudaError_t e = cudaMalloc(&bigger, (size_t)cap * 32);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(ASLAM_E_NOMEM, "loop database cannot grow to %lld rows (%lld live + %d new): %s",
                        (long long)cap, (long long)live, n, cudaGetErrorString(e));
        }
        e = live ? cudaMemcpyAsync(bigger, db->d_rows + new_head * 32, (size_t)live * 32, cudaMemcpyDeviceToDevice, ctx->stream) : cudaSuccess;
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { cudaFree(bigger); return fail(ASLAM_E_CUDA, "loop database grow copy -> %s", cudaGetErrorString(e)); }
        cudaFree(db->d_rows);
        db->d_rows = bigger; db->capacity = cap;
        for (auto &s : db->segs) s.start -= new_head;       // (the evicted entries go away just below)
        db->tail = live; new_head = 0;
    }
    // commit the eviction
    db->row_base += evicted_rows;
    db->segs.erase(db->segs.begin(), db->segs.begin() + (long)evict);
    db->head = new_head;
    if (db->segs.empty()) { db->row_base += db->tail - db->head; db->head = db->tail = 0; }
    if (db->tail + n > db->capacity) {
        // slide the live region to row 0 (regions may overlap: copy in stream order, chunk by chunk)
        const int64_t shift = db->head, rows = db->tail - db->head;
        int64_t done = 0;
        while (done < rows) {
            int64_t chunk = rows - done < shift ? rows - done : shift;
            ASLAM_CUDA(cudaMemcpyAsync(db->d_rows + done * 32, db->d_rows + (shift + done) * 32, (size_t)chunk * 32,
                                       cudaMemcpyDeviceToDevice, ctx->stream));
            done += chunk;
        }
        for (auto &s : db->segs) s.start -= shift;
        db->head = 0; db->tail = rows;
    }
    // Host rows: pageable memory is staged by cudaMemcpyAsync before it returns; PINNED memory is read by the DMA
    // engine later, so the caller keeps it unchanged until the next call that synchronises (any query, aslam_ctx_sync).
    if (n) ASLAM_CUDA(cudaMemcpyAsync(db->d_rows + db->tail * 32, desc, (size_t)n * 32, kind, ctx->stream));
    db->segs.push_back({id, db->tail, n, gbase});
    db->tail += n;
    db_invalidate_tables(db);
    return ASLAM_OK;
}

static int db_query_segmented_device(aslam_loopdb *db, const uint8_t *d_q, int nq, double ratio, int32_t *d_good)
{
    aslam_ctx *ctx = db->ctx;
    const int nseg = (int)db->segs.size();
    const int q_stride = (int)align_up((size_t)nq, 64);
    // keyframes are cut into row ranges until the grid fills the GPU (8 GPUs x 63 keyframes x 4 query blocks would
    // otherwise leave almost half of the resident CTA slots empty)
    const int64_t rows = db->tail - db->head;
    const int splits = choose_splits(ctx, nq, rows / nseg > 0 ? rows / nseg : 1, nseg);
    int rc = db_reserve(db, 0, nseg * splits, (size_t)nseg * q_stride);
    if (rc) return rc;
    if ((rc = db_sync_map(db))) return rc;
    db->tab_splits = 0; db->armed = false;      // the flat query's table and preset slots are overwritten here
    db->last_was_fused = false;
    {
        ProfScope ps(ctx, PK_MISC);
        fill_segment_problems_kernel<<<(nseg * splits + 127) / 128, 128, 0, ctx->stream>>>(db->d_probs, d_q, db->d_rows + db->head * 32,
                                                                                            db->d_lstart, nseg, nq, splits);
        ASLAM_LAUNCH_CHECK();
    }
    rc = launch_knn2_partial(ctx, db->d_probs, nseg * splits, nq, db->d_partial, q_stride, splits);
    if (rc) return rc;
    RatioTest rt; rt.ratio = ratio; rt.as_double = 1; rt.disabled = ratio < 0;
    return launch_knn2_finish_counts(ctx, db->d_probs, db->d_partial, nseg, splits, q_stride, rt, d_good);
}

// Shared body of the flat query: scan this shard, then write d_top2, push to the peers, or run the fused exchange.
static int db_query_global(aslam_loopdb *db, const uint8_t *d_q, int nq, uint64_t *d_top2,
                           uint64_t *const *d_peer_bufs, int nparts, int my_part, uint64_t round = 0)
{
    aslam_ctx *ctx = db->ctx;
    ASLAM_CUDA(cudaSetDevice(ctx->device));
    const int64_t rows = db->tail - db->head;
    if (rows > KNN_MAX_ROWS) return fail(ASLAM_E_INVALID, "shard holds %lld rows; at most %d per shard", (long long)rows, KNN_MAX_ROWS);
    const int q_stride = (int)align_up((size_t)nq, 64);
    const bool fused = d_peer_bufs && round;
    int rc;
    if (db->scattered && (rc = db_sync_map(db))) return rc;
    int splits = 0;
    bool rearm = false;
    if (rows > 0 || fused) {
        splits = rows > 0 ? choose_splits(ctx, nq, rows, 1) : 1;      // an empty shard still takes part in the exchange
        if ((rc = db_reserve(db, 0, splits, (size_t)splits * q_stride))) return rc;
        bool table_rebuilt = false;
        if (db->tab_q != d_q || db->tab_nq != nq || db->tab_splits != splits || db->tab_head != db->head || db->tab_rows != rows) {
            rc = launch_fill_split_problems(ctx, db->d_probs, d_q, db->d_rows + db->head * 32, nullptr, nullptr, nq, (int)rows, splits);
            if (rc) return rc;
            table_rebuilt = true;
            db->tab_q = d_q; db->tab_nq = nq; db->tab_splits = splits; db->tab_head = db->head; db->tab_rows = rows;
        }
        const bool armed = db->armed && db->armed_stride == q_stride;
        db->armed = false;
        if (fused) {
            if (!db->d_arrive) {
                ASLAM_CUDA(cudaMalloc(&db->d_arrive, sizeof(unsigned int) * KNN_FUSED_MAX_QBLOCKS));
                ASLAM_CUDA(cudaMalloc(&db->d_status, sizeof(unsigned int)));
                ASLAM_CUDA(cudaMalloc(&db->d_stamps, sizeof(unsigned long long) * 4 * KNN_FUSED_MAX_QBLOCKS));
                ASLAM_CUDA(cudaMemsetAsync(db->d_arrive, 0, sizeof(unsigned int) * KNN_FUSED_MAX_QBLOCKS, ctx->stream));
                ASLAM_CUDA(cudaMemsetAsync(db->d_status, 0, sizeof(unsigned int), ctx->stream));
                ASLAM_CUDA(cudaMemsetAsync(db->d_stamps, 0, sizeof(unsigned long long) * 4 * KNN_FUSED_MAX_QBLOCKS, ctx->stream));
            }
            if (!armed) ASLAM_CUDA(cudaMemsetAsync(db->d_partial, 0xff, sizeof(uint2) * (size_t)q_stride, ctx->stream));
            ExchangeArgs X;
            X.peer_bufs = (unsigned long long *const *)d_peer_bufs; X.nparts = nparts; X.my_part = my_part; X.nq = nq;
            X.round = round; X.arrive = db->d_arrive; X.out = (unsigned long long *)d_top2; X.stamps = db->d_stamps;
            X.status = db->d_status; X.map = db_row_map(db);
            // pipelined with the previous query only if that was the last thing this database enqueued and the row-range
            // table, the slots and the counters are exactly as its kernel leaves them
            const bool overlap = db->pipelining && db->last_was_fused && armed && !table_rebuilt;
            rc = launch_knn2_scan_exchange(ctx, db->d_probs, splits, nq, db->d_partial, q_stride, X, overlap);
            if (rc == ASLAM_OK) { db->armed = true; db->armed_stride = q_stride; db->last_was_fused = true; }
            return rc;
        }
        db->last_was_fused = false;
        rc = launch_knn2_partial(ctx, db->d_probs, splits, nq, db->d_partial, q_stride, splits, armed);
        if (rc) return rc;
        rearm = splits > 1;                     // a single range writes its slot directly; nothing to preset
    } else {
        if ((rc = db_reserve(db, 0, 1, 1))) return rc;      // an empty shard reports (empty, empty) per query
    }
    const int nslots = rows > 0 ? 1 : 0;
    if (d_peer_bufs)
        rc = launch_knn2_finish_global_push(ctx, db->d_partial, nslots, nq, q_stride, db_row_map(db), d_peer_bufs, nparts, my_part, rearm);
    else
        rc = launch_knn2_finish_global(ctx, db->d_partial, nslots, nq, q_stride, db_row_map(db), d_top2, rearm);
    if (rc == ASLAM_OK && rearm) { db->armed = true; db->armed_stride = q_stride; }
    return rc;
}

// knn2 + ratio of device-resident queries against the resident rows of one keyframe, results to host buffers.
int knn2_resident_to_host(aslam_ctx *ctx, const uint8_t *d_q, int nq, const uint8_t *d_t, int nt, double ratio, int as_double,
                          aslam_match *out, int *nout);

extern "C" {

int aslam_loopdb_create(aslam_ctx *ctx, int64_t capacity_rows, int max_keyframes, aslam_loopdb **out)
{
    ASLAM_REQUIRE(ctx && out, "NULL argument");
    ASLAM_REQUIRE(capacity_rows > 0 && max_keyframes > 0, "capacity and max_keyframes must be positive");
    ASLAM_CUDA(cudaSetDevice(ctx->device));
    aslam_loopdb *db = new aslam_loopdb();
    db->ctx = ctx; db->capacity = capacity_rows; db->max_keyframes = max_keyframes;
    cudaError_t e = cudaMalloc(&db->d_rows, (size_t)capacity_rows * 32);
    if (e != cudaSuccess) { cudaGetLastError(); delete db; return fail(ASLAM_E_NOMEM, "cudaMalloc of %lld rows -> %s", (long long)capacity_rows, cudaGetErrorString(e)); }
    *out = db;
    return ASLAM_OK;
}

void aslam_loopdb_destroy(aslam_loopdb *db)
{
    if (!db) return;
    cudaSetDevice(db->ctx->device);
    cudaStreamSynchronize(db->ctx->stream);
    for (void *p : db->x_opened) cudaIpcCloseMemHandle(p);
    cudaFree(db->x_buf); cudaFree(db->d_x_peers);
    cudaFree(db->d_rows); cudaFree(db->d_q); cudaFree(db->d_probs); cudaFree(db->d_partial);
    cudaFree(db->d_good); cudaFree(db->d_top2); cudaFree(db->d_arrive); cudaFree(db->d_status); cudaFree(db->d_stamps);
    cudaFree(db->d_lstart); cudaFree(db->d_gbase);
    if (db->h_good) cudaFreeHost(db->h_good);
    delete db;
}

int aslam_loopdb_add(aslam_loopdb *db, const uint8_t *desc, int n, uint64_t keyframe_id)
{
    return db_append(db, desc, n, keyframe_id, cudaMemcpyHostToDevice);
}

int aslam_loopdb_add_device(aslam_loopdb *db, const uint8_t *d_desc, int n, uint64_t keyframe_id)
{
    return db_append(db, d_desc, n, keyframe_id, cudaMemcpyDeviceToDevice);
}

int aslam_loopdb_add_at(aslam_loopdb *db, const uint8_t *desc, int n, uint64_t keyframe_id, int64_t global_row_base)
{
    ASLAM_REQUIRE(db, "db is NULL");
    ASLAM_REQUIRE(global_row_base >= 0 && global_row_base < (1LL << 39), "global_row_base out of range");
    if (!db->scattered && !db->segs.empty())
        return fail(ASLAM_E_STATE, "aslam_loopdb_add_at on a database that already holds contiguous keyframes");
    if (!db->segs.empty() && global_row_base < db->segs.back().gbase + db->segs.back().n)
        return fail(ASLAM_E_INVALID, "keyframes of a shard must be added in increasing global row order");
    db->scattered = true;
    return db_append(db, desc, n, keyframe_id, cudaMemcpyHostToDevice, global_row_base);
}

int aslam_loopdb_clear(aslam_loopdb *db)
{
    ASLAM_REQUIRE(db, "db is NULL");
    db->segs.clear(); db->head = db->tail = 0; db->row_base = 0;
    db->scattered = false;
    db_invalidate_tables(db);
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
    if ((int64_t)nkf > db->max_keyframes) {
        fclose(f);
        return fail(ASLAM_E_NOMEM, "%s holds %llu keyframes; this database keeps at most %d",
                    path, (unsigned long long)nkf, db->max_keyframes);
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
    if ((int64_t)rows > db->capacity) {             // the file holds more rows than the current array: grow first
        uint8_t *bigger = nullptr;
        cudaError_t e = cudaMalloc(&bigger, (size_t)rows * 32);
        if (e != cudaSuccess) { cudaGetLastError(); return fail(ASLAM_E_NOMEM, "cudaMalloc of %llu rows -> %s", (unsigned long long)rows, cudaGetErrorString(e)); }
        ASLAM_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(db->d_rows);
        db->d_rows = bigger; db->capacity = (int64_t)rows;
    }
    if (rows) ASLAM_CUDA(cudaMemcpyAsync(db->d_rows, host.data(), host.size(), cudaMemcpyHostToDevice, ctx->stream));
    ASLAM_CUDA(cudaStreamSynchronize(ctx->stream));
    db->segs.clear();
    int64_t at = 0;
    for (uint64_t i = 0; i < nkf; ++i) { db->segs.push_back({table[2 * i], at, (int32_t)table[2 * i + 1], -1}); at += (int64_t)table[2 * i + 1]; }
    db->head = 0; db->tail = (int64_t)rows; db->row_base = (int64_t)hdr[4];
    db_invalidate_tables(db);
    return ASLAM_OK;
}

int aslam_loopdb_num_keyframes(const aslam_loopdb *db) { return db ? (int)db->segs.size() : 0; }
int64_t aslam_loopdb_num_rows(const aslam_loopdb *db) { return db ? db->tail - db->head : 0; }
int64_t aslam_loopdb_capacity_rows(const aslam_loopdb *db) { return db ? db->capacity : 0; }

int aslam_loopdb_set_row_base(aslam_loopdb *db, int64_t row_base)
{
    ASLAM_REQUIRE(db, "db is NULL");
    ASLAM_REQUIRE(row_base >= 0 && row_base < (1LL << 39), "row_base out of range");
    db->row_base = row_base;
    db->map_dirty = true;
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
    int rc = db_reserve(db, nq, nseg, 1);
    if (rc) return rc;
    ASLAM_CUDA(cudaMemcpyAsync(db->d_q, q, (size_t)nq * 32, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = db_query_segmented_device(db, db->d_q, nq, ratio, db->d_good))) return rc;
    ASLAM_CUDA(cudaMemcpyAsync(db->h_good, db->d_good, sizeof(int32_t) * (size_t)nseg, cudaMemcpyDeviceToHost, ctx->stream));
    ASLAM_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(good, db->h_good, sizeof(int32_t) * (size_t)nseg);
    return ASLAM_OK;
}

int aslam_loopdb_query_segmented_device(aslam_loopdb *db, const uint8_t *d_q, int nq, double ratio, int32_t *d_good)
{
    ASLAM_REQUIRE(db, "db is NULL");
    ASLAM_REQUIRE(nq >= 0, "negative nq");
    const int nseg = (int)db->segs.size();
    if (nseg == 0) return ASLAM_OK;
    ASLAM_REQUIRE(d_good && (nq == 0 || d_q), "NULL argument");
    ASLAM_CUDA(cudaSetDevice(db->ctx->device));
    if (nq == 0) {
        ASLAM_CUDA(cudaMemsetAsync(d_good, 0, sizeof(int32_t) * (size_t)nseg, db->ctx->stream));
        return ASLAM_OK;
    }
    return db_query_segmented_device(db, d_q, nq, ratio, d_good);
}

int aslam_loopdb_match_keyframe(aslam_loopdb *db, int kf_index, const uint8_t *q, int nq, double ratio,
                                aslam_match *out, int *nout)
{
    ASLAM_REQUIRE(db && nout, "NULL argument");
    ASLAM_REQUIRE(kf_index >= 0 && kf_index < (int)db->segs.size(), "keyframe index out of range");
    ASLAM_REQUIRE(nq >= 0, "negative nq");
    *nout = 0;
    const aslam_loopdb::Seg &sg = db->segs[(size_t)kf_index];
    if (nq == 0 || sg.n < 2) return ASLAM_OK;            // knn.size() >= 2 (LoopClosure.cpp:124)
    ASLAM_REQUIRE(q && out, "NULL descriptor or output pointer");
    aslam_ctx *ctx = db->ctx;
    ASLAM_CUDA(cudaSetDevice(ctx->device));
    int rc = db_reserve(db, nq, 1, 1);
    if (rc) return rc;
    db->last_was_fused = false;
    ASLAM_CUDA(cudaMemcpyAsync(db->d_q, q, (size_t)nq * 32, cudaMemcpyHostToDevice, ctx->stream));
    return knn2_resident_to_host(ctx, db->d_q, nq, db->d_rows + sg.start * 32, sg.n, ratio, 1, out, nout);
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

int aslam_loopdb_set_pipelining(aslam_loopdb *db, int on)
{
    ASLAM_REQUIRE(db, "db is NULL");
    db->pipelining = on != 0;
    db->last_was_fused = false;
    return ASLAM_OK;
}

int aslam_loopdb_exchange_status(aslam_loopdb *db, uint32_t *status, uint64_t *stamps, int nqblocks, int reset)
{
    ASLAM_REQUIRE(db && status, "NULL argument");
    ASLAM_REQUIRE(nqblocks >= 0 && nqblocks <= KNN_FUSED_MAX_QBLOCKS, "at most 64 query blocks");
    *status = 0;
    if (!db->d_status) return ASLAM_OK;
    ASLAM_CUDA(cudaSetDevice(db->ctx->device));
    ASLAM_CUDA(cudaStreamSynchronize(db->ctx->stream));
    ASLAM_CUDA(cudaMemcpy(status, db->d_status, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (stamps && nqblocks) ASLAM_CUDA(cudaMemcpy(stamps, db->d_stamps, sizeof(uint64_t) * 4 * (size_t)nqblocks, cudaMemcpyDeviceToHost));
    if (reset) ASLAM_CUDA(cudaMemset(db->d_stamps, 0, sizeof(uint64_t) * 4 * KNN_FUSED_MAX_QBLOCKS));
    return ASLAM_OK;
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

// ------------------------------------------------------------------ one process per GPU: peers through CUDA IPC
size_t aslam_loopdb_shard_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

int aslam_loopdb_shard_export(aslam_loopdb *db, int nranks, int max_nq, void *handle_out)
{
    ASLAM_REQUIRE(db && handle_out, "NULL argument");
    ASLAM_REQUIRE(nranks >= 1 && nranks <= 64 && max_nq >= 1 && max_nq <= KNN_FUSED_MAX_QBLOCKS * KNN_QBLOCK, "bad sizes");
    if (db->x_buf) return fail(ASLAM_E_STATE, "this shard already exported an exchange buffer");
    aslam_ctx *ctx = db->ctx;
    ASLAM_CUDA(cudaSetDevice(ctx->device));
    const size_t bytes = aslam_loop_exchange_buffer_bytes(nranks, max_nq);
    ASLAM_CUDA(cudaMalloc(&db->x_buf, bytes));
    ASLAM_CUDA(cudaMemset(db->x_buf, 0, bytes));            // every flag reads "round 0" before any peer can know the handle
    ASLAM_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    ASLAM_CUDA(cudaIpcGetMemHandle(&h, db->x_buf));
    memcpy(handle_out, &h, sizeof(h));
    db->x_nparts = nranks; db->x_max_nq = max_nq;
    return ASLAM_OK;
}

int aslam_loopdb_shard_init(aslam_loopdb *db, int rank, int nranks, const void *handles)
{
    ASLAM_REQUIRE(db && handles, "NULL argument");
    if (!db->x_buf || db->x_nparts != nranks) return fail(ASLAM_E_STATE, "call aslam_loopdb_shard_export with the same nranks first");
    ASLAM_REQUIRE(rank >= 0 && rank < nranks, "rank out of range");
    if (db->d_x_peers) return fail(ASLAM_E_STATE, "this shard is already initialised");
    aslam_ctx *ctx = db->ctx;
    ASLAM_CUDA(cudaSetDevice(ctx->device));
    std::vector<unsigned long long *> peers((size_t)nranks, nullptr);
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) { peers[(size_t)r] = (unsigned long long *)db->x_buf; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const uint8_t *)handles + (size_t)r * sizeof(h), sizeof(h));
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (void *o : db->x_opened) cudaIpcCloseMemHandle(o);
            db->x_opened.clear();
            return fail(ASLAM_E_CUDA, "cudaIpcOpenMemHandle of rank %d's exchange buffer -> %s", r, cudaGetErrorString(e));
        }
        db->x_opened.push_back(p);
        peers[(size_t)r] = (unsigned long long *)p;
    }
    ASLAM_CUDA(cudaMalloc(&db->d_x_peers, sizeof(unsigned long long *) * (size_t)nranks));
    ASLAM_CUDA(cudaMemcpy(db->d_x_peers, peers.data(), sizeof(unsigned long long *) * (size_t)nranks, cudaMemcpyHostToDevice));
    db->x_rank = rank; db->x_round = 0;
    return ASLAM_OK;
}

int aslam_loopdb_query_global_sharded_device(aslam_loopdb *db, const uint8_t *d_q, int nq, uint64_t *d_out)
{
    ASLAM_REQUIRE(db && (nq == 0 || (d_q && d_out)), "NULL argument");
    if (!db->d_x_peers) return fail(ASLAM_E_STATE, "aslam_loopdb_shard_init has not been called on this shard");
    ASLAM_REQUIRE(nq >= 0 && nq <= db->x_max_nq, "nq exceeds the max_nq given to aslam_loopdb_shard_export");
    if (nq == 0) return ASLAM_OK;
    // The buffer layout depends on nq: every rank must pass the same nq in the same round (the queries are replicated).
    return db_query_global(db, d_q, nq, d_out, (uint64_t *const *)db->d_x_peers, db->x_nparts, db->x_rank, ++db->x_round);
}

}  // extern "C"

// ------------------------------------------------------------------ one process, several GPUs: the shard group
// Keyframe k (in insertion order since creation) lives on shard k % n, so the shards stay balanced under the FIFO
// eviction of LoopClosure.cpp:27-30 (the globally oldest keyframe is always the oldest of its shard).  Global row =
// running row count since creation, as in the single database.  Every device can store into every other device's
// exchange buffer (cudaDeviceEnablePeerAccess).
struct aslam_loopdb_group {
    int n = 0;
    std::vector<aslam_ctx *> ctx;
    std::vector<aslam_loopdb *> shard;
    int max_keyframes = 0;
    struct GSeg { uint64_t id; uint64_t seq; int32_t rows; };
    std::deque<GSeg> live;                 // oldest first
    uint64_t next_seq = 0;
    int64_t next_row = 0;
    // query plumbing per shard
    std::vector<uint8_t *> d_q; int q_cap = 0;
    std::vector<uint64_t *> d_out;
    std::vector<int32_t *> d_good; int good_cap = 0;
    std::vector<uint64_t *> x_buf; std::vector<unsigned long long **> d_peers; int x_nq = 0;
    uint64_t round = 0;
    uint8_t *h_q = nullptr; int hq_cap = 0;          // pinned staging of the query block
    uint64_t *h_top2 = nullptr; int32_t *h_good = nullptr; int h_good_cap = 0;
    std::vector<cudaEvent_t> ev;
};

static int group_reserve(aslam_loopdb_group *g, int nq)
{
    if (nq > g->q_cap) {
        const int cap = (int)align_up((size_t)nq, 1024);
        for (int s = 0; s < g->n; ++s) {
            ASLAM_CUDA(cudaSetDevice(g->ctx[(size_t)s]->device));
            ASLAM_CUDA(cudaStreamSynchronize(g->ctx[(size_t)s]->stream));
            if (g->d_q[(size_t)s]) ASLAM_CUDA(cudaFree(g->d_q[(size_t)s]));
            if (g->d_out[(size_t)s]) ASLAM_CUDA(cudaFree(g->d_out[(size_t)s]));
            g->d_q[(size_t)s] = nullptr; g->d_out[(size_t)s] = nullptr;
            ASLAM_CUDA(cudaMalloc(&g->d_q[(size_t)s], (size_t)cap * 32));
            ASLAM_CUDA(cudaMalloc(&g->d_out[(size_t)s], sizeof(uint64_t) * 2 * (size_t)cap));
        }
        if (g->h_q) ASLAM_CUDA(cudaFreeHost(g->h_q));
        if (g->h_top2) ASLAM_CUDA(cudaFreeHost(g->h_top2));
        g->h_q = nullptr; g->h_top2 = nullptr;
        ASLAM_CUDA(cudaMallocHost(&g->h_q, (size_t)cap * 32));
        ASLAM_CUDA(cudaMallocHost(&g->h_top2, sizeof(uint64_t) * 2 * (size_t)cap));
        g->q_cap = cap; g->hq_cap = cap;
    }
    return ASLAM_OK;
}

// The query block goes to every shard (H2D per device, all in flight together).
static int group_upload_query(aslam_loopdb_group *g, const uint8_t *q, int nq)
{
    int rc = group_reserve(g, nq);
    if (rc) return rc;
    memcpy(g->h_q, q, (size_t)nq * 32);
    for (int s = 0; s < g->n; ++s) {
        ASLAM_CUDA(cudaSetDevice(g->ctx[(size_t)s]->device));
        ASLAM_CUDA(cudaMemcpyAsync(g->d_q[(size_t)s], g->h_q, (size_t)nq * 32, cudaMemcpyHostToDevice, g->ctx[(size_t)s]->stream));
    }
    return ASLAM_OK;
}

static int group_exchange_buffers(aslam_loopdb_group *g, int nq)
{
    if (g->x_nq == nq) return ASLAM_OK;
    const size_t bytes = aslam_loop_exchange_buffer_bytes(g->n, nq);
    std::vector<unsigned long long *> peers((size_t)g->n, nullptr);
    for (int s = 0; s < g->n; ++s) {
        ASLAM_CUDA(cudaSetDevice(g->ctx[(size_t)s]->device));
        ASLAM_CUDA(cudaDeviceSynchronize());                 // nobody still reads or writes the old buffers
    }
    for (int s = 0; s < g->n; ++s) {
        ASLAM_CUDA(cudaSetDevice(g->ctx[(size_t)s]->device));
        if (g->x_buf[(size_t)s]) ASLAM_CUDA(cudaFree(g->x_buf[(size_t)s]));
        g->x_buf[(size_t)s] = nullptr;
        ASLAM_CUDA(cudaMalloc(&g->x_buf[(size_t)s], bytes));
        ASLAM_CUDA(cudaMemset(g->x_buf[(size_t)s], 0, bytes));
        ASLAM_CUDA(cudaDeviceSynchronize());
        peers[(size_t)s] = (unsigned long long *)g->x_buf[(size_t)s];
    }
    for (int s = 0; s < g->n; ++s) {
        ASLAM_CUDA(cudaSetDevice(g->ctx[(size_t)s]->device));
        if (!g->d_peers[(size_t)s]) ASLAM_CUDA(cudaMalloc(&g->d_peers[(size_t)s], sizeof(unsigned long long *) * (size_t)g->n));
        ASLAM_CUDA(cudaMemcpy(g->d_peers[(size_t)s], peers.data(), sizeof(unsigned long long *) * (size_t)g->n, cudaMemcpyHostToDevice));
    }
    g->x_nq = nq; g->round = 0;
    return ASLAM_OK;
}

extern "C" {

int aslam_loopdb_shard_group_create(const int *devices, int ndevices, int64_t capacity_rows_per_shard, int max_keyframes,
                                    aslam_loopdb_group **out)
{
    ASLAM_REQUIRE(devices && out, "NULL argument");
    ASLAM_REQUIRE(ndevices >= 1 && ndevices <= 64, "1..64 devices");
    ASLAM_REQUIRE(capacity_rows_per_shard > 0 && max_keyframes > 0, "capacity and max_keyframes must be positive");
    *out = nullptr;
    for (int a = 0; a < ndevices; ++a)
        for (int b = a + 1; b < ndevices; ++b)
            ASLAM_REQUIRE(devices[a] != devices[b], "a device may hold only one shard of a group");
    aslam_loopdb_group *g = new aslam_loopdb_group();
    g->n = ndevices; g->max_keyframes = max_keyframes;
    g->d_q.assign((size_t)ndevices, nullptr); g->d_out.assign((size_t)ndevices, nullptr); g->d_good.assign((size_t)ndevices, nullptr);
    g->x_buf.assign((size_t)ndevices, nullptr); g->d_peers.assign((size_t)ndevices, nullptr);
    int rc = ASLAM_OK;
    for (int s = 0; s < ndevices && rc == ASLAM_OK; ++s) {
        aslam_ctx *c = nullptr;
        rc = aslam_ctx_create(devices[s], nullptr, &c);
        if (rc) break;
        g->ctx.push_back(c);
        aslam_loopdb *db = nullptr;
        // the group evicts explicitly (the globally oldest keyframe), so a shard's own FIFO cap must never bite
        rc = aslam_loopdb_create(c, capacity_rows_per_shard, 0x7fffffff, &db);
        if (rc) break;
        db->scattered = true;
        g->shard.push_back(db);
        cudaEvent_t e;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { rc = fail(ASLAM_E_CUDA, "cudaEventCreate failed"); break; }
        g->ev.push_back(e);
    }
    // every device stores into every other device's exchange buffer
    for (int a = 0; a < ndevices && rc == ASLAM_OK; ++a)
        for (int b = 0; b < ndevices && rc == ASLAM_OK; ++b) {
            if (a == b) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[a], devices[b]);
            if (!can) { rc = fail(ASLAM_E_INVALID, "device %d cannot access device %d (no NVLink / PCIe peer path)", devices[a], devices[b]); break; }
            cudaSetDevice(devices[a]);
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) rc = fail(ASLAM_E_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d) -> %s", devices[a], devices[b], cudaGetErrorString(e));
            cudaGetLastError();
        }
    if (rc) {
        char msg[512];
        snprintf(msg, sizeof(msg), "%s", aslam_last_error());
        aslam_loopdb_shard_group_destroy(g);
        return fail(rc, "%s", msg);
    }
    *out = g;
    return ASLAM_OK;
}

void aslam_loopdb_shard_group_destroy(aslam_loopdb_group *g)
{
    if (!g) return;
    for (size_t s = 0; s < g->ctx.size(); ++s) {
        cudaSetDevice(g->ctx[s]->device);
        cudaDeviceSynchronize();
    }
    for (size_t s = 0; s < g->ctx.size(); ++s) {
        cudaSetDevice(g->ctx[s]->device);
        if (s < g->d_q.size()) { cudaFree(g->d_q[s]); cudaFree(g->d_out[s]); cudaFree(g->d_good[s]); cudaFree(g->x_buf[s]); cudaFree(g->d_peers[s]); }
        if (s < g->ev.size()) cudaEventDestroy(g->ev[s]);
        if (s < g->shard.size()) aslam_loopdb_destroy(g->shard[s]);
        aslam_ctx_destroy(g->ctx[s]);
    }
    if (g->h_q) cudaFreeHost(g->h_q);
    if (g->h_top2) cudaFreeHost(g->h_top2);
    if (g->h_good) cudaFreeHost(g->h_good);
    delete g;
}

int aslam_loopdb_shard_group_num_shards(const aslam_loopdb_group *g) { return g ? g->n : 0; }
int aslam_loopdb_shard_group_num_keyframes(const aslam_loopdb_group *g) { return g ? (int)g->live.size() : 0; }

int64_t aslam_loopdb_shard_group_num_rows(const aslam_loopdb_group *g)
{
    int64_t r = 0;
    if (g) for (const auto &s : g->live) r += s.rows;
    return r;
}

int aslam_loopdb_shard_group_keyframes(const aslam_loopdb_group *g, uint64_t *ids, int32_t *rows, int32_t *shards)
{
    ASLAM_REQUIRE(g, "group is NULL");
    for (size_t i = 0; i < g->live.size(); ++i) {
        if (ids) ids[i] = g->live[i].id;
        if (rows) rows[i] = g->live[i].rows;
        if (shards) shards[i] = (int32_t)(g->live[i].seq % (uint64_t)g->n);
    }
    return ASLAM_OK;
}

int aslam_loopdb_shard_group_add(aslam_loopdb_group *g, const uint8_t *desc, int n, uint64_t keyframe_id)
{
    ASLAM_REQUIRE(g, "group is NULL");
    ASLAM_REQUIRE(n >= 0 && (n == 0 || desc), "bad descriptor block");
    const int s = (int)(g->next_seq % (uint64_t)g->n);
    int rc = db_append(g->shard[(size_t)s], desc, n, keyframe_id, cudaMemcpyHostToDevice, g->next_row);
    if (rc) return rc;
    g->live.push_back({keyframe_id, g->next_seq, n});
    ++g->next_seq;
    g->next_row += n;
    while ((int)g->live.size() > g->max_keyframes) {          // LoopClosure.cpp:27-30
        db_evict_oldest(g->shard[(size_t)(g->live.front().seq % (uint64_t)g->n)]);
        g->live.pop_front();
    }
    return ASLAM_OK;
}

int aslam_loopdb_shard_group_clear(aslam_loopdb_group *g)
{
    ASLAM_REQUIRE(g, "group is NULL");
    for (auto *db : g->shard) { aslam_loopdb_clear(db); db->scattered = true; }
    g->live.clear(); g->next_seq = 0; g->next_row = 0;
    return ASLAM_OK;
}

// Per stored keyframe (global order, oldest first) the ratio-test survivors: every shard counts for its own keyframes
// (findCandidates' loop body, LoopClosure.cpp:86-95), and the only exchange is the int32 count per keyframe.
int aslam_loopdb_shard_group_query_segmented(aslam_loopdb_group *g, const uint8_t *q, int nq, double ratio, int32_t *good)
{
    ASLAM_REQUIRE(g && (good || g->live.empty()), "NULL argument");
    ASLAM_REQUIRE(nq >= 0, "negative nq");
    const int nkf = (int)g->live.size();
    for (int i = 0; i < nkf; ++i) good[i] = 0;
    if (nq == 0 || nkf == 0) return ASLAM_OK;
    ASLAM_REQUIRE(q, "q is NULL");
    int rc = group_upload_query(g, q, nq);
    if (rc) return rc;
    if (nkf > g->h_good_cap) {
        if (g->h_good) ASLAM_CUDA(cudaFreeHost(g->h_good));
        g->h_good = nullptr;
        g->h_good_cap = (int)align_up((size_t)nkf, 512);
        ASLAM_CUDA(cudaMallocHost(&g->h_good, sizeof(int32_t) * (size_t)g->h_good_cap * (size_t)g->n));
    }
    if (nkf > g->good_cap) {
        const int cap = (int)align_up((size_t)nkf, 512);
        for (int s = 0; s < g->n; ++s) {
            ASLAM_CUDA(cudaSetDevice(g->ctx[(size_t)s]->device));
            ASLAM_CUDA(cudaStreamSynchronize(g->ctx[(size_t)s]->stream));
            if (g->d_good[(size_t)s]) ASLAM_CUDA(cudaFree(g->d_good[(size_t)s]));
            g->d_good[(size_t)s] = nullptr;
            ASLAM_CUDA(cudaMalloc(&g->d_good[(size_t)s], sizeof(int32_t) * (size_t)cap));
        }
        g->good_cap = cap;
    }
    for (int s = 0; s < g->n; ++s) {
        aslam_loopdb *db = g->shard[(size_t)s];
        const int local = (int)db->segs.size();
        if (local == 0) continue;
        ASLAM_CUDA(cudaSetDevice(db->ctx->device));
        if ((rc = db_query_segmented_device(db, g->d_q[(size_t)s], nq, ratio, g->d_good[(size_t)s]))) return rc;
        ASLAM_CUDA(cudaMemcpyAsync(g->h_good + (size_t)s * g->h_good_cap, g->d_good[(size_t)s], sizeof(int32_t) * (size_t)local,
                                   cudaMemcpyDeviceToHost, db->ctx->stream));
    }
    for (int s = 0; s < g->n; ++s) {
        ASLAM_CUDA(cudaSetDevice(g->ctx[(size_t)s]->device));
        ASLAM_CUDA(cudaStreamSynchronize(g->ctx[(size_t)s]->stream));
    }
    std::vector<int> at((size_t)g->n, 0);                     // keyframes of a shard are in global order
    for (int i = 0; i < nkf; ++i) {
        const int s = (int)(g->live[(size_t)i].seq % (uint64_t)g->n);
        good[i] = g->h_good[(size_t)s * g->h_good_cap + (size_t)at[(size_t)s]++];
    }
    return ASLAM_OK;
}

// Flat top-2 over all shards: every device runs the fused scan + exchange kernel (stores into all peers' buffers over
// NVLink, flags, merge); the merged keys are identical on every device and device 0's copy is returned.
int aslam_loopdb_shard_group_query_global(aslam_loopdb_group *g, const uint8_t *q, int nq, uint64_t *top2)
{
    ASLAM_REQUIRE(g && (nq == 0 || (q && top2)), "NULL argument");
    ASLAM_REQUIRE(nq >= 0, "negative nq");
    if (nq == 0) return ASLAM_OK;
    ASLAM_REQUIRE(nq <= KNN_FUSED_MAX_QBLOCKS * KNN_QBLOCK, "too many queries for one exchange round; split the batch");
    int rc = group_upload_query(g, q, nq);
    if (rc) return rc;
    if ((rc = group_exchange_buffers(g, nq))) return rc;
    ++g->round;
    for (int s = 0; s < g->n; ++s) {
        aslam_loopdb *db = g->shard[(size_t)s];
        rc = db_query_global(db, g->d_q[(size_t)s], nq, g->d_out[(size_t)s], (uint64_t *const *)g->d_peers[(size_t)s], g->n, s, g->round);
        if (rc) return rc;          // a shard that did not launch leaves the others waiting until their 4 s timeout
    }
    ASLAM_CUDA(cudaSetDevice(g->ctx[0]->device));
    ASLAM_CUDA(cudaMemcpyAsync(g->h_top2, g->d_out[0], sizeof(uint64_t) * 2 * (size_t)nq, cudaMemcpyDeviceToHost, g->ctx[0]->stream));
    unsigned int bad = 0;
    for (int s = 0; s < g->n; ++s) {
        ASLAM_CUDA(cudaSetDevice(g->ctx[(size_t)s]->device));
        ASLAM_CUDA(cudaStreamSynchronize(g->ctx[(size_t)s]->stream));
        unsigned int st = 0;
        ASLAM_CUDA(cudaMemcpy(&st, g->shard[(size_t)s]->d_status, sizeof(st), cudaMemcpyDeviceToHost));
        bad |= st;
    }
    if (bad) return fail(ASLAM_E_CUDA, "a shard's exchange timed out waiting for a peer");
    memcpy(top2, g->h_top2, sizeof(uint64_t) * 2 * (size_t)nq);
    return ASLAM_OK;
}

// Device d's copy of the last flat query's merged keys (all copies are identical; used by the parity tests).
int aslam_loopdb_shard_group_read_result(aslam_loopdb_group *g, int shard, int nq, uint64_t *top2)
{
    ASLAM_REQUIRE(g && top2, "NULL argument");
    ASLAM_REQUIRE(shard >= 0 && shard < g->n && nq >= 0 && nq <= g->q_cap, "bad shard or nq");
    ASLAM_CUDA(cudaSetDevice(g->ctx[(size_t)shard]->device));
    ASLAM_CUDA(cudaStreamSynchronize(g->ctx[(size_t)shard]->stream));
    ASLAM_CUDA(cudaMemcpy(top2, g->d_out[(size_t)shard], sizeof(uint64_t) * 2 * (size_t)nq, cudaMemcpyDeviceToHost));
    return ASLAM_OK;
}

// verifyGeometry's re-match (LoopClosure.cpp:116-130) against the RESIDENT rows of stored keyframe kf_index (global order).
int aslam_loopdb_shard_group_match_keyframe(aslam_loopdb_group *g, int kf_index, const uint8_t *q, int nq, double ratio,
                                            aslam_match *out, int *nout)
{
    ASLAM_REQUIRE(g && nout, "NULL argument");
    ASLAM_REQUIRE(kf_index >= 0 && kf_index < (int)g->live.size(), "keyframe index out of range");
    const int s = (int)(g->live[(size_t)kf_index].seq % (uint64_t)g->n);
    int local = 0;
    for (int i = 0; i < kf_index; ++i)
        if ((int)(g->live[(size_t)i].seq % (uint64_t)g->n) == s) ++local;
    return aslam_loopdb_match_keyframe(g->shard[(size_t)s], local, q, nq, ratio, out, nout);
}

}  // extern "C"
