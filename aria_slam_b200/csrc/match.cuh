// Hamming kNN (k=2) building blocks shared by the matcher, the tracker and the loop database.
#pragma once
#include "common.cuh"

namespace aslam {

// One scan of a query set against a row range of a train set.  Counts may live on the device
// (results of an extraction that has not been read back): *_dev overrides the host value.
struct KnnProblem {
    const uint8_t *q;     // query rows (32 B each, 16-B aligned)
    const uint8_t *t;     // train rows; reported indices are relative to this base
    const int32_t *nq_dev;
    const int32_t *nt_dev;
    int32_t nq, nt;       // host-known counts, or capacities when *_dev is set
    int32_t t_lo, t_hi;   // rows of t this problem scans
};

// Packed neighbour: distance in the top 9 bits, train row in the low 23 (rows < 2^23 per
// problem).  Unsigned order of keys == lexicographic (distance, row); empty == 0xFFFFFFFF.
constexpr int KEY_IDX_BITS = 23;
constexpr uint32_t KEY_EMPTY = 0xFFFFFFFFu;
constexpr int KNN_MAX_ROWS = 1 << KEY_IDX_BITS;

#ifndef ASLAM_KNN_THREADS
#define ASLAM_KNN_THREADS 128
#endif
constexpr int KNN_THREADS = ASLAM_KNN_THREADS;
constexpr int KNN_RQ = 4;                          // queries held in registers per thread
constexpr int KNN_QBLOCK = KNN_THREADS * KNN_RQ;   // queries per CTA

struct RatioTest {
    double ratio;
    int as_double;   // 0: float arithmetic (CudaMatcher.cpp:60); 1: double (LoopClosure.cpp:92)
    int disabled;    // ratio < 0: keep every query that has two neighbours (extension)
};

// The legacy tracking loop's match filter (src/main.cpp:164-175 of the reference): a match is dropped when its query
// keypoint OR its train keypoint lies inside a detection box of a dynamic class (src/main.cpp:29-50).  boxes == NULL
// switches it off.  half_open: 1 = cv::Rect::contains as main.cpp uses it (x1 <= x < x2), 0 = core::Detection::contains
// (x1 <= x <= x2, include/core/Types.hpp:109-111).  nfiltered[group] receives main.cpp's filtered_count.
struct DynFilter {
    const aslam_keypoint *q_kps = nullptr, *t_kps = nullptr;
    const aslam_box *boxes = nullptr;
    int nboxes = 0, half_open = 0;
    int32_t *nfiltered = nullptr;
};

// Local row of a database shard -> global row.  nseg == 0: one contiguous range (row_base + local row).  Otherwise the
// shard's keyframes are scattered over the global row space (round-robin placement of the single-process shard
// group): keyframe i holds local rows [lstart[i], lstart[i+1]) = global rows gbase[i] + 0, 1, ...; local and global
// order agree inside a shard, so ties still resolve to the lowest global row.
struct RowMap {
    long long row_base = 0;
    const long long *gbase = nullptr;
    const int *lstart = nullptr;      // nseg + 1 entries
    int nseg = 0;
};

// Number of row ranges to cut a train set into so the grid covers the GPU a few times.
int choose_splits(const aslam_ctx *ctx, int nq, long long nt, int nsets);
// d_probs[0..splits) = the pair (q, t) cut into equal row ranges (device-side table build).
int launch_fill_split_problems(aslam_ctx *ctx, KnnProblem *d_probs, const uint8_t *q, const uint8_t *t,
                               const int32_t *nq_dev, const int32_t *nt_dev, int nq, int nt, int splits);

// merge_div <= 1: partial[slot * q_stride + q] = (best, second) keys of problem `slot`.
// merge_div  > 1: every merge_div consecutive problems (the row ranges of one pair) are merged by the scan
//                 itself: partial[(slot / merge_div) * q_stride + q]; nprobs must be a multiple of merge_div.
int launch_knn2_partial(aslam_ctx *ctx, const KnnProblem *d_probs, int nprobs, int max_nq,
                        uint2 *d_partial, int q_stride, int merge_div = 1, bool armed = false);
// armed: the merged slots already hold KEY_EMPTY (a finish kernel launched with rearm reset them), no memset.

// One merged slot per group of `splits` consecutive problems (launch_knn2_partial with merge_div = splits);
// survivors of the ratio test are
// written in query order to d_out[group * out_stride ...], their number to d_nout[group].
// d_raw_idx/d_raw_dist (optional, [group][q_stride][2]) receive the merged neighbours.
int launch_knn2_finish_matches(aslam_ctx *ctx, const KnnProblem *d_probs, const uint2 *d_partial,
                               int ngroups, int splits, int q_stride, RatioTest rt,
                               aslam_match *d_out, int out_stride, int32_t *d_nout,
                               int32_t *d_raw_idx, int32_t *d_raw_dist, const DynFilter &filter = DynFilter());

// One count per slot: queries passing the ratio test within that slot alone.  Slot s was merged from the `splits`
// consecutive problems s*splits .. (launch_knn2_partial with merge_div = splits; 1 = one problem per slot).
int launch_knn2_finish_counts(aslam_ctx *ctx, const KnnProblem *d_probs, const uint2 *d_partial,
                              int nslots, int splits, int q_stride, RatioTest rt, int32_t *d_good);

// All slots merged into one top-2 per query, as 64-bit keys (distance << 40 | row_base + row
// + slot_row_offset), where slot rows are relative to the problem's t base and row_base is
// added on top.
int launch_knn2_finish_global(aslam_ctx *ctx, uint2 *d_partial, int nslots, int nq,
                              int q_stride, const RowMap &map, uint64_t *d_top2, bool rearm = false);

// As launch_knn2_finish_global, with the result stored into slot `my_part` of every peer's
// [nparts][nq][2] buffer (peer-mapped pointers in device memory): the all-gather is the epilogue.
int launch_knn2_finish_global_push(aslam_ctx *ctx, uint2 *d_partial, int nslots, int nq,
                                   int q_stride, const RowMap &map, uint64_t *const *d_peer_bufs,
                                   int nparts, int my_part, bool rearm = false);

// The exchange step of the sharded flat query, run as the tail of the scan kernel itself
// (aslam_loopdb_query_global_exchange_device).  Peer buffer layout, identical on all ranks:
//     uint64 data[2][nparts][nq][2];  uint64 flags[2][nparts][qblocks];     qblocks = ceil(nq / KNN_QBLOCK)
struct ExchangeArgs {
    unsigned long long *const *peer_bufs = nullptr;   // device array of nparts peer-mapped buffer addresses
    int nparts = 0, my_part = 0, nq = 0;
    unsigned long long round = 0;
    unsigned int *arrive = nullptr;        // [qblocks] CTA arrival counters, zero between launches
    unsigned long long *out = nullptr;     // [nq][2] merged keys
    unsigned long long *stamps = nullptr;  // [qblocks][4]: %globaltimer of the last push / last all-peers-seen, sum of (seen - push), exchanges
    unsigned int *status = nullptr;        // bit 0: a wait for a peer's flag timed out (the result is then invalid)
    RowMap map;
};
constexpr int KNN_FUSED_MAX_QBLOCKS = 64;  // waiting CTAs (one per query block) must stay far below the resident CTA slots

// Shard scan + exchange in ONE launch: every CTA merges its row range into the per-query slot (atomic minima); the
// last CTA to arrive for a block of KNN_QBLOCK queries turns the merged slots into global keys, stores them into every
// peer's buffer over NVLink, raises this shard's flag for that block on every peer, waits for the peers' flags of the
// same block and round, merges and writes X.out.  The slots must hold KEY_EMPTY on entry and are left so.
// overlap_previous: launched with programmatic stream serialization -- legal only when the kernel before it on the stream
// is the previous query's launch of this same kernel and nothing it reads before its merge (problem table, queries,
// database rows) was written by that kernel.
int launch_knn2_scan_exchange(aslam_ctx *ctx, const KnnProblem *d_probs, int nprobs, int nq, uint2 *d_partial,
                              int q_stride, const ExchangeArgs &X, bool overlap_previous = false);

}  // namespace aslam
