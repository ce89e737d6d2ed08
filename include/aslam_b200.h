/* aslam_b200.h -- C ABI of the B200 (sm_100a) ORB frontend: extraction, Hamming kNN matching and the
 * loop-closure descriptor query of robertteleng/aria-slam.
 *
 * This is the drop-in boundary.  Plain pointers and sizes only; caller owns every host buffer,
 * the library owns all device memory; no exception crosses it.  Every entry point returns an
 * aslam_status (0 = ok) and aslam_last_error() gives the message for the calling thread.
 *
 * Each group cites the reference interface it stands in for (paths relative to the reference
 * repository).  The C++ adapters in aria_slam_b200/adapters/ bind these to the reference's own
 * abstract classes (include/interfaces/IFeatureExtractor.hpp, IMatcher.hpp, ILoopDetector.hpp);
 * INTEGRATION.md shows the binding a maintainer adds.
 *
 * Threading contract (same as the reference adapters, OrbCudaExtractor.hpp:48): one handle,
 * one CUDA stream, one caller thread at a time.
 */
#ifndef ASLAM_B200_H
#define ASLAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ASLAM_API __attribute__((visibility("default")))

typedef enum {
    ASLAM_OK = 0,
    ASLAM_E_INVALID = 1,     /* bad argument */
    ASLAM_E_CUDA = 2,        /* a CUDA runtime call failed; message carries the CUDA error */
    ASLAM_E_NOMEM = 3,
    ASLAM_E_STATE = 4,       /* call order violated (e.g. sync without a pending extract) */
    ASLAM_E_OVERFLOW = 5,    /* results were produced but a capacity clipped them (see call) */
    ASLAM_E_NODEVICE = 6     /* no CUDA device: there is no CPU fallback on this path */
} aslam_status;

typedef struct aslam_ctx aslam_ctx;
typedef struct aslam_orb aslam_orb;
typedef struct aslam_loopdb aslam_loopdb;
typedef struct aslam_loopdb_group aslam_loopdb_group;

/* aria::core::KeyPoint, include/core/Types.hpp:9-15 (24 bytes, same field order) */
typedef struct {
    float x, y;
    float size;
    float angle;
    float response;
    int32_t octave;
} aslam_keypoint;

/* aria::core::Match, include/core/Types.hpp:97-101 */
typedef struct {
    int32_t query_idx;
    int32_t train_idx;
    float distance;
} aslam_match;

/* aria::core::Detection, include/core/Types.hpp:104-112 */
typedef struct {
    float x1, y1, x2, y2;
    float confidence;
    int32_t class_id;
} aslam_box;

/* The argument list of cv::ORB::create as the reference calls it
 * (src/adapters/gpu/OrbCudaExtractor.cpp:35-45, src/euroc_eval.cpp:88).  Only the values the
 * reference uses are implemented: first_level 0, wta_k 2, HARRIS score, patch 31. */
typedef struct {
    int32_t nfeatures;       /* 1000 / 2000 / 4000 */
    float scale_factor;      /* 1.2f */
    int32_t nlevels;         /* 8 (max 8) */
    int32_t edge_threshold;  /* 31 */
    int32_t first_level;     /* 0 */
    int32_t wta_k;           /* 2 */
    int32_t score_type;      /* 0 = HARRIS_SCORE */
    int32_t patch_size;      /* 31 */
    int32_t fast_threshold;  /* 20 */
    int32_t blur_mode;       /* 2 = cv2-wheel hybrid FMA split, 1 = always fused, 0 = never */
    int32_t grid_cell;       /* 0 = off: cv::ORB's plain retainBest (the parity mode, the default).  8..4096 = per-level
                              * grid bucketing of the second cull with square cells of that many level pixels: within
                              * a level every cell gives its best corner (by Harris response) before any cell gives
                              * its second best, until the level's quota is full (ties kept).  An extension named by
                              * BASELINE.json; neither cv::ORB nor the reference buckets. */
} aslam_orb_params;

/* ------------------------------------------------------------------ context
 * Stands in for the adapter constructors' device/stream handling
 * (OrbCudaExtractor.cpp:21-33, CudaMatcher.cpp:7-19; FactoryConfig::cuda_device,
 * include/factory/PipelineFactory.hpp:21).  stream = a cudaStream_t to borrow, or NULL to
 * create (and later destroy) a private one. */
/* 1 when this library is the bounds-checking debug build (-DASLAM_DEBUG_BOUNDS=1: every staged index of every kernel is
 * checked on the device); *violations = checks that failed on the current device so far (after a device sync).  0 and
 * *violations = 0 in the release build, whose kernels carry no checks. */
ASLAM_API int aslam_debug_bounds(uint64_t *violations);
ASLAM_API int aslam_device_count(void);      /* CUDA devices visible to this process (0 without a driver) */
ASLAM_API int aslam_ctx_create(int device, void *stream, aslam_ctx **out);
ASLAM_API void aslam_ctx_destroy(aslam_ctx *ctx);
ASLAM_API int aslam_ctx_sync(aslam_ctx *ctx);
ASLAM_API void *aslam_ctx_stream(aslam_ctx *ctx);
ASLAM_API int aslam_ctx_sm_count(aslam_ctx *ctx);
ASLAM_API const char *aslam_last_error(void);
ASLAM_API const char *aslam_version(void);
/* number of this library's kernels launched by the calling process so far */
ASLAM_API uint64_t aslam_kernel_launches(void);
/* Optional per-kernel timing: when enabled, every launch on the context is bracketed by a CUDA
 * event pair on its stream.  aslam_prof_read waits for the stream and returns the accumulated
 * device time and launch count of one kernel kind (names via aslam_prof_name, 0..count-1). */
ASLAM_API int aslam_prof_enable(aslam_ctx *ctx, int on);
ASLAM_API int aslam_prof_count(void);
ASLAM_API const char *aslam_prof_name(int kind);
ASLAM_API int aslam_prof_read(aslam_ctx *ctx, int kind, double *total_ms, int64_t *launches, int reset);
/* pinned host memory for callers that want truly asynchronous H2D/D2H */
ASLAM_API int aslam_host_alloc(size_t bytes, void **out);
/* The same as write-combined memory (cudaHostAllocWriteCombined): for buffers the host only WRITES (frames on their
 * way to the device); not snooped during the transfer, but very slow for the host to read back. */
ASLAM_API int aslam_host_alloc_wc(size_t bytes, void **out);
ASLAM_API int aslam_host_free(void *p);
/* cudaMemcpyAsync(cudaMemcpyDefault) on the context's stream, for hosts that do not link the CUDA runtime themselves. */
ASLAM_API int aslam_memcpy_async(aslam_ctx *ctx, void *dst, const void *src, size_t bytes);

/* ------------------------------------------------------------------ ORB extraction
 * Replaces OrbCudaExtractor::extract / extractAsync / sync / setMaxFeatures / getMaxFeatures
 * (src/adapters/gpu/OrbCudaExtractor.cpp:64-128,136-171,176-210,212-218) and the legacy
 * Frame constructors (src/legacy/Frame.cpp:6-49).  Results follow the reference's CPU path,
 * cv::ORB::detectAndCompute.
 *
 * A handle is built for one frame size and up to max_batch frames per call. */
ASLAM_API void aslam_orb_default_params(aslam_orb_params *p);
ASLAM_API int aslam_orb_create(aslam_ctx *ctx, const aslam_orb_params *p, int width, int height,
                               int max_batch, aslam_orb **out);
ASLAM_API void aslam_orb_destroy(aslam_orb *orb);
/* setMaxFeatures: like the reference (OrbCudaExtractor.cpp:212-216) every other parameter
 * returns to its default.  Must not be called with an extract pending. */
ASLAM_API int aslam_orb_set_max_features(aslam_orb *orb, int nfeatures);
ASLAM_API int aslam_orb_get_max_features(const aslam_orb *orb);
/* Rows the caller must provide per frame in aslam_orb_sync's kps/desc buffers.  It is
 * nfeatures plus head-room for response ties, which cv::ORB keeps (KeyPointsFilter::retainBest). */
ASLAM_API int aslam_orb_capacity(const aslam_orb *orb);

/* Host-buffer path.  image: grayscale, row_stride bytes per row (IFeatureExtractor.hpp:14 is
 * the tightly packed case row_stride == width).  mask: NULL or width*height bytes, 0 = reject
 * (cv::ORB's mask argument).  boxes: NULL or detections whose dynamic classes
 * (src/main.cpp:29-41) remove keypoints inside them after selection
 * (SlamPipeline::filterDynamicKeypoints, include/pipeline/SlamPipeline.hpp:96-99, using
 * Detection::contains, include/core/Types.hpp:109-111).
 * The pixel buffers must stay valid until aslam_orb_sync returns.
 * The launch chain of a call shape (frames per call, masks or not) is captured into a CUDA graph on first use and replayed
 * by later calls of the same shape (not on the legacy default stream, not while aslam_prof_* timing is on;
 * ASLAM_NO_GRAPH=1 in the environment disables it). */
ASLAM_API int aslam_orb_extract_async(aslam_orb *orb, const uint8_t *image, int width, int height,
                                      int row_stride, const uint8_t *mask,
                                      const aslam_box *boxes, int nboxes);
/* n frames, frame i at images + i*frame_stride.  masks: NULL or n masks, width*height each. */
ASLAM_API int aslam_orb_extract_batch_async(aslam_orb *orb, const uint8_t *images, int n,
                                            size_t frame_stride, int row_stride,
                                            const uint8_t *masks);
/* Waits, then copies out: kps[n][capacity], desc[n][capacity][32], counts[n].
 * ASLAM_E_OVERFLOW: a frame had more tied keypoints than capacity; the first capacity (by level, and
 * within a level tile-major: 62x62-pixel tiles in row-major order, row-major inside a tile) were returned. */
ASLAM_API int aslam_orb_sync(aslam_orb *orb, aslam_keypoint *kps, uint8_t *desc, int32_t *counts);

/* Device-resident path (inputs already in HBM; results stay there). */
ASLAM_API int aslam_orb_extract_batch_device(aslam_orb *orb, const uint8_t *d_images, int n,
                                             size_t frame_stride, int row_stride,
                                             const uint8_t *d_masks);
ASLAM_API int aslam_orb_device_results(aslam_orb *orb, const aslam_keypoint **d_kps,
                                       const uint8_t **d_desc, const int32_t **d_counts);

/* Keeps the descriptors of the first frame of the extract just synced on the device beyond the next extract: they are
 * copied (device to device, 64 KB) into one of two ring slots owned by the handle; *d_desc / *rows describe that slot,
 * which stays valid until the second-next call of this function.  With aslam_match_knn2_resident this chains
 * extract -> match of consecutive frames without a PCIe round trip of descriptors (the reference's
 * OrbCudaExtractor + CudaMatcher::matchGpu pairing).  Once per synced extract. */
ASLAM_API int aslam_orb_retain_descriptors(aslam_orb *orb, const uint8_t **d_desc, int32_t *rows);

/* Extract n frames and match frame i-1 (query) against frame i (train) in the same stream, the
 * direction of the reference's tracking loop, knnMatch(prev_frame, current_frame) (src/main.cpp:158-161,
 * src/euroc_eval.cpp:167-169): matches[i] belongs to frame i, its query_idx indexes the keypoints of
 * frame i-1 and its train_idx those of frame i, and it holds at most as many matches as frame i-1 has
 * keypoints.  Frame 0 is matched against the last frame of the previous call on this handle (no
 * matches on the first call).  One call = the per-frame body of SlamPipeline::processFrame's
 * extract -> match (docs H12:595-601 of the reference) for a whole batch.
 * ratio as in IMatcher::match (include/interfaces/IMatcher.hpp:19-24). */
ASLAM_API int aslam_orb_track_batch_async(aslam_orb *orb, const uint8_t *images, int n,
                                          size_t frame_stride, int row_stride, float ratio,
                                          int images_on_device);
/* As aslam_orb_sync plus matches[n][capacity], match_counts[n].  Any pointer may be NULL to
 * skip that copy. */
ASLAM_API int aslam_orb_track_sync(aslam_orb *orb, aslam_keypoint *kps, uint8_t *desc,
                                   int32_t *counts, aslam_match *matches, int32_t *match_counts);

/* Stage read-back for the parity tests (all after a sync).  level images are W_l x H_l,
 * tightly packed into out. */
ASLAM_API int aslam_orb_level_size(const aslam_orb *orb, int level, int *w, int *h);
ASLAM_API int aslam_orb_read_level(aslam_orb *orb, int frame, int level, int blurred, uint8_t *out);
/* FAST candidates after NMS/border/mask of one level as packed (score<<24 | y<<12 | x);
 * returns how many through *n (cap entries written at most). */
ASLAM_API int aslam_orb_read_corners(aslam_orb *orb, int frame, int level, uint32_t *out, int cap, int *n);

/* ------------------------------------------------------------------ Hamming kNN (k = 2)
 * Replaces CudaMatcher::match (src/adapters/gpu/CudaMatcher.cpp:28-68) and the inlined twins
 * src/main.cpp:158-175, src/euroc_eval.cpp:167-175.  Descriptors are rows of 32 bytes.
 * Results follow cv::BFMatcher(NORM_HAMMING)::knnMatch(k=2): ascending (distance, train index).
 *
 * aslam_match_knn2: ratio test d0 < ratio*d1 evaluated in float exactly as CudaMatcher.cpp:60;
 * out receives at most nq matches in query order; *nout their number.  nt < 2 or nq == 0
 * yields *nout = 0 (the reference emits a match only when both neighbours exist). */
ASLAM_API int aslam_match_knn2(aslam_ctx *ctx, const uint8_t *q, int nq, const uint8_t *t, int nt,
                               float ratio, aslam_match *out, int *nout);
/* The legacy tracking loop's form (src/main.cpp:158-175): as aslam_match_knn2, and a surviving match is dropped when
 * its query keypoint OR its train keypoint lies inside a detection box of a dynamic class (isInDynamicObject,
 * src/main.cpp:29-50).  q_kps[nq] / t_kps[nt] are the keypoints the descriptor rows belong to.
 * half_open = 1: cv::Rect::contains as main.cpp calls it on TRTInference's Detection::box, with x1 = box.x,
 * x2 = box.x + box.width (x1 <= x < x2); half_open = 0: aria::core::Detection::contains (x1 <= x <= x2,
 * include/core/Types.hpp:109-111).  *nfiltered (may be NULL) = matches removed, main.cpp's filtered_count.
 * The predicate runs inside the match compaction kernel; nboxes == 0 is aslam_match_knn2. */
ASLAM_API int aslam_match_knn2_dynamic(aslam_ctx *ctx, const uint8_t *q, const aslam_keypoint *q_kps, int nq,
                                       const uint8_t *t, const aslam_keypoint *t_kps, int nt, float ratio,
                                       const aslam_box *boxes, int nboxes, int half_open,
                                       aslam_match *out, int *nout, int *nfiltered);
/* Raw neighbours: idx[nq][2], dist[nq][2]; -1 where fewer than two train rows exist. */
ASLAM_API int aslam_knn2_raw(aslam_ctx *ctx, const uint8_t *q, int nq, const uint8_t *t, int nt,
                             int32_t *idx, int32_t *dist);
/* One query set against several train sets in one launch: IMatcher::matchMultiple
 * (include/interfaces/IMatcher.hpp:27-37).  t = concatenated train rows, t_start[nsets+1]
 * row offsets; out[nsets][nq], nout[nsets]. */
ASLAM_API int aslam_match_multi(aslam_ctx *ctx, const uint8_t *q, int nq, const uint8_t *t,
                                const int32_t *t_start, int nsets, float ratio,
                                aslam_match *out, int32_t *nout);
/* Device-resident single problem (used by bench `value` and by callers chaining extract ->
 * match in HBM): d_q/d_t device rows; d_out[nq], d_nout device. */
ASLAM_API int aslam_match_knn2_device(aslam_ctx *ctx, const uint8_t *d_q, int nq,
                                      const uint8_t *d_t, int nt, float ratio,
                                      aslam_match *d_out, int32_t *d_nout);

/* CudaMatcher::matchGpu (include/adapters/gpu/CudaMatcher.hpp:22-28: "GPU-to-GPU matching, zero-copy when used with
 * OrbCudaExtractor"): descriptor rows already on the device (aslam_orb_retain_descriptors, aslam_orb_device_results),
 * matches to a host buffer -- the two 64 KB descriptor sets of a frame pair never cross PCIe again. */
ASLAM_API int aslam_match_knn2_resident(aslam_ctx *ctx, const uint8_t *d_q, int nq, const uint8_t *d_t, int nt,
                                        float ratio, aslam_match *out, int *nout);

/* ------------------------------------------------------------------ loop-closure database
 * Replaces the descriptor side of LoopClosureDetector (src/legacy/LoopClosure.cpp): the
 * keyframe store addKeyFrame (:24-31, FIFO cap 500) and findCandidates' per-keyframe
 * knnMatch + 0.7 ratio count (:72-114), as one resident [rows,32] array in HBM
 * (the design of docs/milestones/H14_GPU_LOOPCLOSURE_AUDIT.md:41-78 in the reference).
 * A database may be one shard of a larger one: global row = row_base + local row. */
ASLAM_API int aslam_loopdb_create(aslam_ctx *ctx, int64_t capacity_rows, int max_keyframes,
                                  aslam_loopdb **out);
ASLAM_API void aslam_loopdb_destroy(aslam_loopdb *db);
/* Append one keyframe's descriptors (host rows).  Evicts oldest keyframes while more than
 * max_keyframes are stored (LoopClosure.cpp:27-30).  Like the reference's deque the store never refuses a
 * keyframe: capacity_rows is the initial size and the array doubles when the live rows outgrow it
 * (ASLAM_E_NOMEM only when the device is out of memory; a failed call changes nothing).
 * The copy is asynchronous on the context's stream: pageable host rows are staged before the call returns and
 * may be reused at once; PINNED host rows (aslam_host_alloc) must stay unchanged until the next query or
 * aslam_ctx_sync. */
ASLAM_API int aslam_loopdb_add(aslam_loopdb *db, const uint8_t *desc, int n, uint64_t keyframe_id);
ASLAM_API int aslam_loopdb_add_device(aslam_loopdb *db, const uint8_t *d_desc, int n, uint64_t keyframe_id);
/* A shard whose keyframes are NOT contiguous in the global row space (keyframes dealt round-robin to the shards as they
 * arrive, which is how aslam_loopdb_shard_group places them): this keyframe's rows are global rows global_row_base,
 * global_row_base + 1, ...  Keyframes must arrive in increasing global order; a database is either built with this call
 * or with aslam_loopdb_add (contiguous rows from row_base), not both.  Flat queries translate shard rows to global rows
 * through the keyframe table, so the merged answer of any placement equals the unsharded one. */
ASLAM_API int aslam_loopdb_add_at(aslam_loopdb *db, const uint8_t *desc, int n, uint64_t keyframe_id, int64_t global_row_base);
ASLAM_API int aslam_loopdb_clear(aslam_loopdb *db);
/* Persistence of the keyframe store (the reference keeps it in memory only, LoopClosure.hpp; its H14 notes plan
 * chunked stores): save writes ids, per-keyframe row counts, row_base and the live rows to one file; load
 * replaces the content of db (capacity and max_keyframes of db must suffice). */
ASLAM_API int aslam_loopdb_save(aslam_loopdb *db, const char *path);
ASLAM_API int aslam_loopdb_load(aslam_loopdb *db, const char *path);
ASLAM_API int aslam_loopdb_num_keyframes(const aslam_loopdb *db);
ASLAM_API int64_t aslam_loopdb_num_rows(const aslam_loopdb *db);
ASLAM_API int64_t aslam_loopdb_capacity_rows(const aslam_loopdb *db);
ASLAM_API int aslam_loopdb_set_row_base(aslam_loopdb *db, int64_t row_base);
/* ids[k], rows[k] of the stored keyframes, oldest first. */
ASLAM_API int aslam_loopdb_keyframes(const aslam_loopdb *db, uint64_t *ids, int32_t *rows);
/* Per stored keyframe: how many of the nq queries pass d0 < ratio*d1 (double arithmetic as in
 * LoopClosure.cpp:92) against that keyframe's rows.  good[num_keyframes]. */
ASLAM_API int aslam_loopdb_query_segmented(aslam_loopdb *db, const uint8_t *q, int nq, double ratio,
                                           int32_t *good);
/* The same with the queries already on the device and the counts left there (d_good[num_keyframes]); asynchronous
 * on the context's stream.  The sharded per-keyframe query is this call on every shard followed by an exchange of
 * the counts alone. */
ASLAM_API int aslam_loopdb_query_segmented_device(aslam_loopdb *db, const uint8_t *d_q, int nq, double ratio,
                                                  int32_t *d_good);
/* verifyGeometry's re-match of a loop candidate (LoopClosure.cpp:116-130): knnMatch(query, keyframe, 2) + the
 * d0 < ratio*d1 test in double, against the RESIDENT rows of stored keyframe kf_index (oldest = 0) -- no keyframe
 * descriptors cross PCIe.  out receives at most nq matches in query order (train_idx = row within the keyframe). */
ASLAM_API int aslam_loopdb_match_keyframe(aslam_loopdb *db, int kf_index, const uint8_t *q, int nq, double ratio,
                                          aslam_match *out, int *nout);
/* Flat query: per query the two smallest (distance, global row) over the whole shard as
 * packed keys (distance << 40 | global row); empty slots are UINT64_MAX.  d_q, d_top2 are
 * device pointers (top2[nq][2]); asynchronous on the context's stream. */
ASLAM_API int aslam_loopdb_query_global_device(aslam_loopdb *db, const uint8_t *d_q, int nq,
                                               uint64_t *d_top2);
/* Fused shard scan + all-gather: as aslam_loopdb_query_global_device, but the kernel that finishes the
 * scan stores this shard's keys straight into slot `my_part` of EVERY peer's gather buffer
 * (d_peer_bufs[p] = device pointer of rank p's [nparts][nq][2] buffer, peer-mapped over NVLink; the
 * array itself lives in device memory).  No collective call is needed afterwards, only a barrier
 * before aslam_loop_merge_top2_device reads the local buffer.  Asynchronous on the context's stream. */
ASLAM_API int aslam_loopdb_query_global_push_device(aslam_loopdb *db, const uint8_t *d_q, int nq,
                                                    uint64_t *const *d_peer_bufs, int nparts, int my_part);
/* Shard scan AND exchange in ONE kernel launch (no collective call, no second launch): every CTA of the scan merges its
 * row range into the per-query slots; the last CTA to arrive for a block of 512 queries stores that block's keys into every
 * peer's buffer over NVLink, publishes a per-(shard, block) flag with release semantics, waits (acquire) for the
 * peers' flags of the same block and round, merges, and writes d_out[nq][2].
 * Every peer buffer (peer-mapped: CUDA IPC, cudaDeviceEnablePeerAccess or torch symmetric memory) is laid out as
 *     uint64 data[2][nparts][nq][2];  uint64 flags[2][nparts][ceil(nq/512)];        (zeroed before round 1)
 * `round` = 1, 2, 3, ... identical on all ranks; rounds alternate between the two halves.  Every rank must call this
 * once per round, each on its own GPU, with the same nq <= 32768.  A rank that waits more than 4 s for a peer gives up
 * and sets bit 0 of the status word (aslam_loopdb_exchange_status); its output is then invalid. */
ASLAM_API int aslam_loopdb_query_global_exchange_device(aslam_loopdb *db, const uint8_t *d_q, int nq,
                                                        uint64_t *const *d_peer_bufs, int nparts, int my_part,
                                                        uint64_t round, uint64_t *d_out);
/* Pipelining of consecutive fused queries (off by default).  When on, a fused query launched directly behind another
 * fused query of the same database (same query buffer, same database content) starts its scan while the previous
 * query's last CTAs are still exchanging with the peers (programmatic dependent launch; it holds its own merge until
 * the previous kernel has completed).  The caller's part of the contract: between two such calls no KERNEL on this
 * context's stream writes the query buffer (copies are fine -- they end the overlap by themselves).  A throughput
 * feature for streams of queries; a single query's latency is unchanged. */
ASLAM_API int aslam_loopdb_set_pipelining(aslam_loopdb *db, int on);
/* Waits for the stream; *status = 0 when no exchange on this database ever timed out.  stamps (may be NULL) receives
 * 4 words per query block: %globaltimer (ns) of the last exchange's push to the peers and of the moment all peers'
 * flags had been seen, the SUM of (seen - push) over all exchanges since the last reset -- the time this rank spent
 * waiting for its slowest peer -- and the number of exchanges.  reset != 0 clears the sums. */
ASLAM_API int aslam_loopdb_exchange_status(aslam_loopdb *db, uint32_t *status, uint64_t *stamps, int nqblocks, int reset);
ASLAM_API size_t aslam_loop_exchange_buffer_bytes(int nparts, int nq);
/* Host-buffer form of the flat query (copies in and out, synchronous). */
ASLAM_API int aslam_loopdb_query_global(aslam_loopdb *db, const uint8_t *q, int nq, uint64_t *top2);
/* Merge step after the all-gather of shard results: d_parts[nparts][nq][2] packed keys ->
 * d_out[nq][2], the two smallest per query.  Total order, so any shard count gives the
 * single-GPU answer.  Asynchronous on the context's stream. */
ASLAM_API int aslam_loop_merge_top2_device(aslam_ctx *ctx, const uint64_t *d_parts, int nparts,
                                           int nq, uint64_t *d_out);

/* ------------------------------------------------------------------ sharded loop database
 * the keyframe database is sharded by keyframe over the GPUs of one node, the 64 KB query block is
 * replicated, each GPU answers for its shard and one small exchange over NVLink follows (flat query: 2 keys per query;
 * per-keyframe query: one int32 count per keyframe).
 *
 * (1) One process per GPU.  A plain aslam_loopdb is the rank's shard (row_base = global row of its first row).
 *     aslam_loopdb_shard_export allocates the rank's exchange buffer and returns a CUDA IPC handle of
 *     aslam_loopdb_shard_handle_bytes() bytes; the ranks all-gather the handles by any means (MPI, torch.distributed,
 *     a file) and aslam_loopdb_shard_init opens the peers' buffers.  aslam_loopdb_query_global_sharded_device then is
 *     aslam_loopdb_query_global_exchange_device with those peers and an internal round counter. */
ASLAM_API size_t aslam_loopdb_shard_handle_bytes(void);
ASLAM_API int aslam_loopdb_shard_export(aslam_loopdb *db, int nranks, int max_nq, void *handle_out);
ASLAM_API int aslam_loopdb_shard_init(aslam_loopdb *db, int rank, int nranks, const void *handles /* [nranks][handle_bytes] */);
ASLAM_API int aslam_loopdb_query_global_sharded_device(aslam_loopdb *db, const uint8_t *d_q, int nq, uint64_t *d_out);
/* (2) One process, several GPUs -- the reference's shape: a single-process C++ program whose device knob is
 *     FactoryConfig::cuda_device (include/factory/PipelineFactory.hpp:21) and whose LoopClosureDetector is called from
 *     one thread (src/euroc_eval.cpp:224-242).  The group owns one context + one shard per listed device, enables peer
 *     access between all of them (cudaDeviceEnablePeerAccess; exchange buffers are plain cudaMalloc memory) and
 *     places keyframe k (insertion order) on shard k % ndevices, so the FIFO eviction of LoopClosure.cpp:27-30 keeps
 *     the shards balanced.  Same semantics as the single aslam_loopdb calls of the same name: keyframes are indexed
 *     oldest first across the whole group, global rows count from the group's creation. */
ASLAM_API int aslam_loopdb_shard_group_create(const int *devices, int ndevices, int64_t capacity_rows_per_shard,
                                              int max_keyframes, aslam_loopdb_group **out);
ASLAM_API void aslam_loopdb_shard_group_destroy(aslam_loopdb_group *g);
ASLAM_API int aslam_loopdb_shard_group_num_shards(const aslam_loopdb_group *g);
ASLAM_API int aslam_loopdb_shard_group_add(aslam_loopdb_group *g, const uint8_t *desc, int n, uint64_t keyframe_id);
ASLAM_API int aslam_loopdb_shard_group_clear(aslam_loopdb_group *g);
ASLAM_API int aslam_loopdb_shard_group_num_keyframes(const aslam_loopdb_group *g);
ASLAM_API int64_t aslam_loopdb_shard_group_num_rows(const aslam_loopdb_group *g);
/* ids[k], rows[k], shards[k] (any may be NULL) of the stored keyframes, oldest first. */
ASLAM_API int aslam_loopdb_shard_group_keyframes(const aslam_loopdb_group *g, uint64_t *ids, int32_t *rows, int32_t *shards);
/* findCandidates' counts (LoopClosure.cpp:86-95) with every shard counting for its own keyframes; the exchange is
 * the int32 count per keyframe.  good[num_keyframes], oldest first: equal to aslam_loopdb_query_segmented on one GPU. */
ASLAM_API int aslam_loopdb_shard_group_query_segmented(aslam_loopdb_group *g, const uint8_t *q, int nq, double ratio,
                                                       int32_t *good);
/* Flat top-2 (host buffers): the fused scan + exchange kernel on every device; top2[nq][2] as aslam_loopdb_query_global. */
ASLAM_API int aslam_loopdb_shard_group_query_global(aslam_loopdb_group *g, const uint8_t *q, int nq, uint64_t *top2);
/* Device `shard`'s copy of the last flat query's merged keys (all copies are identical; for parity tests). */
ASLAM_API int aslam_loopdb_shard_group_read_result(aslam_loopdb_group *g, int shard, int nq, uint64_t *top2);
ASLAM_API int aslam_loopdb_shard_group_match_keyframe(aslam_loopdb_group *g, int kf_index, const uint8_t *q, int nq,
                                                      double ratio, aslam_match *out, int *nout);

/* ------------------------------------------------------------------ pipe micro-benchmarks
 * Measured integer-pipe ceilings used as roofline denominators for the matching kernel
 * (SURVEY.md 8d asks for the POPC rate to be measured, not assumed).  which: 0 = POPC,
 * 1 = LOP3, 2 = IADD3, 3 = VIMNMX3.U16x2, 4 = HMNMX2, 5 = VIMNMX3.U16x2 + HMNMX2 interleaved,
 * 6 = IMAD, 7 = PRMT, 8 = FFMA, 9 = FFMA2, 10 = FADD2, 11 = FMUL2 (packed f32x2: two lane operations per
 * instruction and lane; the blur's floor).  Returns giga warp-lane operations per second. */
ASLAM_API int aslam_microbench(aslam_ctx *ctx, int which, double *gops);

#ifdef __cplusplus
}
#endif
#endif /* ASLAM_B200_H */
