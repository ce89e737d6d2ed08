/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the arithmetic that aria-slam's frontend hot path delegates to
 * OpenCV (cv::ORB::detectAndCompute, cv::BFMatcher(NORM_HAMMING)::knnMatch) plus the
 * reference's own candidate ranking (src/legacy/LoopClosure.cpp:72-114).  The algorithm is
 * third-party (OpenCV 4.9.0 pinned in prose by the reference, README.md:957-960; cv2 4.13.0
 * in this image) and is not under /root/reference, so parity is pinned against the live
 * cv2 wheel and the golden vectors generated from it (tests/golden/, tests/test_oracle_*).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  Nothing under aria_slam_b200/ links or imports it.
 */
#ifndef ORB_ORACLE_H
#define ORB_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_LEVELS 16

typedef struct {
    int nfeatures;       /* cv::ORB::create arg 1 (OrbCudaExtractor.cpp:35-45) */
    float scale_factor;  /* 1.2f */
    int nlevels;         /* 8 */
    int edge_threshold;  /* 31 */
    int fast_threshold;  /* 20 */
    int grid_cell;       /* 0 = cv::ORB's plain retainBest; > 0 = per-level grid bucketing of the second cull with
                          * square cells of that many level pixels (an extension named by BASELINE.json's
                          * north_star; cv::ORB and the reference have nothing like it, see .c) */
} orc_params;

typedef struct {
    float x, y, size, angle, response;
    int octave;
} orc_keypoint;          /* same layout as aria::core::KeyPoint (include/core/Types.hpp:9-15) */

typedef struct {
    int x, y, score;
} orc_corner;

/* A1 */
void orc_level_geometry(int W, int H, int nlevels, float scale_factor, int *Wl, int *Hl, float *sf);
void orc_level_quotas(int nfeatures, float scale_factor, int nlevels, int *N);
/* A2: cv::resize(..., INTER_LINEAR_EXACT) for CV_8UC1 */
void orc_resize_coeffs(int src, int dst, int *ofs, int *c1);
void orc_resize_linear_exact(const uint8_t *src, int sw, int sh, int sstride,
                             uint8_t *dst, int dw, int dh, int dstride);
void orc_mask_level(uint8_t *m, int w, int h, int stride);   /* threshold(254, TOZERO) */
/* A3: FAST-9/16 score map (0 where not a corner) and 3x3 strict NMS */
void orc_fast_score_map(const uint8_t *img, int w, int h, int stride, int thr, uint8_t *score);
int orc_fast_nms(const uint8_t *score, int w, int h, orc_corner *out, int cap);
/* Grid-bucketed variant of the second cull (extension, see .c): keep[i] = 1 for the survivors. */
int orc_grid_bucket_keep(const orc_corner *cs, const float *resp, int n, int quota, int cell, uint8_t *keep);
/* A5 / A6 */
float orc_harris(const uint8_t *img, int stride, int x, int y);
float orc_fast_atan2(float y, float x);
float orc_ic_angle(const uint8_t *img, int stride, int x, int y);
/* A8: blur_mode 0 = plain mul/add, 1 = fused multiply-add, 2 = cv2's measured hybrid (see .c) */
void orc_blur7(const uint8_t *src, int w, int h, int sstride, uint8_t *dst, int dstride, int blur_mode);
void orc_blur7_split(const uint8_t *src, int w, int h, int sstride, uint8_t *dst, int dstride, int row_fused_end, int col_fused_end);
/* A9 */
void orc_brief(const uint8_t *blurred, int stride, int xi, int yi, float angle_deg, uint8_t *desc32);

/* Whole ORB (A1..A9).  mask may be NULL (else w*h, 0 = reject).  Output order: level, then y,
 * then x.  Returns the number of keypoints (may exceed nfeatures when responses tie), or
 * -1 if cap is too small. */
int orc_orb_detect_and_compute(const uint8_t *img, int w, int h, int stride, const uint8_t *mask,
                               const orc_params *p, int blur_mode,
                               orc_keypoint *kps, uint8_t *desc, int cap);

/* A10: per query the two smallest (distance, train index) pairs, lexicographic.
 * out_idx/out_dist are nq x 2; missing neighbours (nt < 2) get idx -1, dist -1. */
void orc_knn2_hamming(const uint8_t *q, int nq, const uint8_t *t, int nt, int *out_idx, int *out_dist);
/* Lowe ratio on integer distances: d0 < (num/den) * d1 as num*d1 > den*d0 ... see .c */
int orc_ratio_pass(int d0, int d1, int num, int den);
/* LoopClosure.cpp:72-114 over a flat database split into segments (keyframes):
 * seg_start has nseg+1 entries.  good[s] = number of queries passing the 0.7 ratio test in
 * segment s. */
void orc_loop_segment_counts(const uint8_t *q, int nq, const uint8_t *db, const int *seg_start,
                             int nseg, int num, int den, int *good);

#ifdef __cplusplus
}
#endif
#endif
