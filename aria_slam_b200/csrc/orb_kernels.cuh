// Device kernels of the ORB extractor (sm_100a).  One batch = `nframes` frames of one size; every
// kernel is launched once per batch (the resize once per level) with blockIdx.z / .y = frame.
//
// Per-frame HBM layout (OrbGeom): pyramid levels packed back to back, 16-B aligned rows; the
// blurred pyramid and the optional mask pyramid use the same layout.
#pragma once
#include "common.cuh"
#include "orb_geom.h"

namespace aslam {

struct OrbBuffers {
    uint8_t *pyr;          // [B][pyr_stride]
    uint8_t *blur;         // [B][pyr_stride]
    uint8_t *mask;         // [B][pyr_stride] or NULL
    const uint32_t *tabs;  // resize tables
    const uint32_t *tile_tab;  // [tiles_per_frame] FAST tile -> level | tx << 4 | ty << 16
    const uint32_t *blur_tile_tab;  // [blur_tiles_per_frame] blur tile -> level | tx << 4 | ty << 16
    uint32_t *cand;        // [B][tiles_per_frame][FT_CAP]  score<<24 | y<<12 | x
    uint32_t *cand_cnt;    // [B][tiles_per_frame]
    uint32_t *tile_base;   // [B][tiles_per_frame]  scratch of the selection kernel
    uint32_t *surv;        // [B][tiles_per_frame][FT_CAP] (only the head of each level's range is used)
    float *hresp;          // same shape as surv
    uint32_t *rank;        // same shape as surv; in-cell ranks of the grid-bucketed cull (NULL when grid_cell == 0)
    uint32_t *sel_xy;      // [B][kp_cap]  y<<12 | x, level ranges at LevelGeom::sel_off
    float *sel_resp;       // [B][kp_cap]
    int32_t *sel_cnt;      // [B][ORB_MAX_LEVELS]
    aslam_keypoint *kps;   // [B+1][kp_cap]   slot 0 = carry of the previous batch (tracking)
    uint8_t *desc;         // [B+1][kp_cap][32]
    int32_t *counts;       // [B+1]
    int32_t *overflow;     // [B+1]
    const uint2 *ori_tab;      // [8][32]: (u*mask as s8x4, mask as u8x4) of patch word (row 4*it + lane/8, lane%8)
    const float4 *pattern;     // 256 x (x0, y0, x1, y1): entry j*32 + k = comparison 8*k + j
    // TMA descriptors (CUtensorMap, 128 B each) of every level as a 3-D u8 tensor (x, y, frame): entry 2*l = the level
    // in `pyr` with a 48 x 31 box (orientation patch), entry 2*l + 1 = the level in `blur` with a 64 x 39 box (rBRIEF patch)
    const void *tmaps;
};

int launch_resize_level(aslam_ctx *ctx, const OrbGeom &g, const OrbBuffers &b, int level, int nframes, bool is_mask);
// all levels (and the mask's) of every frame in one launch, a thread-block cluster per frame; for calls with few frames
int launch_pyramid_cluster(aslam_ctx *ctx, const OrbGeom &g, const OrbBuffers &b, int nframes, bool with_mask);
bool pyramid_cluster_wanted(int nframes);
int launch_fast(aslam_ctx *ctx, const OrbGeom &g, const OrbBuffers &b, int nframes);
int launch_blur(aslam_ctx *ctx, const OrbGeom &g, const OrbBuffers &b, int nframes);
int launch_select(aslam_ctx *ctx, const OrbGeom &g, const OrbBuffers &b, int nframes);
int launch_describe(aslam_ctx *ctx, const OrbGeom &g, const OrbBuffers &b, int nframes);
// SlamPipeline::filterDynamicKeypoints: drop keypoints inside boxes of dynamic classes, keeping order.
int launch_box_filter(aslam_ctx *ctx, const OrbGeom &g, const OrbBuffers &b, int frame_slot,
                      const aslam_box *d_boxes, int nboxes, aslam_keypoint *tmp_kps, uint8_t *tmp_desc);

}  // namespace aslam
