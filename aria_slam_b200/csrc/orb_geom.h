// Level geometry of the ORB pyramid and the per-frame memory layout in HBM.
// Arithmetic follows cv::ORB_Impl (OpenCV features2d/orb.cpp) as the reference's CPU path
// reaches it (src/legacy/Frame.cpp:47); see SURVEY.md Appendix A1/A2.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace aslam {

constexpr int ORB_MAX_LEVELS = 8;
// FAST work tile: 62 x FT_OUT_H output pixels from a 64 x (FT_OUT_H+2) score block and a 70 x (FT_OUT_H+8)
// pixel block; FT_OUT_H + 2 must be a multiple of 8 (rows are dealt to 8 warps).
#ifndef ASLAM_FT_OUT_H
#define ASLAM_FT_OUT_H 62
#endif
constexpr int FT_OUT_W = 62, FT_OUT_H = ASLAM_FT_OUT_H;
constexpr int FT_SCORE_W = 64, FT_SCORE_H = FT_OUT_H + 2;
constexpr int FT_PIX_W = 70, FT_PIX_H = FT_OUT_H + 8;
static_assert(FT_SCORE_H % 8 == 0, "score rows are dealt to 8 warps");
// Measured on B200 (ms per 256 frames of 752x480, kernel alone): 30 rows 1.08, 46 rows 0.99, 62 rows 0.96, 94 rows
// 0.92 -- taller tiles amortise the per-CTA prologue and the 8 halo rows.  With 8 frame streams overlapping,
// 62 rows (30 KB of shared memory per CTA) is fastest overall: 94-row CTAs (46 KB) crowd the other streams'
// kernels out of the SMs (frames/s: 30 rows 74.9 k, 46 rows 77.2 k, 62 rows 77.6 k, 94 rows 72.0 k).
// No two 3x3-strict maxima are 8-neighbours, so a tile holds at most 31 * FT_OUT_H/2 corners.
constexpr int FT_CAP = FT_OUT_H <= 32 ? 512 : FT_OUT_H <= 64 ? 1024 : 2048;
static_assert(31 * ((FT_OUT_H + 1) / 2) <= FT_CAP, "candidate slots per tile");
// Blur work tile: BL_W x BL_H output pixels from BL_H + 6 row-pass rows held in shared memory as floats
// (BL_W * (BL_H + 6) * 4 bytes; BL_W a multiple of 8, BL_H + 6 even).
#ifndef ASLAM_BL_W
#define ASLAM_BL_W 128
#endif
#ifndef ASLAM_BL_H
#define ASLAM_BL_H 58
#endif
constexpr int BL_W = ASLAM_BL_W, BL_H = ASLAM_BL_H, BL_ROWS = BL_H + 6;
static_assert(BL_W % 8 == 0 && BL_ROWS % 2 == 0 && BL_W * BL_ROWS * 4 <= 48 * 1024, "blur tile");
// describe CTA = DESC_WARPS warps x DESC_KPW keypoints per warp (the per-lane tables live in registers)
#ifndef ASLAM_DESC_KPW
#define ASLAM_DESC_KPW 8
#endif
#ifndef ASLAM_DESC_TMA
#define ASLAM_DESC_TMA 1       // 1: patches arrive by TMA box loads (the shipped kernel); 0: round 1's LDG staging, kept for A/B runs
#endif
#ifndef ASLAM_DESC_WARPS
// TMA form: warps are independent (own rings, own mbarriers, no CTA barrier), so a CTA is ONE warp: 27 CTAs fit an SM by shared
// memory where six 4-warp CTAs gave 24 warps, and single warps slot in beside other streams' kernels (0.305 -> 0.289 ms; 2 warps 0.296)
#define ASLAM_DESC_WARPS (ASLAM_DESC_TMA ? 1 : 8)
#endif
constexpr int DESC_KPW = ASLAM_DESC_KPW, DESC_WARPS = ASLAM_DESC_WARPS;
static_assert(DESC_KPW >= 1 && DESC_KPW <= 32, "a lane owns one of the warp's keypoints");
constexpr int SEL_TIE_ROOM = 64;   // extra output rows per level for tied Harris responses

struct LevelGeom {
    int w, h, pitch;         // pitch: bytes per row, multiple of 16
    unsigned int off;        // byte offset of the level inside a frame's pyramid block
    int tiles_x, tiles_y;    // FAST tiles covering [edge, w-edge) x [edge, h-edge)
    int tile_off;            // first tile index of this level within the frame
    int blur_tile_off;       // first 128x58 blur tile of this level within the frame
    int quota;               // N_l
    int sel_off, sel_cap;    // rows of this level in the per-frame selection buffer
    int desc_block_off;      // first describe CTA (DESC_WARPS * DESC_KPW keypoints each) of this level within the frame
    int xtab_off, ytab_off;  // resize tables (level l-1 -> l), in uint32 entries
    int stab_off;            // strip table of the marching resize: 2 words per strip of rs_rows output rows (emit mask, first source row)
    int rs_rows;             // output rows per strip of the marching resize
    int rs_fast;             // 1: every strip fits the emit mask (<= 32 source rows, no two output rows with the same upper tap)
    int row_fused_end;       // blur: row-pass columns below this use fused multiply-add
    int col_fused_end;
    float sf;                // scale of the level
    float inv_sf;            // 1.f / sf (float division, as computeOrbDescriptors does)
    float kp_size;           // 31.f * sf
};

struct OrbGeom {
    int nlevels;             // as requested (quotas are spread over all of them)
    int nbuilt;              // levels 0..nbuilt-1 exist in memory (those that can hold keypoints, chained)
    int width, height, edge, fast_thr;
    int grid_cell;           // 0 = cv::ORB's retainBest; > 0 = grid-bucketed second cull with cells of that many level pixels
    int tiles_per_frame;
    int blur_tiles_per_frame;
    int desc_blocks_per_frame;
    int kp_cap;              // keypoint rows per frame in the output buffers
    unsigned int pyr_stride; // bytes per frame of pyramid (and of blurred pyramid, and of mask pyramid)
    LevelGeom lv[ORB_MAX_LEVELS];
};

// Fills g and the resize tables (entry = ofs | c1 << 16; c0 = 256 - c1 applies to p[ofs]).  False: two
// consecutive levels are more than a factor 2 apart in width after rounding (scale_factor 2 on an odd width),
// which the resize kernel's column window cannot express.
bool build_orb_geometry(int width, int height, int nfeatures, float scale_factor, int nlevels,
                        int edge, int fast_thr, int blur_mode, OrbGeom &g, std::vector<uint32_t> &tables);

}  // namespace aslam
