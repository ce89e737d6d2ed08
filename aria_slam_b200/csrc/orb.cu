// ORB extractor handle: memory plan in HBM, launch sequence, host/device entry points and the
// batched extract+match ("track") path.
//
// Stands in for OrbCudaExtractor (src/adapters/gpu/OrbCudaExtractor.cpp) and the legacy Frame
// constructors (src/legacy/Frame.cpp:6-49) of the reference.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <vector>

#include "match.cuh"
#include "orb_kernels.cuh"

using namespace aslam;

static const int8_t k_brief_pattern[256][4] = {
#include "brief_pattern_31.inc"
};

struct aslam_orb {
    aslam_ctx *ctx = nullptr;
    aslam_orb_params prm{};
    int width = 0, height = 0, max_batch = 0;
    OrbGeom g{};
    OrbBuffers b{};
    uint32_t *d_tabs = nullptr;
    uint32_t *d_tile_tab = nullptr;
    float4 *d_pattern = nullptr;
    uint2 *d_ori_tab = nullptr;
    void *d_tmaps = nullptr;               // [ORB_MAX_LEVELS][2] CUtensorMap: (pyr, 48 x 31 box), (blur, 64 x 39 box)
    // tracking
    aslam_match *d_matches = nullptr;      // [B][kp_cap]
    int32_t *d_match_counts = nullptr;     // [B]
    KnnProblem *d_probs = nullptr;
    uint2 *d_partial = nullptr;
    int track_splits = 1;
    bool have_carry = false;
    // box filter scratch
    aslam_box *d_boxes = nullptr; int boxes_cap = 0;
    aslam_keypoint *d_tmp_kps = nullptr; uint8_t *d_tmp_desc = nullptr;
    // staging for host images when row_stride != pitch of level 0 is handled by cudaMemcpy2DAsync
    int pending = 0;        // frames of the call awaiting sync (0 = none)
    bool pending_track = false;
    // pinned result staging
    int32_t *h_counts = nullptr;   // [B+1] counts, [B+1] overflow, [B] match counts
    // aslam_orb_retain_descriptors: two ring slots of kp_cap rows, the first frame's count at the last sync
    uint8_t *d_retained[2] = {nullptr, nullptr};
    int retain_next = 0;
    int last_count = -1;           // -1: nothing synced yet (or already retained)
    // the extraction chain of (n frames, with / without masks) as an instantiated CUDA graph: one launch per call
    // instead of 12 (memset, 7 resize levels, FAST, blur, selection, describe); dropped whenever the buffers are re-planned
    struct ChainGraph { int n; bool with_mask; cudaGraphExec_t exec; int launches; };
    std::vector<ChainGraph> chains;
    // a second stream and two events, used only while a small call's chain is being captured: the blur depends on the
    // pyramid alone, so in the graph it is a branch beside FAST -> selection and joins before describe
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool no_cluster = false;       // a cluster launch was refused once on this device
};

static void orb_drop_graphs(aslam_orb *o)
{
    for (auto &c : o->chains) cudaGraphExecDestroy(c.exec);
    o->chains.clear();
}

static void orb_free_buffers(aslam_orb *o)
{
    orb_drop_graphs(o);
    cudaFree(o->b.pyr); cudaFree(o->b.blur); cudaFree(o->b.mask); cudaFree(o->b.cand); cudaFree(o->b.cand_cnt);
    cudaFree(o->b.tile_base); cudaFree(o->b.surv); cudaFree(o->b.hresp); cudaFree(o->b.rank); cudaFree(o->b.sel_xy);
    cudaFree(o->b.sel_resp); cudaFree(o->b.sel_cnt); cudaFree(o->b.kps); cudaFree(o->b.desc);
    cudaFree(o->b.counts); cudaFree(o->b.overflow); cudaFree(o->d_tabs); cudaFree(o->d_tile_tab);
    cudaFree(o->d_matches); cudaFree(o->d_match_counts); cudaFree(o->d_probs); cudaFree(o->d_partial);
    cudaFree(o->d_tmp_kps); cudaFree(o->d_tmp_desc);
    cudaFree(o->d_retained[0]); cudaFree(o->d_retained[1]);
    cudaFree(o->d_tmaps); o->d_tmaps = nullptr;
    o->d_retained[0] = o->d_retained[1] = nullptr; o->last_count = -1;
    if (o->h_counts) cudaFreeHost(o->h_counts);
    o->b = OrbBuffers();
    o->d_tabs = nullptr; o->d_tile_tab = nullptr; o->d_matches = nullptr; o->d_match_counts = nullptr; o->d_probs = nullptr;
    o->d_partial = nullptr; o->d_tmp_kps = nullptr; o->d_tmp_desc = nullptr; o->h_counts = nullptr;
}

__global__ void fill_track_problems_kernel(KnnProblem *probs, const uint8_t *desc, const int32_t *counts,
                                           int n, int splits, int kp_cap, int rows_per)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * splits) return;
    const int f = i / splits, s = i - f * splits;
    // src/main.cpp:158-161, src/euroc_eval.cpp:167-169: knnMatch(prev_frame, current_frame) -- query = the
    // predecessor, train = the new frame, so query_idx indexes the previous frame's keypoints.
    KnnProblem P;
    P.q = desc + (size_t)f * kp_cap * 32;            // predecessor of frame f (slot 0 = carry of the last call)
    P.t = desc + (size_t)(f + 1) * kp_cap * 32;      // frame f of the batch lives in slot f+1
    P.nq_dev = counts + f; P.nt_dev = counts + f + 1;
    P.nq = kp_cap; P.nt = kp_cap;
    P.t_lo = min(s * rows_per, kp_cap);
    P.t_hi = min((s + 1) * rows_per, kp_cap);
    probs[i] = P;
}

// TMA descriptors of the describe kernel: every built level of `pyr` and `blur` as a 3-D u8 tensor (x = byte in the padded
// row, y = row, z = frame of the batch).  The driver entry point comes through the runtime (no link against libcuda).
static int orb_build_tensor_maps(aslam_orb *o)
{
#if ASLAM_DESC_TMA
    static PFN_cuTensorMapEncodeTiled encode = [] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
        cudaGetLastError();
        return (PFN_cuTensorMapEncodeTiled)fn;
    }();
    if (!encode) return fail(ASLAM_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver (TMA needs CUDA 12 on sm_90+)");
    const OrbGeom &g = o->g;
    alignas(64) CUtensorMap maps[ORB_MAX_LEVELS * 2];
    memset(maps, 0, sizeof(maps));
    for (int l = 0; l < g.nbuilt; ++l) {
        const LevelGeom &L = g.lv[l];
        for (int which = 0; which < 2; ++which) {
            uint8_t *base = (which ? o->b.blur : o->b.pyr) + L.off;
            const cuuint64_t dims[3] = {(cuuint64_t)L.pitch, (cuuint64_t)L.h, (cuuint64_t)o->max_batch};
            const cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)g.pyr_stride};      // bytes, dimensions 1 and 2
            const cuuint32_t box[3] = {which ? 64u : 48u, which ? 39u : 31u, 1u};      // OPB_W x OPB_H, BPB_W x BPB_H of describe_kernel
            const cuuint32_t estr[3] = {1u, 1u, 1u};
            const CUresult r = encode(&maps[2 * l + which], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(ASLAM_E_CUDA, "cuTensorMapEncodeTiled(level %d, %s) -> CUresult %d", l, which ? "blur" : "pyr", (int)r);
        }
    }
    ASLAM_CUDA(cudaMalloc(&o->d_tmaps, sizeof(maps)));
    ASLAM_CUDA(cudaMemcpyAsync(o->d_tmaps, maps, sizeof(maps), cudaMemcpyHostToDevice, o->ctx->stream));
    ASLAM_CUDA(cudaStreamSynchronize(o->ctx->stream));      // maps is a local
    o->b.tmaps = o->d_tmaps;
#endif
    return ASLAM_OK;
}

static int orb_plan(aslam_orb *o)
{
    aslam_ctx *ctx = o->ctx;
    std::vector<uint32_t> tabs;
    if (!build_orb_geometry(o->width, o->height, o->prm.nfeatures, o->prm.scale_factor, o->prm.nlevels,
                            o->prm.edge_threshold, o->prm.fast_threshold, o->prm.blur_mode, o->g, tabs))
        return fail(ASLAM_E_INVALID, "scale_factor %.3f on a %dx%d frame puts two levels more than a factor 2 apart",
                    (double)o->prm.scale_factor, o->width, o->height);
    o->g.grid_cell = o->prm.grid_cell;
    const OrbGeom &g = o->g;
    const size_t B = (size_t)o->max_batch;
    const size_t tiles = (size_t)(g.tiles_per_frame > 0 ? g.tiles_per_frame : 1);
    OrbBuffers &b = o->b;
    ASLAM_CUDA(cudaMalloc(&b.pyr, B * g.pyr_stride + 256));
    ASLAM_CUDA(cudaMalloc(&b.blur, B * g.pyr_stride + 256));
    ASLAM_CUDA(cudaMalloc(&b.cand, sizeof(uint32_t) * B * tiles * FT_CAP));
    ASLAM_CUDA(cudaMalloc(&b.cand_cnt, sizeof(uint32_t) * B * tiles));
    ASLAM_CUDA(cudaMalloc(&b.tile_base, sizeof(uint32_t) * B * tiles));
    ASLAM_CUDA(cudaMalloc(&b.surv, sizeof(uint32_t) * B * tiles * FT_CAP));
    ASLAM_CUDA(cudaMalloc(&b.hresp, sizeof(float) * B * tiles * FT_CAP));
    if (g.grid_cell > 0) ASLAM_CUDA(cudaMalloc(&b.rank, sizeof(uint32_t) * B * tiles * FT_CAP));
    ASLAM_CUDA(cudaMalloc(&b.sel_xy, sizeof(uint32_t) * B * g.kp_cap));
    ASLAM_CUDA(cudaMalloc(&b.sel_resp, sizeof(float) * B * g.kp_cap));
    ASLAM_CUDA(cudaMalloc(&b.sel_cnt, sizeof(int32_t) * B * ORB_MAX_LEVELS));
    ASLAM_CUDA(cudaMalloc(&b.kps, sizeof(aslam_keypoint) * (B + 1) * g.kp_cap));
    ASLAM_CUDA(cudaMalloc(&b.desc, (B + 1) * g.kp_cap * 32));
    ASLAM_CUDA(cudaMalloc(&b.counts, sizeof(int32_t) * (B + 1)));
    ASLAM_CUDA(cudaMalloc(&b.overflow, sizeof(int32_t) * (B + 1)));
    ASLAM_CUDA(cudaMemsetAsync(b.counts, 0, sizeof(int32_t) * (B + 1), ctx->stream));
    ASLAM_CUDA(cudaMemsetAsync(b.overflow, 0, sizeof(int32_t) * (B + 1), ctx->stream));
    ASLAM_CUDA(cudaMemsetAsync(b.sel_cnt, 0, sizeof(int32_t) * B * ORB_MAX_LEVELS, ctx->stream));
    ASLAM_CUDA(cudaMalloc(&o->d_tabs, sizeof(uint32_t) * (tabs.size() + 4)));
    if (!tabs.empty())
        ASLAM_CUDA(cudaMemcpyAsync(o->d_tabs, tabs.data(), sizeof(uint32_t) * tabs.size(), cudaMemcpyHostToDevice, ctx->stream));
    b.tabs = o->d_tabs;
    {
        // FAST and blur work tiles -> level | tx << 4 | ty << 16 (one table: FAST tiles first, blur tiles behind them)
        const size_t btiles = (size_t)(g.blur_tiles_per_frame > 0 ? g.blur_tiles_per_frame : 1);
        std::vector<uint32_t> tt(tiles + btiles, 0u);
        for (int l = 0; l < g.nbuilt; ++l) {
            const LevelGeom &L = g.lv[l];
            for (int t = 0; t < L.tiles_x * L.tiles_y; ++t)
                tt[(size_t)L.tile_off + t] = (uint32_t)l | ((uint32_t)(t % L.tiles_x) << 4) | ((uint32_t)(t / L.tiles_x) << 16);
            if (L.tiles_x * L.tiles_y == 0) continue;               // levels without keypoints are not blurred
            const int btx = (L.w + BL_W - 1) / BL_W, bty = (L.h + BL_H - 1) / BL_H;
            for (int t = 0; t < btx * bty; ++t)
                tt[tiles + (size_t)L.blur_tile_off + t] = (uint32_t)l | ((uint32_t)(t % btx) << 4) | ((uint32_t)(t / btx) << 16);
        }
        ASLAM_CUDA(cudaMalloc(&o->d_tile_tab, sizeof(uint32_t) * tt.size()));
        ASLAM_CUDA(cudaMemcpyAsync(o->d_tile_tab, tt.data(), sizeof(uint32_t) * tt.size(), cudaMemcpyHostToDevice, ctx->stream));
        ASLAM_CUDA(cudaStreamSynchronize(ctx->stream));      // tt is a local
        b.tile_tab = o->d_tile_tab;
        b.blur_tile_tab = o->d_tile_tab + tiles;
    }
    b.pattern = o->d_pattern;
    b.ori_tab = o->d_ori_tab;
    {
        const int rc = orb_build_tensor_maps(o);
        if (rc) return rc;
    }
    // tracking scratch
    o->track_splits = choose_splits(ctx, o->prm.nfeatures, o->prm.nfeatures, o->max_batch);     // expected rows, not the capacity
    ASLAM_CUDA(cudaMalloc(&o->d_matches, sizeof(aslam_match) * B * g.kp_cap));
    ASLAM_CUDA(cudaMalloc(&o->d_match_counts, sizeof(int32_t) * B));
    ASLAM_CUDA(cudaMalloc(&o->d_probs, sizeof(KnnProblem) * B * o->track_splits));
    ASLAM_CUDA(cudaMalloc(&o->d_partial, sizeof(uint2) * B * o->track_splits * g.kp_cap));
    ASLAM_CUDA(cudaMalloc(&o->d_tmp_kps, sizeof(aslam_keypoint) * g.kp_cap));
    ASLAM_CUDA(cudaMalloc(&o->d_tmp_desc, (size_t)g.kp_cap * 32));
    ASLAM_CUDA(cudaMallocHost(&o->h_counts, sizeof(int32_t) * (3 * B + 2)));
    ASLAM_CUDA(cudaStreamSynchronize(ctx->stream));   // tabs is a local
    o->have_carry = false;
    return ASLAM_OK;
}

extern "C" {

void aslam_orb_default_params(aslam_orb_params *p)
{
    if (!p) return;
    p->nfeatures = 1000; p->scale_factor = 1.2f; p->nlevels = 8; p->edge_threshold = 31;
    p->first_level = 0; p->wta_k = 2; p->score_type = 0; p->patch_size = 31; p->fast_threshold = 20;
    p->blur_mode = 2;
    p->grid_cell = 0;
}

static int check_params(const aslam_orb_params *p, int width, int height, int max_batch)
{
    ASLAM_REQUIRE(p->nfeatures >= 0, "nfeatures must be >= 0");
    ASLAM_REQUIRE(p->nlevels >= 1 && p->nlevels <= ORB_MAX_LEVELS, "nlevels must be 1..8");
    ASLAM_REQUIRE(p->scale_factor > 1.0f && p->scale_factor <= 2.0f, "scale_factor must be in (1, 2]");
    ASLAM_REQUIRE(p->edge_threshold >= 19 && p->edge_threshold <= 1024, "edge_threshold must be >= 19 (rBRIEF reach)");
    ASLAM_REQUIRE(p->first_level == 0 && p->wta_k == 2 && p->score_type == 0 && p->patch_size == 31,
                  "only first_level 0, WTA_K 2, HARRIS_SCORE, patch 31 are implemented (the reference's values)");
    ASLAM_REQUIRE(p->fast_threshold >= 1 && p->fast_threshold <= 254, "fast_threshold must be 1..254");
    ASLAM_REQUIRE(p->blur_mode >= 0 && p->blur_mode <= 2, "blur_mode must be 0..2");
    ASLAM_REQUIRE(p->grid_cell == 0 || (p->grid_cell >= 8 && p->grid_cell <= 4096), "grid_cell must be 0 (off) or 8..4096");
    ASLAM_REQUIRE(width >= 1 && height >= 1 && width <= 4095 && height <= 4095, "frame size must be 1..4095 per side");
    ASLAM_REQUIRE(max_batch >= 1 && max_batch <= 4096, "max_batch must be 1..4096");
    return ASLAM_OK;
}

int aslam_orb_create(aslam_ctx *ctx, const aslam_orb_params *p, int width, int height, int max_batch, aslam_orb **out)
{
    ASLAM_REQUIRE(ctx && p && out, "NULL argument");
    *out = nullptr;
    int rc = check_params(p, width, height, max_batch);
    if (rc) return rc;
    ASLAM_CUDA(cudaSetDevice(ctx->device));
    aslam_orb *o = new aslam_orb();
    o->ctx = ctx; o->prm = *p; o->width = width; o->height = height; o->max_batch = max_batch;
    // entry j*32 + k = comparison 8*k + j as (x0, y0, x1, y1): one coalesced 16-byte load per j per warp
    float4 packed[256];
    for (int i = 0; i < 256; ++i)
        packed[(i & 7) * 32 + (i >> 3)] = make_float4((float)k_brief_pattern[i][0], (float)k_brief_pattern[i][1],
                                                      (float)k_brief_pattern[i][2], (float)k_brief_pattern[i][3]);
    cudaError_t e = cudaMalloc(&o->d_pattern, sizeof(packed));
    if (e == cudaSuccess) e = cudaMemcpy(o->d_pattern, packed, sizeof(packed), cudaMemcpyHostToDevice);
    // Orientation weights (ICAngles disc, radius 15) of the patch-aligned word (row, word): byte bt is column
    // u = 4*word + bt - 15 of row v = row - 15; entry it*32 + lane <-> row 4*it + lane/8, word lane%8.
    uint2 ori[8 * 32];
    static const int umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
    for (int it = 0; it < 8; ++it)
        for (int ln = 0; ln < 32; ++ln) {
            const int row = 4 * it + (ln >> 3), word = ln & 7, v = row - 15;
            uint32_t wu = 0, wm = 0;
            for (int bt = 0; bt < 4 && row < 31; ++bt) {
                const int u = 4 * word + bt - 15;
                if (u <= 15 && (u < 0 ? -u : u) <= umax[v < 0 ? -v : v]) { wu |= (uint32_t)(uint8_t)(int8_t)u << (8 * bt); wm |= 1u << (8 * bt); }
            }
            ori[it * 32 + ln] = make_uint2(wu, wm);
        }
    if (e == cudaSuccess) e = cudaMalloc(&o->d_ori_tab, sizeof(ori));
    if (e == cudaSuccess) e = cudaMemcpy(o->d_ori_tab, ori, sizeof(ori), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { delete o; return fail(ASLAM_E_CUDA, "pattern upload -> %s", cudaGetErrorString(e)); }
    rc = orb_plan(o);
    if (rc) { orb_free_buffers(o); cudaFree(o->d_pattern); cudaFree(o->d_ori_tab); delete o; return rc; }
    *out = o;
    return ASLAM_OK;
}

void aslam_orb_destroy(aslam_orb *o)
{
    if (!o) return;
    cudaSetDevice(o->ctx->device);
    cudaStreamSynchronize(o->ctx->stream);
    orb_free_buffers(o);
    cudaFree(o->d_pattern);
    cudaFree(o->d_ori_tab);
    cudaFree(o->d_boxes);
    if (o->ev_fork) cudaEventDestroy(o->ev_fork);
    if (o->ev_join) cudaEventDestroy(o->ev_join);
    if (o->side) cudaStreamDestroy(o->side);
    delete o;
}

int aslam_orb_set_max_features(aslam_orb *o, int nfeatures)
{
    ASLAM_REQUIRE(o, "orb is NULL");
    ASLAM_REQUIRE(nfeatures >= 0, "nfeatures must be >= 0");
    if (o->pending) return fail(ASLAM_E_STATE, "set_max_features with an extract pending");
    ASLAM_CUDA(cudaSetDevice(o->ctx->device));
    ASLAM_CUDA(cudaStreamSynchronize(o->ctx->stream));
    // OrbCudaExtractor.cpp:212-216 re-creates the detector with only nfeatures set: defaults elsewhere.
    // blur_mode and grid_cell are this library's own knobs, not cv::ORB::create arguments: they stay.
    const aslam_orb_params before = o->prm;
    aslam_orb_default_params(&o->prm);
    o->prm.nfeatures = nfeatures;
    o->prm.blur_mode = before.blur_mode;
    o->prm.grid_cell = before.grid_cell;
    orb_free_buffers(o);
    const int rc = orb_plan(o);
    if (rc == ASLAM_OK) return rc;
    // the new plan could not be allocated: go back to the previous one; if even that fails the handle refuses all work
    char msg[512];
    snprintf(msg, sizeof(msg), "%s", aslam_last_error());
    orb_free_buffers(o);
    o->prm = before;
    if (orb_plan(o) != ASLAM_OK) { orb_free_buffers(o); o->max_batch = 0; }
    return fail(rc, "%s", msg);
}

int aslam_orb_get_max_features(const aslam_orb *o) { return o ? o->prm.nfeatures : 0; }
int aslam_orb_capacity(const aslam_orb *o) { return o ? o->g.kp_cap : 0; }

int aslam_orb_level_size(const aslam_orb *o, int level, int *w, int *h)
{
    ASLAM_REQUIRE(o && w && h, "NULL argument");
    ASLAM_REQUIRE(level >= 0 && level < o->g.nlevels, "level out of range");
    *w = o->g.lv[level].w; *h = o->g.lv[level].h;
    return ASLAM_OK;
}

}  // extern "C"

// Enqueue the whole extraction of n frames whose level-0 pixels (and masks) are already in place.
static int orb_enqueue_launches(aslam_orb *o, int n, bool with_mask, bool fork_blur = false)
{
    aslam_ctx *ctx = o->ctx;
    const OrbGeom &g = o->g;
    OrbBuffers b = o->b;
    if (!with_mask) b.mask = nullptr;
    int rc;
    ASLAM_CUDA(cudaMemsetAsync(o->b.overflow + 1, 0, sizeof(int32_t) * n, ctx->stream));
    bool levels_done = false;
    if (g.nbuilt > 1 && pyramid_cluster_wanted(n) && !o->no_cluster) {
        if (launch_pyramid_cluster(ctx, g, b, n, with_mask) == ASLAM_OK) levels_done = true;
        else { cudaGetLastError(); o->no_cluster = true; }      // no cluster launches here (e.g. a partitioned GPU): level by level from now on
    }
    if (!levels_done) {
        for (int l = 1; l < g.nbuilt; ++l) {
            if ((rc = launch_resize_level(ctx, g, b, l, n, false))) return rc;
            if (with_mask && (rc = launch_resize_level(ctx, g, b, l, n, true))) return rc;
        }
    }
    if (fork_blur) {
        cudaStream_t main_stream = ctx->stream;
        ASLAM_CUDA(cudaEventRecord(o->ev_fork, main_stream));
        ASLAM_CUDA(cudaStreamWaitEvent(o->side, o->ev_fork, 0));
        ctx->stream = o->side;
        rc = launch_blur(ctx, g, b, n);
        ctx->stream = main_stream;
        if (rc) return rc;
        ASLAM_CUDA(cudaEventRecord(o->ev_join, o->side));
        if ((rc = launch_fast(ctx, g, b, n))) return rc;
        if ((rc = launch_select(ctx, g, b, n))) return rc;
        ASLAM_CUDA(cudaStreamWaitEvent(main_stream, o->ev_join, 0));
        return launch_describe(ctx, g, b, n);
    }
    if ((rc = launch_fast(ctx, g, b, n))) return rc;
    if ((rc = launch_blur(ctx, g, b, n))) return rc;
    if ((rc = launch_select(ctx, g, b, n))) return rc;
    if ((rc = launch_describe(ctx, g, b, n))) return rc;
    return ASLAM_OK;
}

// The chain's launches depend only on (n, masks or not) once the handle is planned -- every kernel argument is a value
// or a pointer into the handle's own buffers -- so the first call with a given shape is captured into a CUDA graph and
// later calls replay it: per-frame callers (the reference runs one frame per call) pay one launch instead of twelve.
// Not used while per-kernel profiling is on (the event pairs belong between the launches), on the legacy default stream
// (it cannot be captured), or with ASLAM_NO_GRAPH=1 (A/B runs).
static int orb_enqueue(aslam_orb *o, int n, bool with_mask)
{
    aslam_ctx *ctx = o->ctx;
    static const bool no_graph = [] { const char *e = getenv("ASLAM_NO_GRAPH"); return e && atoi(e) != 0; }();
    if (no_graph || ctx->prof_on || ctx->stream == nullptr || ctx->stream == cudaStreamLegacy || ctx->stream == cudaStreamPerThread)
        return orb_enqueue_launches(o, n, with_mask);
    for (auto &c : o->chains)
        if (c.n == n && c.with_mask == with_mask) {
            ASLAM_CUDA(cudaGraphLaunch(c.exec, ctx->stream));
            aslam::count_launch(c.launches);
            return ASLAM_OK;
        }
    if (with_mask && !o->b.mask) return orb_enqueue_launches(o, n, with_mask);
    // calls that do not fill the GPU for long are latency-bound: their blur runs as a branch of the graph beside
    // FAST -> selection.  Measured (752 x 480, device input, wall per call, branch off / on): 4 frames 99.2 / 96.1 us,
    // 8: 126.5 / 120.6, 32: 289 / 280, 64: 497 / 483, 256: no difference; bench.py's 8 concurrent streams: no difference.
    static const int fork_max = [] { const char *e = getenv("ASLAM_FORK_BLUR_MAX"); return e ? atoi(e) : 64; }();
    bool fork = n <= fork_max;
    if (fork && !o->side) {
        if (cudaStreamCreateWithFlags(&o->side, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&o->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&o->ev_join, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            fork = false;
        }
    }
    cudaGraph_t graph = nullptr;
    const uint64_t before = aslam::g_launches.load(std::memory_order_relaxed);
    if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
        cudaGetLastError();
        return orb_enqueue_launches(o, n, with_mask);
    }
    const int rc = orb_enqueue_launches(o, n, with_mask, fork && o->side && o->ev_fork && o->ev_join);
    const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
    const int launches = (int)(aslam::g_launches.load(std::memory_order_relaxed) - before);
    aslam::g_launches.fetch_sub((uint64_t)launches, std::memory_order_relaxed);      // recorded, not launched
    cudaGraphExec_t exec = nullptr;
    if (rc != ASLAM_OK || ce != cudaSuccess || !graph || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        // whatever went wrong while recording (a refused launch invalidates the capture and every launch after it): issue
        // the chain directly; a genuine failure shows up again there and is reported
        return orb_enqueue_launches(o, n, with_mask);
    }
    cudaGraphDestroy(graph);
    if (o->chains.size() >= 4) { cudaGraphExecDestroy(o->chains.front().exec); o->chains.erase(o->chains.begin()); }
    o->chains.push_back({n, with_mask, exec, launches});
    ASLAM_CUDA(cudaGraphLaunch(exec, ctx->stream));
    aslam::count_launch(launches);
    return ASLAM_OK;
}

static int orb_upload(aslam_orb *o, uint8_t *dst_base, const uint8_t *src, int n, size_t frame_stride,
                      int row_stride, cudaMemcpyKind kind)
{
    const OrbGeom &g = o->g;
    const LevelGeom &L0 = g.lv[0];
    if (g.nbuilt == 0) return ASLAM_OK;
    // level 0 of frame i sits at dst_base + i*pyr_stride (off 0); one 2-D copy per frame keeps
    // the caller's row stride and ours independent.  A single 3-D-shaped copy is possible when
    // both strides are uniform, which cudaMemcpy2DAsync expresses when rows are contiguous.
    if ((size_t)row_stride == (size_t)L0.pitch && frame_stride == (size_t)L0.pitch * L0.h && g.pyr_stride == frame_stride) {
        ASLAM_CUDA(cudaMemcpyAsync(dst_base, src, frame_stride * n, kind, o->ctx->stream));
        return ASLAM_OK;
    }
    if ((size_t)row_stride == (size_t)L0.pitch && frame_stride >= (size_t)L0.pitch * L0.h) {
        // whole frames as "rows" of a 2-D copy
        ASLAM_CUDA(cudaMemcpy2DAsync(dst_base, g.pyr_stride, src, frame_stride, (size_t)L0.pitch * L0.h, n, kind, o->ctx->stream));
        return ASLAM_OK;
    }
    for (int i = 0; i < n; ++i)
        ASLAM_CUDA(cudaMemcpy2DAsync(dst_base + (size_t)i * g.pyr_stride, L0.pitch, src + (size_t)i * frame_stride,
                                     row_stride, L0.w, L0.h, kind, o->ctx->stream));
    return ASLAM_OK;
}

static int orb_prepare_mask(aslam_orb *o)
{
    if (!o->b.mask) ASLAM_CUDA(cudaMalloc(&o->b.mask, (size_t)o->max_batch * o->g.pyr_stride + 256));
    return ASLAM_OK;
}

static int orb_track_enqueue(aslam_orb *o, int n, float ratio);

static int orb_extract_common(aslam_orb *o, const uint8_t *images, int n, size_t frame_stride, int row_stride,
                              const uint8_t *masks, cudaMemcpyKind kind)
{
    ASLAM_REQUIRE(o && images, "NULL argument");
    ASLAM_REQUIRE(n >= 1 && n <= o->max_batch, "n must be 1..max_batch");
    ASLAM_REQUIRE(row_stride >= o->width, "row_stride smaller than width");
    ASLAM_REQUIRE(n == 1 || frame_stride >= (size_t)row_stride * (o->height - 1) + o->width, "frame_stride too small");
    if (o->pending) return fail(ASLAM_E_STATE, "an extract is already pending on this handle; call sync first");
    ASLAM_CUDA(cudaSetDevice(o->ctx->device));
    int rc;
    if ((rc = orb_upload(o, o->b.pyr, images, n, frame_stride, row_stride, kind))) return rc;
    if (masks) {
        if ((rc = orb_prepare_mask(o))) return rc;
        if ((rc = orb_upload(o, o->b.mask, masks, n, (size_t)o->width * o->height, o->width, kind))) return rc;
    }
    if ((rc = orb_enqueue(o, n, masks != nullptr))) return rc;
    o->pending = n;
    o->pending_track = false;
    return ASLAM_OK;
}

extern "C" {

int aslam_orb_extract_async(aslam_orb *o, const uint8_t *image, int width, int height, int row_stride,
                            const uint8_t *mask, const aslam_box *boxes, int nboxes)
{
    ASLAM_REQUIRE(o, "orb is NULL");
    ASLAM_REQUIRE(width == o->width && height == o->height, "frame size differs from the handle's");
    ASLAM_REQUIRE(nboxes >= 0 && (nboxes == 0 || boxes), "bad boxes");
    int rc = orb_extract_common(o, image, 1, 0, row_stride, mask, cudaMemcpyHostToDevice);
    if (rc) return rc;
    if (nboxes > 0) {
        rc = [&]() -> int {
            aslam_ctx *ctx = o->ctx;
            if (nboxes > o->boxes_cap) {
                if (o->d_boxes) ASLAM_CUDA(cudaFree(o->d_boxes));
                o->d_boxes = nullptr; o->boxes_cap = 0;
                ASLAM_CUDA(cudaMalloc(&o->d_boxes, sizeof(aslam_box) * (size_t)(nboxes + 64)));
                o->boxes_cap = nboxes + 64;
            }
            ASLAM_CUDA(cudaMemcpyAsync(o->d_boxes, boxes, sizeof(aslam_box) * (size_t)nboxes, cudaMemcpyHostToDevice, ctx->stream));
            return launch_box_filter(ctx, o->g, o->b, 1, o->d_boxes, nboxes, o->d_tmp_kps, o->d_tmp_desc);
        }();
        if (rc) o->pending = 0;         // the call failed as a whole: nothing is pending
    }
    return rc;
}

int aslam_orb_extract_batch_async(aslam_orb *o, const uint8_t *images, int n, size_t frame_stride, int row_stride,
                                  const uint8_t *masks)
{
    return orb_extract_common(o, images, n, frame_stride, row_stride, masks, cudaMemcpyHostToDevice);
}

int aslam_orb_extract_batch_device(aslam_orb *o, const uint8_t *d_images, int n, size_t frame_stride, int row_stride,
                                   const uint8_t *d_masks)
{
    return orb_extract_common(o, d_images, n, frame_stride, row_stride, d_masks, cudaMemcpyDeviceToDevice);
}

int aslam_orb_device_results(aslam_orb *o, const aslam_keypoint **d_kps, const uint8_t **d_desc, const int32_t **d_counts)
{
    ASLAM_REQUIRE(o, "orb is NULL");
    if (d_kps) *d_kps = o->b.kps + o->g.kp_cap;
    if (d_desc) *d_desc = o->b.desc + (size_t)o->g.kp_cap * 32;
    if (d_counts) *d_counts = o->b.counts + 1;
    o->pending = 0;     // the caller reads in stream order on the device
    return ASLAM_OK;
}

static int orb_collect(aslam_orb *o, aslam_keypoint *kps, uint8_t *desc, int32_t *counts,
                       aslam_match *matches, int32_t *match_counts)
{
    ASLAM_REQUIRE(o, "orb is NULL");
    if (!o->pending) return fail(ASLAM_E_STATE, "sync without a pending extract");
    aslam_ctx *ctx = o->ctx;
    const int n = o->pending;
    // Whatever happens below, the extract is no longer pending: a CUDA error here must not leave the handle refusing
    // every later extract with "already pending" (the failed results are lost, the handle stays usable).
    o->pending = 0;
    o->last_count = -1;
    ASLAM_CUDA(cudaSetDevice(ctx->device));
    const size_t cap = (size_t)o->g.kp_cap;
    const size_t B = (size_t)o->max_batch;
    int32_t *h_cnt = o->h_counts, *h_ovf = o->h_counts + (B + 1), *h_mc = o->h_counts + 2 * (B + 1);
    ASLAM_CUDA(cudaMemcpyAsync(h_cnt, o->b.counts + 1, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    ASLAM_CUDA(cudaMemcpyAsync(h_ovf, o->b.overflow + 1, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    // Per frame only the first `head` = nfeatures rows are copied up front: a frame holds more than that only when
    // responses tie at a level's cut (rows up to the capacity); those tails are fetched in a second pass
    // below, which almost never runs.  Rows beyond a frame's count are not part of the result.
    const size_t head = std::min(cap, (size_t)std::max(o->prm.nfeatures, 1));
    auto copy_rows = [&](void *dst, const void *src, size_t row_bytes, size_t first, size_t nrows) -> cudaError_t {
        return cudaMemcpy2DAsync(static_cast<uint8_t *>(dst) + first * row_bytes, cap * row_bytes,
                                 static_cast<const uint8_t *>(src) + first * row_bytes, cap * row_bytes,
                                 nrows * row_bytes, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream);
    };
    const bool with_matches = o->pending_track && matches;
    if (kps) ASLAM_CUDA(copy_rows(kps, o->b.kps + cap, sizeof(aslam_keypoint), 0, head));
    if (desc) ASLAM_CUDA(copy_rows(desc, o->b.desc + cap * 32, 32, 0, head));
    if (o->pending_track) {
        ASLAM_CUDA(cudaMemcpyAsync(h_mc, o->d_match_counts, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
        if (with_matches) ASLAM_CUDA(copy_rows(matches, o->d_matches, sizeof(aslam_match), 0, head));
    }
    ASLAM_CUDA(cudaStreamSynchronize(ctx->stream));
    if (head < cap) {
        int most = 0, most_m = 0;
        for (int i = 0; i < n; ++i) { most = std::max(most, h_cnt[i]); if (o->pending_track) most_m = std::max(most_m, h_mc[i]); }
        bool again = false;
        if ((size_t)most > head) {
            if (kps) ASLAM_CUDA(copy_rows(kps, o->b.kps + cap, sizeof(aslam_keypoint), head, (size_t)most - head));
            if (desc) ASLAM_CUDA(copy_rows(desc, o->b.desc + cap * 32, 32, head, (size_t)most - head));
            again = kps || desc;
        }
        if (with_matches && (size_t)most_m > head) {
            ASLAM_CUDA(copy_rows(matches, o->d_matches, sizeof(aslam_match), head, (size_t)most_m - head));
            again = true;
        }
        if (again) ASLAM_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    o->last_count = h_cnt[0];
    int ovf = 0;
    for (int i = 0; i < n; ++i) {
        if (counts) counts[i] = h_cnt[i];
        if (o->pending_track && match_counts) match_counts[i] = h_mc[i];
        ovf |= h_ovf[i];
    }
    if (ovf) return fail(ASLAM_E_OVERFLOW, "a pyramid level had more tied keypoints than its %d spare rows; output clipped", SEL_TIE_ROOM);
    return ASLAM_OK;
}

int aslam_orb_sync(aslam_orb *o, aslam_keypoint *kps, uint8_t *desc, int32_t *counts)
{
    return orb_collect(o, kps, desc, counts, nullptr, nullptr);
}

int aslam_orb_track_batch_async(aslam_orb *o, const uint8_t *images, int n, size_t frame_stride, int row_stride,
                                float ratio, int images_on_device)
{
    ASLAM_REQUIRE(o, "orb is NULL");
    int rc = orb_extract_common(o, images, n, frame_stride, row_stride, nullptr,
                                images_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice);
    if (rc) return rc;
    rc = orb_track_enqueue(o, n, ratio);
    if (rc) o->pending = 0;             // the call failed as a whole: nothing is pending
    return rc;
}

}  // extern "C"

static int orb_track_enqueue(aslam_orb *o, int n, float ratio)
{
    int rc;
    aslam_ctx *ctx = o->ctx;
    const OrbGeom &g = o->g;
    const int cap = g.kp_cap, splits = o->track_splits;
    if (!o->have_carry) ASLAM_CUDA(cudaMemsetAsync(o->b.counts, 0, sizeof(int32_t), ctx->stream));
    const int rows_per = (cap + splits - 1) / splits;
    {
    ProfScope ps(ctx, PK_MISC);
    fill_track_problems_kernel<<<(n * splits + 127) / 128, 128, 0, ctx->stream>>>(o->d_probs, o->b.desc, o->b.counts,
                                                                                  n, splits, cap, rows_per);
    }
    ASLAM_LAUNCH_CHECK();
    if ((rc = launch_knn2_partial(ctx, o->d_probs, n * splits, cap, o->d_partial, cap, splits))) return rc;
    RatioTest rt; rt.ratio = (double)ratio; rt.as_double = 0; rt.disabled = ratio < 0;
    if ((rc = launch_knn2_finish_matches(ctx, o->d_probs, o->d_partial, n, splits, cap, rt, o->d_matches, cap,
                                         o->d_match_counts, nullptr, nullptr))) return rc;
    // carry the last frame of this batch into slot 0 for the next call
    ASLAM_CUDA(cudaMemcpyAsync(o->b.desc, o->b.desc + (size_t)n * cap * 32, (size_t)cap * 32, cudaMemcpyDeviceToDevice, ctx->stream));
    ASLAM_CUDA(cudaMemcpyAsync(o->b.kps, o->b.kps + (size_t)n * cap, sizeof(aslam_keypoint) * cap, cudaMemcpyDeviceToDevice, ctx->stream));
    ASLAM_CUDA(cudaMemcpyAsync(o->b.counts, o->b.counts + n, sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    o->have_carry = true;
    o->pending_track = true;
    return ASLAM_OK;
}

extern "C" {

int aslam_orb_retain_descriptors(aslam_orb *o, const uint8_t **d_desc, int32_t *rows)
{
    ASLAM_REQUIRE(o && d_desc && rows, "NULL argument");
    if (o->pending) return fail(ASLAM_E_STATE, "retain_descriptors with an extract pending; call sync first");
    if (o->last_count < 0) return fail(ASLAM_E_STATE, "retain_descriptors needs a synced extract whose descriptors were not retained yet");
    aslam_ctx *ctx = o->ctx;
    ASLAM_CUDA(cudaSetDevice(ctx->device));
    const int slot = o->retain_next;
    if (!o->d_retained[slot]) ASLAM_CUDA(cudaMalloc(&o->d_retained[slot], (size_t)o->g.kp_cap * 32));
    const int n = std::min(o->last_count, o->g.kp_cap);
    if (n) ASLAM_CUDA(cudaMemcpyAsync(o->d_retained[slot], o->b.desc + (size_t)o->g.kp_cap * 32, (size_t)n * 32,
                                      cudaMemcpyDeviceToDevice, ctx->stream));
    *d_desc = o->d_retained[slot];
    *rows = n;
    o->retain_next = slot ^ 1;
    o->last_count = -1;
    return ASLAM_OK;
}

int aslam_orb_track_sync(aslam_orb *o, aslam_keypoint *kps, uint8_t *desc, int32_t *counts,
                         aslam_match *matches, int32_t *match_counts)
{
    return orb_collect(o, kps, desc, counts, matches, match_counts);
}

int aslam_orb_read_level(aslam_orb *o, int frame, int level, int blurred, uint8_t *out)
{
    ASLAM_REQUIRE(o && out, "NULL argument");
    ASLAM_REQUIRE(frame >= 0 && frame < o->max_batch, "frame out of range");
    ASLAM_REQUIRE(level >= 0 && level < o->g.nbuilt, "level not built (it cannot hold keypoints)");
    const LevelGeom &L = o->g.lv[level];
    ASLAM_CUDA(cudaSetDevice(o->ctx->device));
    ASLAM_CUDA(cudaStreamSynchronize(o->ctx->stream));
    const uint8_t *src = (blurred ? o->b.blur : o->b.pyr) + (size_t)frame * o->g.pyr_stride + L.off;
    ASLAM_CUDA(cudaMemcpy2D(out, L.w, src, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost));
    return ASLAM_OK;
}

int aslam_orb_read_corners(aslam_orb *o, int frame, int level, uint32_t *out, int cap, int *n)
{
    ASLAM_REQUIRE(o && n && (out || cap == 0), "NULL argument");
    ASLAM_REQUIRE(frame >= 0 && frame < o->max_batch, "frame out of range");
    ASLAM_REQUIRE(level >= 0 && level < o->g.nlevels, "level out of range");
    *n = 0;
    const LevelGeom &L = o->g.lv[level];
    const int ntiles = level < o->g.nbuilt ? L.tiles_x * L.tiles_y : 0;
    if (ntiles == 0) return ASLAM_OK;
    ASLAM_CUDA(cudaSetDevice(o->ctx->device));
    ASLAM_CUDA(cudaStreamSynchronize(o->ctx->stream));
    std::vector<uint32_t> cnt((size_t)ntiles), tile((size_t)FT_CAP);
    const size_t tile0 = (size_t)frame * o->g.tiles_per_frame + L.tile_off;
    ASLAM_CUDA(cudaMemcpy(cnt.data(), o->b.cand_cnt + tile0, sizeof(uint32_t) * ntiles, cudaMemcpyDeviceToHost));
    int total = 0;
    for (int t = 0; t < ntiles; ++t) {
        const int c = (int)cnt[t];
        if (c > FT_CAP) return fail(ASLAM_E_STATE, "tile %d reports %d corners", t, c);
        if (c) ASLAM_CUDA(cudaMemcpy(tile.data(), o->b.cand + (tile0 + t) * FT_CAP, sizeof(uint32_t) * c, cudaMemcpyDeviceToHost));
        for (int i = 0; i < c; ++i) {
            if (total < cap) out[total] = tile[i];
            ++total;
        }
    }
    *n = total;
    return ASLAM_OK;
}

}  // extern "C"
