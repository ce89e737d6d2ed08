// ORB extraction kernels for sm_100a.  Results follow cv::ORB::detectAndCompute, the reference's
// CPU path (src/legacy/Frame.cpp:47, parameters of src/adapters/gpu/OrbCudaExtractor.cpp:35-45);
// the arithmetic of each stage is SURVEY.md Appendix A2..A9.
//
// Bit-exactness notes: every float expression that OpenCV evaluates without fused multiply-add
// is spelled with __fmul_rn/__fadd_rn/__fsub_rn here (nvcc contracts a*b+c by default);
// cvRound is __float2int_rn (half to even).
#include <cfloat>

#include "orb_kernels.cuh"

namespace aslam {

// ------------------------------------------------------------------ exact bilinear resize (A2)
// One thread = 4 consecutive output pixels of one row (one 32-bit store).  The two source rows
// and the per-column (offset, weight) table entries come through L1; the table is 16-B loaded.
__global__ void __launch_bounds__(256)
resize_kernel(const uint8_t *srcb, uint8_t *dstb, unsigned int stride,
              LevelGeom S, LevelGeom D, const uint32_t *__restrict__ tabs, int thresh)
{
    const int x4 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x4 >= D.w || y >= D.h) return;
    const uint8_t *src = srcb + (size_t)blockIdx.z * stride + S.off;
    uint8_t *dst = dstb + (size_t)blockIdx.z * stride + D.off;
    const uint32_t yt = __ldg(tabs + D.ytab_off + y);
    const uint8_t *r0 = src + (size_t)(yt & 0xffffu) * S.pitch;
    const uint8_t *r1 = r0 + S.pitch;
    const uint32_t wy1 = yt >> 16, wy0 = 256u - wy1;
    const uint4 xt4 = __ldg(reinterpret_cast<const uint4 *>(tabs + D.xtab_off + x4));
    const uint32_t xt[4] = {xt4.x, xt4.y, xt4.z, xt4.w};
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (x4 + i < D.w) {
            const uint32_t o = xt[i] & 0xffffu, wx1 = xt[i] >> 16, wx0 = 256u - wx1;
            const uint32_t h0 = __ldg(r0 + o) * wx0 + __ldg(r0 + o + 1) * wx1;   // 8.8
            const uint32_t h1 = __ldg(r1 + o) * wx0 + __ldg(r1 + o + 1) * wx1;
            uint32_t v = (h0 * wy0 + h1 * wy1 + 32768u) >> 16;                   // 16.16 -> u8
            if (thresh && v <= 254u) v = 0;   // mask levels: threshold(254, THRESH_TOZERO)
            out |= v << (8 * i);
        }
    }
    *reinterpret_cast<uint32_t *>(dst + (size_t)y * D.pitch + x4) = out;
}

int launch_resize_level(aslam_ctx *ctx, const OrbGeom &g, const OrbBuffers &b, int level, int nframes, bool is_mask)
{
    const LevelGeom &D = g.lv[level], &S = g.lv[level - 1];
    dim3 grid((D.w + 127) / 128, (D.h + 7) / 8, nframes), block(32, 8);
    uint8_t *base = is_mask ? b.mask : b.pyr;
    ProfScope ps(ctx, PK_RESIZE);
    resize_kernel<<<grid, block, 0, ctx->stream>>>(base, base, g.pyr_stride, S, D, b.tabs, is_mask ? 1 : 0);
    ASLAM_LAUNCH_CHECK();
    return ASLAM_OK;
}

// ------------------------------------------------------------------ block scan helper
template <int THREADS>
__device__ __forceinline__ int block_excl_scan(int v, int *warp_sums, int &total)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    __syncthreads();
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < THREADS / 32; ++i) {
        const int s = warp_sums[i];
        if (i < w) base += s;
        tot += s;
    }
    total = tot;
    return base + inc - v;
}

// ------------------------------------------------------------------ FAST-9/16 + NMS (A3, A4 filters)
// One CTA = one 62x30 output tile.  Pixels are staged in shared memory twice as u16 (one copy
// shifted by a pixel) so that any horizontally adjacent pixel pair is one aligned 32-bit word:
// the ring test then runs on u16x2 lanes with the 3-input min/max instructions (VIMNMX3.U16x2),
// two pixels per thread per instruction.
//
// With d_k = v - ring_k:  A = max_arcs min_k d_k = v - min_arcs max_k ring_k, and
// B = max_arcs min_k(-d_k) = max_arcs min_k ring_k - v, so only min/max of raw ring values over
// the sixteen 9-arcs are needed.  Corner iff max(A,B) > t; score = max(A,B) - 1.
constexpr int FA_PITCH = 72;

template <int DX, int DY>
__device__ __forceinline__ uint32_t ring_pair(const uint16_t (*A)[FA_PITCH], const uint16_t (*Bc)[FA_PITCH], int r, int p)
{
    if (DX & 1) return reinterpret_cast<const uint32_t *>(A[r + 3 + DY])[(3 + DX) / 2 + p];
    return reinterpret_cast<const uint32_t *>(Bc[r + 3 + DY])[(2 + DX) / 2 + p];
}

__device__ __forceinline__ uint32_t max3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_u16x2(a, b, c); }
__device__ __forceinline__ uint32_t min3(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u16x2(a, b, c); }

__global__ void __launch_bounds__(256)
fast_kernel(OrbGeom g, OrbBuffers b)
{
    __shared__ __align__(16) uint16_t A[FT_PIX_H][FA_PITCH];
    __shared__ __align__(16) uint16_t Bc[FT_PIX_H][FA_PITCH];
    __shared__ __align__(16) uint8_t S[FT_SCORE_H][FT_SCORE_W];
    __shared__ int warp_sums[8];

    const int tile = blockIdx.x, frame = blockIdx.y, tid = threadIdx.x;
    int l = 0;
#pragma unroll
    for (int i = 1; i < ORB_MAX_LEVELS; ++i)
        if (i < g.nbuilt && tile >= g.lv[i].tile_off) l = i;
    const LevelGeom L = g.lv[l];
    const int tl = tile - L.tile_off;
    const int ty = tl / L.tiles_x, tx = tl - ty * L.tiles_x;
    const int ox = g.edge + tx * FT_OUT_W, oy = g.edge + ty * FT_OUT_H;   // first output pixel
    const int px0 = ox - 4, py0 = oy - 4;
    const uint8_t *img = b.pyr + (size_t)frame * g.pyr_stride + L.off;

    for (int i = tid; i < FT_PIX_H * FT_PIX_W; i += 256) {
        const int r = i / FT_PIX_W, c = i - r * FT_PIX_W;
        const int x = min(px0 + c, L.w - 1), y = min(py0 + r, L.h - 1);
        const uint16_t v = __ldg(img + (size_t)y * L.pitch + x);
        A[r][c] = v;
        if (c >= 1) Bc[r][c - 1] = v;
    }
    __syncthreads();

    const int warp = tid >> 5, p = tid & 31;
    const int thr = g.fast_thr;
#pragma unroll 1
    for (int r = warp; r < FT_SCORE_H; r += 8) {
        uint32_t v[16];
        v[0] = ring_pair<0, 3>(A, Bc, r, p);    v[1] = ring_pair<1, 3>(A, Bc, r, p);
        v[2] = ring_pair<2, 2>(A, Bc, r, p);    v[3] = ring_pair<3, 1>(A, Bc, r, p);
        v[4] = ring_pair<3, 0>(A, Bc, r, p);    v[5] = ring_pair<3, -1>(A, Bc, r, p);
        v[6] = ring_pair<2, -2>(A, Bc, r, p);   v[7] = ring_pair<1, -3>(A, Bc, r, p);
        v[8] = ring_pair<0, -3>(A, Bc, r, p);   v[9] = ring_pair<-1, -3>(A, Bc, r, p);
        v[10] = ring_pair<-2, -2>(A, Bc, r, p); v[11] = ring_pair<-3, -1>(A, Bc, r, p);
        v[12] = ring_pair<-3, 0>(A, Bc, r, p);  v[13] = ring_pair<-3, 1>(A, Bc, r, p);
        v[14] = ring_pair<-2, 2>(A, Bc, r, p);  v[15] = ring_pair<-1, 3>(A, Bc, r, p);
        const uint32_t c = ring_pair<0, 0>(A, Bc, r, p);

        uint32_t hi3[16], lo3[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            hi3[k] = max3(v[k], v[(k + 1) & 15], v[(k + 2) & 15]);
            lo3[k] = min3(v[k], v[(k + 1) & 15], v[(k + 2) & 15]);
        }
        uint32_t arc_max[16], arc_min[16];      // max / min of ring over the 9-arc starting at k
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            arc_max[k] = max3(hi3[k], hi3[(k + 3) & 15], hi3[(k + 6) & 15]);
            arc_min[k] = min3(lo3[k], lo3[(k + 3) & 15], lo3[(k + 6) & 15]);
        }
        uint32_t mn = min3(min3(arc_max[0], arc_max[1], arc_max[2]), min3(arc_max[3], arc_max[4], arc_max[5]),
                           min3(arc_max[6], arc_max[7], arc_max[8]));
        mn = min3(mn, min3(arc_max[9], arc_max[10], arc_max[11]), min3(arc_max[12], arc_max[13], arc_max[14]));
        mn = __vminu2(mn, arc_max[15]);
        uint32_t mx = max3(max3(arc_min[0], arc_min[1], arc_min[2]), max3(arc_min[3], arc_min[4], arc_min[5]),
                           max3(arc_min[6], arc_min[7], arc_min[8]));
        mx = max3(mx, max3(arc_min[9], arc_min[10], arc_min[11]), max3(arc_min[12], arc_min[13], arc_min[14]));
        mx = __vmaxu2(mx, arc_min[15]);

        const int c0 = (int)(c & 0xffffu), c1 = (int)(c >> 16);
        const int s0 = max(c0 - (int)(mn & 0xffffu), (int)(mx & 0xffffu) - c0);
        const int s1 = max(c1 - (int)(mn >> 16), (int)(mx >> 16) - c1);
        // cv::FAST evaluates pixels [3, w-3) x [3, h-3) only
        const int x = ox - 1 + 2 * p, y = oy - 1 + r;
        const bool yin = y >= 3 && y < L.h - 3;
        const uint32_t o0 = (s0 > thr && yin && x >= 3 && x < L.w - 3) ? (uint32_t)(s0 - 1) : 0u;
        const uint32_t o1 = (s1 > thr && yin && x + 1 >= 3 && x + 1 < L.w - 3) ? (uint32_t)(s1 - 1) : 0u;
        reinterpret_cast<uint16_t *>(S[r])[p] = (uint16_t)(o0 | (o1 << 8));
    }
    __syncthreads();

    // 3x3 strict non-max suppression + border/mask filter; 8 consecutive columns per thread.
    uint32_t found[4];
    int nfound = 0;
    if (tid < 240) {
        const int row = 1 + tid / 8, cbase = (tid & 7) * 8;
        const uint8_t *mrow = b.mask ? b.mask + (size_t)frame * g.pyr_stride + L.off : nullptr;
        const int y = oy - 1 + row;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
            const int col = cbase + cc;
            if (col < 1 || col > FT_OUT_W) continue;
            const int s = S[row][col];
            if (s == 0) continue;
            const bool is_max = s > S[row - 1][col - 1] && s > S[row - 1][col] && s > S[row - 1][col + 1] &&
                                s > S[row][col - 1] && s > S[row][col + 1] &&
                                s > S[row + 1][col - 1] && s > S[row + 1][col] && s > S[row + 1][col + 1];
            const int x = ox - 1 + col;
            if (!is_max || x >= L.w - g.edge || y >= L.h - g.edge) continue;
            if (mrow && __ldg(mrow + (size_t)y * L.pitch + x) == 0) continue;
            if (nfound < 4) found[nfound] = ((uint32_t)s << 24) | ((uint32_t)y << 12) | (uint32_t)x;
            ++nfound;
        }
    }
    int total;
    const int pos = block_excl_scan<256>(nfound, warp_sums, total);
    uint32_t *out = b.cand + ((size_t)frame * g.tiles_per_frame + tile) * FT_CAP;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (i < nfound) out[pos + i] = found[i];
    if (tid == 0) b.cand_cnt[(size_t)frame * g.tiles_per_frame + tile] = (uint32_t)total;
}

int launch_fast(aslam_ctx *ctx, const OrbGeom &g, const OrbBuffers &b, int nframes)
{
    if (g.tiles_per_frame == 0) return ASLAM_OK;
    dim3 grid(g.tiles_per_frame, nframes);
    ProfScope ps(ctx, PK_FAST);
    fast_kernel<<<grid, 256, 0, ctx->stream>>>(g, b);
    ASLAM_LAUNCH_CHECK();
    return ASLAM_OK;
}

// ------------------------------------------------------------------ 7x7 Gaussian blur (A8)
// Separable float filter as ORB reaches it through cv::GaussianBlur on a pyramid sub-matrix:
// row pass accumulates taps left to right, column pass centre first then the symmetric pairs as
// k*(a+b); BORDER_REFLECT_101.  Columns below row_fused_end / col_fused_end accumulate with
// fused multiply-add, the rest round the product first (the split the cv2 wheel shows).
struct BlurTaps { float k[4]; };   // k[0..3] = taps 0..3 (tap 3 is the centre)

__device__ __forceinline__ int reflect101(int i, int n)
{
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return min(max(i, 0), n - 1);
}

constexpr int BL_W = 128, BL_H = 16;

__global__ void __launch_bounds__(256)
blur_kernel(const uint8_t *__restrict__ pyr, uint8_t *__restrict__ blur, unsigned int stride, LevelGeom L, BlurTaps T)
{
    __shared__ uint8_t P[BL_H + 6][BL_W + 8];
    __shared__ float R[BL_H + 6][BL_W];
    const int x0 = blockIdx.x * BL_W, y0 = blockIdx.y * BL_H, tid = threadIdx.x;
    const uint8_t *img = pyr + (size_t)blockIdx.z * stride + L.off;
    uint8_t *dst = blur + (size_t)blockIdx.z * stride + L.off;

    for (int i = tid; i < (BL_H + 6) * (BL_W + 6); i += 256) {
        const int r = i / (BL_W + 6), c = i - r * (BL_W + 6);
        const int x = reflect101(x0 - 3 + c, L.w), y = reflect101(y0 - 3 + r, L.h);
        P[r][c] = __ldg(img + (size_t)y * L.pitch + x);
    }
    __syncthreads();
    for (int i = tid; i < (BL_H + 6) * BL_W; i += 256) {
        const int r = i / BL_W, c = i - r * BL_W;
        const bool fused = x0 + c < L.row_fused_end;
        float s = __fmul_rn(T.k[0], (float)P[r][c]);
#pragma unroll
        for (int k = 1; k < 7; ++k) {
            const float kk = T.k[k < 4 ? k : 6 - k], pv = (float)P[r][c + k];
            s = fused ? __fmaf_rn(kk, pv, s) : __fadd_rn(s, __fmul_rn(kk, pv));
        }
        R[r][c] = s;
    }
    __syncthreads();
    for (int t = tid; t < BL_H * BL_W / 4; t += 256) {
        const int r = t / (BL_W / 4), c4 = (t - r * (BL_W / 4)) * 4;
        const int y = y0 + r;
        if (y >= L.h || x0 + c4 >= L.w) continue;
        uint32_t out = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = c4 + i;
            const bool fused = x0 + c < L.col_fused_end;
            float s = __fmul_rn(T.k[3], R[r + 3][c]);
#pragma unroll
            for (int k = 1; k <= 3; ++k) {
                const float pr = __fadd_rn(R[r + 3 + k][c], R[r + 3 - k][c]);
                s = fused ? __fmaf_rn(T.k[3 - k], pr, s) : __fadd_rn(s, __fmul_rn(T.k[3 - k], pr));
            }
            const int v = min(max(__float2int_rn(s), 0), 255);
            out |= (uint32_t)v << (8 * i);
        }
        *reinterpret_cast<uint32_t *>(dst + (size_t)y * L.pitch + x0 + c4) = out;
    }
}

int launch_blur(aslam_ctx *ctx, const OrbGeom &g, const OrbBuffers &b, int nframes)
{
    BlurTaps T;
    double t[7], sum = 0;
    for (int i = 0; i < 7; ++i) { const double x = i - 3; t[i] = exp(-0.125 * x * x); sum += t[i]; }
    for (int i = 0; i < 4; ++i) T.k[i] = (float)(t[i] / sum);
    for (int l = 0; l < g.nbuilt; ++l) {
        const LevelGeom &L = g.lv[l];
        if (L.tiles_x == 0) continue;   // no keypoints possible on this level
        dim3 grid((L.w + BL_W - 1) / BL_W, (L.h + BL_H - 1) / BL_H, nframes);
        ProfScope ps(ctx, PK_BLUR);
        blur_kernel<<<grid, 256, 0, ctx->stream>>>(b.pyr, b.blur, g.pyr_stride, L, T);
        ASLAM_LAUNCH_CHECK();
    }
    return ASLAM_OK;
}

// ------------------------------------------------------------------ selection (A4, A5)
// One CTA per (frame, level): retainBest(2N) on the FAST score via a 256-bin histogram,
// Harris response for the survivors, retainBest(N) via a radix select on the float bits; ties
// with the cut-off are all kept, as KeyPointsFilter::retainBest does.  Survivor order is the
// candidate order (tile-major, row-major inside a tile): deterministic, no atomics on positions.
constexpr int SEL_THREADS = 512;

// HarrisResponses(blockSize 7, k 0.04f): integer sums, then a fixed unfused float expression.
__device__ __forceinline__ float harris_response(const uint8_t *__restrict__ img, int pitch, int x, int y)
{
    int a = 0, b = 0, c = 0;
    int r0[9], r1[9], r2[9];
    const uint8_t *p = img + (size_t)(y - 4) * pitch + (x - 4);
#pragma unroll
    for (int i = 0; i < 9; ++i) { r0[i] = __ldg(p + i); r1[i] = __ldg(p + pitch + i); }
#pragma unroll
    for (int dy = 0; dy < 7; ++dy) {
        const uint8_t *q = p + (size_t)(dy + 2) * pitch;
#pragma unroll
        for (int i = 0; i < 9; ++i) r2[i] = __ldg(q + i);
#pragma unroll
        for (int i = 1; i < 8; ++i) {
            const int Ix = (r1[i + 1] - r1[i - 1]) * 2 + (r0[i + 1] - r0[i - 1]) + (r2[i + 1] - r2[i - 1]);
            const int Iy = (r2[i] - r0[i]) * 2 + (r2[i - 1] - r0[i - 1]) + (r2[i + 1] - r0[i + 1]);
            a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) { r0[i] = r1[i]; r1[i] = r2[i]; }
    }
    const float scale = 1.f / (4 * 7 * 255.f);
    const float s4 = __fmul_rn(__fmul_rn(__fmul_rn(scale, scale), scale), scale);
    const float fa = (float)a, fb = (float)b, fc = (float)c;
    const float sum = __fadd_rn(fa, fb);
    const float det = __fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc));
    const float r = __fsub_rn(det, __fmul_rn(__fmul_rn(0.04f, sum), sum));
    return __fmul_rn(r, s4);
}

__device__ __forceinline__ uint32_t float_order_key(float f)
{
    const uint32_t u = __float_as_uint(f == 0.f ? 0.f : f);   // -0 and +0 compare equal
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(SEL_THREADS)
select_kernel(OrbGeom g, OrbBuffers b)
{
    __shared__ int hist[256];
    __shared__ int warp_sums[SEL_THREADS / 32];
    __shared__ int s_thr;
    __shared__ uint32_t s_prefix;
    __shared__ int s_k;

    const int l = blockIdx.x, frame = blockIdx.y, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NWARPS = SEL_THREADS / 32;
    const LevelGeom L = g.lv[l];
    const int ntiles = L.tiles_x * L.tiles_y;
    int32_t *sel_cnt = b.sel_cnt + (size_t)frame * ORB_MAX_LEVELS + l;
    if (ntiles == 0) {
        if (tid == 0) *sel_cnt = 0;
        return;
    }
    const size_t tile0 = (size_t)frame * g.tiles_per_frame + L.tile_off;
    const uint32_t *cand = b.cand + tile0 * FT_CAP;
    const uint32_t *ccnt = b.cand_cnt + tile0;
    uint32_t *tbase = b.tile_base + tile0;
    uint32_t *surv = b.surv + tile0 * FT_CAP;
    float *hresp = b.hresp + tile0 * FT_CAP;
    const uint8_t *img = b.pyr + (size_t)frame * g.pyr_stride + L.off;

    // --- retainBest(2N) threshold on the integer FAST score
    for (int i = tid; i < 256; i += SEL_THREADS) hist[i] = 0;
    __syncthreads();
    for (int t = warp; t < ntiles; t += NWARPS) {
        const int n = (int)ccnt[t];
        for (int i = lane; i < n; i += 32) atomicAdd(&hist[cand[(size_t)t * FT_CAP + i] >> 24], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int total = 0;
        for (int i = 0; i < 256; ++i) total += hist[i];
        const int K1 = 2 * L.quota;
        int thr = 0;
        if (total > K1) {
            int acc = 0;
            thr = 256;
            for (int s = 255; s >= 0; --s) {
                acc += hist[s];
                if (acc >= K1) { thr = s; break; }
            }
        }
        s_thr = thr;
    }
    __syncthreads();
    const uint32_t thr = (uint32_t)s_thr;

    // --- ordered compaction of the survivors: per-tile counts, scan over tiles, ranked writes
    for (int t = warp; t < ntiles; t += NWARPS) {
        const int n = (int)ccnt[t];
        int cnt = 0;
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            const bool pass = i < n && (cand[(size_t)t * FT_CAP + i] >> 24) >= thr;
            cnt += __popc(__ballot_sync(0xffffffffu, pass));
        }
        if (lane == 0) tbase[t] = (uint32_t)cnt;
    }
    __syncthreads();
    int running = 0;
    for (int t0 = 0; t0 < ntiles; t0 += SEL_THREADS) {
        const int t = t0 + tid;
        const int v = t < ntiles ? (int)tbase[t] : 0;
        int total;
        const int ex = block_excl_scan<SEL_THREADS>(v, warp_sums, total);
        if (t < ntiles) tbase[t] = (uint32_t)(running + ex);
        running += total;
    }
    const int M1 = running;
    __syncthreads();
    for (int t = warp; t < ntiles; t += NWARPS) {
        const int n = (int)ccnt[t];
        int base = (int)tbase[t];
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            const uint32_t e = i < n ? cand[(size_t)t * FT_CAP + i] : 0u;
            const bool pass = i < n && (e >> 24) >= thr;
            const uint32_t bal = __ballot_sync(0xffffffffu, pass);
            if (pass) surv[base + __popc(bal & ((1u << lane) - 1u))] = e;
            base += __popc(bal);
        }
    }
    __syncthreads();

    // --- Harris response of every survivor
    for (int i = tid; i < M1; i += SEL_THREADS) {
        const uint32_t e = surv[i];
        hresp[i] = harris_response(img, L.pitch, (int)(e & 0xfffu), (int)((e >> 12) & 0xfffu));
    }
    __syncthreads();

    // --- retainBest(N): K-th largest response by 4 x 8-bit radix passes over the ordered key
    uint32_t cut = 0;   // keep keys >= cut
    const int K2 = L.quota;
    if (M1 > K2) {
        if (tid == 0) { s_prefix = 0; s_k = K2; }
        uint32_t maskbits = 0;
        for (int pass = 3; pass >= 0; --pass) {
            for (int i = tid; i < 256; i += SEL_THREADS) hist[i] = 0;
            __syncthreads();
            const uint32_t prefix = s_prefix;
            for (int i = tid; i < M1; i += SEL_THREADS) {
                const uint32_t key = float_order_key(hresp[i]);
                if ((key & maskbits) == prefix) atomicAdd(&hist[(key >> (8 * pass)) & 0xffu], 1);
            }
            __syncthreads();
            if (tid == 0) {
                int acc = 0, k = s_k;
                for (int bin = 255; bin >= 0; --bin) {
                    if (acc + hist[bin] >= k) {
                        s_k = k - acc;
                        s_prefix = prefix | ((uint32_t)bin << (8 * pass));
                        break;
                    }
                    acc += hist[bin];
                }
            }
            maskbits |= 0xffu << (8 * pass);
            __syncthreads();
        }
        cut = s_prefix;
    }

    // --- ordered output
    uint32_t *sel_xy = b.sel_xy + (size_t)frame * g.kp_cap + L.sel_off;
    float *sel_resp = b.sel_resp + (size_t)frame * g.kp_cap + L.sel_off;
    int written = 0;
    for (int i0 = 0; i0 < M1; i0 += SEL_THREADS) {
        const int i = i0 + tid;
        float r = 0.f;
        bool pass = false;
        if (i < M1) { r = hresp[i]; pass = float_order_key(r) >= cut; }
        int total;
        const int ex = block_excl_scan<SEL_THREADS>(pass ? 1 : 0, warp_sums, total);
        const int pos = written + ex;
        if (pass && pos < L.sel_cap) {
            sel_xy[pos] = surv[i] & 0xffffffu;
            sel_resp[pos] = r;
        }
        written += total;
    }
    if (tid == 0) {
        *sel_cnt = min(written, L.sel_cap);
        if (written > L.sel_cap) atomicOr(b.overflow + frame + 1, 1);
    }
}

int launch_select(aslam_ctx *ctx, const OrbGeom &g, const OrbBuffers &b, int nframes)
{
    if (g.nbuilt == 0) return ASLAM_OK;
    dim3 grid(g.nbuilt, nframes);
    ProfScope ps(ctx, PK_SELECT);
    select_kernel<<<grid, SEL_THREADS, 0, ctx->stream>>>(g, b);
    ASLAM_LAUNCH_CHECK();
    return ASLAM_OK;
}

// ------------------------------------------------------------------ orientation + rBRIEF (A6, A7, A9)
// One warp per keypoint.  Orientation: lanes 0..30 are the columns u = -15..15 of the radius-15
// disc, rows are walked together (one coalesced 31-byte read per row), integer moments reduced
// by shuffles.  Descriptor: lane k produces byte k (8 comparisons on the blurred level).
__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    // cv::fastAtan2 (scalar atan_f32): coefficients pre-scaled by (float)(180/pi), plain float ops.
    const float S = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * S, p3 = -0.3258083974640975f * S;
    const float p5 = 0.1555786518463281f * S, p7 = -0.04432655554792128f * S;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, (float)DBL_EPSILON));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, (float)DBL_EPSILON));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0.f) a = __fsub_rn(180.f, a);
    if (y < 0.f) a = __fsub_rn(360.f, a);
    return a;
}

__device__ const int8_t c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

__global__ void __launch_bounds__(256)
describe_kernel(OrbGeom g, OrbBuffers b)
{
    const int frame = blockIdx.y, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * 8 + (threadIdx.x >> 5);     // keypoint index within the frame
    const int32_t *sc = b.sel_cnt + (size_t)frame * ORB_MAX_LEVELS;
    int l = -1, first = 0, total = 0;
#pragma unroll
    for (int i = 0; i < ORB_MAX_LEVELS; ++i) {
        const int n = i < g.nbuilt ? sc[i] : 0;
        if (l < 0 && slot < total + n) { l = i; first = total; }
        total += n;
    }
    const size_t fslot = (size_t)frame + 1;    // output slot 0 is the tracker's carry
    if (blockIdx.x == 0 && threadIdx.x == 0) b.counts[fslot] = min(total, g.kp_cap);
    if (l < 0 || slot >= g.kp_cap) return;

    const LevelGeom L = g.lv[l];
    const int si = L.sel_off + (slot - first);
    const uint32_t xy = b.sel_xy[(size_t)frame * g.kp_cap + si];
    const float resp = b.sel_resp[(size_t)frame * g.kp_cap + si];
    const int x = (int)(xy & 0xfffu), y = (int)(xy >> 12);
    const uint8_t *img = b.pyr + (size_t)frame * g.pyr_stride + L.off;
    const uint8_t *blr = b.blur + (size_t)frame * g.pyr_stride + L.off;

    // ICAngles: m10 = sum u*I, m01 = sum v*I over |u| <= umax[|v|]
    const int u = lane - 15;
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const uint8_t *c = img + (size_t)y * L.pitch + x + u;
        const int au = u < 0 ? -u : u;
#pragma unroll 1
        for (int v = -15; v <= 15; ++v) {
            const int av = v < 0 ? -v : v;
            if (au <= c_umax[av]) {
                const int val = __ldg(c + (ptrdiff_t)v * L.pitch);
                m10 += u * val;
                m01 += v * val;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // computeKeyPoints scales the point to level-0 coordinates; computeOrbDescriptors scales it back.
    const float kx = __fmul_rn((float)x, L.sf), ky = __fmul_rn((float)y, L.sf);
    const float inv = __fdiv_rn(1.f, L.sf);
    const int xi = __float2int_rn(__fmul_rn(kx, inv)), yi = __float2int_rn(__fmul_rn(ky, inv));
    const float ang = __fmul_rn(angle, (float)(3.14159265358979323846 / 180.f));
    const float ca = (float)cos((double)ang), sa = (float)sin((double)ang);
    const uint8_t *center = blr + (size_t)yi * L.pitch + xi;
    uint32_t byte = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t pk = __ldg(b.pattern + lane * 8 + j);
        const float x0 = (float)(int8_t)(pk & 0xff), y0 = (float)(int8_t)((pk >> 8) & 0xff);
        const float x1 = (float)(int8_t)((pk >> 16) & 0xff), y1 = (float)(int8_t)(pk >> 24);
        const int ix0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, ca), __fmul_rn(y0, sa)));
        const int iy0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, sa), __fmul_rn(y0, ca)));
        const int ix1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, ca), __fmul_rn(y1, sa)));
        const int iy1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, sa), __fmul_rn(y1, ca)));
        const int t0 = __ldg(center + (ptrdiff_t)iy0 * L.pitch + ix0);
        const int t1 = __ldg(center + (ptrdiff_t)iy1 * L.pitch + ix1);
        byte |= (uint32_t)(t0 < t1) << j;
    }
    b.desc[(fslot * g.kp_cap + slot) * 32 + lane] = (uint8_t)byte;
    if (lane == 0) {
        aslam_keypoint k;
        k.x = kx; k.y = ky; k.size = __fmul_rn(31.f, L.sf); k.angle = angle; k.response = resp; k.octave = l;
        b.kps[fslot * g.kp_cap + slot] = k;
    }
}

int launch_describe(aslam_ctx *ctx, const OrbGeom &g, const OrbBuffers &b, int nframes)
{
    dim3 grid((g.kp_cap + 7) / 8, nframes);
    ProfScope ps(ctx, PK_DESCRIBE);
    describe_kernel<<<grid, 256, 0, ctx->stream>>>(g, b);
    ASLAM_LAUNCH_CHECK();
    return ASLAM_OK;
}

// ------------------------------------------------------------------ dynamic-object filter (a4)
// src/main.cpp:29-50 DYNAMIC_CLASSES {0,1,2,3,5,6,7,14,15,16}; Detection::contains is closed on
// floats (include/core/Types.hpp:109-111).  One CTA compacts one frame's keypoints in order.
__device__ __forceinline__ bool is_dynamic_class(int c)
{
    return c == 0 || c == 1 || c == 2 || c == 3 || c == 5 || c == 6 || c == 7 || c == 14 || c == 15 || c == 16;
}

__global__ void __launch_bounds__(256)
box_filter_kernel(aslam_keypoint *kps, uint8_t *desc, int32_t *count, int cap,
                  const aslam_box *__restrict__ boxes, int nboxes, aslam_keypoint *tmp_kps, uint8_t *tmp_desc)
{
    __shared__ int warp_sums[8];
    const int n = min(*count, cap);
    int written = 0;
    for (int i0 = 0; i0 < n; i0 += 256) {
        const int i = i0 + threadIdx.x;
        bool keep = false;
        aslam_keypoint k;
        if (i < n) {
            k = kps[i];
            keep = true;
            for (int j = 0; j < nboxes; ++j) {
                const aslam_box bx = boxes[j];
                if (is_dynamic_class(bx.class_id) && k.x >= bx.x1 && k.x <= bx.x2 && k.y >= bx.y1 && k.y <= bx.y2) { keep = false; break; }
            }
        }
        int total;
        const int ex = block_excl_scan<256>(keep ? 1 : 0, warp_sums, total);
        if (keep) {
            tmp_kps[written + ex] = k;
            const uint4 *s = reinterpret_cast<const uint4 *>(desc + (size_t)i * 32);
            uint4 *d = reinterpret_cast<uint4 *>(tmp_desc + (size_t)(written + ex) * 32);
            d[0] = s[0]; d[1] = s[1];
        }
        written += total;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < written; i += 256) {
        kps[i] = tmp_kps[i];
        const uint4 *s = reinterpret_cast<const uint4 *>(tmp_desc + (size_t)i * 32);
        uint4 *d = reinterpret_cast<uint4 *>(desc + (size_t)i * 32);
        d[0] = s[0]; d[1] = s[1];
    }
    if (threadIdx.x == 0) *count = written;
}

int launch_box_filter(aslam_ctx *ctx, const OrbGeom &g, const OrbBuffers &b, int frame_slot,
                      const aslam_box *d_boxes, int nboxes, aslam_keypoint *tmp_kps, uint8_t *tmp_desc)
{
    if (nboxes <= 0) return ASLAM_OK;
    ProfScope ps(ctx, PK_MISC);
    box_filter_kernel<<<1, 256, 0, ctx->stream>>>(b.kps + (size_t)frame_slot * g.kp_cap,
                                                   b.desc + (size_t)frame_slot * g.kp_cap * 32,
                                                   b.counts + frame_slot, g.kp_cap, d_boxes, nboxes, tmp_kps, tmp_desc);
    ASLAM_LAUNCH_CHECK();
    return ASLAM_OK;
}

}  // namespace aslam
