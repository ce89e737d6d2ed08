// ORB extraction kernels for sm_100a.  Results follow cv::ORB::detectAndCompute, the reference's
// CPU path (src/legacy/Frame.cpp:47, parameters of src/adapters/gpu/OrbCudaExtractor.cpp:35-45);
// the arithmetic of each stage is SURVEY.md Appendix A2..A9.
//
// Bit-exactness notes: every float expression that OpenCV evaluates without fused multiply-add
// is spelled with __fmul_rn/__fadd_rn/__fsub_rn here (nvcc contracts a*b+c by default);
// cvRound is __float2int_rn (half to even).
#include <algorithm>
#include <cfloat>
#include <cstdlib>

#include "orb_kernels.cuh"

namespace aslam {

// ------------------------------------------------------------------ mbarrier / bulk-copy (TMA) primitives
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n"
                 ".reg .pred P1;\n"
                 "LAB_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
                 "@P1 bra DONE;\n"
                 "bra LAB_WAIT;\n"
                 "DONE:\n"
                 "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// one box of a 3-D tensor (x, y, frame) -> shared memory; completion bytes are counted on `bar`.  x must be a multiple of 16.
__device__ __forceinline__ void tma_load_box(uint32_t dst, const void *tmap, int x, int y, int z, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}

// rows of bytes global -> shared without a tensor map (cp.async.bulk, UBLKCP): addresses and size multiples of 16 bytes
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ------------------------------------------------------------------ exact bilinear resize (A2)
// Per source row and group of 4 output columns:
//   3 aligned 32-bit loads cover the <= 6-byte window of the 4 outputs, a funnel shift aligns it,
//   and each 8.8 value is ONE PRMT + ONE DP4A:  (a, b, a, .) . (255-c1, c1, 1, 0) = a*(256-c1) + b*c1.
// Vertical: v = hA*256 + (hB-hA)*wy1 + 2^15, byte 2 of v is the pixel (>>16); 3 PRMTs pack 4 pixels.
// Column tables (window word, shift, 4 selectors, 4 weight words) are built on the host with the
// double arithmetic OpenCV uses and depend only on the level sizes.
struct HRow { uint32_t h[4]; };

__device__ __forceinline__ HRow resize_hrow(uint32_t w0, uint32_t w1, uint32_t w2, const uint4 t0, const uint4 t1)
{
    const uint32_t lo = __funnelshift_r(w0, w1, t0.y), hi = __funnelshift_r(w1, w2, t0.y);
    HRow r;
    r.h[0] = __dp4a(__byte_perm(lo, hi, t0.z & 0xffffu), t1.x, 0u);
    r.h[1] = __dp4a(__byte_perm(lo, hi, t0.z >> 16), t1.y, 0u);
    r.h[2] = __dp4a(__byte_perm(lo, hi, t0.w & 0xffffu), t1.z, 0u);
    r.h[3] = __dp4a(__byte_perm(lo, hi, t0.w >> 16), t1.w, 0u);
    return r;
}

// One thread = 4 consecutive output columns of a strip of RS_ROWS output rows, marched top to bottom:
// the column tables stay in registers, every source row of the strip goes through the horizontal pass
// exactly once (the two latest results are kept), and the raw words of the next source row are
// fetched one step ahead.  Source rows advance by 0, 1, 2 (or more) per output row -- the march emits whatever
// output rows have the current source row as lower tap -- and the advance depends on y only, so the control
// flow is uniform over the CTA.  Columns are windowed: the taps of a group of 4 outputs must lie within 8 bytes
// of the first one, i.e. a level-to-level factor <= 2 (build_orb_geometry refuses anything else).
#ifndef ASLAM_RS_ROWS
#define ASLAM_RS_ROWS 16
#endif
constexpr int RS_ROWS = ASLAM_RS_ROWS;
#ifndef ASLAM_RS_DEPTH
#define ASLAM_RS_DEPTH 4        // source rows in flight per thread: 2 or 4
#endif
static_assert(ASLAM_RS_DEPTH == 2 || ASLAM_RS_DEPTH == 4, "resize prefetch depth");

// THRESH: the mask pyramid's levels are thresholded (threshold(254, TOZERO)); image levels skip those 12 instructions per
// group of 4 pixels (the kernel is bound by instruction issue: 73 % issue-active at 44 % ALU and 34 % DRAM on level 1).
// NC: the source level is read-only for the whole kernel (ld.global.nc); false inside pyramid_cluster_kernel, where a level
// is written by the same launch that reads it.
template <bool NC>
__device__ __forceinline__ uint32_t rs_load(const uint32_t *p)
{
    if (NC) return __ldg(p);
    uint32_t v;         // an ordinary cached load that the compiler cannot turn into ld.global.nc or move across the cluster barrier
    asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

template <bool THRESH, bool NC>
__device__ __forceinline__ void resize_strip_task(unsigned long long sbase, unsigned long long dbase, const LevelGeom &S, const LevelGeom &D,
                                                  const uint32_t *__restrict__ tabs, int cg, int y0, int rows)
{
    // CTA-uniform bases + 32-bit per-thread offsets (a frame's pyramid is far below 4 GB)
    // The frame's base pointers are CTA-uniform.  Passed through an empty asm they stay in ordinary registers; left to the
    // compiler they are re-derived from blockIdx.z and the parameters on the uniform datapath for every source row and
    // every emitted row (8 + 10 of the loop's ~60 instructions per group of 4 pixels).
    asm volatile("" : "+l"(sbase), "+l"(dbase));
    const uint32_t *src = reinterpret_cast<const uint32_t *>(sbase);
    uint32_t *dst = reinterpret_cast<uint32_t *>(dbase);
    const uint4 *ct = reinterpret_cast<const uint4 *>(tabs + D.xtab_off) + 2 * cg;
    const uint4 t0 = __ldg(ct), t1 = __ldg(ct + 1);
    const uint32_t *ytab = tabs + D.ytab_off;
    const uint32_t yend = (uint32_t)min(y0 + rows, D.h);
    const uint32_t pw = (uint32_t)S.pitch >> 2, dpw = (uint32_t)D.pitch >> 2;
    const uint32_t so_last = t0.x + (uint32_t)(S.h - 1) * pw;     // window offset in the last source row

    // Source-row driven march with the two horizontal results in alternating registers (no copies):
    // row s goes through the horizontal pass, then every output row whose lower tap is s is emitted.
    uint32_t e = __ldg(ytab + y0);
    int s_row = (int)(e & 0xffffu);                    // upper tap of the strip's first output row
    // raw words of the next RS_DEPTH source rows are always in flight (the kernel waits on load latency -- 47 % of its
    // stall cycles are long-scoreboard -- not on instruction issue)
    uint32_t so = t0.x + (uint32_t)s_row * pw;
    uint32_t a0 = rs_load<NC>(src + so), a1 = rs_load<NC>(src + so + 1), a2 = rs_load<NC>(src + so + 2);
    so = min(so + pw, so_last);
    uint32_t b0 = rs_load<NC>(src + so), b1 = rs_load<NC>(src + so + 1), b2 = rs_load<NC>(src + so + 2);
#if ASLAM_RS_DEPTH == 4
    so = min(so + pw, so_last);
    uint32_t c0 = rs_load<NC>(src + so), c1 = rs_load<NC>(src + so + 1), c2 = rs_load<NC>(src + so + 2);
    so = min(so + pw, so_last);
    uint32_t d0 = rs_load<NC>(src + so), d1 = rs_load<NC>(src + so + 1), d2 = rs_load<NC>(src + so + 2);
#endif
    // per-thread running pointers (the destination word and the row-table entry advance by constants): without them the
    // compiler rebuilds both base addresses on the uniform datapath for every emitted row -- 17 of an emit's 42 instructions
    uint32_t *dp = dst + ((size_t)y0 * dpw + (size_t)cg);
    const uint32_t *yp = ytab + y0;
    asm volatile("" : "+l"(yp));            // (same reason: a running pointer in ordinary registers)
    uint32_t y = (uint32_t)y0;
#define ASLAM_RS_HROW(H, N0, N1, N2)                                        \
    do {                                                                    \
        H = resize_hrow(N0, N1, N2, t0, t1);                                \
        so = min(so + pw, so_last);                                         \
        ASLAM_BOUNDS((size_t)(so + 3) * 4 <= (size_t)S.pitch * S.h + 256);  /* the third word may touch the 256-B level padding */ \
        N0 = rs_load<NC>(src + so); N1 = rs_load<NC>(src + so + 1); N2 = rs_load<NC>(src + so + 2); \
    } while (0)
#define ASLAM_RS_EMIT(OLD, NEW)                                             \
    while ((int)(e & 0xffffu) + 1 == s_row) {                               \
        const uint32_t wy1 = e >> 16, wy0 = 256u - wy1;                     \
        uint32_t v[4];                                                      \
        _Pragma("unroll")                                                   \
        for (int i = 0; i < 4; ++i) {                                       \
            v[i] = NEW.h[i] * wy1 + (OLD.h[i] * wy0 + 32768u);              /* 16.16, byte 2 = rounded pixel */ \
            if (THRESH && (v[i] >> 16) <= 254u) v[i] = 0;                   /* mask levels: threshold(254, TOZERO) */ \
        }                                                                   \
        const uint32_t p01 = __byte_perm(v[0], v[1], 0x0062), p23 = __byte_perm(v[2], v[3], 0x0062); \
        ASLAM_BOUNDS(dp >= dst && dp + 1 <= dst + (size_t)dpw * D.h && y < (uint32_t)D.h); \
        *dp = __byte_perm(p01, p23, 0x5410);                                \
        dp += dpw;                                                          \
        if (++y == yend) return;                                            \
        e = __ldg(++yp);                                                    \
    }
    HRow h0, h1;
    ASLAM_RS_HROW(h1, a0, a1, a2);
#pragma unroll 1
    for (;;) {
        ++s_row;
        ASLAM_RS_HROW(h0, b0, b1, b2);
        ASLAM_RS_EMIT(h1, h0)
#if ASLAM_RS_DEPTH == 4
        ++s_row;
        ASLAM_RS_HROW(h1, c0, c1, c2);
        ASLAM_RS_EMIT(h0, h1)
        ++s_row;
        ASLAM_RS_HROW(h0, d0, d1, d2);
        ASLAM_RS_EMIT(h1, h0)
#endif
        ++s_row;
        ASLAM_RS_HROW(h1, a0, a1, a2);
        ASLAM_RS_EMIT(h0, h1)
    }
#undef ASLAM_RS_HROW
#undef ASLAM_RS_EMIT
}

template <bool THRESH>
__global__ void __launch_bounds__(256)
resize_kernel(const uint8_t *srcb, uint8_t *dstb, unsigned int stride,
              LevelGeom S, LevelGeom D, const uint32_t *__restrict__ tabs)
{
    // a CTA's warps sit side by side on the same strip of rows, so the CTA reads and writes whole row segments
    const int cg = blockIdx.x * blockDim.x + threadIdx.x;
    const int y0 = blockIdx.y * RS_ROWS;
    if (cg * 4 >= D.w || y0 >= D.h) return;
    resize_strip_task<THRESH, true>(reinterpret_cast<unsigned long long>(srcb + (size_t)blockIdx.z * stride + S.off),
                                    reinterpret_cast<unsigned long long>(dstb + (size_t)blockIdx.z * stride + D.off), S, D, tabs, cg, y0, RS_ROWS);
}

// ---- round 2, second form: the same march with its control flow taken off the memory system.  resize_kernel decides
// after every source row whether to emit by comparing against a row-table entry it has just loaded (a global load, two
// logic instructions and a branch in the loop-carried path, and ~13 instructions of pointer upkeep per emitted row).  Here
// the emit decisions of a strip are one 32-bit mask built on the host (LevelGeom::stab_off: bit i = "source row
// s_first + 1 + i is the lower tap of the next output row"), the weight of the next output row is fetched right after the
// previous emit, and the loop ends with the mask.  Levels whose strips do not fit the mask (rs_fast == 0: more than 32
// source rows per strip, or two output rows with the same upper tap) keep resize_kernel.
// MEASURED SLOWER and therefore off by default (ASLAM_RS_MARCH=0): 26 % fewer executed instructions (32.2 M -> 23.7 M warp
// instructions on level 1) but 0.189 ms per 256 frames against 0.176 (0.198 with two frames per thread, 0.187 / 0.225 /
// 0.240 with 6 / 8 / 12 source rows in flight, 0.213 with the mask in ordinary registers).  ncu, level 1, two frames per
// thread: issue-active 43 % instead of 77 %, long-scoreboard stall 8.6 instead of 4.4 warps per issue, achieved occupancy
// 34 % instead of 57 % -- the pass is bound by the latency of its DRAM reads (L2 hit rate 18 %), and the instructions that
// were removed had been issuing in that shadow (profiles/r02_resize_march_ncu.csv).
// NF frames per thread (same column group, same strip): the tables, the emit decisions, the weights and the address
// arithmetic are shared, and the frames' independent loads and multiplies fill each other's latencies.
#ifndef ASLAM_RM_NF
#define ASLAM_RM_NF 1
#endif
#ifndef ASLAM_RM_DEPTH
#define ASLAM_RM_DEPTH 8        // source rows in flight per thread and frame (even)
#endif
#ifndef ASLAM_RM_VMASK
#define ASLAM_RM_VMASK 0
#endif
template <bool THRESH, int NF, int DEPTH>
__global__ void __launch_bounds__(256)
resize_march_kernel(const uint8_t *srcb, uint8_t *dstb, unsigned int stride,
                    LevelGeom S, LevelGeom D, const uint32_t *__restrict__ tabs, int nframes)
{
    static_assert(DEPTH >= 2 && DEPTH % 2 == 0, "the two horizontal results alternate");
    const int cg = blockIdx.x * blockDim.x + threadIdx.x;
    const int y0 = blockIdx.y * D.rs_rows;
    if (cg * 4 >= D.w) return;
    // frame f of this thread; a batch that is not a multiple of NF repeats its last frame (same bytes written twice)
    const uint32_t *src[NF];
    uint32_t *dp[NF];
    const uint4 *ct = reinterpret_cast<const uint4 *>(tabs + D.xtab_off) + 2 * cg;
    const uint4 t0 = __ldg(ct), t1 = __ldg(ct + 1);
    // (ASLAM_RM_VMASK: fetch the CTA-uniform mask through an address ptxas cannot prove uniform, so that it is tested in
    // ordinary registers instead of on the uniform datapath -- measured slower, 0.213 vs 0.189 ms)
    const uint2 st = __ldg(reinterpret_cast<const uint2 *>(tabs + D.stab_off) + blockIdx.y + (ASLAM_RM_VMASK ? (cg >> 28) : 0));
    uint32_t mask = st.x;
    const uint32_t pw = (uint32_t)S.pitch >> 2, dpw = (uint32_t)D.pitch >> 2;
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        const int z = min((int)blockIdx.z * NF + f, nframes - 1);
        unsigned long long sb = reinterpret_cast<unsigned long long>(srcb + (size_t)z * stride + S.off);
        unsigned long long db = reinterpret_cast<unsigned long long>(dstb + (size_t)z * stride + D.off);
        asm volatile("" : "+l"(sb), "+l"(db));
        src[f] = reinterpret_cast<const uint32_t *>(sb);
        dp[f] = reinterpret_cast<uint32_t *>(db) + ((size_t)y0 * dpw + (size_t)cg);
    }
    const uint32_t so_last = t0.x + (uint32_t)(S.h - 1) * pw;     // window offset in the last source row
    uint32_t so = t0.x + st.y * pw;
    uint32_t raw[DEPTH][NF][3];                                   // raw words of the next DEPTH source rows, per frame
    int rows_left = 33 - __clz((int)st.x);                        // the strip reads source rows s_first .. s_first + 1 + highest mask bit
#pragma unroll
    for (int j = 0; j < DEPTH; ++j) {
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            raw[j][f][0] = raw[j][f][1] = raw[j][f][2] = 0u;
            if (rows_left > 0) { raw[j][f][0] = __ldg(src[f] + so); raw[j][f][1] = __ldg(src[f] + so + 1); raw[j][f][2] = __ldg(src[f] + so + 2); }
        }
        --rows_left;
        if (j + 1 < DEPTH) so = min(so + pw, so_last);
    }
    uint32_t yi = (uint32_t)(D.ytab_off + y0);
    uint32_t e = __ldg(tabs + yi);                                // weight word of the next output row
    uint32_t doff = 0;                                            // words from dp[f] to the next output row
    HRow h[2][NF];
    // source row j of the ring through the horizontal pass into h[j & 1]; its registers take the row DEPTH further on
    auto hrow = [&](int j) {
#pragma unroll
        for (int f = 0; f < NF; ++f) h[j & 1][f] = resize_hrow(raw[j][f][0], raw[j][f][1], raw[j][f][2], t0, t1);
        so = min(so + pw, so_last);
        ASLAM_BOUNDS((size_t)(so + 3) * 4 <= (size_t)S.pitch * S.h + 256);
        if (rows_left > 0) {                                       // (CTA-uniform) no loads past the strip's last source row
#pragma unroll
            for (int f = 0; f < NF; ++f) { raw[j][f][0] = __ldg(src[f] + so); raw[j][f][1] = __ldg(src[f] + so + 1); raw[j][f][2] = __ldg(src[f] + so + 2); }
        }
        --rows_left;
    };
    hrow(0);
#pragma unroll 1
    for (;;) {
#pragma unroll
        for (int jj = 1; jj <= DEPTH; ++jj) {
            const int j = jj % DEPTH;                              // rows 1 .. DEPTH-1, then 0 again
            hrow(j);
            if (mask & 1u) {                                       // row j is the lower tap of the next output row
                const uint32_t wy1 = e >> 16, wy0 = 256u - wy1;
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    uint32_t v[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        v[i] = h[j & 1][f].h[i] * wy1 + (h[(j & 1) ^ 1][f].h[i] * wy0 + 32768u);     // 16.16, byte 2 = rounded pixel
                        if (THRESH && (v[i] >> 16) <= 254u) v[i] = 0;                              // mask levels: threshold(254, TOZERO)
                    }
                    const uint32_t p01 = __byte_perm(v[0], v[1], 0x0062), p23 = __byte_perm(v[2], v[3], 0x0062);
                    ASLAM_BOUNDS((size_t)y0 * dpw + cg + doff < (size_t)dpw * D.h);
                    dp[f][doff] = __byte_perm(p01, p23, 0x5410);
                }
                if (mask == 1u) return;
                doff += dpw;
                e = __ldg(tabs + ++yi);
            }
            mask >>= 1;
        }
    }
}

// ---- shared-memory form (the shipped kernel).  One CTA = a strip of RS_ROWS output rows over the level's full width.
// The source rows the strip needs (about 1.2 * RS_ROWS + 2 of them) arrive by bulk asynchronous copies -- one
// cp.async.bulk per row, issued by the first lanes of the CTA, completion on an mbarrier -- so every source byte crosses
// L2 -> SM once per strip in full lines and no thread waits on its own global loads; the march then reads its 3-word
// windows from shared memory.  Same arithmetic, same tables, same one-thread-per-4-columns march as resize_kernel.
#ifndef ASLAM_RS_GROUP
#define ASLAM_RS_GROUP 64       // rows per arrival barrier: 64 = the whole strip on one barrier (0.184 ms); 6: 0.192, 3: 0.191
#endif
constexpr int RS_GROUP = ASLAM_RS_GROUP;
constexpr int RS_NGROUPS = (2 * RS_ROWS + 3 + RS_GROUP - 1) / RS_GROUP;      // a strip reads at most 2 * RS_ROWS + 3 source rows (scale <= 2)
template <bool THRESH>
__global__ void __launch_bounds__(1024)
resize_smem_kernel(const uint8_t *srcb, uint8_t *dstb, unsigned int stride,
                   LevelGeom S, LevelGeom D, const uint32_t *__restrict__ tabs)
{
    extern __shared__ __align__(128) uint8_t rs_rows[];          // [rows of the strip][S.pitch]
    // the strip's rows arrive in groups of RS_GROUP, one mbarrier each: the march starts when the first group is in and
    // the later groups land while it computes
    __shared__ __align__(8) unsigned long long bar[RS_NGROUPS];
    const int tid = threadIdx.x, cg = tid;
    const int y0 = blockIdx.x * RS_ROWS;
    const uint32_t *ytab = tabs + D.ytab_off;
    const int yend = min(y0 + RS_ROWS, D.h);
    const int s_first = (int)(__ldg(ytab + y0) & 0xffffu);                                   // upper tap of the first output row
    const int s_last = min((int)(__ldg(ytab + yend - 1) & 0xffffu) + 1, S.h - 1);          // lower tap of the last one
    const int nrows = s_last - s_first + 1;
    const uint32_t bar_a = smem_u32(&bar[0]);
    if (tid < RS_NGROUPS) {
        mbar_init(bar_a + 8u * tid, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const int in_group = min(RS_GROUP, nrows - tid * RS_GROUP);
        if (in_group > 0) mbar_arrive_expect_tx(bar_a + 8u * tid, (uint32_t)in_group * (uint32_t)S.pitch);
    }
    __syncthreads();
    const uint8_t *sbase = srcb + (size_t)blockIdx.z * stride + S.off;
    for (int r = tid; r < nrows; r += (int)blockDim.x)      // thread r fetches source row s_first + r
        bulk_load(smem_u32(rs_rows + (size_t)r * S.pitch), sbase + (size_t)(s_first + r) * S.pitch, (uint32_t)S.pitch,
                  bar_a + 8u * (uint32_t)(r / RS_GROUP));
    if (cg * 4 >= D.w) return;
    const uint4 *ct = reinterpret_cast<const uint4 *>(tabs + D.xtab_off) + 2 * cg;
    const uint4 t0 = __ldg(ct), t1 = __ldg(ct + 1);
    const uint32_t pw = (uint32_t)S.pitch >> 2, dpw = (uint32_t)D.pitch >> 2;
    unsigned long long dbase = reinterpret_cast<unsigned long long>(dstb + (size_t)blockIdx.z * stride + D.off);
    asm volatile("" : "+l"(dbase));
    uint32_t *dp = reinterpret_cast<uint32_t *>(dbase) + ((size_t)y0 * dpw + (size_t)cg);
    const uint32_t *yp = ytab + y0;
    uint32_t e = __ldg(yp);
    int s_row = s_first;
    int y = y0;
    // window of this thread in row r of the strip: words t0.x .. t0.x + 2 (the third may lie in the row's padding)
    const uint32_t *win = reinterpret_cast<const uint32_t *>(rs_rows) + t0.x;
    const uint32_t last = (uint32_t)(nrows - 1) * pw;                 // rows beyond the level's last are its last (clamped taps)
    uint32_t ro = 0;
    int ri = 0;                                                       // row of the strip the next horizontal pass reads
#define ASLAM_RS_HROW_S(H)                                                  \
    do {                                                                    \
        if (ri % RS_GROUP == 0 && ri < nrows) mbar_wait(bar_a + 8u * (uint32_t)(ri / RS_GROUP), 0u); \
        ++ri;                                                               \
        H = resize_hrow(win[ro], win[ro + 1], win[ro + 2], t0, t1);         \
        ro = min(ro + pw, last);                                            \
    } while (0)
#define ASLAM_RS_EMIT_S(OLD, NEW)                                           \
    while ((int)(e & 0xffffu) + 1 == s_row) {                               \
        const uint32_t wy1 = e >> 16, wy0 = 256u - wy1;                     \
        uint32_t v[4];                                                      \
        _Pragma("unroll")                                                   \
        for (int i = 0; i < 4; ++i) {                                       \
            v[i] = NEW.h[i] * wy1 + (OLD.h[i] * wy0 + 32768u);              /* 16.16, byte 2 = rounded pixel */ \
            if (THRESH && (v[i] >> 16) <= 254u) v[i] = 0;                   /* mask levels: threshold(254, TOZERO) */ \
        }                                                                   \
        const uint32_t p01 = __byte_perm(v[0], v[1], 0x0062), p23 = __byte_perm(v[2], v[3], 0x0062); \
        *dp = __byte_perm(p01, p23, 0x5410);                                \
        dp += dpw;                                                          \
        if (++y == yend) return;                                            \
        e = __ldg(++yp);                                                    \
    }
    HRow h0, h1;
    ASLAM_RS_HROW_S(h1);
#pragma unroll 1
    for (;;) {
        ++s_row;
        ASLAM_RS_HROW_S(h0);
        ASLAM_RS_EMIT_S(h1, h0)
        ++s_row;
        ASLAM_RS_HROW_S(h1);
        ASLAM_RS_EMIT_S(h0, h1)
    }
#undef ASLAM_RS_HROW_S
#undef ASLAM_RS_EMIT_S
}

// Measured on B200 (ms per 256 frames, all 7 levels): direct loads with 4 source rows in flight 0.176 (2 rows: 0.189); source
// rows staged in shared memory by bulk asynchronous copies (resize_smem_kernel) 0.184 with 16-row strips, 0.185 with 8,
// 0.197 with 32, 0.191 when the rows arrive in groups on several barriers.
#ifndef ASLAM_RS_MARCH
#define ASLAM_RS_MARCH 0        // 1: resize_march_kernel (host-built emit masks) wherever a level's strips fit it -- measured slower, see there
#endif
#ifndef ASLAM_RS_SMEM
#define ASLAM_RS_SMEM 0         // 1: the staged form (kept for A/B runs and for GPUs where it wins); 0: direct loads
#endif
// ---- the whole level chain of a frame in ONE launch: a thread-block cluster per frame walks the levels, a cluster barrier
// (release / acquire at cluster scope: the level just written is visible to, and its stale L1 lines dropped by, every CTA of
// the cluster) between them.  For calls with few frames, where seven dependent launches of a few microseconds each are most
// of the pyramid's latency (one 752 x 480 frame: 60 us of the 117 us of kernel time with per-level launches).  Same strip
// march as resize_kernel (resize_strip_task), the strip height chosen per level so that a level is about one pass of the
// cluster's threads; the mask pyramid's levels are further tasks of the same pass.
constexpr int PYR_CLUSTER = 8, PYR_THREADS = 512;
template <bool WITH_MASK>
__global__ void __launch_bounds__(PYR_THREADS)
pyramid_cluster_kernel(OrbGeom g, OrbBuffers b)
{
    const int frame = blockIdx.y;
    const int tid_c = (int)blockIdx.x * PYR_THREADS + (int)threadIdx.x, nthr = PYR_CLUSTER * PYR_THREADS;
    const unsigned long long ibase = reinterpret_cast<unsigned long long>(b.pyr + (size_t)frame * g.pyr_stride);
    const unsigned long long mbase = WITH_MASK ? reinterpret_cast<unsigned long long>(b.mask + (size_t)frame * g.pyr_stride) : 0ull;
#pragma unroll 1
    for (int l = 1; l < g.nbuilt; ++l) {
        const LevelGeom S = g.lv[l - 1], D = g.lv[l];
        const int ncg = (D.w + 3) >> 2;
        // rows per strip: a multiple of 4 in [4, RS_ROWS] that gives the cluster's threads about one task each
        int rows = (int)(((long long)D.h * ncg * (WITH_MASK ? 2 : 1) + nthr - 1) / nthr);
        rows = min(RS_ROWS, max(4, (rows + 3) & ~3));
        const int nstrips = (D.h + rows - 1) / rows, ntasks = ncg * nstrips;
        for (int t = tid_c; t < (WITH_MASK ? 2 : 1) * ntasks; t += nthr) {
            const bool is_mask = WITH_MASK && t >= ntasks;
            const int tt = is_mask ? t - ntasks : t;
            const int strip = tt / ncg, cg = tt - strip * ncg;
            if (is_mask) resize_strip_task<true, false>(mbase + S.off, mbase + D.off, S, D, b.tabs, cg, strip * rows, rows);
            else resize_strip_task<false, false>(ibase + S.off, ibase + D.off, S, D, b.tabs, cg, strip * rows, rows);
        }
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
}

#ifndef ASLAM_RS_WARPS
#define ASLAM_RS_WARPS 8           // warps per resize CTA (side by side on a strip of rows)
#endif
#ifndef ASLAM_PYR_CLUSTER_MAX
#define ASLAM_PYR_CLUSTER_MAX 8     // calls with at most this many frames build the pyramid with pyramid_cluster_kernel (0 = never).  Alone on the GPU the
                                    // cluster form wins up to ~32 frames per call (resize 76 -> 38 us at 32), but under 8 concurrent frame streams of 32
                                    // frames its CTAs idle at the barriers while other streams wait for SM slots: 88.4 k vs 91.9 k frames/s in bench.py
#endif
int launch_pyramid_cluster(aslam_ctx *ctx, const OrbGeom &g, const OrbBuffers &b, int nframes, bool with_mask)
{
    // test hook: behave like a device that refuses cluster launches (tests/test_gpu_errors.py checks the level-by-level path takes over)
    static const bool refuse = [] { const char *e = getenv("ASLAM_TEST_REFUSE_CLUSTER"); return e && atoi(e) != 0; }();
    if (refuse) return fail(ASLAM_E_CUDA, "cluster launch refused (ASLAM_TEST_REFUSE_CLUSTER)");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(PYR_CLUSTER, nframes, 1);
    cfg.blockDim = dim3(PYR_THREADS, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PYR_CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    ProfScope ps(ctx, PK_RESIZE);
    if (with_mask) ASLAM_CUDA(cudaLaunchKernelEx(&cfg, pyramid_cluster_kernel<true>, g, b));
    else ASLAM_CUDA(cudaLaunchKernelEx(&cfg, pyramid_cluster_kernel<false>, g, b));
    ASLAM_LAUNCH_CHECK();
    return ASLAM_OK;
}
bool pyramid_cluster_wanted(int nframes)
{
    static const int lim = [] { const char *e = getenv("ASLAM_PYR_CLUSTER_MAX"); return e ? atoi(e) : ASLAM_PYR_CLUSTER_MAX; }();
    return nframes <= lim;
}

int launch_resize_level(aslam_ctx *ctx, const OrbGeom &g, const OrbBuffers &b, int level, int nframes, bool is_mask)
{
    const LevelGeom &D = g.lv[level], &S = g.lv[level - 1];
    const int cgs = (D.w + 3) / 4;
    uint8_t *base = is_mask ? b.mask : b.pyr;
#if ASLAM_RS_SMEM
    // a strip of RS_ROWS output rows reads at most ceil(RS_ROWS * S.h / D.h) + 2 source rows
    const int max_rows = (int)(((long long)RS_ROWS * S.h + D.h - 1) / D.h) + 2;
    const size_t smem = (size_t)max_rows * S.pitch + 16;        // + the last window's third word
    const int threads = (cgs + 31) / 32 * 32;
    if (threads <= 1024 && smem <= 200 * 1024) {
        if (smem > 48 * 1024 && (size_t)ctx->resize_smem_set < smem) {       // opt in once per device (and size)
            ASLAM_CUDA(cudaFuncSetAttribute(resize_smem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            ASLAM_CUDA(cudaFuncSetAttribute(resize_smem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            ctx->resize_smem_set = 200 * 1024;
        }
        dim3 grid((D.h + RS_ROWS - 1) / RS_ROWS, 1, nframes);
        ProfScope ps(ctx, PK_RESIZE);
        if (is_mask) resize_smem_kernel<true><<<grid, threads, smem, ctx->stream>>>(base, base, g.pyr_stride, S, D, b.tabs);
        else resize_smem_kernel<false><<<grid, threads, smem, ctx->stream>>>(base, base, g.pyr_stride, S, D, b.tabs);
        ASLAM_LAUNCH_CHECK();
        return ASLAM_OK;
    }
#endif
    const int warps = min(ASLAM_RS_WARPS, (cgs + 31) / 32);
#if ASLAM_RS_MARCH
    if (D.rs_fast) {
        // several frames per thread only once there are enough of them to fill the GPU anyway
        const int nf = nframes >= 16 ? ASLAM_RM_NF : 1;
        dim3 grid((cgs + 32 * warps - 1) / (32 * warps), (D.h + D.rs_rows - 1) / D.rs_rows, (nframes + nf - 1) / nf), block(32 * warps);
        ProfScope ps(ctx, PK_RESIZE);
        if (nf == 1) {
            if (is_mask) resize_march_kernel<true, 1, ASLAM_RM_DEPTH><<<grid, block, 0, ctx->stream>>>(base, base, g.pyr_stride, S, D, b.tabs, nframes);
            else resize_march_kernel<false, 1, ASLAM_RM_DEPTH><<<grid, block, 0, ctx->stream>>>(base, base, g.pyr_stride, S, D, b.tabs, nframes);
        } else {
            if (is_mask) resize_march_kernel<true, ASLAM_RM_NF, ASLAM_RM_DEPTH><<<grid, block, 0, ctx->stream>>>(base, base, g.pyr_stride, S, D, b.tabs, nframes);
            else resize_march_kernel<false, ASLAM_RM_NF, ASLAM_RM_DEPTH><<<grid, block, 0, ctx->stream>>>(base, base, g.pyr_stride, S, D, b.tabs, nframes);
        }
        ASLAM_LAUNCH_CHECK();
        return ASLAM_OK;
    }
#endif
    dim3 grid((cgs + 32 * warps - 1) / (32 * warps), (D.h + RS_ROWS - 1) / RS_ROWS, nframes), block(32 * warps);
    ProfScope ps(ctx, PK_RESIZE);
    if (is_mask) resize_kernel<true><<<grid, block, 0, ctx->stream>>>(base, base, g.pyr_stride, S, D, b.tabs);
    else resize_kernel<false><<<grid, block, 0, ctx->stream>>>(base, base, g.pyr_stride, S, D, b.tabs);
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
// One CTA = one tile of 62 x FT_OUT_H output pixels = a 64 x (FT_OUT_H+2) block of scores = a
// (64+6) x (FT_OUT_H+8) block of pixels.  Pixels are staged in shared memory twice as u16 -- copy A at even alignment, copy Bc
// shifted by one pixel -- so that ANY horizontally adjacent pixel pair is one aligned 32-bit word.
// The ring test then runs on u16x2 lanes with the 3-input min/max instruction (VIMNMX3.U16x2):
// two pixels per thread per instruction.
//
// With d_k = v - ring_k:  A = max_arcs min_k d_k = v - min_arcs max_k ring_k, and
// B = max_arcs min_k(-d_k) = max_arcs min_k ring_k - v, so only min/max of raw ring values over
// the sixteen 9-arcs are needed.  Corner iff max(A,B) > t; score = max(A,B) - 1.
//
// Score block origin: sx0 = ((edge-1) & ~1) + 62*tx (even, so a pair never straddles a word),
// sy0 = edge-1 + 30*ty; outputs are its interior columns 1..62 / rows 1..30.
constexpr int FA_PITCH = 80;      // u16 per staged row (76 used), 160 B
constexpr int FA_WORDS = 19;      // aligned 32-bit global words per staged row

// The two staged copies are one array (copy 0 = A, copy 1 = Bc) so that every pixel pair of a ring is ONE shared-memory
// address plus a compile-time offset: `at` = address of the word (row r, pair column gp - 2) of copy A, where r is the
// score row (= staged row of the window's top) -- all offsets below are >= 0.
constexpr int FA_COPY_BYTES = FT_PIX_H * FA_PITCH * 2;
template <int OFF>
__device__ __forceinline__ uint32_t lds_u32_off(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
    return v;
}
template <int DX, int DY>
__device__ __forceinline__ uint32_t ring_pair(uint32_t at)
{
    constexpr int word = (DX & 1) == 0 ? DX / 2 : (DX - 1) / 2;
    constexpr int off = (DY + 3) * FA_PITCH * 2 + 4 * (word + 2) + ((DX & 1) == 0 ? 0 : FA_COPY_BYTES);
    static_assert(off >= 0, "ring offsets are non-negative");
    return lds_u32_off<off>(at);
}

__device__ __forceinline__ uint32_t max3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_u16x2(a, b, c); }
__device__ __forceinline__ uint32_t min3(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u16x2(a, b, c); }

#ifndef ASLAM_FAST_MINB
#define ASLAM_FAST_MINB 5
#endif
#ifndef ASLAM_FAST_COMPACT
#define ASLAM_FAST_COMPACT 1
#endif
#ifndef ASLAM_FAST_ZERO_S
#define ASLAM_FAST_ZERO_S 1     // 1: the score block is zeroed by 16-byte stores before phase A instead of per warp and row
#endif
#ifndef ASLAM_FAST_A_UNROLL
#define ASLAM_FAST_A_UNROLL 1   // rows of the quick test per loop trip
#endif
constexpr int FAST_A_UNROLL = ASLAM_FAST_A_UNROLL;
#ifndef ASLAM_FAST_Q32
#define ASLAM_FAST_Q32 1        // queue entries as 32-bit words: the loop variable is stored as it is (no 16-bit twin of it)
#endif
#if ASLAM_FAST_Q32
typedef uint32_t FastQ;
#else
typedef uint16_t FastQ;
#endif
#ifndef ASLAM_FAST_NET
#define ASLAM_FAST_NET 1     // 1: the 72-instruction score network (shared eight-pixel windows); 0: round 1's 80
#endif
struct FastMul { uint32_t one, minus1, bias; };      // 1, 0xffffffff, 0x01000100

__device__ __forceinline__ uint32_t imad32(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// Corner score of the pixel pair at score row r, pair column gp - g0/2: relu(max(A, B) - t) per u16 lane.
__device__ __forceinline__ uint32_t fast_score_pair(uint32_t at, uint32_t neg_thr, const FastMul M)
{
    uint32_t v[16];
    v[0] = ring_pair<0, 3>(at);    v[1] = ring_pair<1, 3>(at);
    v[2] = ring_pair<2, 2>(at);    v[3] = ring_pair<3, 1>(at);
    v[4] = ring_pair<3, 0>(at);    v[5] = ring_pair<3, -1>(at);
    v[6] = ring_pair<2, -2>(at);   v[7] = ring_pair<1, -3>(at);
    v[8] = ring_pair<0, -3>(at);   v[9] = ring_pair<-1, -3>(at);
    v[10] = ring_pair<-2, -2>(at); v[11] = ring_pair<-3, -1>(at);
    v[12] = ring_pair<-3, 0>(at);  v[13] = ring_pair<-3, 1>(at);
    v[14] = ring_pair<-2, 2>(at);  v[15] = ring_pair<-1, 3>(at);
    const uint32_t c = ring_pair<0, 0>(at);

#if ASLAM_FAST_NET == 1
    // 72 min/max instead of 80.  The arcs starting at 2j and 2j+1 share the eight pixels 2j+1 .. 2j+8 (W) and differ in
    // one end pixel: max(min(W, v[2j]), min(W, v[2j+9])) = min(W, max(v[2j], v[2j+9])) -- the form of cv::cornerScore's
    // loop -- and W is two of the eight 4-windows that start at odd pixels, each two of the eight odd-aligned pairs.
    uint32_t plo[8], phi[8], qlo[8], qhi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        plo[j] = __vminu2(v[(2 * j + 1) & 15], v[(2 * j + 2) & 15]);
        phi[j] = __vmaxu2(v[(2 * j + 1) & 15], v[(2 * j + 2) & 15]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        qlo[j] = __vminu2(plo[j], plo[(j + 1) & 7]);
        qhi[j] = __vmaxu2(phi[j], phi[(j + 1) & 7]);
    }
    uint32_t rmin[8], rmax[8];      // rmin[j] = max over the arcs starting at 2j, 2j+1 of the arc's minimum; rmax alike
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        rmin[j] = min3(qlo[j], qlo[(j + 2) & 7], __vmaxu2(v[2 * j], v[(2 * j + 9) & 15]));
        rmax[j] = max3(qhi[j], qhi[(j + 2) & 7], __vminu2(v[2 * j], v[(2 * j + 9) & 15]));
    }
    const uint32_t mx = __vmaxu2(max3(max3(rmin[0], rmin[1], rmin[2]), max3(rmin[3], rmin[4], rmin[5]), rmin[6]), rmin[7]);
    const uint32_t mn = __vminu2(min3(min3(rmax[0], rmax[1], rmax[2]), min3(rmax[3], rmax[4], rmax[5]), rmax[6]), rmax[7]);
#else
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

#endif

    // per u16 lane: s = max(c - mn, mx - c); result = max(s - t, 0).  The two differences are taken with a
    // +256 bias per lane (no borrow between lanes: every lane stays in [1, 511]) as 32-bit IMADs whose
    // multipliers are kernel parameters -- FMA pipe; the ALU pipe is this kernel's bottleneck.
    const uint32_t cb = imad32(c, M.one, M.bias);              // c + 256
    const uint32_t cn = imad32(c, M.minus1, M.bias);           // 256 - c
    const uint32_t a = imad32(mn, M.minus1, cb);               // c - mn + 256
    const uint32_t bq = imad32(mx, M.one, cn);                 // mx - c + 256
    return __viaddmax_s16x2_relu(__vmaxu2(a, bq), neg_thr, 0u);
}

// NMS rows without a single positive score are skipped by the whole warp (one vote per row): on camera-like
// frames most rows of a tile are such rows; on the dense synthetic frames almost none are.
#ifndef ASLAM_FAST_NMS_SKIP
#define ASLAM_FAST_NMS_SKIP 1
#endif
// ASLAM_FAST_NMS_QUEUE: the NMS runs only on the pixel pairs that hold a positive score (a second, in-place
// compaction of the phase-B queue), on dense warps, instead of on every pair of every row.
#ifndef ASLAM_FAST_NMS_QUEUE
#define ASLAM_FAST_NMS_QUEUE 1
#endif
#if ASLAM_FAST_NMS_QUEUE && !ASLAM_FAST_COMPACT
#error "ASLAM_FAST_NMS_QUEUE needs the two-phase score (ASLAM_FAST_COMPACT)"
#endif
#if ASLAM_FAST_NMS_SKIP
#define NMS_ROW_HAS_SCORES(word) __any_sync(0xffffffffu, (word) != 0u)
#else
#define NMS_ROW_HAS_SCORES(word) true
#endif

__global__ void __launch_bounds__(256, ASLAM_FAST_MINB)
fast_kernel(OrbGeom g, OrbBuffers b, const FastMul M)
{
    __shared__ __align__(16) uint16_t AB[2][FT_PIX_H][FA_PITCH];
    uint16_t (*const A)[FA_PITCH] = AB[0], (*const Bc)[FA_PITCH] = AB[1];
    static_assert(sizeof(AB[0]) == FA_COPY_BYTES, "copy stride");
    __shared__ __align__(16) uint32_t S[FT_SCORE_H][FT_SCORE_W / 2];      // two u16 scores per word
    __shared__ int warp_tot[8];
#if ASLAM_FAST_COMPACT
    __shared__ FastQ Q[8][FT_SCORE_H / 8 * 32 + 2];         // per warp: queued (row << 5 | lane) of pairs to score (+ a sink slot)
#endif

    const int tile = blockIdx.x, frame = blockIdx.y, tid = threadIdx.x;
    // tile -> (level, tx, ty) from a table built on the host (level | tx << 4 | ty << 16) instead of an 8-way level search
    // and an integer division per CTA: 0.845 -> 0.832 ms in the A/B of round 2
    const uint32_t tt = __ldg(b.tile_tab + tile);
    const int l = (int)(tt & 15u), tx = (int)((tt >> 4) & 0xfffu), ty = (int)(tt >> 16);
    const LevelGeom L = g.lv[l];
    const int sx0 = ((g.edge - 1) & ~1) + tx * FT_OUT_W, sy0 = g.edge - 1 + ty * FT_OUT_H;
    const int gx0 = (sx0 - 3) & ~3;          // first staged column (aligned), staged row 0 is y = sy0-3
    const int g0 = sx0 - gx0;                // 4 or 6
    const uint8_t *img = b.pyr + (size_t)frame * g.pyr_stride + L.off;

    // stage 38 rows x 19 words; each word becomes four u16 in A, and (shifted by one pixel) in Bc
    // (passes of 13 rows x 19 words: the word column, its clamp and the shared-memory slot are per-thread
    // constants, a pass costs two loads, four PRMTs and two stores)
    // rows below y = h - edge feed no output (the NMS of row y reads y-1 .. y+1 and emits y < h - edge only)
    const int r_end = min(FT_SCORE_H, L.h - g.edge - sy0 + 1);
    {
        constexpr int PASSES = (FT_PIX_H + 12) / 13;
        const int rr = tid / FA_WORDS, jj = tid - rr * FA_WORDS;
        const int x = min(gx0 + 4 * jj, L.pitch - 4);
        const int xn = min(x + 4, L.pitch - 1);
        if (rr < 13) {
#pragma unroll
            for (int k = 0; k < PASSES; ++k) {
                const int r = rr + 13 * k;
                if (13 * (k + 1) <= FT_PIX_H || r < FT_PIX_H) {
                    ASLAM_BOUNDS(r < FT_PIX_H && 4 * jj + 3 < FA_PITCH && x >= 0 && x + 4 <= L.pitch && xn < L.pitch && sy0 - 3 + r >= 0);
                    const uint8_t *row = img + (size_t)min(sy0 - 3 + r, L.h - 1) * L.pitch;
                    const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(row + x));
                    const uint32_t nx = __ldg(row + xn);
                    uint2 a, c;
                    a.x = __byte_perm(w, 0, 0x4140); a.y = __byte_perm(w, 0, 0x4342);
                    c.x = __byte_perm(w, 0, 0x4241); c.y = __byte_perm(w, nx, 0x7473);
                    *reinterpret_cast<uint2 *>(&A[r][4 * jj]) = a;
                    *reinterpret_cast<uint2 *>(&Bc[r][4 * jj]) = c;       // Bc[i] = pixel i+1
                }
            }
        }
    }
#if ASLAM_FAST_ZERO_S
    // the score block starts out as zeros (two 16-byte stores per thread here instead of a store with its addressing per
    // warp and row in phase A); phase B fills in the pairs that pass the quick test
    {
        uint4 *sz = reinterpret_cast<uint4 *>(&S[0][0]);
        static_assert(sizeof(S) % sizeof(uint4) == 0, "score block is a whole number of 16-byte stores");
#pragma unroll
        for (int i = tid; i < (int)(sizeof(S) / sizeof(uint4)); i += 256) sz[i] = make_uint4(0u, 0u, 0u, 0u);
    }
#endif
    __syncthreads();

    const int warp = tid >> 5, p = tid & 31;
    const int gp = g0 / 2 + p;
    // cv::FAST evaluates pixels [3, w-3) x [3, h-3) only; x >= 3 and y >= 3 always hold here (edge >= 19).
    // The score block holds s - t (>= 1 for corners, 0 otherwise): same order for the NMS, and the
    // response max(A,B) - 1 is restored at emit time.  neg_thr: -t per valid u16 lane, -30000 otherwise.
    const int px_x = sx0 + 2 * p;
    const uint32_t neg_thr = ((uint32_t)(px_x < L.w - 3 ? -g.fast_thr - 256 : -30000) & 0xffffu) |
                             ((uint32_t)(px_x + 1 < L.w - 3 ? -g.fast_thr - 256 : -30000) << 16);     // incl. the bias below
#if ASLAM_FAST_NMS_QUEUE
    int n2 = 0;                 // pairs of this warp's rows with a positive score (queued in Q[warp] after phase B)
#endif
#if ASLAM_FAST_COMPACT
    // Two phases per warp.  A: a cheap necessary test -- a 9-arc contains one pixel of every opposite ring pair,
    // so a bright (dark) corner needs max (min) of each pair above c + t (below c - t); four of the eight
    // pairs are tested, 14 ALU instructions instead of 98.  Pairs that pass are queued as (row, lane).
    // B: the full score runs on dense warps of queued pairs only (35 % of the pairs on the synthetic frames,
    // a few per cent on natural images); everything else scores 0, exactly as the full test would find.
    {
        FastQ *queue = Q[warp];
        const uint32_t kq = (uint32_t)(32767 - g.fast_thr) * 0x00010001u;
        const uint32_t colbits = (px_x < L.w - 3 ? 0x8000u : 0u) | (px_x + 1 < L.w - 3 ? 0x80000000u : 0u);
        int nq = 0;
        // a warp owns FT_SCORE_H / 8 consecutive rows, so its queue -- and with it the emitted order -- is row-major.
        // The loop runs on the queue entry itself ((r << 5) | lane, + 32 per row) and on one shared-memory address.
        const int r_first = warp * (FT_SCORE_H / 8), r_last = min(r_end, (warp + 1) * (FT_SCORE_H / 8));
        const uint32_t val_end = (uint32_t)r_last << 5;
        const uint32_t lt_mask = (1u << p) - 1u;
        uint32_t at = smem_u32(&AB[0][r_first][0]) + 4u * (uint32_t)(gp - 2);
#pragma unroll FAST_A_UNROLL
        for (uint32_t val = ((uint32_t)r_first << 5) | (uint32_t)p; val < val_end; val += 32u, at += FA_PITCH * 2) {
            const uint32_t c = ring_pair<0, 0>(at);
            const uint32_t p0 = ring_pair<0, 3>(at), p8 = ring_pair<0, -3>(at);
            const uint32_t p4 = ring_pair<3, 0>(at), p12 = ring_pair<-3, 0>(at);
            const uint32_t p2 = ring_pair<2, 2>(at), p10 = ring_pair<-2, -2>(at);
            const uint32_t p6 = ring_pair<2, -2>(at), p14 = ring_pair<-2, 2>(at);
            const uint32_t mb = __vminu2(min3(__vmaxu2(p0, p8), __vmaxu2(p4, p12), __vmaxu2(p2, p10)), __vmaxu2(p6, p14));
            const uint32_t md = __vmaxu2(max3(__vminu2(p0, p8), __vminu2(p4, p12), __vminu2(p2, p10)), __vminu2(p6, p14));
            // per lane: mb - c - t + 32767 and c - t - md + 32767 (lanes stay positive: no borrow); bit 15 <=> > 0
            const uint32_t ub = imad32(c, M.minus1, imad32(mb, M.one, kq));
            const uint32_t ud = imad32(md, M.minus1, imad32(c, M.one, kq));
            const bool pass = ((ub | ud) & colbits) != 0;
            ASLAM_BOUNDS((int)(val >> 5) >= 0 && (int)(val >> 5) + 6 < FT_PIX_H && gp - 2 >= 0 && 2 * (gp + 2) < FA_PITCH && nq <= FT_SCORE_H / 8 * 32);
#if !ASLAM_FAST_ZERO_S
            S[val >> 5][p] = 0u;
#endif
            const uint32_t bal = __ballot_sync(0xffffffffu, pass);
            queue[pass ? nq + __popc(bal & lt_mask) : FT_SCORE_H / 8 * 32] = (FastQ)val;      // branch-free: slot 256 is a sink
            nq += __popc(bal);
        }
        __syncwarp();
#if ASLAM_FAST_NMS_QUEUE
        n2 = 0;
#endif
        const uint32_t ab_base = smem_u32(&AB[0][0][0]) + 4u * (uint32_t)(g0 / 2 - 2);
#pragma unroll 1
        for (int i0 = 0; i0 < nq; i0 += 32) {
            const int i = i0 + p;
            const int id = queue[min(i, nq - 1)];
            const int rr = id >> 5, pp = id & 31;
            const uint32_t thr = __shfl_sync(0xffffffffu, neg_thr, pp);
            const uint32_t sc = fast_score_pair(ab_base + (uint32_t)rr * (FA_PITCH * 2) + (uint32_t)pp * 4u, thr, M);
            ASLAM_BOUNDS(rr >= 0 && rr < FT_SCORE_H && pp >= 0 && pp < 32 && n2 <= i0);
            if (i < nq) S[rr][pp] = sc;
#if ASLAM_FAST_NMS_QUEUE
            // keep the pairs with a positive score, in place: n2 <= i0, so a slot written here was read already
            const bool nz = i < nq && sc != 0u;
            const uint32_t bal = __ballot_sync(0xffffffffu, nz);
            __syncwarp();
            if (nz) queue[n2 + __popc(bal & ((1u << p) - 1u))] = (FastQ)id;
            n2 += __popc(bal);
#endif
        }
    }
#else
#pragma unroll 1
    for (int r = warp; r < r_end; r += 8) S[r][p] = fast_score_pair(smem_u32(&AB[0][r][0]) + 4u * (uint32_t)(gp - 2), neg_thr, M);
#endif
    __syncthreads();

    // 3x3 strict non-max suppression in the same pair layout.  Warp w owns RPW consecutive rows, so the
    // emitted order is row-major; positions come from ballots (no atomics, deterministic).
    // For a column x: C3 = max of the three rows (centre row included), V = max of rows above/below.
    // neighbours(x) = max(V[x], C3[x-1], C3[x+1]); columns x-1 / x+1 of a pair come from the lanes beside it.
    const uint8_t *mrow = b.mask ? b.mask + (size_t)frame * g.pyr_stride + L.off : nullptr;
#if ASLAM_FAST_NMS_QUEUE
    // One lane = one queued pair (row rr, pair column pp): the 3x3 block of score words around it comes straight from
    // S (no shuffles), the arithmetic is the row-wise form's.  Results are parked per warp in the pixel tile's memory
    // (dead after the barrier above) in queue order = row-major, then copied out behind the warps before.
    uint32_t *stage = reinterpret_cast<uint32_t *>(&A[0][0]) + warp * (FT_SCORE_H / 8 * 32);
    static_assert(sizeof(uint32_t) * FT_SCORE_H * 32 <= sizeof(uint16_t) * FT_PIX_H * FA_PITCH, "candidate staging fits the pixel tile");
    const uint32_t resp_bias = (uint32_t)(g.fast_thr - 1);
    int local = 0;
    {
        const FastQ *queue = Q[warp];
        // which of a pair column's two pixels may be emitted depends on the column only: lane p keeps the mask of column p
        // and a queued pair fetches its column's by a shuffle (it was ~30 ALU instructions per trip when recomputed)
        const uint32_t col_ok = ((p >= 1 && px_x >= g.edge && px_x < L.w - g.edge) ? 0xffffu : 0u) |
                                ((p <= 30 && px_x + 1 >= g.edge && px_x + 1 < L.w - g.edge) ? 0xffff0000u : 0u);
        const int r_hi = min(FT_OUT_H, L.h - g.edge - sy0 - 1);      // last score row that is an output row of the image
#pragma unroll 1
        for (int i0 = 0; i0 < n2; i0 += 32) {
            const int i = i0 + p;
            const int id = queue[min(i, n2 - 1)];
            const int r = id >> 5, pp = id & 31;
            const uint32_t ok = __shfl_sync(0xffffffffu, col_ok, pp);
            uint32_t e = 0;
            if (i < n2 && r >= 1 && r <= r_hi) {
                const int pl = max(pp - 1, 0), pr = min(pp + 1, FT_SCORE_W / 2 - 1);     // clamped words feed masked lanes only
                const uint32_t up = S[r - 1][pp], cur = S[r][pp], dn = S[r + 1][pp];
                const uint32_t V = __vmaxu2(up, dn), C3 = max3(up, cur, dn);
                const uint32_t Lw = max3(S[r - 1][pl], S[r][pl], S[r + 1][pl]), Rw = max3(S[r - 1][pr], S[r][pr], S[r + 1][pr]);
                const uint32_t nb = max3(V, __byte_perm(Lw, C3, 0x5432), __byte_perm(C3, Rw, 0x5432));
                const int x0 = sx0 + 2 * pp, y = sy0 + r;
                const uint32_t d = __vimax_s16x2_relu(__vsub2(cur, nb), 0u) & ok;
                if (d) {
                    const bool second = (d & 0xffffu) == 0;
                    const int x = x0 + (second ? 1 : 0);
                    const uint32_t sc = second ? cur >> 16 : cur & 0xffffu;
                    if (!(mrow && __ldg(mrow + (size_t)y * L.pitch + x) == 0))
                        e = ((sc + resp_bias) << 24) | ((uint32_t)y << 12) | (uint32_t)x;
                }
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, e != 0);
            ASLAM_BOUNDS(local + 32 <= FT_SCORE_H / 8 * 32 + 32 && (e == 0 || (((e >> 12) & 0xfffu) < (uint32_t)L.h && (e & 0xfffu) < (uint32_t)L.w)));
            if (e) stage[local + __popc(bal & ((1u << p) - 1u))] = e;
            local += __popc(bal);
        }
    }
    if (p == 0) warp_tot[warp] = local;
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int t = warp_tot[i];
        if (i < warp) base += t;
        total += t;
    }
    uint32_t *out = b.cand + ((size_t)frame * g.tiles_per_frame + tile) * FT_CAP;
    ASLAM_BOUNDS(total <= FT_CAP && base + local <= FT_CAP);
    for (int i = p; i < local; i += 32) out[base + i] = stage[i];
    if (tid == 0) b.cand_cnt[(size_t)frame * g.tiles_per_frame + tile] = (uint32_t)total;
#else
    constexpr int RPW = FT_SCORE_H / 8;
    uint32_t found[RPW];
    int off[RPW];
    int local = 0;
    const uint32_t col_ok = ((p >= 1 && px_x >= g.edge && px_x < L.w - g.edge) ? 0xffffu : 0u) |
                            ((p <= 30 && px_x + 1 >= g.edge && px_x + 1 < L.w - g.edge) ? 0xffff0000u : 0u);
    const uint32_t resp_bias = (uint32_t)(g.fast_thr - 1);
#pragma unroll
    for (int k = 0; k < RPW; ++k) {
        const int r = 1 + RPW * warp + k;
        uint32_t e = 0;
        if (r <= FT_OUT_H && sy0 + r < L.h - g.edge && NMS_ROW_HAS_SCORES(S[r][p])) {
            const uint32_t up = S[r - 1][p], cur = S[r][p], dn = S[r + 1][p];
            const uint32_t V = __vmaxu2(up, dn), C3 = max3(up, cur, dn);
            const uint32_t Lw = __shfl_up_sync(0xffffffffu, C3, 1), Rw = __shfl_down_sync(0xffffffffu, C3, 1);
            const uint32_t nb = max3(V, __byte_perm(Lw, C3, 0x5432), __byte_perm(C3, Rw, 0x5432));
            const int y = sy0 + r;
            // strict maximum per lane: cur - nb > 0; adjacent pixels are never both maxima
            uint32_t d = __vimax_s16x2_relu(__vsub2(cur, nb), 0u) & col_ok;
            if (d) {
                const bool second = (d & 0xffffu) == 0;
                const int x = px_x + (second ? 1 : 0);
                const uint32_t sc = second ? cur >> 16 : cur & 0xffffu;
                if (!(mrow && __ldg(mrow + (size_t)y * L.pitch + x) == 0))
                    e = ((sc + resp_bias) << 24) | ((uint32_t)y << 12) | (uint32_t)x;
            }
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, e != 0);
        found[k] = e;
        off[k] = local + __popc(bal & ((1u << p) - 1u));
        local += __popc(bal);
    }
    if (p == 0) warp_tot[warp] = local;
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int t = warp_tot[i];
        if (i < warp) base += t;
        total += t;
    }
    uint32_t *out = b.cand + ((size_t)frame * g.tiles_per_frame + tile) * FT_CAP;
#pragma unroll
    for (int k = 0; k < RPW; ++k)
        if (found[k]) out[base + off[k]] = found[k];
    if (tid == 0) b.cand_cnt[(size_t)frame * g.tiles_per_frame + tile] = (uint32_t)total;
#endif
}

int launch_fast(aslam_ctx *ctx, const OrbGeom &g, const OrbBuffers &b, int nframes)
{
    if (g.tiles_per_frame == 0) return ASLAM_OK;
    dim3 grid(g.tiles_per_frame, nframes);
    ProfScope ps(ctx, PK_FAST);
    fast_kernel<<<grid, 256, 0, ctx->stream>>>(g, b, FastMul{1u, 0xffffffffu, 0x01000100u});
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

// v4: one launch covers every level (tile table in LevelGeom::blur_tile_off); a CTA produces a
// BL_W x BL_H tile (128x58; 64x122 and 32x250 measured slower) from BL_H + 6 staged rows.
//  * Row pass: a thread takes 8 outputs of TWO adjacent rows.  It reads the 16 aligned bytes around
//    them straight from global memory (L1-resident), converts with the exponent trick (u8 -> f32
//    without the conversion unit) and runs the 7 taps on packed f32x2 lanes (FFMA2: one lane per
//    row), writing float row results to shared memory.
//  * Column pass: 4 columns x 4 rows per thread, packed f32x2 over adjacent columns; rounding to
//    nearest even by the 1.5*2^23 trick, whose low byte IS the pixel (no clamp: the taps sum to
//    1 within float rounding and inputs are in [0,255]); one PRMT tree packs 4 pixels per store.
//  * Columns at or beyond row_fused_end / col_fused_end (the build's scalar tail) take a scalar
//    path with separately rounded products.  (ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2,
//    so packed arithmetic is used only where fusion is the wanted behaviour.)

__device__ __forceinline__ uint32_t u8_magic(uint32_t word, int k)
{
    return __byte_perm(word, 0x4B000000u, 0x7650 + k);      // bits of 8388608.0f + byte k
}

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }

// Row pass of one task: 8 outputs at columns x..x+7 of two rows.  SLOW: the 14-byte window crosses the
// image border (reflect per byte); FUSED: taps accumulate with FFMA2, else separately rounded products.
template <bool SLOW, bool FUSED>
__device__ __forceinline__ void blur_row_task(const uint8_t *__restrict__ rowa, const uint8_t *__restrict__ rowb, int x, int w,
                                              const BlurTaps &T, float *__restrict__ outa, float *__restrict__ outb)
{
    float2 v[14];       // .x = first row, .y = second row; v[k] = pixel x-3+k
    if (!SLOW) {
        const uint32_t *pa = reinterpret_cast<const uint32_t *>(rowa + x - 4), *pb = reinterpret_cast<const uint32_t *>(rowb + x - 4);
        const uint32_t aw[4] = {__ldg(pa), __ldg(pa + 1), __ldg(pa + 2), __ldg(pa + 3)};
        const uint32_t bw[4] = {__ldg(pb), __ldg(pb + 1), __ldg(pb + 2), __ldg(pb + 3)};
        const float2 M = f2(-8388608.0f, -8388608.0f);
#pragma unroll
        for (int k = 0; k < 14; ++k) {
            const int byte = k + 1;
            v[k] = __fadd2_rn(f2(__uint_as_float(u8_magic(aw[byte >> 2], byte & 3)), __uint_as_float(u8_magic(bw[byte >> 2], byte & 3))), M);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 14; ++k) {
            const int xx = reflect101(x - 3 + k, w);
            v[k] = f2((float)__ldg(rowa + xx), (float)__ldg(rowb + xx));
        }
    }
    float2 r[8];
    if (FUSED) {
        const float2 K0 = f2(T.k[0], T.k[0]), K1 = f2(T.k[1], T.k[1]), K2 = f2(T.k[2], T.k[2]), K3 = f2(T.k[3], T.k[3]);
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            float2 s = __fmul2_rn(K0, v[o]);
            s = __ffma2_rn(K1, v[o + 1], s);
            s = __ffma2_rn(K2, v[o + 2], s);
            s = __ffma2_rn(K3, v[o + 3], s);
            s = __ffma2_rn(K2, v[o + 4], s);
            s = __ffma2_rn(K1, v[o + 5], s);
            s = __ffma2_rn(K0, v[o + 6], s);
            r[o] = s;
        }
    } else {
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            float sa = __fmul_rn(T.k[0], v[o].x), sb = __fmul_rn(T.k[0], v[o].y);
#pragma unroll
            for (int k = 1; k < 7; ++k) {
                const float kk = T.k[k < 4 ? k : 6 - k];
                sa = __fadd_rn(sa, __fmul_rn(kk, v[o + k].x));
                sb = __fadd_rn(sb, __fmul_rn(kk, v[o + k].y));
            }
            r[o] = f2(sa, sb);
        }
    }
    *reinterpret_cast<float4 *>(outa) = make_float4(r[0].x, r[1].x, r[2].x, r[3].x);
    *reinterpret_cast<float4 *>(outa + 4) = make_float4(r[4].x, r[5].x, r[6].x, r[7].x);
    *reinterpret_cast<float4 *>(outb) = make_float4(r[0].y, r[1].y, r[2].y, r[3].y);
    *reinterpret_cast<float4 *>(outb + 4) = make_float4(r[4].y, r[5].y, r[6].y, r[7].y);
}

// Column pass of one task: 4 columns x up to 4 rows from 10 staged rows.
template <bool FUSED>
__device__ __forceinline__ void blur_col_task(const float (*R)[BL_W], int c4, int r0, int nrows, const BlurTaps &T, uint8_t *__restrict__ dst, int pitch)
{
    float4 w[10];
#pragma unroll
    ASLAM_BOUNDS(r0 >= 0 && r0 < BL_ROWS && c4 >= 0 && c4 + 4 <= BL_W);
    for (int k = 0; k < 10; ++k) w[k] = *reinterpret_cast<const float4 *>(&R[min(r0 + k, BL_ROWS - 1)][c4]);
    const float2 K0 = f2(T.k[0], T.k[0]), K1 = f2(T.k[1], T.k[1]), K2 = f2(T.k[2], T.k[2]), K3 = f2(T.k[3], T.k[3]);
    const float2 MG = f2(12582912.0f, 12582912.0f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j >= nrows) break;
        uint32_t b0, b1, b2, b3;         // float bit patterns whose low byte is the rounded pixel
        if (FUSED) {
            float2 sl = __fmul2_rn(K3, f2(w[j + 3].x, w[j + 3].y)), sh = __fmul2_rn(K3, f2(w[j + 3].z, w[j + 3].w));
            sl = __ffma2_rn(K2, __fadd2_rn(f2(w[j + 4].x, w[j + 4].y), f2(w[j + 2].x, w[j + 2].y)), sl);
            sh = __ffma2_rn(K2, __fadd2_rn(f2(w[j + 4].z, w[j + 4].w), f2(w[j + 2].z, w[j + 2].w)), sh);
            sl = __ffma2_rn(K1, __fadd2_rn(f2(w[j + 5].x, w[j + 5].y), f2(w[j + 1].x, w[j + 1].y)), sl);
            sh = __ffma2_rn(K1, __fadd2_rn(f2(w[j + 5].z, w[j + 5].w), f2(w[j + 1].z, w[j + 1].w)), sh);
            sl = __ffma2_rn(K0, __fadd2_rn(f2(w[j + 6].x, w[j + 6].y), f2(w[j].x, w[j].y)), sl);
            sh = __ffma2_rn(K0, __fadd2_rn(f2(w[j + 6].z, w[j + 6].w), f2(w[j].z, w[j].w)), sh);
            sl = __fadd2_rn(sl, MG); sh = __fadd2_rn(sh, MG);
            b0 = __float_as_uint(sl.x); b1 = __float_as_uint(sl.y); b2 = __float_as_uint(sh.x); b3 = __float_as_uint(sh.y);
        } else {
            const float *c = reinterpret_cast<const float *>(&w[j + 3]);
            uint32_t bb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float s = __fmul_rn(T.k[3], c[i]);
#pragma unroll
                for (int k = 1; k <= 3; ++k) {
                    const float pr = __fadd_rn(reinterpret_cast<const float *>(&w[j + 3 + k])[i], reinterpret_cast<const float *>(&w[j + 3 - k])[i]);
                    s = __fadd_rn(s, __fmul_rn(T.k[3 - k], pr));
                }
                bb[i] = __float_as_uint(__fadd_rn(s, 12582912.0f));
            }
            b0 = bb[0]; b1 = bb[1]; b2 = bb[2]; b3 = bb[3];
        }
        const uint32_t p01 = __byte_perm(b0, b1, 0x0040), p23 = __byte_perm(b2, b3, 0x0040);
        *reinterpret_cast<uint32_t *>(dst + (size_t)j * pitch) = __byte_perm(p01, p23, 0x5410);
    }
}

#ifndef ASLAM_BLUR_MINB
#define ASLAM_BLUR_MINB 4
#endif
__global__ void __launch_bounds__(256, ASLAM_BLUR_MINB)
blur_kernel(OrbGeom g, OrbBuffers b, BlurTaps T)
{
    __shared__ __align__(16) float R[BL_ROWS][BL_W];
    const int tile = blockIdx.x, tid = threadIdx.x;
    int l = 0;
#pragma unroll
    for (int i = 1; i < ORB_MAX_LEVELS; ++i)
        if (i < g.nbuilt && tile >= g.lv[i].blur_tile_off) l = i;
    const LevelGeom L = g.lv[l];
    const int tl = tile - L.blur_tile_off;
    const int btx = (L.w + BL_W - 1) / BL_W;
    const int ty = tl / btx, tx = tl - ty * btx;
    const int x0 = tx * BL_W, y0 = ty * BL_H;
    const uint8_t *img = b.pyr + (size_t)blockIdx.y * g.pyr_stride + L.off;
    uint8_t *dst = b.blur + (size_t)blockIdx.y * g.pyr_stride + L.off;

    // Column groups of 8 of this tile: [0, ng) lie inside the image; a group is "plain" when its 16-byte
    // window is inside the image and it is left of row_fused_end.  Plain groups go through a branch-free
    // main loop; the few others (left border, right border, unfused tail) are handled afterwards by a
    // handful of threads, so no full warp pays for the slow paths.
    const int ng = min(BL_W / 8, (L.w - x0 + 7) / 8);
    const int left = (x0 == 0 && ng > 0) ? 1 : 0;                         // group 0 of the first tile needs reflection
    const int lim = min(L.row_fused_end, L.w - 11);                       // plain <=> x >= 4 and x < lim
    const int gs = max(left, min(ng, max(0, (lim - x0 + 7) / 8)));        // first non-plain group on the right
    const int nspecial = left + (ng - gs);

    // ---- row pass, plain groups: 32 row pairs x 16 groups of 8 outputs
#pragma unroll 1
    for (int i = tid; i < (BL_ROWS / 2) * (BL_W / 8); i += 256) {
        const int rp = i / (BL_W / 8), gi = i - rp * (BL_W / 8);
        if (gi < left || gi >= gs) continue;
        const uint8_t *rowa = img + (size_t)reflect101(y0 - 3 + 2 * rp, L.h) * L.pitch;
        const uint8_t *rowb = img + (size_t)reflect101(y0 - 2 + 2 * rp, L.h) * L.pitch;
        ASLAM_BOUNDS(2 * rp + 1 < BL_ROWS && gi * 8 + 8 <= BL_W && x0 + gi * 8 >= 4 && x0 + gi * 8 + 12 <= L.pitch);
        blur_row_task<false, true>(rowa, rowb, x0 + gi * 8, L.w, T, &R[2 * rp][gi * 8], &R[2 * rp + 1][gi * 8]);
    }
    // ---- row pass, special groups
#pragma unroll 1
    for (int e = tid; e < (BL_ROWS / 2) * nspecial; e += 256) {
        const int rp = e / nspecial, si = e - rp * nspecial;
        const int gi = si < left ? 0 : gs + si - left;
        const int x = x0 + gi * 8;
        const uint8_t *rowa = img + (size_t)reflect101(y0 - 3 + 2 * rp, L.h) * L.pitch;
        const uint8_t *rowb = img + (size_t)reflect101(y0 - 2 + 2 * rp, L.h) * L.pitch;
        const bool slow = !(x >= 4 && x + 12 <= L.w), fused = x < L.row_fused_end;
        float *oa = &R[2 * rp][gi * 8], *ob = &R[2 * rp + 1][gi * 8];
        if (slow && fused) blur_row_task<true, true>(rowa, rowb, x, L.w, T, oa, ob);
        else if (slow) blur_row_task<true, false>(rowa, rowb, x, L.w, T, oa, ob);
        else blur_row_task<false, false>(rowa, rowb, x, L.w, T, oa, ob);
    }
    __syncthreads();

    // ---- column pass: 32 groups of 4 columns x 15 runs of 4 rows (the last run has 2); fused groups first
    const int nc = min(BL_W / 4, (L.w - x0 + 3) / 4);                     // column groups inside the image
    const int cf = min(nc, max(0, (L.col_fused_end - x0) / 4));           // [0, cf) fused, [cf, nc) unfused
#pragma unroll 1
    for (int t = tid; t < (BL_W / 4) * ((BL_H + 3) / 4); t += 256) {
        const int ci = t % (BL_W / 4), r0 = (t / (BL_W / 4)) * 4;
        if (ci >= cf || y0 + r0 >= L.h) continue;
        blur_col_task<true>(R, ci * 4, r0, min(4, min(BL_H - r0, L.h - y0 - r0)), T, dst + (size_t)(y0 + r0) * L.pitch + x0 + ci * 4, L.pitch);
    }
#pragma unroll 1
    for (int t = tid; t < (nc - cf) * ((BL_H + 3) / 4); t += 256) {
        const int ci = cf + t % (nc - cf), r0 = (t / (nc - cf)) * 4;
        if (y0 + r0 >= L.h) continue;
        blur_col_task<false>(R, ci * 4, r0, min(4, min(BL_H - r0, L.h - y0 - r0)), T, dst + (size_t)(y0 + r0) * L.pitch + x0 + ci * 4, L.pitch);
    }
}

// ---- round 2: the blur as a register march (the shipped kernel; blur_kernel above stays for A/B runs, ASLAM_BLUR_MARCH=0).
// One thread = 4 adjacent columns of a strip of rows, marched top to bottom two rows per step; nothing goes through
// shared memory.  Per step a thread
//   * reads the three aligned words around its columns in two source rows (the loads of the NEXT step are issued first),
//     patches them at the image's left/right border with per-thread PRMT selectors (BORDER_REFLECT_101; identity selectors
//     in the interior, so the instruction stream is the same for every thread),
//   * converts 10 bytes per row with the exponent trick and runs the 7 row taps on packed f32x2 lanes, lane .x = the
//     upper row, .y = the lower one: P = (R[2m], R[2m+1]) for each of its 4 columns,
//   * keeps the last four P and the three "odd" pairs Q = (R[2m-1], R[2m]) (two register moves per column and step), which
//     are exactly the operands of the column pass for the output rows centred on R[2m-3] and R[2m-2]:
//       K3*Q[m-1] + K2*(P[m-2] + P[m-1]) + K1*(Q[m-2] + Q[m]) + K0*(P[m-3] + P[m]),
//     again on packed lanes (.x = upper output row, .y = lower) -- 8 packed instructions per column for two pixels,
//   * rounds by the 1.5*2^23 trick, packs 4 pixels per row with 3 PRMTs and stores one word per row.
// The rings of P and Q are register names rotated by unrolling the march four steps.  Threads of the unfused column tail
// (x >= row_fused_end / col_fused_end, the split of the cv2 build) take product-then-sum forms that are still packed:
// s = fma(fmul(k, v), ONE, s) with ONE = 1.0f arriving as a kernel parameter rounds the product first and ptxas cannot
// contract it.  ~130 instructions per 8 pixels against ~270 in blur_kernel; the FMA pipe (70 packed instructions per step
// at one per two cycles) is the floor.
struct BlurPlan {
    int cta_off[ORB_MAX_LEVELS + 1];   // first CTA of each level within a frame
    int strip_h[ORB_MAX_LEVELS];       // rows per strip (even)
    int nstrips[ORB_MAX_LEVELS];
    // ASLAM_BM_SPLIT: a level's column groups g in [1, gs) are "bulk" -- no image border inside their 12-byte window, both
    // passes fused -- and come first, in CTAs of their own (bulk_ctas); group 0 and the groups from gs on follow.
    int bulk_ctas[ORB_MAX_LEVELS];
    int gs[ORB_MAX_LEVELS];
};
struct BlurTapsM { float k[4]; float one; };
// one warp per CTA: a finished warp frees its registers at once instead of waiting for three siblings with other strip
// lengths (0.287 -> 0.269 ms; 64 threads 0.282; bench.py 96.9 k -> 97.6 k frames/s)
#ifndef ASLAM_BM_THREADS
#define ASLAM_BM_THREADS 32
#endif
constexpr int BM_THREADS = ASLAM_BM_THREADS;

// Packed f32x2 values live in 64-bit registers from birth to death (inline PTX with "l" operands): held as float2 the
// halves are two 32-bit values to the register allocator, which then assembles an aligned pair with MOVs for every packed
// operand that crosses a step of the march (measured: 39 MOVs per step, 18 % of the executed instructions).
typedef unsigned long long pk2;
__device__ __forceinline__ pk2 pk_make(uint32_t lo, uint32_t hi) { pk2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ pk2 pk_dup(float v) { return pk_make(__float_as_uint(v), __float_as_uint(v)); }
__device__ __forceinline__ uint32_t pk_lo(pk2 v) { [[maybe_unused]] uint32_t lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); return lo; }
__device__ __forceinline__ uint32_t pk_hi(pk2 v) { [[maybe_unused]] uint32_t lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); return hi; }
__device__ __forceinline__ pk2 pk_fma(pk2 a, pk2 b, pk2 c) { pk2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ pk2 pk_mul(pk2 a, pk2 b) { pk2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ pk2 pk_add(pk2 a, pk2 b) { pk2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
// round(k*v) first, then round(. + s): ONE = 1.0f arrives as a kernel parameter, so ptxas cannot contract the pair
__device__ __forceinline__ pk2 pk_fma_unfused(pk2 k, pk2 v, pk2 s, pk2 one) { return pk_fma(pk_mul(k, v), one, s); }

#ifndef ASLAM_BM_DEPTH
#define ASLAM_BM_DEPTH 2        // source-row pairs in flight per thread (1, 2 or 4: the ring must divide the 4-step unrolling)
#endif
#ifndef ASLAM_BM_PF
#define ASLAM_BM_PF 0           // > 0: also prefetch the rows of step m + PF into L2
#endif
constexpr int BM_DEPTH = ASLAM_BM_DEPTH;
static_assert(BM_DEPTH == 1 || BM_DEPTH == 2 || BM_DEPTH == 4, "blur march prefetch depth");
struct BlurMarch {
    pk2 P[4][4], Q[4][4];              // [ring slot][column]
    uint32_t wa[BM_DEPTH][3], wb[BM_DEPTH][3];     // words of the two source rows of steps m .. m + BM_DEPTH - 1 (slot = step % BM_DEPTH)
    const uint8_t *col;                // level base + x0 (frame's pyramid)
    uint8_t *dp;                       // destination of the next output row pair
    int h, pitch, two_h_m2;
    uint32_t selL, selM, selR1, selR2;
    bool not_left, row_unf, col_unf;
    pk2 K0, K1, K2, K3, ONE;

    __device__ __forceinline__ const uint8_t *row_ptr(int v) const
    {
        int a = abs(v);
        a = max(min(a, two_h_m2 - a), 0);               // BORDER_REFLECT_101 (rows past 2h-2 belong to no output)
        ASLAM_BOUNDS(a >= 0 && a < h);
        return col + (unsigned long long)(unsigned)a * (unsigned)pitch;
    }
    __device__ __forceinline__ void load(int v, uint32_t (&w)[3]) const
    {
        const uint32_t *p = reinterpret_cast<const uint32_t *>(row_ptr(v));
        if (not_left) w[0] = __ldg(p - 1);
        w[1] = __ldg(p);
        w[2] = __ldg(p + 1);
    }
    __device__ __forceinline__ void prefetch_l2(int v) const
    {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(row_ptr(v)));
    }
};

template <int SLOT, bool OUT>
__device__ __forceinline__ void blur_march_step(BlurMarch &S, int v_next, int y_out)
{
    // borders: l' = columns x0-4..x0-1, m' = x0..x0+3, r' = x0+4..x0+7 after reflection (identity in the interior)
    constexpr int WS = SLOT % BM_DEPTH;
    pk2 v[10];
    {
        uint32_t (&wa)[3] = S.wa[WS], (&wb)[3] = S.wb[WS];
        const uint32_t la = __byte_perm(wa[0], wa[1], S.selL), lb = __byte_perm(wb[0], wb[1], S.selL);
        const uint32_t ma = __byte_perm(wa[1], wa[0], S.selM), mb = __byte_perm(wb[1], wb[0], S.selM);
        const uint32_t ra = __byte_perm(__byte_perm(wa[2], wa[1], S.selR1), wa[0], S.selR2);
        const uint32_t rb = __byte_perm(__byte_perm(wb[2], wb[1], S.selR1), wb[0], S.selR2);
        // the words are consumed: their registers take the rows of step m + BM_DEPTH, whose latency hides behind the
        // arithmetic of BM_DEPTH steps
        S.load(v_next, wa);
        S.load(v_next + 1, wb);
#if ASLAM_BM_PF > 0
        S.prefetch_l2(v_next + 2 * (ASLAM_BM_PF - BM_DEPTH));
        S.prefetch_l2(v_next + 2 * (ASLAM_BM_PF - BM_DEPTH) + 1);
#endif
        const uint32_t A[3] = {la, ma, ra}, B[3] = {lb, mb, rb};
        const pk2 M = pk_dup(-8388608.0f);
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            const int byte = k + 1;                     // v[k] = column x0 - 3 + k = byte k + 1 of the 12-byte window
            v[k] = pk_add(pk_make(u8_magic(A[byte >> 2], byte & 3), u8_magic(B[byte >> 2], byte & 3)), M);
        }
    }
    // row pass -> P[SLOT]
    if (!S.row_unf) {
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            pk2 s = pk_mul(S.K0, v[o]);
            s = pk_fma(S.K1, v[o + 1], s);
            s = pk_fma(S.K2, v[o + 2], s);
            s = pk_fma(S.K3, v[o + 3], s);
            s = pk_fma(S.K2, v[o + 4], s);
            s = pk_fma(S.K1, v[o + 5], s);
            s = pk_fma(S.K0, v[o + 6], s);
            S.P[SLOT][o] = s;
        }
    } else {
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            pk2 s = pk_mul(S.K0, v[o]);
            s = pk_fma_unfused(S.K1, v[o + 1], s, S.ONE);
            s = pk_fma_unfused(S.K2, v[o + 2], s, S.ONE);
            s = pk_fma_unfused(S.K3, v[o + 3], s, S.ONE);
            s = pk_fma_unfused(S.K2, v[o + 4], s, S.ONE);
            s = pk_fma_unfused(S.K1, v[o + 5], s, S.ONE);
            s = pk_fma_unfused(S.K0, v[o + 6], s, S.ONE);
            S.P[SLOT][o] = s;
        }
    }
    constexpr int S1 = (SLOT + 1) & 3, S2 = (SLOT + 2) & 3, S3 = (SLOT + 3) & 3;     // steps m-3, m-2, m-1
#pragma unroll
    for (int o = 0; o < 4; ++o) S.Q[SLOT][o] = pk_make(pk_hi(S.P[S3][o]), pk_lo(S.P[SLOT][o]));       // (R[2m-1], R[2m])

    if (OUT) {
        uint32_t bx[4], by[4];
        const pk2 MG = pk_dup(12582912.0f);
        if (!S.col_unf) {
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                pk2 s = pk_mul(S.K3, S.Q[S3][o]);
                s = pk_fma(S.K2, pk_add(S.P[S3][o], S.P[S2][o]), s);
                s = pk_fma(S.K1, pk_add(S.Q[SLOT][o], S.Q[S2][o]), s);
                s = pk_fma(S.K0, pk_add(S.P[SLOT][o], S.P[S1][o]), s);
                s = pk_add(s, MG);
                bx[o] = pk_lo(s); by[o] = pk_hi(s);
            }
        } else {
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                pk2 s = pk_mul(S.K3, S.Q[S3][o]);
                s = pk_fma_unfused(S.K2, pk_add(S.P[S3][o], S.P[S2][o]), s, S.ONE);
                s = pk_fma_unfused(S.K1, pk_add(S.Q[SLOT][o], S.Q[S2][o]), s, S.ONE);
                s = pk_fma_unfused(S.K0, pk_add(S.P[SLOT][o], S.P[S1][o]), s, S.ONE);
                s = pk_add(s, MG);
                bx[o] = pk_lo(s); by[o] = pk_hi(s);
            }
        }
        if (y_out < S.h) {
            ASLAM_BOUNDS(y_out >= 0);
            *reinterpret_cast<uint32_t *>(S.dp) = __byte_perm(__byte_perm(bx[0], bx[1], 0x0040), __byte_perm(bx[2], bx[3], 0x0040), 0x5410);
            if (y_out + 1 < S.h)
                *reinterpret_cast<uint32_t *>(S.dp + S.pitch) = __byte_perm(__byte_perm(by[0], by[1], 0x0040), __byte_perm(by[2], by[3], 0x0040), 0x5410);
        }
        S.dp += 2 * (size_t)S.pitch;
    }
}

#ifndef ASLAM_BM_SPLIT
#define ASLAM_BM_SPLIT 1        // 1: bulk column groups take blur_bulk_march (below); 0: every thread takes the general march
#endif
// The march of the bulk threads (interior columns, fused taps: > 95 % of a level).  Same arithmetic as the general march,
// different bookkeeping: the row pass still works on (upper row, lower row) lane pairs, but its LAST tap is two scalar
// FFMAs per column whose results are born in column-packed pairs C[row][h] = (R[row][2h], R[row][2h+1]).  The column pass
// then reads only aligned pairs of a ring of 8 rows (16 pairs; the general march keeps 16 P and 16 Q pairs and builds Q
// with register moves), has no border selectors and no fused / unfused branches, so nothing is copied at a join.
struct BlurBulk {
    pk2 C[8][2];
    uint32_t wa[BM_DEPTH][3], wb[BM_DEPTH][3];
    const uint8_t *col;
    uint8_t *dp;
    int h, pitch, two_h_m2;
    pk2 K0, K1, K2, K3;
    float k0;

    __device__ __forceinline__ const uint32_t *row_ptr(int v) const
    {
        int a = abs(v);
        a = max(min(a, two_h_m2 - a), 0);
        ASLAM_BOUNDS(a >= 0 && a < h);
        return reinterpret_cast<const uint32_t *>(col + (unsigned long long)(unsigned)a * (unsigned)pitch);
    }
    __device__ __forceinline__ void load(int v, uint32_t (&w)[3]) const
    {
        const uint32_t *p = row_ptr(v);
        w[0] = __ldg(p - 1); w[1] = __ldg(p); w[2] = __ldg(p + 1);
    }
};

template <int STEP, bool OUT>
__device__ __forceinline__ void blur_bulk_step(BlurBulk &S, int v_next, int y_out)
{
    constexpr int WS = STEP % BM_DEPTH;
    pk2 v[10];
    {
        const uint32_t A[3] = {S.wa[WS][0], S.wa[WS][1], S.wa[WS][2]}, B[3] = {S.wb[WS][0], S.wb[WS][1], S.wb[WS][2]};
        S.load(v_next, S.wa[WS]);
        S.load(v_next + 1, S.wb[WS]);
        const pk2 M = pk_dup(-8388608.0f);
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            const int byte = k + 1;
            v[k] = pk_add(pk_make(u8_magic(A[byte >> 2], byte & 3), u8_magic(B[byte >> 2], byte & 3)), M);
        }
    }
    uint32_t ra[4], rb[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        pk2 s = pk_mul(S.K0, v[o]);
        s = pk_fma(S.K1, v[o + 1], s);
        s = pk_fma(S.K2, v[o + 2], s);
        s = pk_fma(S.K3, v[o + 3], s);
        s = pk_fma(S.K2, v[o + 4], s);
        s = pk_fma(S.K1, v[o + 5], s);
        ra[o] = __float_as_uint(__fmaf_rn(S.k0, __uint_as_float(pk_lo(v[o + 6])), __uint_as_float(pk_lo(s))));
        rb[o] = __float_as_uint(__fmaf_rn(S.k0, __uint_as_float(pk_hi(v[o + 6])), __uint_as_float(pk_hi(s))));
    }
    constexpr int RA = (2 * STEP) & 7, RB = (2 * STEP + 1) & 7;        // ring slots of this step's two rows
    S.C[RA][0] = pk_make(ra[0], ra[1]); S.C[RA][1] = pk_make(ra[2], ra[3]);
    S.C[RB][0] = pk_make(rb[0], rb[1]); S.C[RB][1] = pk_make(rb[2], rb[3]);

    if (OUT) {
        const pk2 MG = pk_dup(12582912.0f);
        uint32_t px[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            constexpr int dummy = 0; (void)dummy;
            const int c = (2 * STEP - 3 + r) & 7;                      // centre row of output row r (compile-time after unrolling)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                pk2 s = pk_mul(S.K3, S.C[c][hh]);
                s = pk_fma(S.K2, pk_add(S.C[(c + 7) & 7][hh], S.C[(c + 1) & 7][hh]), s);
                s = pk_fma(S.K1, pk_add(S.C[(c + 6) & 7][hh], S.C[(c + 2) & 7][hh]), s);
                s = pk_fma(S.K0, pk_add(S.C[(c + 5) & 7][hh], S.C[(c + 3) & 7][hh]), s);
                s = pk_add(s, MG);
                px[r][2 * hh] = pk_lo(s); px[r][2 * hh + 1] = pk_hi(s);
            }
        }
        if (y_out < S.h) {
            ASLAM_BOUNDS(y_out >= 0);
            *reinterpret_cast<uint32_t *>(S.dp) = __byte_perm(__byte_perm(px[0][0], px[0][1], 0x0040), __byte_perm(px[0][2], px[0][3], 0x0040), 0x5410);
            if (y_out + 1 < S.h)
                *reinterpret_cast<uint32_t *>(S.dp + S.pitch) = __byte_perm(__byte_perm(px[1][0], px[1][1], 0x0040), __byte_perm(px[1][2], px[1][3], 0x0040), 0x5410);
        }
        S.dp += 2 * (size_t)S.pitch;
    }
}

__device__ __forceinline__ void blur_bulk_march(const uint8_t *col, uint8_t *dp, int h, int pitch, int ys, int nsteps, const BlurTapsM &T)
{
    BlurBulk S;
    S.h = h; S.pitch = pitch; S.two_h_m2 = 2 * h - 2;
    S.col = col; S.dp = dp;
    S.K0 = pk_dup(T.k[0]); S.K1 = pk_dup(T.k[1]); S.K2 = pk_dup(T.k[2]); S.K3 = pk_dup(T.k[3]);
    S.k0 = T.k[0];
#pragma unroll
    for (int d = 0; d < BM_DEPTH; ++d) {
        S.load(ys - 3 + 2 * d, S.wa[d]);
        S.load(ys - 2 + 2 * d, S.wb[d]);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) { S.C[r][0] = 0ull; S.C[r][1] = 0ull; }
    constexpr int VN = 2 * BM_DEPTH - 3;
    blur_bulk_step<0, false>(S, ys + VN, 0);
    blur_bulk_step<1, false>(S, ys + VN + 2, 0);
    blur_bulk_step<2, false>(S, ys + VN + 4, 0);
    int m = 3;
#pragma unroll 1
    for (;;) {
        blur_bulk_step<3, true>(S, ys + VN + 2 * m, ys + 2 * m - 6);
        if (++m >= nsteps) break;
        blur_bulk_step<0, true>(S, ys + VN + 2 * m, ys + 2 * m - 6);
        if (++m >= nsteps) break;
        blur_bulk_step<1, true>(S, ys + VN + 2 * m, ys + 2 * m - 6);
        if (++m >= nsteps) break;
        blur_bulk_step<2, true>(S, ys + VN + 2 * m, ys + 2 * m - 6);
        if (++m >= nsteps) break;
    }
}

#ifndef ASLAM_BM_MINB
#define ASLAM_BM_MINB (640 / ASLAM_BM_THREADS)     // 96 registers (8 bytes of spill) and 20 warps per SM; 16 warps (120 registers): 0.303 ms, 24 (80): 0.291
#endif
__global__ void __launch_bounds__(BM_THREADS, ASLAM_BM_MINB)
blur_march_kernel(OrbGeom g, OrbBuffers b, BlurTapsM T, BlurPlan plan)
{
    int l = 0;
#pragma unroll
    for (int i = 1; i < ORB_MAX_LEVELS; ++i)
        if (i < g.nbuilt && (int)blockIdx.x >= plan.cta_off[i]) l = i;
    const LevelGeom L = g.lv[l];
    const int ncg = (L.w + 3) >> 2;                                   // column groups per row
    const int hs = plan.strip_h[l];
    const int nsteps = hs / 2 + 3;                                    // steps m = 0 .. hs/2 + 2; outputs from m = 3
#if ASLAM_BM_SPLIT
    const int cta = (int)blockIdx.x - plan.cta_off[l], gs = plan.gs[l];
    int strip, x0;
    if (cta < plan.bulk_ctas[l]) {                                    // CTA-uniform
        const int nb = gs - 1;
        const int t = cta * BM_THREADS + (int)threadIdx.x;
        strip = t / nb;
        x0 = (1 + t - strip * nb) * 4;
        if (strip >= plan.nstrips[l]) return;
        ASLAM_BOUNDS(x0 >= 4 && x0 + 7 < L.w && x0 + 3 < L.row_fused_end && x0 + 3 < L.col_fused_end);
        const int ys = strip * hs;
        blur_bulk_march(b.pyr + (size_t)blockIdx.y * g.pyr_stride + L.off + x0,
                        b.blur + (size_t)blockIdx.y * g.pyr_stride + L.off + (size_t)ys * L.pitch + x0, L.h, L.pitch, ys, nsteps, T);
        return;
    }
    {
        const int nsp = ncg - (gs - 1);                               // group 0 and groups gs .. ncg-1
        const int t = (cta - plan.bulk_ctas[l]) * BM_THREADS + (int)threadIdx.x;
        strip = t / nsp;
        const int k = t - strip * nsp;
        x0 = (k == 0 ? 0 : gs + k - 1) * 4;
        if (strip >= plan.nstrips[l]) return;
    }
#else
    const int t = ((int)blockIdx.x - plan.cta_off[l]) * BM_THREADS + (int)threadIdx.x;
    const int strip = t / ncg, x0 = (t - strip * ncg) * 4;
    if (strip >= plan.nstrips[l]) return;
#endif
    const int ys = strip * hs;

    BlurMarch S;
    S.h = L.h; S.pitch = L.pitch; S.two_h_m2 = 2 * L.h - 2;
    S.col = b.pyr + (size_t)blockIdx.y * g.pyr_stride + L.off + x0;
    S.dp = b.blur + (size_t)blockIdx.y * g.pyr_stride + L.off + (size_t)ys * L.pitch + x0;
    S.not_left = x0 != 0;
    S.row_unf = x0 >= L.row_fused_end;
    S.col_unf = x0 >= L.col_fused_end;
    S.K0 = pk_dup(T.k[0]); S.K1 = pk_dup(T.k[1]); S.K2 = pk_dup(T.k[2]); S.K3 = pk_dup(T.k[3]);
    S.ONE = pk_dup(T.one);
    // border selectors.  Window byte j (j = 0..11) is column x0 - 4 + j; the pixel wanted there is reflect101 of it.
    {
        S.selL = x0 == 0 ? 0x5670u : 0x3210u;                         // columns -3, -2, -1 -> 3, 2, 1 (all in m)
        const int d = L.w - x0;                                       // columns x0 + j, j >= d, lie right of the image
        uint32_t sm = 0, s1 = 0, s2 = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int jj = j < d ? j : 2 * d - 2 - j;                       // source column relative to x0
            jj = max(jj, -4);
            if (j < 4) {                                              // m' = PRMT(m, l): m = bytes 0..3, l = bytes 4..7
                sm |= (uint32_t)(jj >= 0 ? jj : 8 + jj) << (4 * j);
            } else {                                                  // t = PRMT(r, m): r = 0..3, m = 4..7; r' = PRMT(t, l)
                const int q = j - 4;
                s1 |= (uint32_t)(jj >= 4 ? jj - 4 : jj >= 0 ? 4 + jj : 0) << (4 * q);
                s2 |= (uint32_t)(jj >= 0 ? q : 8 + jj) << (4 * q);
            }
        }
        S.selM = sm; S.selR1 = s1; S.selR2 = s2;
    }
#pragma unroll
    for (int d = 0; d < BM_DEPTH; ++d) {
        S.wa[d][0] = S.wb[d][0] = 0u;
        S.load(ys - 3 + 2 * d, S.wa[d]);
        S.load(ys - 2 + 2 * d, S.wb[d]);
    }
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int o = 0; o < 4; ++o) { S.P[s][o] = 0ull; S.Q[s][o] = 0ull; }

    // step m reads virtual rows ys - 3 + 2m, + 1 (fetched by step m - BM_DEPTH), fetches those of step m + BM_DEPTH and
    // writes rows ys + 2m - 6, + 1
    constexpr int VN = 2 * BM_DEPTH - 3;
    blur_march_step<0, false>(S, ys + VN, 0);
    blur_march_step<1, false>(S, ys + VN + 2, 0);
    blur_march_step<2, false>(S, ys + VN + 4, 0);
    int m = 3;
#pragma unroll 1
    for (;;) {
        blur_march_step<3, true>(S, ys + VN + 2 * m, ys + 2 * m - 6);
        if (++m >= nsteps) break;
        blur_march_step<0, true>(S, ys + VN + 2 * m, ys + 2 * m - 6);
        if (++m >= nsteps) break;
        blur_march_step<1, true>(S, ys + VN + 2 * m, ys + 2 * m - 6);
        if (++m >= nsteps) break;
        blur_march_step<2, true>(S, ys + VN + 2 * m, ys + 2 * m - 6);
        if (++m >= nsteps) break;
    }
}

#ifndef ASLAM_BLUR_MARCH
#define ASLAM_BLUR_MARCH 1      // 1: blur_march_kernel (registers only); 0: round 1's shared-memory tiles
#endif
int launch_blur(aslam_ctx *ctx, const OrbGeom &g, const OrbBuffers &b, int nframes)
{
    if (g.blur_tiles_per_frame == 0) return ASLAM_OK;
    double t[7], sum = 0;
    for (int i = 0; i < 7; ++i) { const double x = i - 3; t[i] = exp(-0.125 * x * x); sum += t[i]; }
#if ASLAM_BLUR_MARCH
    BlurTapsM T;
    for (int i = 0; i < 4; ++i) T.k[i] = (float)(t[i] / sum);
    T.one = 1.0f;
    // strips: tall for big batches (6 halo rows per strip), short when a few frames must still fill 148 SMs
    static const int env_h = [] { const char *e = getenv("ASLAM_BLUR_STRIP"); return e ? atoi(e) : 0; }();
    const int hmax = env_h > 0 ? env_h : nframes >= 32 ? 64 : nframes >= 8 ? 32 : 16;
    BlurPlan plan;
    int cta = 0;
    for (int l = 0; l < ORB_MAX_LEVELS; ++l) {
        plan.cta_off[l] = cta;
        plan.strip_h[l] = 2; plan.nstrips[l] = 0;
        plan.bulk_ctas[l] = 0; plan.gs[l] = 1;
        if (l >= g.nbuilt) continue;
        const LevelGeom &L = g.lv[l];
        const int ns = (L.h + hmax - 1) / hmax;
        const int hs = ((L.h + ns - 1) / ns + 1) & ~1;
        plan.nstrips[l] = (L.h + hs - 1) / hs;
        plan.strip_h[l] = hs;
        const int ncg = (L.w + 3) / 4;
#if ASLAM_BM_SPLIT
        // first group that is not bulk: its window touches the right border (4g + 7 >= w) or its taps are not all fused
        int gs = std::min(std::min(L.row_fused_end, L.col_fused_end) / 4, (L.w - 4) / 4);
        if (gs < 2) gs = 1;                                           // no bulk group
        plan.gs[l] = gs;
        const long long bulk = (long long)(gs - 1) * plan.nstrips[l], special = (long long)(ncg - (gs - 1)) * plan.nstrips[l];
        plan.bulk_ctas[l] = (int)((bulk + BM_THREADS - 1) / BM_THREADS);
        cta += plan.bulk_ctas[l] + (int)((special + BM_THREADS - 1) / BM_THREADS);
#else
        const long long threads = (long long)ncg * plan.nstrips[l];
        cta += (int)((threads + BM_THREADS - 1) / BM_THREADS);
#endif
    }
    plan.cta_off[ORB_MAX_LEVELS] = cta;
    dim3 grid(cta, nframes);
    ProfScope ps(ctx, PK_BLUR);
    blur_march_kernel<<<grid, BM_THREADS, 0, ctx->stream>>>(g, b, T, plan);
#else
    BlurTaps T;
    for (int i = 0; i < 4; ++i) T.k[i] = (float)(t[i] / sum);
    dim3 grid(g.blur_tiles_per_frame, nframes);
    ProfScope ps(ctx, PK_BLUR);
    blur_kernel<<<grid, 256, 0, ctx->stream>>>(g, b, T);
#endif
    ASLAM_LAUNCH_CHECK();
    return ASLAM_OK;
}

// ------------------------------------------------------------------ selection (A4, A5)
// One CTA per (frame, level): retainBest(2N) on the FAST score via a 256-bin histogram,
// Harris response for the survivors, retainBest(N) via a radix select on the float bits; ties
// with the cut-off are all kept, as KeyPointsFilter::retainBest does.  Survivor order is the
// candidate order (tile-major, row-major inside a tile): deterministic, no atomics on positions.

// HarrisResponses(blockSize 7, k 0.04f): integer sums, then a fixed unfused float expression.
#ifndef ASLAM_HARRIS_DP4A
#define ASLAM_HARRIS_DP4A 1     // 1: rows by three aligned word loads, Sobel sums by DP4A; 0: round 1's 81 byte loads
#endif
__device__ __forceinline__ int dp4a_px(uint32_t px, uint32_t w, int c)       // unsigned pixels x signed weights
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(px), "r"(w), "r"(c));
    return d;
}
__device__ __forceinline__ float harris_response(const uint8_t *__restrict__ img, int pitch, int x, int y)
{
    int a = 0, b = 0, c = 0;
#if ASLAM_HARRIS_DP4A
    // The 9 x 9 window (columns x-4 .. x+4) as, per row, the seven words win[k] = columns x-4+k .. x-1+k (k = 0..6):
    // three aligned loads (rows are 16-byte padded, so the third word lies inside the row whenever x+4 does), three
    // funnel shifts by the window's byte offset, five more by constants.  Sobel x / y at window column i = k+1 are dot
    // products of win[k] of the three rows with (-1,0,1,0) / (-2,0,2,0) and (-1,-2,-1,0) / (1,2,1,0): exact integers.
    const int xs = x - 4, sh = (xs & 3) * 8;
    const uint8_t *p = img + (size_t)(y - 4) * pitch + (xs & ~3);
    constexpr uint32_t WX1 = 0x000100ffu, WX2 = 0x000200feu, WY0 = 0x00fffeffu, WY2 = 0x00010201u;
    uint32_t r0[7], r1[7], r2[7];
    auto load_row = [&](int j, uint32_t (&win)[7]) {
        const uint32_t *q = reinterpret_cast<const uint32_t *>(p + (size_t)j * pitch);
        const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2);
        const uint32_t s0 = __funnelshift_r(w0, w1, sh), s1 = __funnelshift_r(w1, w2, sh), s2 = w2 >> sh;
        win[0] = s0; win[4] = s1;
        win[1] = __funnelshift_r(s0, s1, 8); win[2] = __funnelshift_r(s0, s1, 16); win[3] = __funnelshift_r(s0, s1, 24);
        win[5] = __funnelshift_r(s1, s2, 8); win[6] = __funnelshift_r(s1, s2, 16);
    };
    load_row(0, r0);
    load_row(1, r1);
#pragma unroll
    for (int dy = 0; dy < 7; ++dy) {
        load_row(dy + 2, r2);
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const int Ix = dp4a_px(r2[k], WX1, dp4a_px(r1[k], WX2, dp4a_px(r0[k], WX1, 0)));
            const int Iy = dp4a_px(r2[k], WY2, dp4a_px(r0[k], WY0, 0));
            a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
        }
#pragma unroll
        for (int k = 0; k < 7; ++k) { r0[k] = r1[k]; r1[k] = r2[k]; }
    }
#else
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
#endif
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

// Called by one full warp.  hist[256]: returns the largest bin B with sum(hist[B..255]) >= k and, through
// `above`, sum(hist[B+1..255]); -1 if the whole histogram holds fewer than k.
__device__ __forceinline__ int warp_kth_bin_from_top(const int *hist, int k, int &above)
{
    const int lane = threadIdx.x & 31;
    int h[8], c = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { h[i] = hist[lane * 8 + i]; c += h[i]; }
    int suf = c;                                   // inclusive suffix sum over lanes >= lane
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_down_sync(0xffffffffu, suf, o);
        if (lane + o < 32) suf += n;
    }
    const uint32_t ok = __ballot_sync(0xffffffffu, suf >= k);
    if (ok == 0) { above = 0; return -1; }
    const int owner = 31 - __clz(ok);              // suffix sums decrease with the lane index
    int bin = -1, ab = 0;
    if (lane == owner) {
        int acc = suf - c;                         // everything in higher lanes
#pragma unroll
        for (int i = 7; i >= 0; --i) {
            if (bin < 0 && acc + h[i] >= k) { bin = lane * 8 + i; ab = acc; }
            acc += h[i];
        }
    }
    above = __shfl_sync(0xffffffffu, ab, owner);
    return __shfl_sync(0xffffffffu, bin, owner);
}

// (round 1's form of this kernel -- per-tile counts through global memory, a scan over tiles, three barriers per radix pass,
// 512-thread CTAs only -- measured 0.157 ms per 256 frames against 0.148 for this one at 512 threads and 0.121 at 256.)
// K-th largest of the 32-bit keys key_of(i, key) (false = element i does not take part), i in [0, M), by
// 4 x 8-bit radix passes; at least K >= 1 elements must take part.  Called by the whole CTA.  Returns that key;
// k_left = how many of the elements equal to it are needed to reach K (>= 1).
// One barrier per pass: each pass has a histogram of its own (hist4[pass][256], zeroed together up front), the first two
// keys of a thread stay in registers, and EVERY warp finds the pass's bin from the finished histogram by itself, so the
// prefix and the remaining count never travel through shared memory.
template <int THREADS, class KeyFn>
__device__ __forceinline__ uint32_t block_kth_largest(int M, int K, KeyFn key_of, int *hist4, int &k_left)
{
    const int tid = threadIdx.x;
    for (int i = tid; i < 4 * 256; i += THREADS) hist4[i] = 0;
    uint32_t kreg[2] = {0u, 0u};
    bool kval[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int i = tid + j * THREADS;
        kval[j] = i < M && key_of(i, kreg[j]);
    }
    __syncthreads();
    uint32_t maskbits = 0, prefix = 0;
    int k = K;
#pragma unroll 1
    for (int pass = 3; pass >= 0; --pass) {
        int *h = hist4 + pass * 256;
#pragma unroll
        for (int j = 0; j < 2; ++j)
            if (kval[j] && (kreg[j] & maskbits) == prefix) atomicAdd(&h[(kreg[j] >> (8 * pass)) & 0xffu], 1);
        for (int i = tid + 2 * THREADS; i < M; i += THREADS) {
            uint32_t key;
            if (key_of(i, key) && (key & maskbits) == prefix) atomicAdd(&h[(key >> (8 * pass)) & 0xffu], 1);
        }
        __syncthreads();
        int above;
        const int bin = warp_kth_bin_from_top(h, k, above);       // exists: the prefix class holds >= k keys
        k -= above;
        prefix |= (uint32_t)bin << (8 * pass);
        maskbits |= 0xffu << (8 * pass);
    }
    k_left = k;
    __syncthreads();          // all warps are done with hist4 before a next call zeroes it
    return prefix;
}

// A warp's share of the candidate tiles [t_begin, t_end): visit(e, valid) is called for every slot in candidate order,
// 32 slots per call, with the same trip count for all lanes.  The first 128 slots of a tile are loaded together with its
// count (no dependent load; the slots exist whatever the count is) and the next tile's are requested before this one is
// visited.
template <class Visit>
__device__ __forceinline__ void for_each_candidate(const uint32_t *__restrict__ cand, const uint32_t *__restrict__ ccnt,
                                                   int t_begin, int t_end, int lane, Visit visit)
{
    if (t_begin >= t_end) return;
    uint32_t e[4], ne[4] = {0u, 0u, 0u, 0u};
    int n, nn = 0;
    {
        const uint32_t *ct = cand + (size_t)t_begin * FT_CAP + lane;
#pragma unroll
        for (int k = 0; k < 4; ++k) e[k] = __ldg(ct + 32 * k);
        n = (int)__ldg(ccnt + t_begin);
    }
    for (int t = t_begin; t < t_end; ++t) {
        if (t + 1 < t_end) {
            const uint32_t *ct = cand + (size_t)(t + 1) * FT_CAP + lane;
#pragma unroll
            for (int k = 0; k < 4; ++k) ne[k] = __ldg(ct + 32 * k);
            nn = (int)__ldg(ccnt + t + 1);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (32 * k < n) visit(e[k], lane + 32 * k < n);
        for (int i0 = 128; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            visit(i < n ? __ldg(cand + (size_t)t * FT_CAP + i) : 0u, i < n);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) e[k] = ne[k];
        n = nn;
    }
}

// BUCKET = the grid-bucketed second cull (g.grid_cell > 0), an extension cv::ORB does not have; its definition is
// in include/aslam_b200.h (aslam_orb_params::grid_cell) and DESIGN.md section 4a.
// Round 2: every warp owns a CONTIGUOUS range of the level's tiles, so the ordered compaction needs one exchange of 16
// warp totals instead of per-tile counts through global memory and a scan over the tiles; the cut-off bins are found by
// every warp for itself; the output pass hands every warp a contiguous range of survivors in the same way.
template <bool BUCKET, int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
select_kernel(OrbGeom g, OrbBuffers b)
{
    __shared__ int hist4[4 * 256];
    __shared__ int warp_sums[THREADS / 32];
    __shared__ int warp_sums2[THREADS / 32];

    const int l = blockIdx.x, frame = blockIdx.y, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NWARPS = THREADS / 32;
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
    uint32_t *surv = b.surv + tile0 * FT_CAP;
    float *hresp = b.hresp + tile0 * FT_CAP;
    const uint8_t *img = b.pyr + (size_t)frame * g.pyr_stride + L.off;
    const int tpw = (ntiles + NWARPS - 1) / NWARPS;
    const int t_begin = min(warp * tpw, ntiles), t_end = min(t_begin + tpw, ntiles);

    // --- retainBest(2N) threshold on the integer FAST score
    int *hist = hist4;
    for (int i = tid; i < 256; i += THREADS) hist[i] = 0;
    __syncthreads();
    for_each_candidate(cand, ccnt, t_begin, t_end, lane, [&](uint32_t e, bool valid) {
        if (valid) atomicAdd(&hist[e >> 24], 1);
    });
    __syncthreads();
    uint32_t thr = 0;
    {
        // keep everything if there are at most K1 candidates, else cut at the K1-th largest score (ties kept)
        int above;
        const int bin = warp_kth_bin_from_top(hist, 2 * L.quota + 1, above);      // >= K1+1 exist <=> total > K1
        if (bin >= 0) thr = (uint32_t)warp_kth_bin_from_top(hist, 2 * L.quota, above);
    }

    // --- ordered compaction of the survivors: count, exchange the warp totals, ranked writes
    int cnt = 0;
    for_each_candidate(cand, ccnt, t_begin, t_end, lane, [&](uint32_t e, bool valid) {
        cnt += __popc(__ballot_sync(0xffffffffu, valid && (e >> 24) >= thr));
    });
    if (lane == 0) warp_sums[warp] = cnt;
    __syncthreads();
    int base = 0, M1 = 0;
#pragma unroll
    for (int i = 0; i < NWARPS; ++i) {
        const int v = warp_sums[i];
        if (i < warp) base += v;
        M1 += v;
    }
    for_each_candidate(cand, ccnt, t_begin, t_end, lane, [&](uint32_t e, bool valid) {
        const bool pass = valid && (e >> 24) >= thr;
        const uint32_t bal = __ballot_sync(0xffffffffu, pass);
        ASLAM_BOUNDS(!pass || base + 32 <= ntiles * FT_CAP + 32);
        if (pass) surv[base + __popc(bal & ((1u << lane) - 1u))] = e;
        base += __popc(bal);
    });
    __syncthreads();

    // --- Harris response of every survivor
    for (int i = tid; i < M1; i += THREADS) {
        const uint32_t e = surv[i];
        ASLAM_BOUNDS((int)(e & 0xfffu) >= 4 && (int)(e & 0xfffu) + 4 < L.w && (int)((e >> 12) & 0xfffu) >= 4 && (int)((e >> 12) & 0xfffu) + 4 < L.h &&
                     ((((int)(e & 0xfffu) - 4) & ~3) + 12 <= L.pitch));      // the three aligned words of a window row lie inside the padded row
        hresp[i] = harris_response(img, L.pitch, (int)(e & 0xfffu), (int)((e >> 12) & 0xfffu));
    }
    __syncthreads();

    // --- retainBest(N): K-th largest response by a radix select over the ordered key
    uint32_t cut = 0;   // keep keys >= cut
    const int K2 = L.quota;
    int k_left;
    [[maybe_unused]] uint32_t rank_cut = 0xffffffffu;     // BUCKET: keep ranks < rank_cut, and rank == rank_cut with key >= cut
    [[maybe_unused]] uint32_t *rank = nullptr;
    if constexpr (!BUCKET) {
        if (M1 > K2)
            cut = block_kth_largest<THREADS>(M1, K2, [&](int i, uint32_t &key) { key = float_order_key(hresp[i]); return true; }, hist4, k_left);
    } else {
        // Grid-bucketed cull: rank of a survivor = how many survivors of its cell have a greater response; order by
        // (rank ascending, response descending), keep the first K2 and everything tied with the K2-th.
        rank = b.rank + tile0 * FT_CAP;
        if (M1 > K2) {
            __shared__ uint2 s_ck[THREADS];        // (cell id, ordered response key) of a block of survivors
            const int cell = g.grid_cell;
            auto cell_of = [&](uint32_t e) { return ((e & 0xfffu) / (uint32_t)cell) | ((((e >> 12) & 0xfffu) / (uint32_t)cell) << 12); };
            for (int i0 = 0; i0 < M1; i0 += THREADS) {
                const int i = i0 + tid;
                const uint32_t mycell = i < M1 ? cell_of(surv[i]) : 0xfffffffeu;
                const uint32_t mykey = i < M1 ? float_order_key(hresp[i]) : 0u;
                uint32_t r = 0;
                for (int j0 = 0; j0 < M1; j0 += THREADS) {
                    __syncthreads();
                    const int j = j0 + tid;
                    s_ck[tid] = j < M1 ? make_uint2(cell_of(surv[j]), float_order_key(hresp[j])) : make_uint2(0xffffffffu, 0u);
                    __syncthreads();
                    const int n = min(THREADS, M1 - j0);        // padding entries match no cell: whole groups of 4 are safe
#pragma unroll 4
                    for (int k = 0; k < ((n + 3) & ~3); ++k) {
                        const uint2 o = s_ck[k];                    // one broadcast 64-bit load
                        r += (o.x == mycell) & (o.y > mykey);
                    }
                }
                if (i < M1) rank[i] = r;
            }
            __syncthreads();
            rank_cut = ~block_kth_largest<THREADS>(M1, K2, [&](int i, uint32_t &key) { key = ~rank[i]; return true; }, hist4, k_left);
            const uint32_t R = rank_cut;
            cut = block_kth_largest<THREADS>(M1, k_left, [&](int i, uint32_t &key) { key = float_order_key(hresp[i]); return rank[i] == R; },
                                    hist4, k_left);
        }
    }

    // --- ordered output: warp w takes survivors [w * spw, (w + 1) * spw)
    uint32_t *sel_xy = b.sel_xy + (size_t)frame * g.kp_cap + L.sel_off;
    float *sel_resp = b.sel_resp + (size_t)frame * g.kp_cap + L.sel_off;
    const int spw = (M1 + NWARPS - 1) / NWARPS;
    const int s_begin = min(warp * spw, M1), s_end = min(s_begin + spw, M1);
    auto keeps = [&](int i, float &r) {
        r = hresp[i];
        bool pass = float_order_key(r) >= cut;
        if constexpr (BUCKET)
            if (M1 > K2) pass = rank[i] < rank_cut || (rank[i] == rank_cut && pass);
        return pass;
    };
    int kept = 0;
    for (int i0 = s_begin; i0 < s_end; i0 += 32) {
        const int i = i0 + lane;
        float r;
        kept += __popc(__ballot_sync(0xffffffffu, i < s_end && keeps(i, r)));
    }
    if (lane == 0) warp_sums2[warp] = kept;
    __syncthreads();
    int pos = 0, written = 0;
#pragma unroll
    for (int i = 0; i < NWARPS; ++i) {
        const int v = warp_sums2[i];
        if (i < warp) pos += v;
        written += v;
    }
    for (int i0 = s_begin; i0 < s_end; i0 += 32) {
        const int i = i0 + lane;
        float r = 0.f;
        const bool pass = i < s_end && keeps(i, r);
        const uint32_t bal = __ballot_sync(0xffffffffu, pass);
        const int at = pos + __popc(bal & ((1u << lane) - 1u));
        if (pass && at < L.sel_cap) {
            sel_xy[at] = surv[i] & 0xffffffu;
            sel_resp[at] = r;
        }
        pos += __popc(bal);
    }
    if (tid == 0) {
        *sel_cnt = min(written, L.sel_cap);
        if (written > L.sel_cap) atomicOr(b.overflow + frame + 1, 1);
    }
}

#ifndef ASLAM_SEL_SMALL_THREADS
#define ASLAM_SEL_SMALL_THREADS 128
#endif
int launch_select(aslam_ctx *ctx, const OrbGeom &g, const OrbBuffers &b, int nframes)
{
    if (g.nbuilt == 0) return ASLAM_OK;
    dim3 grid(g.nbuilt, nframes);
    ProfScope ps(ctx, PK_SELECT);
    // CTA size by call size.  A level's CTA is a chain of short phases; with enough (frame, level) pairs to fill the GPU,
    // many small CTAs per SM interleave their phases better than two of 512 threads (0.149 -> 0.121 ms per 256 frames with 256
    // threads, 0.123 with 128) and hand their slots to other streams' kernels sooner (bench.py, eight concurrent streams of 32
    // frames: 91.6 k / 94.7 k / 96.5 k frames/s with 512 / 256 / 128 threads), while a call with few frames wants the most
    // threads per level (one frame: 21 us with 512 threads, 25 with 256, 35 with 128; 1024 threads gain nothing: 24 us).
    static const int small_from = [] { const char *e = getenv("ASLAM_SEL_SMALL_FROM"); return e ? atoi(e) : 16; }();
    // ... and only while a level's survivors are a few rounds of a small CTA (2 x the largest quota <= 1024, i.e. up to ~2300
    // keypoints per frame): at 1408 x 1408 / 4000 keypoints in calls of 16 frames the small CTAs lose 5 % (12 % with grid bucketing)
    int max_quota = 0;
    for (int l = 0; l < g.nbuilt; ++l) max_quota = max_quota > g.lv[l].quota ? max_quota : g.lv[l].quota;
    const bool small = nframes >= small_from && 2 * max_quota <= 1024;
    if (g.grid_cell > 0) {
        if (small) select_kernel<true, ASLAM_SEL_SMALL_THREADS><<<grid, ASLAM_SEL_SMALL_THREADS, 0, ctx->stream>>>(g, b);
        else select_kernel<true, 512><<<grid, 512, 0, ctx->stream>>>(g, b);
    } else {
        if (small) select_kernel<false, ASLAM_SEL_SMALL_THREADS><<<grid, ASLAM_SEL_SMALL_THREADS, 0, ctx->stream>>>(g, b);
        else select_kernel<false, 512><<<grid, 512, 0, ctx->stream>>>(g, b);
    }
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


#if !ASLAM_DESC_TMA
constexpr int DP_ROWS = 39, DP_WORDS = 12;      // blurred patch: +-19 rows of 12 aligned words
constexpr int OP_ROWS = 31;                     // unblurred patch for the moments: +-15 rows, same row format

// Stage `rows` rows of 12 aligned words starting at `src` (row 0, word 0) into smem: lanes 0..23 take two
// rows per step, one 64-bit pointer bump per step, stores at compile-time offsets.
template <int ROWS>
__device__ __forceinline__ void stage_rows_load(const uint8_t *src, int pitch, int col0, int lane, uint32_t (&w)[(ROWS + 1) / 2])
{
    const int hi = lane >= 12 ? 1 : 0, wd = lane - 12 * hi;
    const uint8_t *p = src + (size_t)hi * pitch + min(col0 + 4 * wd, pitch - 4);
    const size_t step = 2 * (size_t)pitch;
#pragma unroll
    for (int it = 0; it < (ROWS + 1) / 2; ++it) {
        w[it] = 0;
        if (lane < 24 && 2 * it + hi < ROWS) w[it] = __ldg(reinterpret_cast<const uint32_t *>(p));
        p += step;
    }
}

template <int ROWS>
__device__ __forceinline__ void stage_rows_store(uint32_t *ps, int lane, const uint32_t (&w)[(ROWS + 1) / 2])
{
#pragma unroll
    for (int it = 0; it < (ROWS + 1) / 2; ++it)
        if (lane < 24 && 24 * it + lane < ROWS * DP_WORDS) ps[24 * it + lane] = w[it];
}

#endif

__device__ __forceinline__ int dp4a_u8_s8(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// sin and cos of a double in [0, 2*pi + eps] to < 1 ulp (double): quadrant reduction with a two-part
// pi/2 (k <= 4, so k*PIO2_HI is exact) and the fdlibm kernel polynomials on [-pi/4, pi/4].  OpenCV
// evaluates cos/sin in double and keeps the float cast; a sub-ulp double error changes that cast
// only when the true value sits within ~1e-16 relative of a float rounding boundary.
__device__ __forceinline__ void sincos_0_2pi(double t, double &s, double &c)
{
    const double k = rint(t * 0.63661977236758134308);            // 2/pi
    double r = fma(-k, 1.57079632679489655800e+00, t);            // pi/2 high part
    r = fma(-k, 6.12323399573676603587e-17, r);                   // pi/2 low part
    const double z = r * r;
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06);
    ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03);
    ps = fma(z, ps, -1.66666666666666324348e-01);
    const double sr = fma(z * r, ps, r);
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07);
    pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03);
    pc = fma(z, pc, 4.16666666666666019037e-02);
    const double cr = fma(z * z, pc, fma(z, -0.5, 1.0));
    const int q = (int)k & 3;
    const double s0 = (q & 1) ? cr : sr, c0 = (q & 1) ? sr : cr;
    s = (q & 2) ? -s0 : s0;
    c = ((q + 1) & 2) ? -c0 : c0;
}

#if !ASLAM_DESC_TMA
// grid.x = describe CTAs of all levels (LevelGeom::desc_block_off): the level is CTA-uniform.  A warp
// handles DESC_KPW keypoints one after another, so its per-lane constants are fetched once and stay
// in registers: the 8 rBRIEF comparisons of lane k (bits of descriptor byte k) as (x,y) float pairs,
// and the orientation weights of the lane's 8 patch words.
__global__ void __launch_bounds__(32 * DESC_WARPS)
describe_kernel(OrbGeom g, OrbBuffers b)
{
    __shared__ uint32_t blur_s[DESC_WARPS][DP_ROWS * DP_WORDS];
    __shared__ uint32_t orig_s[DESC_WARPS][(OP_ROWS + 1) * DP_WORDS];      // +1 row: the lanes of row 31 read (weight 0)
    const int frame = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int l = 0;
#pragma unroll
    for (int i = 1; i < ORB_MAX_LEVELS; ++i)
        if (i < g.nbuilt && (int)blockIdx.x >= g.lv[i].desc_block_off) l = i;
    const LevelGeom L = g.lv[l];
    const int32_t *sc = b.sel_cnt + (size_t)frame * ORB_MAX_LEVELS;
    const size_t fslot = (size_t)frame + 1;     // output slot 0 is the tracker's carry
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int total = 0;
        for (int i = 0; i < g.nbuilt; ++i) total += sc[i];
        b.counts[fslot] = min(total, g.kp_cap);
    }
    const int cnt = sc[l];
    const int k0 = (((int)blockIdx.x - L.desc_block_off) * DESC_WARPS + warp) * DESC_KPW;
    if (k0 >= cnt) return;
    int level_base = 0;                         // position of the level's first keypoint in the frame
    for (int i = 0; i < l; ++i) level_base += sc[i];

    // per-lane constants
    float2 P0[8], P1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 pt = __ldg(b.pattern + j * 32 + lane);        // (x0, y0, x1, y1) of comparison 8*lane + j
        P0[j] = make_float2(pt.x, pt.y);
        P1[j] = make_float2(pt.z, pt.w);
    }
    uint2 OW[8];                                                   // (u*mask as s8x4, mask as u8x4) of patch word (row 4*it + lane/8, lane%8)
#pragma unroll
    for (int it = 0; it < 8; ++it) OW[it] = __ldg(b.ori_tab + it * 32 + lane);
    const int vl = (lane >> 3) - 15;

    const uint8_t *img = b.pyr + (size_t)frame * g.pyr_stride + L.off;
    const uint8_t *blr = b.blur + (size_t)frame * g.pyr_stride + L.off;
    uint32_t *os = orig_s[warp], *ps = blur_s[warp];

#pragma unroll 1
    for (int kk = 0; kk < DESC_KPW; ++kk) {
        const int k = k0 + kk;
        if (k >= cnt) break;
        const int slot = level_base + k;
        if (slot >= g.kp_cap) break;
        const int si = L.sel_off + k;
        const uint32_t xy = b.sel_xy[(size_t)frame * g.kp_cap + si];
        const float resp = b.sel_resp[(size_t)frame * g.kp_cap + si];
        const int x = (int)(xy & 0xfffu), y = (int)(xy >> 12);

        // computeKeyPoints scales the point to level-0 coordinates; computeOrbDescriptors scales it back.
        const float kx = __fmul_rn((float)x, L.sf), ky = __fmul_rn((float)y, L.sf);
        const int xi = __float2int_rn(__fmul_rn(kx, L.inv_sf)), yi = __float2int_rn(__fmul_rn(ky, L.inv_sf));

        // Both patches are fetched with aligned 32-bit loads into registers (36 in flight), then parked in
        // shared memory, so every later access has a compile-time row stride.
        const int ax = (xi - 19) & ~3, ox = (x - 15) & ~3;
        {
            uint32_t ow[(OP_ROWS + 1) / 2], bw[(DP_ROWS + 1) / 2];
            stage_rows_load<OP_ROWS>(img + (size_t)(y - 15) * L.pitch, L.pitch, ox, lane, ow);
            stage_rows_load<DP_ROWS>(blr + (size_t)(yi - 19) * L.pitch, L.pitch, ax, lane, bw);
            __syncwarp();                       // the previous keypoint's reads are done
            stage_rows_store<OP_ROWS>(os, lane, ow);
            stage_rows_store<DP_ROWS>(ps, lane, bw);
        }
        __syncwarp();

        // ICAngles: m10 = sum u*I, m01 = sum v*I over the radius-15 disc.  The lane's word of row
        // 4*it + lane/8 is re-aligned to the patch (funnel shift of two staged words), so its column
        // weights do not depend on the keypoint: two DP4A per word.
        int m10 = 0, m01 = 0;
        {
            const uint32_t sh = (uint32_t)((x - 15) & 3) * 8u;
            const uint32_t *row = os + (lane >> 3) * DP_WORDS + (lane & 7);
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const uint32_t px = __funnelshift_r(row[it * 4 * DP_WORDS], row[it * 4 * DP_WORDS + 1], sh);
                m10 = dp4a_u8_s8(px, OW[it].x, m10);
                m01 += (vl + 4 * it) * (int)__dp4a(px, OW[it].y, 0u);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m10 += __shfl_xor_sync(0xffffffffu, m10, o);
            m01 += __shfl_xor_sync(0xffffffffu, m01, o);
        }
        const float angle = fast_atan2_deg((float)m01, (float)m10);

        const float ang = __fmul_rn(angle, (float)(3.14159265358979323846 / 180.f));
        double sd, cd;
        sincos_0_2pi((double)ang, sd, cd);
        const float ca = (float)cd, sa = (float)sd;
        // Rotation: (px*ca, py*ca) and (px*sa, py*sa) by FMUL2, then the two exactly rounded sums
        //   X = px*ca - py*sa,  Y = px*sa + py*ca   as OpenCV's unfused expression evaluates them.
        // cvRound by the 1.5*2^23 trick: bits(v + 1.5*2^23) = 0x4B400000 + round_half_even(v); the constant
        // is folded into the patch base: byte = (iy+19)*48 + (xi-ax) + ix.
        const float2 CC = make_float2(ca, ca), SS = make_float2(sa, sa), MG = make_float2(12582912.0f, 12582912.0f);
        const uint8_t *pb = reinterpret_cast<const uint8_t *>(ps) +
                            ((uint32_t)(19 * DP_WORDS * 4 + (xi - ax)) - 0x4B400000u * (uint32_t)(DP_WORDS * 4 + 1));
        uint32_t byte = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float2 a1 = __fmul2_rn(P0[j], CC), a2 = __fmul2_rn(P0[j], SS);
            const float2 c1 = __fmul2_rn(P1[j], CC), c2 = __fmul2_rn(P1[j], SS);
            const float2 r0 = __fadd2_rn(make_float2(__fsub_rn(a1.x, a2.y), __fadd_rn(a2.x, a1.y)), MG);
            const float2 r1 = __fadd2_rn(make_float2(__fsub_rn(c1.x, c2.y), __fadd_rn(c2.x, c1.y)), MG);
            const int t0 = pb[__float_as_uint(r0.y) * (uint32_t)(DP_WORDS * 4) + __float_as_uint(r0.x)];
            const int t1 = pb[__float_as_uint(r1.y) * (uint32_t)(DP_WORDS * 4) + __float_as_uint(r1.x)];
            byte |= (uint32_t)(t0 < t1) << j;
        }
        b.desc[(fslot * g.kp_cap + slot) * 32 + lane] = (uint8_t)byte;
        if (lane == 0) {
            aslam_keypoint kp;
            kp.x = kx; kp.y = ky; kp.size = L.kp_size; kp.angle = angle; kp.response = resp; kp.octave = l;
            b.kps[fslot * g.kp_cap + slot] = kp;
        }
    }
}

int launch_describe(aslam_ctx *ctx, const OrbGeom &g, const OrbBuffers &b, int nframes)
{
    if (g.desc_blocks_per_frame == 0) {
        ASLAM_CUDA(cudaMemsetAsync(b.counts + 1, 0, sizeof(int32_t) * nframes, ctx->stream));
        return ASLAM_OK;
    }
    dim3 grid(g.desc_blocks_per_frame, nframes);
    ProfScope ps(ctx, PK_DESCRIBE);
    describe_kernel<<<grid, 32 * DESC_WARPS, 0, ctx->stream>>>(g, b);
    ASLAM_LAUNCH_CHECK();
    return ASLAM_OK;
}

#else
// ---- TMA form (the shipped kernel).  grid.x = describe CTAs of all levels (LevelGeom::desc_block_off): the level is
// CTA-uniform.  A warp owns DESC_KPW keypoints, lane kk holds keypoint kk's coordinates, angle and rotation.
//  * Patches arrive by TMA box loads (cp.async.bulk.tensor.3d over an (x, y, frame) u8 tensor per level): no staging
//    loads or stores and no per-load addressing arithmetic.  Measured on B200: the box's first byte must be 16-byte
//    aligned in global memory (an unaligned innermost coordinate of a u8 tensor raises "illegal instruction",
//    tools/probes/tma_probe.cu), so a box starts at the patch's first column rounded down to 16 and is 48 x 31 bytes
//    for the orientation patch (31 columns + up to 15 of slack) and 64 x 39 for the rBRIEF patch (39 + 15).
//    A lane issues the loads of its own keypoint; both kinds go through small rings of buffers (DESC_ORING / DESC_RING
//    slots per warp) refilled as keypoints finish; completion is signalled on per-warp mbarriers.
//  * Two passes: (1) integer moments of every keypoint (2 DP4A per patch word), then fastAtan2 and the double-precision
//    sin/cos ONCE for the whole warp -- lane kk computes keypoint kk's -- instead of once per keypoint on all lanes;
//    (2) the 256 rotated comparisons of each keypoint, lane k = descriptor byte k.
constexpr int OPB_W = 48, OPB_H = 31, OPB_BYTES = OPB_W * OPB_H, OPB_STRIDE = 1536;     // orientation box; smem slots are 128-B multiples
constexpr int BPB_W = 64, BPB_H = 39, BPB_BYTES = BPB_W * BPB_H, BPB_STRIDE = 2560;     // rBRIEF box
// Measured on B200 (ms per 256 frames x 2000 keypoints; round 1's LDG staging: 0.510): rings of 2 + 2 slots 0.322,
// 4 + 3 slots 0.350, 4 + 4 or 8-warp CTAs 0.40 (fewer resident warps), 1 + 1 (no prefetch) 0.404.
#ifndef ASLAM_DESC_RING
#define ASLAM_DESC_RING 2
#endif
#ifndef ASLAM_DESC_ORING
#define ASLAM_DESC_ORING 2
#endif
constexpr int DESC_RING = ASLAM_DESC_RING < DESC_KPW ? ASLAM_DESC_RING : DESC_KPW;
constexpr int DESC_ORING = ASLAM_DESC_ORING < DESC_KPW ? ASLAM_DESC_ORING : DESC_KPW;
constexpr int DESC_WARP_SMEM = DESC_ORING * OPB_STRIDE + DESC_RING * BPB_STRIDE;
constexpr size_t DESC_SMEM_BYTES = (size_t)DESC_WARPS * DESC_WARP_SMEM + 128;           // + slack to align the base to 128 B
static_assert(OPB_STRIDE >= (OPB_H + 1) * OPB_W && OPB_STRIDE % 128 == 0 && BPB_STRIDE >= BPB_BYTES && BPB_STRIDE % 128 == 0, "patch slots");

__global__ void __launch_bounds__(32 * DESC_WARPS)
describe_kernel(OrbGeom g, OrbBuffers b)
{
    extern __shared__ uint8_t desc_smem_raw[];
    __shared__ __align__(8) unsigned long long bars[DESC_WARPS][DESC_ORING + DESC_RING];
    const int frame = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int l = 0;
#pragma unroll
    for (int i = 1; i < ORB_MAX_LEVELS; ++i)
        if (i < g.nbuilt && (int)blockIdx.x >= g.lv[i].desc_block_off) l = i;
    const LevelGeom L = g.lv[l];
    const int32_t *sc = b.sel_cnt + (size_t)frame * ORB_MAX_LEVELS;
    const size_t fslot = (size_t)frame + 1;     // output slot 0 is the tracker's carry
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int total = 0;
        for (int i = 0; i < g.nbuilt; ++i) total += sc[i];
        b.counts[fslot] = min(total, g.kp_cap);
    }
    const int cnt = sc[l];
    const int k0 = (((int)blockIdx.x - L.desc_block_off) * DESC_WARPS + warp) * DESC_KPW;
    int level_base = 0;                         // position of the level's first keypoint in the frame
    for (int i = 0; i < l; ++i) level_base += sc[i];
    const int nk = min(DESC_KPW, min(cnt - k0, g.kp_cap - level_base - k0));      // this warp's keypoints
    if (nk <= 0) return;                        // warps are independent from here on (no CTA-wide barrier below)

    uint8_t *wsm = desc_smem_raw + ((128u - (smem_u32(desc_smem_raw) & 127u)) & 127u) + (size_t)warp * DESC_WARP_SMEM;
    uint8_t *bp_ring = wsm + DESC_ORING * OPB_STRIDE;
    const uint32_t obar = smem_u32(&bars[warp][0]), bbar = obar + 8u * DESC_ORING;
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < DESC_ORING + DESC_RING; ++r) mbar_init(obar + 8u * r, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    // lane kk < nk: keypoint kk of this warp
    int x = 0, y = 0, xi = 0, yi = 0;
    float resp = 0.f, kx = 0.f, ky = 0.f;
    const uint8_t *tm = reinterpret_cast<const uint8_t *>(b.tmaps) + (size_t)l * 256;
    if (lane < nk) {
        const int si = L.sel_off + k0 + lane;
        const uint32_t xy = b.sel_xy[(size_t)frame * g.kp_cap + si];
        resp = b.sel_resp[(size_t)frame * g.kp_cap + si];
        x = (int)(xy & 0xfffu); y = (int)(xy >> 12);
        // computeKeyPoints scales the point to level-0 coordinates; computeOrbDescriptors scales it back.
        kx = __fmul_rn((float)x, L.sf); ky = __fmul_rn((float)y, L.sf);
        xi = __float2int_rn(__fmul_rn(kx, L.inv_sf)); yi = __float2int_rn(__fmul_rn(ky, L.inv_sf));
        ASLAM_BOUNDS(x >= 15 && y >= 15 && y + 15 < L.h && x + 15 < L.w && xi >= 19 && yi >= 19 && yi + 19 < L.h && xi + 19 < L.w &&
                     level_base + k0 + lane < g.kp_cap);
        if (lane < DESC_ORING) {
            mbar_arrive_expect_tx(obar + 8u * lane, OPB_BYTES);
            tma_load_box(smem_u32(wsm + lane * OPB_STRIDE), tm, (x - 15) & ~15, y - 15, frame, obar + 8u * lane);
        }
        if (lane < DESC_RING) {
            mbar_arrive_expect_tx(bbar + 8u * lane, BPB_BYTES);
            tma_load_box(smem_u32(bp_ring + lane * BPB_STRIDE), tm + 128, (xi - 19) & ~15, yi - 19, frame, bbar + 8u * lane);
        }
    }

    // ---- pass 1: ICAngles moments m10 = sum u*I, m01 = sum v*I over the radius-15 disc.  Lane = (row 4*it + lane/8,
    // word lane%8) of the patch; its word is re-aligned to the patch's first column (funnel shift of two staged words),
    // so byte bt is column u = 4*(lane%8) + bt - 15 whatever the keypoint: the weights are per-lane constants.
    int my_m10 = 0, my_m01 = 0;
    {
        uint2 OW[8];                            // (u*mask as s8x4, mask as u8x4) of patch word (row 4*it + lane/8, lane%8)
#pragma unroll
        for (int it = 0; it < 8; ++it) OW[it] = __ldg(b.ori_tab + it * 32 + lane);
        const int vl = (lane >> 3) - 15;
#pragma unroll 1
        for (int kk = 0; kk < nk; ++kk) {
            const int slot = kk % DESC_ORING;
            const int off = __shfl_sync(0xffffffffu, (x - 15) & 15, kk);           // patch's first column inside the box
            mbar_wait(obar + 8u * slot, (uint32_t)(kk / DESC_ORING) & 1u);
            const uint32_t *row = reinterpret_cast<const uint32_t *>(wsm + slot * OPB_STRIDE) + (lane >> 3) * (OPB_W / 4) + (lane & 7) + (off >> 2);
            const uint32_t sh = (uint32_t)(off & 3) * 8u;
            int m10 = 0, m01 = 0;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const uint32_t px = __funnelshift_r(row[it * 4 * (OPB_W / 4)], row[it * 4 * (OPB_W / 4) + 1], sh);
                m10 = dp4a_u8_s8(px, OW[it].x, m10);
                m01 += (vl + 4 * it) * (int)__dp4a(px, OW[it].y, 0u);
            }
            __syncwarp();                       // every lane has read this slot
            if (lane == kk + DESC_ORING && lane < nk) {     // the slot's next tenant is this lane's keypoint
                mbar_arrive_expect_tx(obar + 8u * slot, OPB_BYTES);
                tma_load_box(smem_u32(wsm + slot * OPB_STRIDE), tm, (x - 15) & ~15, y - 15, frame, obar + 8u * slot);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                m10 += __shfl_xor_sync(0xffffffffu, m10, o);
                m01 += __shfl_xor_sync(0xffffffffu, m01, o);
            }
            if (lane == kk) { my_m10 = m10; my_m01 = m01; }
        }
    }
    // orientation and rotation of all keypoints of the warp in one go (lane kk: keypoint kk)
    const float angle = fast_atan2_deg((float)my_m01, (float)my_m10);
    float my_ca, my_sa;
    {
        const float ang = __fmul_rn(angle, (float)(3.14159265358979323846 / 180.f));
        double sd, cd;
        sincos_0_2pi((double)ang, sd, cd);
        my_ca = (float)cd; my_sa = (float)sd;
    }
    if (lane < nk) {
        aslam_keypoint kp;
        kp.x = kx; kp.y = ky; kp.size = L.kp_size; kp.angle = angle; kp.response = resp; kp.octave = l;
        b.kps[fslot * g.kp_cap + level_base + k0 + lane] = kp;
    }

    // ---- pass 2: rBRIEF, lane k = descriptor byte k (comparisons 8k .. 8k+7 as (x,y) float pairs in registers)
    float2 P0[8], P1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 pt = __ldg(b.pattern + j * 32 + lane);        // (x0, y0, x1, y1) of comparison 8*lane + j
        P0[j] = make_float2(pt.x, pt.y);
        P1[j] = make_float2(pt.z, pt.w);
    }
    uint8_t *dout = b.desc + (fslot * g.kp_cap + level_base + k0) * 32 + lane;
#pragma unroll 1
    for (int kk = 0; kk < nk; ++kk) {
        const int slot = kk % DESC_RING;
        const float ca = __shfl_sync(0xffffffffu, my_ca, kk), sa = __shfl_sync(0xffffffffu, my_sa, kk);
        const int off = __shfl_sync(0xffffffffu, (xi - 19) & 15, kk);              // patch's first column inside the box
        mbar_wait(bbar + 8u * slot, (uint32_t)(kk / DESC_RING) & 1u);
        // Rotation: (px*ca, py*ca) and (px*sa, py*sa) by FMUL2, then the two exactly rounded sums
        //   X = px*ca - py*sa,  Y = px*sa + py*ca   as OpenCV's unfused expression evaluates them.
        // cvRound by the 1.5*2^23 trick: bits(v + 1.5*2^23) = 0x4B400000 + round_half_even(v); the constant
        // is folded into the patch base: byte = (iy + 19) * 64 + off + 19 + ix (the box starts at row yi - 19).
        // (32-bit shared-window address: the subtraction wraps and the additions below wrap back)
        const float2 CC = make_float2(ca, ca), SS = make_float2(sa, sa), MG = make_float2(12582912.0f, 12582912.0f);
        const uint32_t pb = smem_u32(bp_ring + slot * BPB_STRIDE) + (uint32_t)(19 * BPB_W + 19 + off) - 0x4B400000u * (uint32_t)(BPB_W + 1);
        uint32_t byte = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float2 a1 = __fmul2_rn(P0[j], CC), a2 = __fmul2_rn(P0[j], SS);
            const float2 c1 = __fmul2_rn(P1[j], CC), c2 = __fmul2_rn(P1[j], SS);
            const float2 r0 = __fadd2_rn(make_float2(__fsub_rn(a1.x, a2.y), __fadd_rn(a2.x, a1.y)), MG);
            const float2 r1 = __fadd2_rn(make_float2(__fsub_rn(c1.x, c2.y), __fadd_rn(c2.x, c1.y)), MG);
            ASLAM_BOUNDS(pb + __float_as_uint(r0.y) * (uint32_t)BPB_W + __float_as_uint(r0.x) - smem_u32(bp_ring + slot * BPB_STRIDE) < (uint32_t)BPB_BYTES &&
                         pb + __float_as_uint(r1.y) * (uint32_t)BPB_W + __float_as_uint(r1.x) - smem_u32(bp_ring + slot * BPB_STRIDE) < (uint32_t)BPB_BYTES);
            const uint32_t t0 = lds_u8(pb + __float_as_uint(r0.y) * (uint32_t)BPB_W + __float_as_uint(r0.x));
            const uint32_t t1 = lds_u8(pb + __float_as_uint(r1.y) * (uint32_t)BPB_W + __float_as_uint(r1.x));
            byte |= (uint32_t)(t0 < t1) << j;
        }
        dout[(size_t)kk * 32] = (uint8_t)byte;
        __syncwarp();                           // every lane has read this slot
        if (lane == kk + DESC_RING && lane < nk) {      // the slot's next tenant is this lane's keypoint
            mbar_arrive_expect_tx(bbar + 8u * slot, BPB_BYTES);
            tma_load_box(smem_u32(bp_ring + slot * BPB_STRIDE), tm + 128, (xi - 19) & ~15, yi - 19, frame, bbar + 8u * slot);
        }
    }
}

int launch_describe(aslam_ctx *ctx, const OrbGeom &g, const OrbBuffers &b, int nframes)
{
    if (g.desc_blocks_per_frame == 0) {
        ASLAM_CUDA(cudaMemsetAsync(b.counts + 1, 0, sizeof(int32_t) * nframes, ctx->stream));
        return ASLAM_OK;
    }
    if (!b.tmaps) return fail(ASLAM_E_STATE, "describe_kernel: the handle has no TMA descriptors");
    if (!ctx->desc_attr_set) {                  // per device, once
        ASLAM_CUDA(cudaFuncSetAttribute(describe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DESC_SMEM_BYTES));
        ctx->desc_attr_set = true;
    }
    dim3 grid(g.desc_blocks_per_frame, nframes);
    ProfScope ps(ctx, PK_DESCRIBE);
    describe_kernel<<<grid, 32 * DESC_WARPS, DESC_SMEM_BYTES, ctx->stream>>>(g, b);
    ASLAM_LAUNCH_CHECK();
    return ASLAM_OK;
}
#endif      // ASLAM_DESC_TMA

// ------------------------------------------------------------------ dynamic-object filter (a4)
// src/main.cpp:29-50 DYNAMIC_CLASSES {0,1,2,3,5,6,7,14,15,16}; Detection::contains is closed on
// floats (include/core/Types.hpp:109-111): in_dynamic_box of common.cuh.  One CTA compacts one frame's keypoints in order.
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
            keep = !in_dynamic_box(k.x, k.y, boxes, nboxes, 0);
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
