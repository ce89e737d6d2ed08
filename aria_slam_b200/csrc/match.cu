// Brute-force Hamming kNN (k = 2) over 256-bit descriptors: XOR + POPC on the integer pipes,
// queries register-blocked, train rows streamed through shared memory and broadcast to the warp.
//
// Replaces cv::cuda::DescriptorMatcher::knnMatch as the reference calls it
// (src/adapters/gpu/CudaMatcher.cpp:56, src/main.cpp:159, src/euroc_eval.cpp:168) and the CPU
// cv::BFMatcher::knnMatch of src/legacy/LoopClosure.cpp:87; results follow the CPU matcher:
// ascending (distance, train index).
#include <cmath>
#include <cstdlib>

#include "match.cuh"

namespace aslam {

constexpr int KNN_TILE = KNN_THREADS;   // train rows per shared-memory stage (48 B each: one staged row per thread)

// popcount(q ^ t) over 256 bits.  POPC issues at 16 lanes/clk/SM (measured 4.38 Tpopc/s), LOP3 at 64:
// carry-save adders (2 LOP3 each) compress three equal-weight words into a sum and a carry word and so
// trade one quarter-rate POPC for two full-rate logic ops.  With Qa = q0^q1^q2 kept per query and
// Ta = t0^t1^t2 computed once per train row, the adder sum s0 = Qa^Ta needs no x2, and the carry
// maj(x0,x1,x2) = f(x0,x1,s0) is still one LOP3: three such adders + one over their carries leave
// four words to count (weights 1, 1, 2, 4): 13 logic ops + 4 POPC per pair instead of 8 + 8.
__device__ __forceinline__ void csa(uint32_t a, uint32_t b, uint32_t c, uint32_t &sum, uint32_t &carry)
{
    sum = a ^ b ^ c;
    carry = (a & b) | (c & (a ^ b));
}

// carry of a full adder given two inputs and the SUM of all three: x2 = s ^ x0 ^ x1
__device__ __forceinline__ uint32_t carry_from_sum(uint32_t x0, uint32_t x1, uint32_t s)
{
    uint32_t r;      // one LOP3: (x0 & x1) | ((s ^ x0 ^ x1) & (x0 ^ x1))
    asm("lop3.b32 %0, %1, %2, %3, 0xd4;" : "=r"(r) : "r"(x0), "r"(x1), "r"(s));
    return r;
}

struct QueryParity { uint32_t a, b, c; };      // q0^q1^q2, q3^q4^q5, q0^..^q6

// ---- scan kernel (ASLAM_KNN_CSA 7: 5 POPC per pair, 8: 4 POPC per pair)
// Everything that is not logic or POPC is kept off the ALU pipe:
//  * the three parity words of a train row are computed once when the row is staged (a third
//    16-byte word per row in shared memory) instead of by every thread;
//  * popcounts are combined and keys are built with IMADs whose multipliers are kernel
//    parameters (the FMA pipe is idle in this kernel; literal powers of two would be turned back
//    into ALU shifts/adds by the compiler);
//  * inside a 128-row tile a neighbour is the 16-bit key distance << 7 | row-in-tile, two queries
//    per register, so one VIMNMX.U16x2 triple updates the best-two of two queries; tile results are
//    widened to 32-bit (distance, row) keys and merged after every tile.  Order and tie-breaking
//    are unchanged: the unsigned order of both key formats is lexicographic (distance, row).
struct KnnMul { uint32_t two, four, pack16, shl7; };      // 2, 4, 65536, 128

__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

template <bool POPC4>
__device__ __forceinline__ uint32_t hamming256_v2(const uint32_t (&q)[8], const QueryParity qp, const uint4 a, const uint4 b,
                                                 const uint4 par, const KnnMul M)
{
    const uint32_t x0 = q[0] ^ a.x, x1 = q[1] ^ a.y, s0 = qp.a ^ par.x, c0 = carry_from_sum(x0, x1, s0);
    const uint32_t x3 = q[3] ^ a.w, x4 = q[4] ^ b.x, s1 = qp.b ^ par.y, c1 = carry_from_sum(x3, x4, s1);
    const uint32_t s2 = qp.c ^ par.z, c2 = carry_from_sum(s0, s1, s2);
    const uint32_t x7 = q[7] ^ b.w;
    if (POPC4) {
        uint32_t s3, c3;
        csa(c0, c1, c2, s3, c3);
        return imad(__popc(c3), M.four, imad(__popc(s3), M.two, __popc(s2) + __popc(x7)));
    }
    return imad(__popc(c0) + __popc(c1) + __popc(c2), M.two, __popc(s2) + __popc(x7));
}

// tile-local 16-bit key -> global 32-bit key
__device__ __forceinline__ uint32_t widen_key(uint32_t k16, uint32_t t0)
{
    return k16 == 0xffffu ? KEY_EMPTY : ((k16 >> 7) << KEY_IDX_BITS) + t0 + (k16 & 127u);
}

// ---- peer-memory primitives of the fused exchange
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// 32-bit shard key (distance << 23 | local row) -> 64-bit global key (distance << 40 | global row)
__device__ __forceinline__ unsigned long long global_key(uint32_t k, const RowMap &m)
{
    if (k == KEY_EMPTY) return ~0ull;
    const int local = (int)(k & (KNN_MAX_ROWS - 1));
    long long row;
    if (m.nseg == 0) {
        row = m.row_base + local;
    } else {
        int lo = 0, hi = m.nseg - 1;              // last keyframe whose first local row is <= local
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (m.lstart[mid] <= local) lo = mid; else hi = mid - 1;
        }
        row = m.gbase[lo] + (local - m.lstart[lo]);
    }
    return ((unsigned long long)(k >> KEY_IDX_BITS) << 40) | (unsigned long long)row;
}

#ifndef ASLAM_KNN_MINB
#define ASLAM_KNN_MINB 1
#endif
template <bool POPC4, bool FUSED>
__global__ void __launch_bounds__(KNN_THREADS, ASLAM_KNN_MINB)
knn2_partial_v2_kernel(const KnnProblem *__restrict__ probs, uint2 *partial, int q_stride, const KnnMul M, int merge_div,
                       const ExchangeArgs X)
{
    static_assert(KNN_TILE == KNN_THREADS && KNN_TILE <= 128 && KNN_TILE % 32 == 0 && (KNN_RQ & 1) == 0, "one staged row per thread, 7-bit row-in-tile");
    __shared__ uint4 tile[2][KNN_TILE * 3];

    // Pipelined queries (programmatic dependent launch): the next query's kernel may be scheduled while this one's
    // last CTAs are still in their exchange tail; it scans at once and holds its merge (griddepcontrol.wait below)
    // until this kernel has completed.  Both instructions do nothing in a launch without the attribute.
    if (FUSED) asm volatile("griddepcontrol.launch_dependents;");
    const KnnProblem P = probs[blockIdx.y];
    const int nq = P.nq_dev ? min(*P.nq_dev, P.nq) : P.nq;
    const int nt = P.nt_dev ? min(*P.nt_dev, P.nt) : P.nt;
    const int qbase = blockIdx.x * KNN_QBLOCK;
    if (!FUSED && qbase >= nq) return;
    const int t_lo = P.t_lo, t_hi = min(P.t_hi, nt);
    const int tid = threadIdx.x;

    uint32_t qd[KNN_RQ][8];
    uint32_t b1[KNN_RQ], b2[KNN_RQ];
    QueryParity qp[KNN_RQ];
#pragma unroll
    for (int j = 0; j < KNN_RQ; ++j) {
        const int qi = qbase + j * KNN_THREADS + tid;
        uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
        if (qi < nq) {
            const uint4 *src = reinterpret_cast<const uint4 *>(P.q) + (size_t)qi * 2;
            lo = __ldg(src); hi = __ldg(src + 1);
        }
        qd[j][0] = lo.x; qd[j][1] = lo.y; qd[j][2] = lo.z; qd[j][3] = lo.w;
        qd[j][4] = hi.x; qd[j][5] = hi.y; qd[j][6] = hi.z; qd[j][7] = hi.w;
        b1[j] = KEY_EMPTY; b2[j] = KEY_EMPTY;
        qp[j].a = qd[j][0] ^ qd[j][1] ^ qd[j][2];
        qp[j].b = qd[j][3] ^ qd[j][4] ^ qd[j][5];
        qp[j].c = qp[j].a ^ qp[j].b ^ qd[j][6];
    }

    const uint4 *trows = reinterpret_cast<const uint4 *>(P.t);
    uint4 pre0 = make_uint4(0, 0, 0, 0), pre1 = pre0;      // both halves of train row t0 + tid
    if (t_lo + tid < t_hi) { pre0 = __ldg(trows + 2 * (size_t)(t_lo + tid)); pre1 = __ldg(trows + 2 * (size_t)(t_lo + tid) + 1); }
    int buf = 0;
    for (int t0 = t_lo; t0 < t_hi; t0 += KNN_TILE, buf ^= 1) {
        {
            uint4 par;
            par.x = pre0.x ^ pre0.y ^ pre0.z; par.y = pre0.w ^ pre1.x ^ pre1.y; par.z = par.x ^ par.y ^ pre1.z; par.w = 0;
            ASLAM_BOUNDS(3 * tid + 2 < KNN_TILE * 3 && t0 >= t_lo && t0 < t_hi && t_hi <= nt);
            tile[buf][3 * tid] = pre0; tile[buf][3 * tid + 1] = pre1; tile[buf][3 * tid + 2] = par;
        }
        __syncthreads();
        if (t0 + KNN_TILE + tid < t_hi) {
            pre0 = __ldg(trows + 2 * (size_t)(t0 + KNN_TILE + tid)); pre1 = __ldg(trows + 2 * (size_t)(t0 + KNN_TILE + tid) + 1);
        }
        const int rows = min(KNN_TILE, t_hi - t0);
        const uint4 *tb = tile[buf];
        uint32_t l1[KNN_RQ / 2], l2[KNN_RQ / 2];           // tile-local best / second, two queries per register
#pragma unroll
        for (int j = 0; j < KNN_RQ / 2; ++j) { l1[j] = 0xffffffffu; l2[j] = 0xffffffffu; }
        uint32_t rr = 0;                                   // row-in-tile in both halves
#pragma unroll 4
        for (int r = 0; r < rows; ++r) {
            const uint4 a = tb[3 * r], b = tb[3 * r + 1], par = tb[3 * r + 2];
#pragma unroll
            for (int j = 0; j < KNN_RQ / 2; ++j) {
                const uint32_t da = hamming256_v2<POPC4>(qd[2 * j], qp[2 * j], a, b, par, M);
                const uint32_t db = hamming256_v2<POPC4>(qd[2 * j + 1], qp[2 * j + 1], a, b, par, M);
                const uint32_t key = imad(imad(db, M.pack16, da), M.shl7, rr);
                l2[j] = __vminu2(l2[j], __vmaxu2(l1[j], key));
                l1[j] = __vminu2(l1[j], key);
            }
            rr += 0x00010001u;
        }
#pragma unroll
        for (int j = 0; j < KNN_RQ; ++j) {
            const int sh = (j & 1) * 16;
            const uint32_t k1 = widen_key((l1[j >> 1] >> sh) & 0xffffu, (uint32_t)t0);
            const uint32_t k2 = widen_key((l2[j >> 1] >> sh) & 0xffffu, (uint32_t)t0);
            const uint32_t n2 = min(max(b1[j], k1), min(b2[j], k2));      // k1 <= k2
            b1[j] = min(b1[j], k1);
            b2[j] = n2;
        }
    }
    // (pipelined launch: the previous query's kernel re-arms the merged slots and the arrival counters in its tail)
    if (FUSED) asm volatile("griddepcontrol.wait;" ::: "memory");
    // Row ranges of one (query set, train set) pair are consecutive problems: with merge_div > 1 they are merged
    // here, into one (best, second) per query and group, by three order-independent atomic minima (the slot
    // was preset to KEY_EMPTY): B1 <- min(B1, b1) returns the previous best; whichever of the two lost goes to
    // B2, as does b2.  Every key except the overall minimum thus reaches B2, so B2 ends as the second smallest.
#pragma unroll
    for (int j = 0; j < KNN_RQ; ++j) {
        const int qi = qbase + j * KNN_THREADS + tid;
        if (qi >= nq) continue;
        if (!FUSED && merge_div <= 1) {
            partial[(size_t)blockIdx.y * q_stride + qi] = make_uint2(b1[j], b2[j]);
        } else {
            unsigned int *slot = reinterpret_cast<unsigned int *>(partial + (size_t)(FUSED ? 0 : blockIdx.y / merge_div) * q_stride + qi);
            if (b1[j] != KEY_EMPTY) {
                const uint32_t old = atomicMin(slot, b1[j]);
                const uint32_t lost = max(old, b1[j]);
                if (lost != KEY_EMPTY) atomicMin(slot + 1, lost);
                if (b2[j] != KEY_EMPTY) atomicMin(slot + 1, b2[j]);
            }
        }
    }
    if (!FUSED) return;

    // ---- fused exchange (one shard of a sharded database; every problem of the launch is a row range of the same pair).
    // The last CTA to arrive for this block of queries owns the block from here on.
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int prev = atomicAdd(X.arrive + blockIdx.x, 1u);
        s_last = prev == gridDim.y - 1;
        if (s_last) X.arrive[blockIdx.x] = 0;               // ready for the next launch (stream-ordered)
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int half = (int)(X.round & 1ull);
    const int nqb = gridDim.x;
    const size_t flags_off = (size_t)2 * X.nparts * X.nq * 2;                 // per peer buffer, after the data
    // phase 1: this shard's top-2 of the block's queries, as global keys, into every peer's gather buffer (NVLink stores)
#pragma unroll
    for (int j = 0; j < KNN_RQ; ++j) {
        const int qi = qbase + j * KNN_THREADS + tid;
        if (qi >= nq) continue;
        const uint2 p = __ldcg(partial + qi);
        partial[qi] = make_uint2(KEY_EMPTY, KEY_EMPTY);     // the merged slot is ready for the next scan
        ulonglong2 v;
        v.x = global_key(p.x, X.map);
        v.y = global_key(p.y, X.map);
        const size_t at = (((size_t)half * X.nparts + X.my_part) * X.nq + qi) * 2;
        for (int pr = 0; pr < X.nparts; ++pr) *reinterpret_cast<ulonglong2 *>(X.peer_bufs[pr] + at) = v;
    }
    __threadfence_system();
    __syncthreads();
    // publish: this shard's flag of this block and round, on every peer
    for (int pr = tid; pr < X.nparts; pr += KNN_THREADS)
        st_release_sys(X.peer_bufs[pr] + flags_off + ((size_t)half * X.nparts + X.my_part) * nqb + blockIdx.x, X.round);
    const unsigned long long t_push = global_timer_ns();
    // phase 2: wait for every shard's flag of this block, then merge the block's queries from the LOCAL buffer
    unsigned long long *mine = X.peer_bufs[X.my_part];
    for (int pr = tid; pr < X.nparts; pr += KNN_THREADS) {
        const unsigned long long *f = mine + flags_off + ((size_t)half * X.nparts + pr) * nqb + blockIdx.x;
        while (ld_acquire_sys(f) < X.round) {
            __nanosleep(32);
            if (global_timer_ns() - t_push > 4000000000ull) { atomicOr(X.status, 1u); break; }      // a peer never came: 4 s
        }
    }
    __syncthreads();
    if (tid == 0 && X.stamps) {               // per query block: last push, last all-seen, sum of waits, exchanges
        const unsigned long long t_seen = global_timer_ns();
        unsigned long long *st = X.stamps + 4 * blockIdx.x;
        st[0] = t_push; st[1] = t_seen; st[2] += t_seen - t_push; st[3] += 1;
    }
#pragma unroll
    for (int j = 0; j < KNN_RQ; ++j) {
        const int qi = qbase + j * KNN_THREADS + tid;
        if (qi >= nq) continue;
        unsigned long long k1 = ~0ull, k2 = ~0ull;
        for (int pr = 0; pr < X.nparts; ++pr) {
            const ulonglong2 ab = __ldcg(reinterpret_cast<const ulonglong2 *>(mine + (((size_t)half * X.nparts + pr) * X.nq + qi) * 2));
            const unsigned long long n2 = min(max(k1, ab.x), min(k2, ab.y));
            k1 = min(k1, ab.x);
            k2 = n2;
        }
        X.out[2 * (size_t)qi] = k1;
        X.out[2 * (size_t)qi + 1] = k2;
    }
}

// Merge `splits` partial (best, second) pairs of one query: keys are distinct across splits
// (disjoint row ranges), so the two smallest of the union are the answer.
__device__ __forceinline__ void merge_pair(uint32_t &k1, uint32_t &k2, const uint2 p)
{
    // p.x <= p.y always
    const uint32_t n2 = min(max(k1, p.x), min(k2, p.y));
    k1 = min(k1, p.x);
    k2 = n2;
}

__device__ __forceinline__ bool ratio_pass(uint32_t d0, uint32_t d1, const RatioTest rt)
{
    if (rt.disabled) return true;
    if (rt.as_double) return (double)d0 < __dmul_rn(rt.ratio, (double)d1);
    return (float)d0 < __fmul_rn((float)rt.ratio, (float)d1);
}

constexpr int FIN_THREADS = 256;

__device__ __forceinline__ int block_exclusive_scan(int v, int *warp_sums, int &total)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    __syncthreads();   // protect warp_sums from the previous use
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < FIN_THREADS / 32; ++i) {
        int s = warp_sums[i];
        if (i < w) base += s;
        tot += s;
    }
    total = tot;
    return base + inc - v;
}

__global__ void __launch_bounds__(FIN_THREADS)
knn2_finish_matches_kernel(const KnnProblem *__restrict__ probs, const uint2 *__restrict__ partial,
                           int splits, int q_stride, RatioTest rt,
                           aslam_match *__restrict__ out, int out_stride, int32_t *__restrict__ nout,
                           int32_t *__restrict__ raw_idx, int32_t *__restrict__ raw_dist, const DynFilter F)
{
    __shared__ int warp_sums[FIN_THREADS / 32];
    __shared__ int s_filtered;
    if (threadIdx.x == 0) s_filtered = 0;
    __syncthreads();
    const int g = blockIdx.x;
    const KnnProblem P = probs[g * splits];
    const int nq = P.nq_dev ? min(*P.nq_dev, P.nq) : P.nq;
    int written = 0;
    for (int q0 = 0; q0 < nq; q0 += FIN_THREADS) {
        const int q = q0 + threadIdx.x;
        uint32_t k1 = KEY_EMPTY, k2 = KEY_EMPTY;
        if (q < nq) merge_pair(k1, k2, partial[(size_t)g * q_stride + q]);      // the scan merged the group's row ranges
        const uint32_t d0 = k1 >> KEY_IDX_BITS, d1 = k2 >> KEY_IDX_BITS;
        const int i0 = (int)(k1 & (KNN_MAX_ROWS - 1)), i1 = (int)(k2 & (KNN_MAX_ROWS - 1));
        if (q < nq && raw_idx) {
            const size_t o = ((size_t)g * q_stride + q) * 2;
            raw_idx[o] = k1 == KEY_EMPTY ? -1 : i0;      raw_dist[o] = k1 == KEY_EMPTY ? -1 : (int)d0;
            raw_idx[o + 1] = k2 == KEY_EMPTY ? -1 : i1;  raw_dist[o + 1] = k2 == KEY_EMPTY ? -1 : (int)d1;
        }
        int pass = (q < nq && k2 != KEY_EMPTY && ratio_pass(d0, d1, rt)) ? 1 : 0;
        if (pass && F.boxes) {          // src/main.cpp:166-173: either endpoint on a dynamic object drops the match
            const aslam_keypoint a = F.q_kps[q], b = F.t_kps[i0];
            if (in_dynamic_box(a.x, a.y, F.boxes, F.nboxes, F.half_open) || in_dynamic_box(b.x, b.y, F.boxes, F.nboxes, F.half_open)) {
                pass = 0;
                atomicAdd(&s_filtered, 1);
            }
        }
        int total;
        const int pos = block_exclusive_scan(pass, warp_sums, total);
        if (pass && out) {
            aslam_match m;
            m.query_idx = q; m.train_idx = i0; m.distance = (float)d0;
            ASLAM_BOUNDS(i0 >= 0 && i0 < P.nt && q < nq);
            if (written + pos < out_stride) out[(size_t)g * out_stride + written + pos] = m;
        }
        written += total;
    }
    if (threadIdx.x == 0 && nout) nout[g] = written;
    if (threadIdx.x == 0 && F.nfiltered) F.nfiltered[g] = s_filtered;      // ordered after the loop's last barrier
}

__global__ void __launch_bounds__(FIN_THREADS)
knn2_finish_counts_kernel(const KnnProblem *__restrict__ probs, const uint2 *__restrict__ partial,
                          int splits, int q_stride, RatioTest rt, int32_t *__restrict__ good)
{
    __shared__ int warp_sums[FIN_THREADS / 32];
    const int s = blockIdx.x;
    const KnnProblem P = probs[s * splits];
    const int nq = P.nq_dev ? min(*P.nq_dev, P.nq) : P.nq;
    int cnt = 0;
    for (int q = threadIdx.x; q < nq; q += FIN_THREADS) {
        const uint2 p = partial[(size_t)s * q_stride + q];
        if (p.y != KEY_EMPTY && ratio_pass(p.x >> KEY_IDX_BITS, p.y >> KEY_IDX_BITS, rt)) ++cnt;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < FIN_THREADS / 32; ++i) t += warp_sums[i];
        good[s] = t;
    }
}

__global__ void __launch_bounds__(FIN_THREADS)
knn2_finish_global_kernel(uint2 *__restrict__ partial, int nslots, int nq, int q_stride,
                          const RowMap map, unsigned long long *__restrict__ top2, int rearm)
{
    const int q = blockIdx.x * FIN_THREADS + threadIdx.x;
    if (q >= nq) return;
    uint32_t k1 = KEY_EMPTY, k2 = KEY_EMPTY;
    for (int s = 0; s < nslots; ++s) merge_pair(k1, k2, partial[(size_t)s * q_stride + q]);
    if (rearm && nslots) partial[q] = make_uint2(KEY_EMPTY, KEY_EMPTY);       // the merged slot is ready for the next scan
    top2[2 * (size_t)q] = global_key(k1, map);
    top2[2 * (size_t)q + 1] = global_key(k2, map);
}

// Fused finish + all-gather: the merged keys of this shard go directly into every peer's gather buffer
// with plain 16-byte stores to peer-mapped addresses (NVLink / NVSwitch); one thread = one query.
__global__ void __launch_bounds__(FIN_THREADS)
knn2_finish_global_push_kernel(uint2 *__restrict__ partial, int nslots, int nq, int q_stride,
                               const RowMap map, unsigned long long *const *__restrict__ peer_bufs,
                               int nparts, int my_part, int rearm)
{
    const int q = blockIdx.x * FIN_THREADS + threadIdx.x;
    if (q >= nq) return;
    uint32_t k1 = KEY_EMPTY, k2 = KEY_EMPTY;
    for (int s = 0; s < nslots; ++s) merge_pair(k1, k2, partial[(size_t)s * q_stride + q]);
    if (rearm && nslots) partial[q] = make_uint2(KEY_EMPTY, KEY_EMPTY);       // the merged slot is ready for the next scan
    ulonglong2 v;
    v.x = global_key(k1, map);
    v.y = global_key(k2, map);
    const size_t at = ((size_t)my_part * nq + q) * 2;
    for (int p = 0; p < nparts; ++p)
        *reinterpret_cast<ulonglong2 *>(peer_bufs[p] + at) = v;
}

// Merge of all-gathered shard results: parts[nparts][nq][2] packed 64-bit keys.
__global__ void __launch_bounds__(FIN_THREADS)
merge_top2_kernel(const unsigned long long *__restrict__ parts, int nparts, int nq,
                  unsigned long long *__restrict__ out)
{
    const int q = blockIdx.x * FIN_THREADS + threadIdx.x;
    if (q >= nq) return;
    unsigned long long k1 = ~0ull, k2 = ~0ull;
    for (int p = 0; p < nparts; ++p) {
        const unsigned long long a = parts[((size_t)p * nq + q) * 2], b = parts[((size_t)p * nq + q) * 2 + 1];
        const unsigned long long n2 = min(max(k1, a), min(k2, b));
        k1 = min(k1, a);
        k2 = n2;
    }
    out[2 * (size_t)q] = k1;
    out[2 * (size_t)q + 1] = k2;
}

// Problem table of one (q, t) pair cut into `splits` row ranges, written on the device so that
// fully asynchronous callers never depend on a host staging buffer staying untouched.
__global__ void fill_split_problems_kernel(KnnProblem *probs, const uint8_t *q, const uint8_t *t,
                                           const int32_t *nq_dev, const int32_t *nt_dev,
                                           int nq, int nt, int splits, int rows_per)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= splits) return;
    KnnProblem P;
    P.q = q; P.t = t; P.nq_dev = nq_dev; P.nt_dev = nt_dev; P.nq = nq; P.nt = nt;
    P.t_lo = min(s * rows_per, nt);
    P.t_hi = min((s + 1) * rows_per, nt);
    probs[s] = P;
}

int launch_fill_split_problems(aslam_ctx *ctx, KnnProblem *d_probs, const uint8_t *q, const uint8_t *t,
                               const int32_t *nq_dev, const int32_t *nt_dev, int nq, int nt, int splits)
{
    const int rows_per = (nt + splits - 1) / splits;
    ProfScope ps(ctx, PK_MISC);
    fill_split_problems_kernel<<<(splits + 127) / 128, 128, 0, ctx->stream>>>(d_probs, q, t, nq_dev, nt_dev,
                                                                              nq, nt, splits, rows_per);
    ASLAM_LAUNCH_CHECK();
    return ASLAM_OK;
}

int launch_knn2_partial(aslam_ctx *ctx, const KnnProblem *d_probs, int nprobs, int max_nq,
                        uint2 *d_partial, int q_stride, int merge_div, bool armed)
{
    if (nprobs <= 0 || max_nq <= 0) return ASLAM_OK;
    dim3 grid((max_nq + KNN_QBLOCK - 1) / KNN_QBLOCK, nprobs);
    static const int csa = [] { const char *e = getenv("ASLAM_KNN_CSA"); return e ? atoi(e) : ASLAM_KNN_CSA; }();
    if (merge_div > 1 && !armed)
        ASLAM_CUDA(cudaMemsetAsync(d_partial, 0xff, sizeof(uint2) * (size_t)(nprobs / merge_div) * q_stride, ctx->stream));
    const KnnMul M{2u, 4u, 65536u, 128u};
    // ASLAM_KNN_SMEM_PAD (A/B runs): unused dynamic shared memory per CTA, to cap the scan's resident CTAs per SM and leave
    // room for the kernels of other frame streams (3 CTAs per SM fit by registers; 64 KB of padding leaves 2)
    static const int pad = [] { const char *e = getenv("ASLAM_KNN_SMEM_PAD"); return e ? atoi(e) : 0; }();
    if (pad > 0 && !ctx->knn_pad_set) {
        ASLAM_CUDA(cudaFuncSetAttribute(knn2_partial_v2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad));
        ASLAM_CUDA(cudaFuncSetAttribute(knn2_partial_v2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad));
        ctx->knn_pad_set = true;
    }
    ProfScope ps(ctx, PK_KNN_PARTIAL);
    if (csa == 7) knn2_partial_v2_kernel<false, false><<<grid, KNN_THREADS, pad, ctx->stream>>>(d_probs, d_partial, q_stride, M, merge_div, ExchangeArgs());
    else knn2_partial_v2_kernel<true, false><<<grid, KNN_THREADS, pad, ctx->stream>>>(d_probs, d_partial, q_stride, M, merge_div, ExchangeArgs());
    ASLAM_LAUNCH_CHECK();
    return ASLAM_OK;
}

int launch_knn2_scan_exchange(aslam_ctx *ctx, const KnnProblem *d_probs, int nprobs, int nq, uint2 *d_partial,
                              int q_stride, const ExchangeArgs &X, bool overlap_previous)
{
    if (nprobs <= 0 || nq <= 0) return fail(ASLAM_E_INVALID, "fused exchange needs at least one row range and one query");
    const int nqb = (nq + KNN_QBLOCK - 1) / KNN_QBLOCK;
    // One CTA per query block ends up waiting for the peers: they must never fill the GPU while CTAs of other
    // blocks have not started (those would then never push, and the peers would wait for ever).
    if (nqb > KNN_FUSED_MAX_QBLOCKS)
        return fail(ASLAM_E_INVALID, "fused exchange takes at most %d queries per call (got %d); split the batch",
                    KNN_FUSED_MAX_QBLOCKS * KNN_QBLOCK, nq);
    if (X.nparts < 1 || X.nparts > 64 || X.my_part < 0 || X.my_part >= X.nparts || X.nq != nq || X.round < 1)
        return fail(ASLAM_E_INVALID, "bad exchange arguments");
    dim3 grid(nqb, nprobs);
    static const int csa = [] { const char *e = getenv("ASLAM_KNN_CSA"); return e ? atoi(e) : ASLAM_KNN_CSA; }();
    const KnnMul M{2u, 4u, 65536u, 128u};
    ProfScope ps(ctx, PK_KNN_PARTIAL);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(KNN_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = overlap_previous ? 1 : 0;
    if (csa == 7) ASLAM_CUDA(cudaLaunchKernelEx(&cfg, knn2_partial_v2_kernel<false, true>, d_probs, d_partial, q_stride, M, nprobs, X));
    else ASLAM_CUDA(cudaLaunchKernelEx(&cfg, knn2_partial_v2_kernel<true, true>, d_probs, d_partial, q_stride, M, nprobs, X));
    ASLAM_LAUNCH_CHECK();
    return ASLAM_OK;
}

int launch_knn2_finish_matches(aslam_ctx *ctx, const KnnProblem *d_probs, const uint2 *d_partial,
                               int ngroups, int splits, int q_stride, RatioTest rt,
                               aslam_match *d_out, int out_stride, int32_t *d_nout,
                               int32_t *d_raw_idx, int32_t *d_raw_dist, const DynFilter &filter)
{
    if (ngroups <= 0) return ASLAM_OK;
    ProfScope ps(ctx, PK_KNN_FINISH);
    knn2_finish_matches_kernel<<<ngroups, FIN_THREADS, 0, ctx->stream>>>(
        d_probs, d_partial, splits, q_stride, rt, d_out, out_stride, d_nout, d_raw_idx, d_raw_dist, filter);
    ASLAM_LAUNCH_CHECK();
    return ASLAM_OK;
}

int launch_knn2_finish_counts(aslam_ctx *ctx, const KnnProblem *d_probs, const uint2 *d_partial,
                              int nslots, int splits, int q_stride, RatioTest rt, int32_t *d_good)
{
    if (nslots <= 0) return ASLAM_OK;
    ProfScope ps(ctx, PK_KNN_FINISH);
    knn2_finish_counts_kernel<<<nslots, FIN_THREADS, 0, ctx->stream>>>(d_probs, d_partial, splits, q_stride, rt, d_good);
    ASLAM_LAUNCH_CHECK();
    return ASLAM_OK;
}

int launch_knn2_finish_global(aslam_ctx *ctx, uint2 *d_partial, int nslots, int nq,
                              int q_stride, const RowMap &map, uint64_t *d_top2, bool rearm)
{
    if (nq <= 0) return ASLAM_OK;
    ProfScope ps(ctx, PK_KNN_FINISH);
    knn2_finish_global_kernel<<<(nq + FIN_THREADS - 1) / FIN_THREADS, FIN_THREADS, 0, ctx->stream>>>(
        d_partial, nslots, nq, q_stride, map, (unsigned long long *)d_top2, rearm ? 1 : 0);
    ASLAM_LAUNCH_CHECK();
    return ASLAM_OK;
}

int launch_knn2_finish_global_push(aslam_ctx *ctx, uint2 *d_partial, int nslots, int nq,
                                   int q_stride, const RowMap &map, uint64_t *const *d_peer_bufs,
                                   int nparts, int my_part, bool rearm)
{
    if (nq <= 0) return ASLAM_OK;
    ProfScope ps(ctx, PK_KNN_FINISH);
    knn2_finish_global_push_kernel<<<(nq + FIN_THREADS - 1) / FIN_THREADS, FIN_THREADS, 0, ctx->stream>>>(
        d_partial, nslots, nq, q_stride, map, (unsigned long long *const *)d_peer_bufs, nparts, my_part, rearm ? 1 : 0);
    ASLAM_LAUNCH_CHECK();
    return ASLAM_OK;
}

static RatioTest make_ratio(double ratio, int as_double)
{
    RatioTest rt;
    rt.ratio = ratio; rt.as_double = as_double; rt.disabled = ratio < 0 ? 1 : 0;
    return rt;
}

// How many row ranges to cut a train set of nt rows into.  The scan kernel keeps a few CTAs of equal work
// resident per SM (occupancy API), so the grid should be a whole number of such waves (a 1.4-wave grid runs as long as
// a 2-wave grid): among the cuts that leave at least 128 rows per range, take the one with the best
// wave efficiency, preferring fewer ranges on ties.
int choose_splits(const aslam_ctx *ctx, int nq, long long nt, int nsets)
{
    const long long units = (long long)((nq + KNN_QBLOCK - 1) / KNN_QBLOCK) * nsets;      // CTAs per range
    static const int ctas_per_sm = [] {                // resident scan CTAs per SM, as the compiler's register count allows
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, knn2_partial_v2_kernel<true, false>, KNN_THREADS, 0) != cudaSuccess || n < 1) n = 4;
        return n;
    }();
    const double slots = (double)ctas_per_sm * ctx->sm_count;
    const long long max_s = nt / 128 < 1 ? 1 : (nt / 128 > 4096 ? 4096 : nt / 128);
    static const int forced = [] { const char *e = getenv("ASLAM_KNN_SPLITS"); return e ? atoi(e) : 0; }();
    if (forced > 0) return (int)(forced < max_s ? forced : max_s);
    // first choice: the coarsest cut with at least ASLAM_KNN_WAVES (10) nearly full waves -- CTAs of several
    // waves run out of phase, which keeps the POPC pipe busier than one wave of CTAs in lock step
    static const int min_waves = [] { const char *e = getenv("ASLAM_KNN_WAVES"); return e ? atoi(e) : 10; }();
    int best = 1;
    double best_score = -1.0;
    for (long long s = 1; s <= max_s; ++s) {
        const double waves = (double)(units * s) / slots;
        const double eff = waves / ceil(waves);
        if (waves >= (double)min_waves && eff >= 0.95) return (int)s;
        const double score = eff * (waves < 4.0 ? 0.8 + 0.05 * waves : 1.0);
        if (score > best_score + 1e-9) { best_score = score; best = (int)s; }
    }
    return best;
}

}  // namespace aslam

using namespace aslam;

// ------------------------------------------------------------------ host-buffer entry points
// Scratch layout (device): [problems][q rows][t rows][partial][matches][nout][raw idx][raw dist]
struct HostDynFilter {          // host buffers of aslam_match_knn2_dynamic
    const aslam_keypoint *q_kps, *t_kps;
    const aslam_box *boxes;
    int nboxes, half_open;
    int *nfiltered;
};

static int run_single(aslam_ctx *ctx, const uint8_t *q, int nq, const uint8_t *t, int nt,
                      bool inputs_on_device, RatioTest rt, aslam_match *out, int *nout,
                      int32_t *raw_idx, int32_t *raw_dist, aslam_match *d_out_user, int32_t *d_nout_user,
                      const HostDynFilter *hf = nullptr)
{
    ASLAM_CUDA(cudaSetDevice(ctx->device));
    if (nt > KNN_MAX_ROWS) return fail(ASLAM_E_INVALID, "train set has %d rows; at most %d per call", nt, KNN_MAX_ROWS);
    const int splits = choose_splits(ctx, nq, nt, 1);
    const int q_stride = (int)align_up((size_t)nq, 64);
    const size_t o_prob = 0;
    const size_t o_q = align_up(o_prob + sizeof(KnnProblem) * splits, 256);
    const size_t o_t = align_up(o_q + (inputs_on_device ? 0 : (size_t)nq * 32), 256);
    const size_t o_part = align_up(o_t + (inputs_on_device ? 0 : (size_t)nt * 32), 256);
    const size_t o_match = align_up(o_part + sizeof(uint2) * (size_t)splits * q_stride, 256);
    const size_t o_nout = align_up(o_match + sizeof(aslam_match) * (size_t)nq, 256);
    const size_t o_ridx = align_up(o_nout + 256, 256);
    const size_t o_rdist = align_up(o_ridx + sizeof(int32_t) * 2 * (size_t)q_stride, 256);
    const size_t o_qkp = align_up(o_rdist + sizeof(int32_t) * 2 * (size_t)q_stride, 256);
    const size_t o_tkp = align_up(o_qkp + (hf ? sizeof(aslam_keypoint) * (size_t)nq : 0), 256);
    const size_t o_box = align_up(o_tkp + (hf ? sizeof(aslam_keypoint) * (size_t)nt : 0), 256);
    const size_t total = align_up(o_box + (hf ? sizeof(aslam_box) * (size_t)hf->nboxes : 0), 256);
    const size_t h_need = 256;
    int rc = ensure_scratch(ctx, total, h_need);
    if (rc) return rc;
    uint8_t *base = (uint8_t *)ctx->d_scratch;
    const uint8_t *dq = inputs_on_device ? q : base + o_q;
    const uint8_t *dt = inputs_on_device ? t : base + o_t;
    if (!inputs_on_device) {
        ASLAM_CUDA(cudaMemcpyAsync(base + o_q, q, (size_t)nq * 32, cudaMemcpyHostToDevice, ctx->stream));
        ASLAM_CUDA(cudaMemcpyAsync(base + o_t, t, (size_t)nt * 32, cudaMemcpyHostToDevice, ctx->stream));
    }
    KnnProblem *dp = (KnnProblem *)(base + o_prob);
    rc = launch_fill_split_problems(ctx, dp, dq, dt, nullptr, nullptr, nq, nt, splits);
    if (rc) return rc;
    uint2 *dpart = (uint2 *)(base + o_part);
    rc = launch_knn2_partial(ctx, dp, splits, nq, dpart, q_stride, splits);
    if (rc) return rc;
    aslam_match *dm = d_out_user ? d_out_user : (aslam_match *)(base + o_match);
    int32_t *dn = d_nout_user ? d_nout_user : (int32_t *)(base + o_nout);
    int32_t *dri = raw_idx ? (int32_t *)(base + o_ridx) : nullptr;
    int32_t *drd = raw_idx ? (int32_t *)(base + o_rdist) : nullptr;
    DynFilter F;
    if (hf) {
        ASLAM_CUDA(cudaMemcpyAsync(base + o_qkp, hf->q_kps, sizeof(aslam_keypoint) * (size_t)nq, cudaMemcpyHostToDevice, ctx->stream));
        ASLAM_CUDA(cudaMemcpyAsync(base + o_tkp, hf->t_kps, sizeof(aslam_keypoint) * (size_t)nt, cudaMemcpyHostToDevice, ctx->stream));
        ASLAM_CUDA(cudaMemcpyAsync(base + o_box, hf->boxes, sizeof(aslam_box) * (size_t)hf->nboxes, cudaMemcpyHostToDevice, ctx->stream));
        F.q_kps = (const aslam_keypoint *)(base + o_qkp); F.t_kps = (const aslam_keypoint *)(base + o_tkp);
        F.boxes = (const aslam_box *)(base + o_box); F.nboxes = hf->nboxes; F.half_open = hf->half_open;
        F.nfiltered = dn + 1;           // second word of the 256-byte count slot
    }
    rc = launch_knn2_finish_matches(ctx, dp, dpart, 1, splits, q_stride, rt, dm, nq, dn, dri, drd, F);
    if (rc) return rc;
    if (d_out_user) return ASLAM_OK;     // fully asynchronous device form
    if (raw_idx) {
        ASLAM_CUDA(cudaMemcpyAsync(raw_idx, dri, sizeof(int32_t) * 2 * (size_t)nq, cudaMemcpyDeviceToHost, ctx->stream));
        ASLAM_CUDA(cudaMemcpyAsync(raw_dist, drd, sizeof(int32_t) * 2 * (size_t)nq, cudaMemcpyDeviceToHost, ctx->stream));
    }
    int32_t *hn = (int32_t *)ctx->h_scratch;
    ASLAM_CUDA(cudaMemcpyAsync(hn, dn, sizeof(int32_t) * (hf ? 2 : 1), cudaMemcpyDeviceToHost, ctx->stream));
    ASLAM_CUDA(cudaStreamSynchronize(ctx->stream));
    if (nout) *nout = *hn;
    if (hf && hf->nfiltered) *hf->nfiltered = hn[1];
    if (out && *hn > 0)
        ASLAM_CUDA(cudaMemcpy(out, dm, sizeof(aslam_match) * (size_t)*hn, cudaMemcpyDeviceToHost));
    return ASLAM_OK;
}

// Device-resident rows on both sides, results to host buffers (the loop database's candidate re-match).
int knn2_resident_to_host(aslam_ctx *ctx, const uint8_t *d_q, int nq, const uint8_t *d_t, int nt, double ratio, int as_double,
                          aslam_match *out, int *nout)
{
    return run_single(ctx, d_q, nq, d_t, nt, true, make_ratio(ratio, as_double), out, nout, nullptr, nullptr, nullptr, nullptr);
}

extern "C" int aslam_match_knn2(aslam_ctx *ctx, const uint8_t *q, int nq, const uint8_t *t, int nt,
                                float ratio, aslam_match *out, int *nout)
{
    ASLAM_REQUIRE(ctx && nout, "NULL argument");
    ASLAM_REQUIRE(nq >= 0 && nt >= 0, "negative row count");
    *nout = 0;
    // CudaMatcher.cpp:35-37 returns early on an empty side; :60 needs two neighbours.
    if (nq == 0 || nt < 2) return ASLAM_OK;
    ASLAM_REQUIRE(q && t && out, "NULL descriptor or output pointer");
    return run_single(ctx, q, nq, t, nt, false, make_ratio((double)ratio, 0), out, nout, nullptr, nullptr, nullptr, nullptr);
}

extern "C" int aslam_match_knn2_dynamic(aslam_ctx *ctx, const uint8_t *q, const aslam_keypoint *q_kps, int nq,
                                        const uint8_t *t, const aslam_keypoint *t_kps, int nt, float ratio,
                                        const aslam_box *boxes, int nboxes, int half_open,
                                        aslam_match *out, int *nout, int *nfiltered)
{
    ASLAM_REQUIRE(ctx && nout, "NULL argument");
    ASLAM_REQUIRE(nq >= 0 && nt >= 0 && nboxes >= 0, "negative count");
    *nout = 0;
    if (nfiltered) *nfiltered = 0;
    if (nq == 0 || nt < 2) return ASLAM_OK;
    ASLAM_REQUIRE(q && t && out, "NULL descriptor or output pointer");
    if (nboxes == 0)
        return run_single(ctx, q, nq, t, nt, false, make_ratio((double)ratio, 0), out, nout, nullptr, nullptr, nullptr, nullptr);
    ASLAM_REQUIRE(q_kps && t_kps && boxes, "NULL keypoints or boxes");
    const HostDynFilter hf{q_kps, t_kps, boxes, nboxes, half_open ? 1 : 0, nfiltered};
    return run_single(ctx, q, nq, t, nt, false, make_ratio((double)ratio, 0), out, nout, nullptr, nullptr, nullptr, nullptr, &hf);
}

extern "C" int aslam_knn2_raw(aslam_ctx *ctx, const uint8_t *q, int nq, const uint8_t *t, int nt,
                              int32_t *idx, int32_t *dist)
{
    ASLAM_REQUIRE(ctx && (nq == 0 || (idx && dist)), "NULL argument");
    ASLAM_REQUIRE(nq >= 0 && nt >= 0, "negative row count");
    if (nq == 0) return ASLAM_OK;
    if (nt == 0) {
        for (int i = 0; i < 2 * nq; ++i) { idx[i] = -1; dist[i] = -1; }
        return ASLAM_OK;
    }
    ASLAM_REQUIRE(q && t, "NULL descriptor pointer");
    int n = 0;
    return run_single(ctx, q, nq, t, nt, false, make_ratio(-1.0, 0), nullptr, &n, idx, dist, nullptr, nullptr);
}

extern "C" int aslam_match_knn2_device(aslam_ctx *ctx, const uint8_t *d_q, int nq, const uint8_t *d_t, int nt,
                                       float ratio, aslam_match *d_out, int32_t *d_nout)
{
    ASLAM_REQUIRE(ctx && d_out && d_nout, "NULL argument");
    ASLAM_REQUIRE(nq >= 0 && nt >= 0, "negative row count");
    if (nq == 0 || nt < 2) {
        ASLAM_CUDA(cudaMemsetAsync(d_nout, 0, sizeof(int32_t), ctx->stream));
        return ASLAM_OK;
    }
    ASLAM_REQUIRE(d_q && d_t, "NULL descriptor pointer");
    return run_single(ctx, d_q, nq, d_t, nt, true, make_ratio((double)ratio, 0), nullptr, nullptr, nullptr, nullptr, d_out, d_nout);
}

extern "C" int aslam_match_knn2_resident(aslam_ctx *ctx, const uint8_t *d_q, int nq, const uint8_t *d_t, int nt,
                                         float ratio, aslam_match *out, int *nout)
{
    ASLAM_REQUIRE(ctx && nout, "NULL argument");
    ASLAM_REQUIRE(nq >= 0 && nt >= 0, "negative row count");
    *nout = 0;
    if (nq == 0 || nt < 2) return ASLAM_OK;
    ASLAM_REQUIRE(d_q && d_t && out, "NULL descriptor or output pointer");
    return run_single(ctx, d_q, nq, d_t, nt, true, make_ratio((double)ratio, 0), out, nout, nullptr, nullptr, nullptr, nullptr);
}

extern "C" int aslam_match_multi(aslam_ctx *ctx, const uint8_t *q, int nq, const uint8_t *t,
                                 const int32_t *t_start, int nsets, float ratio,
                                 aslam_match *out, int32_t *nout)
{
    ASLAM_REQUIRE(ctx && nout && t_start, "NULL argument");
    ASLAM_REQUIRE(nq >= 0 && nsets >= 0, "negative count");
    for (int s = 0; s < nsets; ++s) nout[s] = 0;
    if (nq == 0 || nsets == 0) return ASLAM_OK;
    const long long rows = t_start[nsets];
    ASLAM_REQUIRE(q && out && (rows == 0 || t), "NULL descriptor or output pointer");
    for (int s = 0; s < nsets; ++s) {
        ASLAM_REQUIRE(t_start[s + 1] >= t_start[s], "t_start must be non-decreasing");
        if (t_start[s + 1] - t_start[s] > KNN_MAX_ROWS) return fail(ASLAM_E_INVALID, "train set %d too large", s);
    }
    ASLAM_CUDA(cudaSetDevice(ctx->device));
    const int q_stride = (int)align_up((size_t)nq, 64);
    const size_t o_q = align_up(sizeof(KnnProblem) * (size_t)nsets, 256);
    const size_t o_t = align_up(o_q + (size_t)nq * 32, 256);
    const size_t o_part = align_up(o_t + (size_t)rows * 32, 256);
    const size_t o_match = align_up(o_part + sizeof(uint2) * (size_t)nsets * q_stride, 256);
    const size_t o_nout = align_up(o_match + sizeof(aslam_match) * (size_t)nsets * nq, 256);
    const size_t total = align_up(o_nout + sizeof(int32_t) * (size_t)nsets, 256);
    int rc = ensure_scratch(ctx, total, sizeof(KnnProblem) * (size_t)nsets);
    if (rc) return rc;
    uint8_t *base = (uint8_t *)ctx->d_scratch;
    ASLAM_CUDA(cudaMemcpyAsync(base + o_q, q, (size_t)nq * 32, cudaMemcpyHostToDevice, ctx->stream));
    if (rows) ASLAM_CUDA(cudaMemcpyAsync(base + o_t, t, (size_t)rows * 32, cudaMemcpyHostToDevice, ctx->stream));
    KnnProblem *hp = (KnnProblem *)ctx->h_scratch;
    for (int s = 0; s < nsets; ++s) {
        hp[s].q = base + o_q; hp[s].t = base + o_t + (size_t)t_start[s] * 32;
        hp[s].nq_dev = nullptr; hp[s].nt_dev = nullptr;
        hp[s].nq = nq; hp[s].nt = t_start[s + 1] - t_start[s];
        hp[s].t_lo = 0; hp[s].t_hi = hp[s].nt;
    }
    KnnProblem *dp = (KnnProblem *)base;
    ASLAM_CUDA(cudaMemcpyAsync(dp, hp, sizeof(KnnProblem) * (size_t)nsets, cudaMemcpyHostToDevice, ctx->stream));
    uint2 *dpart = (uint2 *)(base + o_part);
    rc = launch_knn2_partial(ctx, dp, nsets, nq, dpart, q_stride);
    if (rc) return rc;
    rc = launch_knn2_finish_matches(ctx, dp, dpart, nsets, 1, q_stride, make_ratio((double)ratio, 0),
                                    (aslam_match *)(base + o_match), nq, (int32_t *)(base + o_nout), nullptr, nullptr);
    if (rc) return rc;
    ASLAM_CUDA(cudaMemcpyAsync(nout, base + o_nout, sizeof(int32_t) * (size_t)nsets, cudaMemcpyDeviceToHost, ctx->stream));
    ASLAM_CUDA(cudaMemcpyAsync(out, base + o_match, sizeof(aslam_match) * (size_t)nsets * nq, cudaMemcpyDeviceToHost, ctx->stream));
    ASLAM_CUDA(cudaStreamSynchronize(ctx->stream));
    return ASLAM_OK;
}

extern "C" size_t aslam_loop_exchange_buffer_bytes(int nparts, int nq)
{
    if (nparts < 1 || nq < 0) return 0;
    const size_t nqb = ((size_t)nq + KNN_QBLOCK - 1) / KNN_QBLOCK;
    return sizeof(uint64_t) * ((size_t)2 * nparts * nq * 2 + (size_t)2 * nparts * nqb);
}

extern "C" int aslam_loop_merge_top2_device(aslam_ctx *ctx, const uint64_t *d_parts, int nparts, int nq, uint64_t *d_out)
{
    ASLAM_REQUIRE(ctx && d_parts && d_out, "NULL argument");
    ASLAM_REQUIRE(nparts >= 1 && nq >= 0, "bad sizes");
    if (nq == 0) return ASLAM_OK;
    ASLAM_CUDA(cudaSetDevice(ctx->device));
    ProfScope ps(ctx, PK_MERGE);
    merge_top2_kernel<<<(nq + FIN_THREADS - 1) / FIN_THREADS, FIN_THREADS, 0, ctx->stream>>>(
        (const unsigned long long *)d_parts, nparts, nq, (unsigned long long *)d_out);
    ASLAM_LAUNCH_CHECK();
    return ASLAM_OK;
}
