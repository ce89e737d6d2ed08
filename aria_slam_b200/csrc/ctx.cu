// Context, error state, pinned-memory helpers and the integer-pipe micro-benchmarks.
#include <cuda_fp16.h>

#include "common.cuh"

namespace aslam {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
#if ASLAM_DEBUG_BOUNDS
static std::vector<int (*)(unsigned long long *)> &bounds_readers()
{
    static std::vector<int (*)(unsigned long long *)> r;
    return r;
}
void register_bounds_reader(int (*reader)(unsigned long long *)) { bounds_readers().push_back(reader); }
#endif

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int ensure_scratch(aslam_ctx *ctx, size_t dev_bytes, size_t host_bytes)
{
    if (dev_bytes > ctx->scratch_bytes) {
        if (ctx->d_scratch) { ASLAM_CUDA(cudaStreamSynchronize(ctx->stream)); ASLAM_CUDA(cudaFree(ctx->d_scratch)); }
        ctx->d_scratch = nullptr; ctx->scratch_bytes = 0;
        size_t want = align_up(dev_bytes + dev_bytes / 4, 1 << 20);
        ASLAM_CUDA(cudaMalloc(&ctx->d_scratch, want));
        ctx->scratch_bytes = want;
    }
    if (host_bytes > ctx->h_scratch_bytes) {
        if (ctx->h_scratch) { ASLAM_CUDA(cudaStreamSynchronize(ctx->stream)); ASLAM_CUDA(cudaFreeHost(ctx->h_scratch)); }
        ctx->h_scratch = nullptr; ctx->h_scratch_bytes = 0;
        size_t want = align_up(host_bytes + host_bytes / 4, 1 << 20);
        ASLAM_CUDA(cudaMallocHost(&ctx->h_scratch, want));
        ctx->h_scratch_bytes = want;
    }
    return ASLAM_OK;
}

static cudaEvent_t prof_event(aslam_ctx *c)
{
    if (!c->prof_pool.empty()) { cudaEvent_t e = c->prof_pool.back(); c->prof_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

ProfScope::ProfScope(aslam_ctx *ctx, int k) : c(ctx), kind(k)
{
    if (!c->prof_on) return;
    e0 = prof_event(c); e1 = prof_event(c);
    cudaEventRecord(e0, c->stream);
}

ProfScope::~ProfScope()
{
    if (!e0) return;
    cudaEventRecord(e1, c->stream);
    c->prof_pending.push_back({kind, e0, e1});
}

}  // namespace aslam

extern "C" {

static const char *k_prof_names[aslam::PK_COUNT] = {"resize", "fast_nms", "blur", "select_harris", "orient_brief",
                                                    "knn2_partial", "knn2_finish", "merge_top2", "misc"};

int aslam_prof_enable(aslam_ctx *ctx, int on)
{
    ASLAM_REQUIRE(ctx, "ctx is NULL");
    ctx->prof_on = on != 0;
    return ASLAM_OK;
}

int aslam_prof_count(void) { return aslam::PK_COUNT; }
const char *aslam_prof_name(int kind) { return kind >= 0 && kind < aslam::PK_COUNT ? k_prof_names[kind] : ""; }

int aslam_prof_read(aslam_ctx *ctx, int kind, double *total_ms, int64_t *launches, int reset)
{
    ASLAM_REQUIRE(ctx && kind >= 0 && kind < aslam::PK_COUNT, "bad argument");
    ASLAM_CUDA(cudaStreamSynchronize(ctx->stream));
    for (auto &r : ctx->prof_pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) { ctx->prof_ms[r.kind] += ms; ctx->prof_n[r.kind] += 1; }
        ctx->prof_pool.push_back(r.e0); ctx->prof_pool.push_back(r.e1);
    }
    ctx->prof_pending.clear();
    if (total_ms) *total_ms = ctx->prof_ms[kind];
    if (launches) *launches = ctx->prof_n[kind];
    if (reset) { ctx->prof_ms[kind] = 0; ctx->prof_n[kind] = 0; }
    return ASLAM_OK;
}

const char *aslam_last_error(void) { return aslam::g_err; }
const char *aslam_version(void) { return "aslam_b200 0.2 (sm_100a)"; }

int aslam_debug_bounds(uint64_t *violations)
{
#if ASLAM_DEBUG_BOUNDS
    if (violations) {
        unsigned long long sum = 0;
        ASLAM_CUDA(cudaDeviceSynchronize());
        for (auto reader : aslam::bounds_readers()) {
            unsigned long long v = 0;
            if (reader(&v)) return aslam::fail(ASLAM_E_CUDA, "cannot read a bounds counter");
            sum += v;
        }
        *violations = sum;
    }
    return 1;
#else
    if (violations) *violations = 0;
    return 0;
#endif
}

int aslam_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
uint64_t aslam_kernel_launches(void) { return aslam::g_launches.load(); }

int aslam_ctx_create(int device, void *stream, aslam_ctx **out)
{
    if (!out) return aslam::fail(ASLAM_E_INVALID, "aslam_ctx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        (void)cudaGetLastError();
        return aslam::fail(ASLAM_E_NODEVICE, "no CUDA device visible (%s); this path has no CPU fallback",
                           e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= ndev) return aslam::fail(ASLAM_E_INVALID, "device %d out of range [0,%d)", device, ndev);
    ASLAM_CUDA(cudaSetDevice(device));
    aslam_ctx *c = new aslam_ctx();
    c->device = device;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        cudaError_t se = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (se != cudaSuccess) { delete c; return aslam::fail(ASLAM_E_CUDA, "cudaStreamCreate -> %s", cudaGetErrorString(se)); }
        c->owns_stream = true;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    *out = c;
    return ASLAM_OK;
}

void aslam_ctx_destroy(aslam_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    if (ctx->h_scratch) cudaFreeHost(ctx->h_scratch);
    for (auto &r : ctx->prof_pending) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    for (auto e : ctx->prof_pool) cudaEventDestroy(e);
    if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int aslam_ctx_sync(aslam_ctx *ctx)
{
    ASLAM_REQUIRE(ctx, "ctx is NULL");
    ASLAM_CUDA(cudaStreamSynchronize(ctx->stream));
    return ASLAM_OK;
}

void *aslam_ctx_stream(aslam_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int aslam_ctx_sm_count(aslam_ctx *ctx) { return ctx ? ctx->sm_count : 0; }

int aslam_host_alloc(size_t bytes, void **out)
{
    ASLAM_REQUIRE(out, "out is NULL");
    ASLAM_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
    return ASLAM_OK;
}

int aslam_host_alloc_wc(size_t bytes, void **out)
{
    ASLAM_REQUIRE(out, "out is NULL");
    ASLAM_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocWriteCombined));
    return ASLAM_OK;
}

int aslam_memcpy_async(aslam_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    ASLAM_REQUIRE(ctx && (bytes == 0 || (dst && src)), "NULL argument");
    ASLAM_CUDA(cudaSetDevice(ctx->device));
    if (bytes) ASLAM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->stream));
    return ASLAM_OK;
}

int aslam_host_free(void *p)
{
    if (p) ASLAM_CUDA(cudaFreeHost(p));
    return ASLAM_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ pipe micro-benchmarks
// Eight independent dependency chains per thread so the measured rate is the pipe's, not the
// latency's.  The result is written so the compiler cannot drop the loop.
__device__ __forceinline__ uint32_t hmax2_bits(uint32_t a, uint32_t b)
{
    __half2 x = *reinterpret_cast<__half2 *>(&a), y = *reinterpret_cast<__half2 *>(&b);
    __half2 r = __hmax2(x, y);
    return *reinterpret_cast<uint32_t *>(&r);
}

template <int WHICH>
__global__ void __launch_bounds__(256) pipe_bench_kernel(uint32_t *out, int iters, uint32_t seed)
{
    uint32_t a[8], h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u;
        h[i] = 0x64006400u | ((a[i] >> 3) & 0x00ff00ffu);     // two halves in [1024, 1279]
    }
    uint32_t b = seed ^ 0x5bd1e995u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (WHICH == 0) a[i] = __popc(a[i]) + b;                 // POPC (+1 IADD to keep a live chain)
                else if (WHICH == 1) a[i] = (a[i] & b) ^ a[(i + 1) & 7];       // LOP3
                else if (WHICH == 2) a[i] = a[i] + b + a[(i + 5) & 7];   // IADD3
                else if (WHICH == 3) a[i] = __vimax3_u16x2(a[i], b, a[(i + 3) & 7]);     // VIMNMX3.U16x2
                else if (WHICH == 4) h[i] = hmax2_bits(h[i], h[(i + 3) & 7]);          // HMNMX2
                else if (WHICH == 5) {                                   // both, independent chains
                    a[i] = __vimax3_u16x2(a[i], b, a[(i + 3) & 7]);
                    h[i] = hmax2_bits(h[i], h[(i + 3) & 7]);
                } else if (WHICH == 6) a[i] = a[i] * b + a[(i + 3) & 7]; // IMAD
                else if (WHICH == 7) a[i] = __byte_perm(a[i], a[(i + 3) & 7], b);        // PRMT
                else if (WHICH == 8) a[i] = __float_as_uint(__fmaf_rn(__uint_as_float(a[i]), __uint_as_float(b), __uint_as_float(a[(i + 3) & 7])));   // FFMA
                else if (WHICH >= 9 && (i & 1) == 0) {                   // packed f32x2 on register pairs (a[i], a[i+1]): 4 chains
                    unsigned long long x, y, z;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "r"(a[i]), "r"(a[i + 1]));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "r"(a[(i + 2) & 7]), "r"(a[(i + 3) & 7]));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(z) : "r"(b), "r"(b));
                    if (WHICH == 9) asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(x) : "l"(x), "l"(z), "l"(y));          // FFMA2
                    else if (WHICH == 10) asm("add.rn.f32x2 %0, %1, %2;" : "=l"(x) : "l"(x), "l"(y));               // FADD2
                    else asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(x) : "l"(x), "l"(z));                              // FMUL2
                    asm("mov.b64 {%0, %1}, %2;" : "=r"(a[i]), "=r"(a[i + 1]) : "l"(x));
                }
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i] ^ h[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" int aslam_microbench(aslam_ctx *ctx, int which, double *gops)
{
    ASLAM_REQUIRE(ctx && gops, "NULL argument");
    ASLAM_REQUIRE(which >= 0 && which <= 11, "which must be 0..11");
    ASLAM_CUDA(cudaSetDevice(ctx->device));
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 4096;
    uint32_t *d = nullptr;
    ASLAM_CUDA(cudaMalloc(&d, sizeof(uint32_t) * blocks * threads));
    cudaEvent_t e0, e1;
    ASLAM_CUDA(cudaEventCreate(&e0));
    ASLAM_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        ASLAM_CUDA(cudaEventRecord(e0, ctx->stream));
        switch (which) {
        case 0: pipe_bench_kernel<0><<<blocks, threads, 0, ctx->stream>>>(d, iters, 12345u + rep); break;
        case 1: pipe_bench_kernel<1><<<blocks, threads, 0, ctx->stream>>>(d, iters, 12345u + rep); break;
        case 2: pipe_bench_kernel<2><<<blocks, threads, 0, ctx->stream>>>(d, iters, 12345u + rep); break;
        case 3: pipe_bench_kernel<3><<<blocks, threads, 0, ctx->stream>>>(d, iters, 12345u + rep); break;
        case 4: pipe_bench_kernel<4><<<blocks, threads, 0, ctx->stream>>>(d, iters, 12345u + rep); break;
        case 5: pipe_bench_kernel<5><<<blocks, threads, 0, ctx->stream>>>(d, iters, 12345u + rep); break;
        case 6: pipe_bench_kernel<6><<<blocks, threads, 0, ctx->stream>>>(d, iters, 12345u + rep); break;
        case 7: pipe_bench_kernel<7><<<blocks, threads, 0, ctx->stream>>>(d, iters, 12345u + rep); break;
        case 8: pipe_bench_kernel<8><<<blocks, threads, 0, ctx->stream>>>(d, iters, 12345u + rep); break;
        case 9: pipe_bench_kernel<9><<<blocks, threads, 0, ctx->stream>>>(d, iters, 12345u + rep); break;
        case 10: pipe_bench_kernel<10><<<blocks, threads, 0, ctx->stream>>>(d, iters, 12345u + rep); break;
        default: pipe_bench_kernel<11><<<blocks, threads, 0, ctx->stream>>>(d, iters, 12345u + rep); break;
        }
        ASLAM_LAUNCH_CHECK();
        ASLAM_CUDA(cudaEventRecord(e1, ctx->stream));
        ASLAM_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        ASLAM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    // kinds 9..11 issue 4 packed instructions (8 lane operations) where the others issue 8 instructions
    double ops = (double)blocks * threads * (double)iters * 32.0 * (which == 5 ? 2.0 : 1.0);
    *gops = ops / (best * 1e-3) * 1e-9;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return ASLAM_OK;
}
