// Shared host-side plumbing of libaslam_b200: context, error reporting, launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/aslam_b200.h"

struct aslam_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    int sm_count = 148;
    // scratch for the host-buffer matching entry points (grown on demand)
    void *d_scratch = nullptr;
    size_t scratch_bytes = 0;
    void *h_scratch = nullptr;     // pinned
    size_t h_scratch_bytes = 0;
    // optional per-kernel timing (aslam_prof_*): CUDA event pairs around every launch
    bool prof_on = false;
    struct ProfRec { int kind; cudaEvent_t e0, e1; };
    std::vector<ProfRec> prof_pending;
    std::vector<cudaEvent_t> prof_pool;
    int resize_smem_set = 0;       // resize_smem_kernel's dynamic shared memory opt-in made on this device (bytes)
    bool knn_pad_set = false;
    bool desc_attr_set = false;    // describe_kernel's dynamic shared memory opt-in was made on this device
    double prof_ms[16] = {0};
    long long prof_n[16] = {0};
};

namespace aslam {

extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;

int fail(int code, const char *fmt, ...);

inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int ensure_scratch(aslam_ctx *ctx, size_t dev_bytes, size_t host_bytes);

enum ProfKind { PK_RESIZE = 0, PK_FAST, PK_BLUR, PK_SELECT, PK_DESCRIBE, PK_KNN_PARTIAL, PK_KNN_FINISH,
                PK_MERGE, PK_MISC, PK_COUNT };

// Brackets one kernel launch with an event pair when profiling is enabled on the context.
struct ProfScope {
    aslam_ctx *c;
    int kind;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ProfScope(aslam_ctx *ctx, int k);
    ~ProfScope();
};

#ifdef __CUDACC__
// Debug build (-DASLAM_DEBUG_BOUNDS=1, libaslam_b200_dbg.so): every staged index -- shared-memory tiles, queues, candidate
// slots, patch bytes, source and destination rows -- is checked and violations are counted (compute-sanitizer is closed
// on the GPU pool this was developed on, so the parity tests over ragged sizes run once against this build as well).
// Release builds compile the checks away.
#if ASLAM_DEBUG_BOUNDS
// (no relocatable device code in this build: every translation unit counts into its own copy and registers a reader)
static __device__ unsigned long long g_bounds_violations;
void register_bounds_reader(int (*reader)(unsigned long long *));
static int tu_bounds_read(unsigned long long *v)
{
    return cudaMemcpyFromSymbol(v, g_bounds_violations, sizeof(*v)) == cudaSuccess ? 0 : 1;
}
[[maybe_unused]] static const int tu_bounds_registered = (register_bounds_reader(tu_bounds_read), 0);
#define ASLAM_BOUNDS(cond) do { if (!(cond)) atomicAdd(&aslam::g_bounds_violations, 1ull); } while (0)
#else
#define ASLAM_BOUNDS(cond) do { } while (0)
#endif
// DYNAMIC_CLASSES of the reference (src/main.cpp:29-41): person, bicycle, car, motorcycle, bus, train, truck, bird, cat, dog
__device__ __forceinline__ bool is_dynamic_class(int c)
{
    return c == 0 || c == 1 || c == 2 || c == 3 || c == 5 || c == 6 || c == 7 || c == 14 || c == 15 || c == 16;
}

// isInDynamicObject (src/main.cpp:43-50).  half_open: cv::Rect::contains (x1 <= x < x2); else Detection::contains (closed).
__device__ __forceinline__ bool in_dynamic_box(float x, float y, const aslam_box *__restrict__ boxes, int nboxes, int half_open)
{
    for (int j = 0; j < nboxes; ++j) {
        const aslam_box bx = boxes[j];
        if (!is_dynamic_class(bx.class_id) || x < bx.x1 || y < bx.y1) continue;
        if (half_open ? (x < bx.x2 && y < bx.y2) : (x <= bx.x2 && y <= bx.y2)) return true;
    }
    return false;
}
#endif

}  // namespace aslam

#define ASLAM_CUDA(expr)                                                                      \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return aslam::fail(ASLAM_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,    \
                               cudaGetErrorString(e__));                                      \
    } while (0)

#define ASLAM_LAUNCH_CHECK()                                                                  \
    do {                                                                                      \
        aslam::count_launch();                                                                \
        cudaError_t e__ = cudaGetLastError();                                                 \
        if (e__ != cudaSuccess)                                                               \
            return aslam::fail(ASLAM_E_CUDA, "%s:%d kernel launch -> %s", __FILE__, __LINE__, \
                               cudaGetErrorString(e__));                                      \
    } while (0)

#define ASLAM_REQUIRE(cond, msg)                                                              \
    do {                                                                                      \
        if (!(cond)) return aslam::fail(ASLAM_E_INVALID, "%s: %s", __func__, msg);            \
    } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
