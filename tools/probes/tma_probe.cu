// Bisects what the B200 accepts for mbarrier / bulk / TMA box loads (tools only; not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu && for m in 0 1 2 3 4 5 6; do ./tma_probe $m; done
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait0(uint32_t b)
{
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(b), "r"(0) : "memory");
}

// mode 0: mbarrier only (init, plain arrive, wait)
// mode 1: 1-D bulk copy global -> shared (no tensor map)
// mode 2: 2-D tensor map, param;  mode 3: 3-D tensor map, param;  mode 4: 3-D, map in global memory
// mode 5: 3-D param, 4 lanes issue their own box;  mode 6: 3-D param without fence.mbarrier_init
__global__ void probe(const __grid_constant__ CUtensorMap map2, const __grid_constant__ CUtensorMap map3, const void *gmap,
                      const uint8_t *img, uint8_t *out, int x, int y, int z, int mode)
{
    __shared__ __align__(128) uint8_t buf[4][1024];
    __shared__ __align__(8) unsigned long long bar;
    const int lane = threadIdx.x;
    const uint32_t b = smem_u32(&bar);
    const int n = mode == 5 ? 4 : 1;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(n) : "memory");
        if (mode != 6) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (mode == 0) {
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory");
        wait0(b);
        if (lane == 0) out[0] = 77;
        return;
    }
    if (lane < n) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(992) : "memory");
        if (mode == 1)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(buf[0])), "l"(img), "r"(992), "r"(b) : "memory");
        else if (mode == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(buf[0])), "l"(&map2), "r"(x), "r"(y), "r"(b) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_u32(buf[lane])), "l"(mode == 4 ? gmap : (const void *)&map3), "r"(x + lane), "r"(y), "r"(z), "r"(b) : "memory");
    }
    wait0(b);
    for (int k = 0; k < n; ++k)
        for (int i = lane; i < 992; i += 32) out[k * 992 + i] = buf[k][i];
}

int main(int argc, char **argv)
{
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int H = 480, B = 3, pitch = 752, stride = 752 * 480 + 256;
    std::vector<uint8_t> img((size_t)B * stride);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)(i * 2654435761u >> 24);
    uint8_t *d_img, *d_out; void *d_map;
    cudaMalloc(&d_img, img.size()); cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&d_out, 4 * 992); cudaMalloc(&d_map, 128);
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto encode = (PFN_cuTensorMapEncodeTiled)fn;
    alignas(64) CUtensorMap m2, m3;
    cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)H, (cuuint64_t)B}, strides[2] = {(cuuint64_t)pitch, (cuuint64_t)stride};
    cuuint32_t box[3] = {32, 31, 1}, es[3] = {1, 1, 1};
    CUresult r2 = encode(&m2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d_img, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r3 = encode(&m3, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d_img, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cudaMemcpy(d_map, &m3, 128, cudaMemcpyHostToDevice);
    const int x = argc > 2 ? atoi(argv[2]) : 101, y = 57, z = mode == 2 ? 0 : 2;
    cudaMemset(d_out, 0, 4 * 992);
    probe<<<1, 32>>>(m2, m3, d_map, d_img, d_out, x, y, z, mode);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<uint8_t> out(4 * 992);
    cudaMemcpy(out.data(), d_out, out.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    const int n = mode == 5 ? 4 : 1;
    if (mode == 0) bad = out[0] != 77;
    else if (mode == 1) { for (int i = 0; i < 992; ++i) bad += out[i] != img[i]; }
    else
        for (int k = 0; k < n; ++k)
            for (int yy = 0; yy < 31; ++yy)
                for (int xx = 0; xx < 32; ++xx)
                    if (out[k * 992 + yy * 32 + xx] != img[(size_t)z * stride + (size_t)(y + yy) * pitch + x + k + xx]) ++bad;
    printf("x %d mode %d: encode %d/%d, %s, %d mismatches\n", x, mode, (int)r2, (int)r3, cudaGetErrorString(e), bad);
    return e != cudaSuccess;
}
