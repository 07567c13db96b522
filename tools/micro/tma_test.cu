// Stand-alone check of the 3-D u8 tiled TMA load used by k_fast_nms / k_blur.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
struct TmapSet { CUtensorMap m[16]; };
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ TmapSet tm, int level, int x, int y, int z, uint8_t* out)
{
    __shared__ __align__(128) uint8_t buf[40 * 80];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&bar)), "r"(3200) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_addr(buf)), "l"(&tm.m[level]), "r"(x), "r"(y), "r"(z), "r"(smem_addr(&bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DN;\n\tbra WL;\n\tDN:\n\t}" ::"r"(smem_addr(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < 3200; i += blockDim.x) out[i] = buf[i];
}
__global__ void k1(const __grid_constant__ CUtensorMap tm, int x, int y, int z, uint8_t* out, int bytes)
{
    __shared__ __align__(128) uint8_t buf[40 * 128];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_addr(buf)), "l"(&tm), "r"(x), "r"(y), "r"(z), "r"(smem_addr(&bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DN;\n\tbra WL;\n\tDN:\n\t}" ::"r"(smem_addr(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = buf[i];
}
typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main()
{
    const int stride = 784, rows = 512, frames = 3; const size_t fb = 1 << 20;
    std::vector<uint8_t> h(fb * frames);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 7 + (i >> 9));
    uint8_t *d, *o; cudaMalloc(&d, h.size()); cudaMalloc(&o, 3200); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry point: %d %d %p\n", (int)e, (int)q, p);
    TmapSet tm; memset(&tm, 0, sizeof tm);
    cuuint64_t dims[3] = { (cuuint64_t)stride, (cuuint64_t)rows, (cuuint64_t)frames };
    cuuint64_t strides[2] = { (cuuint64_t)stride, (cuuint64_t)fb };
    cuuint32_t es[3] = { 1, 1, 1 }, box[3] = { 80, 40, 1 };
    for (int l = 0; l < 2; l++) {
        CUresult r = ((PFN)p)(&tm.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d + 256 * l, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode level %d -> %d\n", l, (int)r);
    }
    for (int inner : {64, 80, 128}) {
        CUtensorMap one; cuuint32_t bx[3] = { (cuuint32_t)inner, 40, 1 };
        CUresult r = ((PFN)p)(&one, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        uint8_t* o2; cudaMalloc(&o2, 40 * 128);
        k1<<<1, 128>>>(one, 24, 12, 1, o2, inner * 40);
        e = cudaDeviceSynchronize();
        printf("single map inner=%d encode=%d run: %s\n", inner, (int)r, cudaGetErrorString(e));
        if (e != cudaSuccess) { cudaDeviceReset(); return 2; }
    }
    for (int t = 0; t < 3; t++) {
        int level = t & 1, x = 24 + 64 * t, y = 12 + 30 * t, z = t;
        k<<<1, 128>>>(tm, level, x, y, z, o);
        e = cudaDeviceSynchronize();
        printf("run %d: %s\n", t, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<uint8_t> r(3200); cudaMemcpy(r.data(), o, 3200, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int yy = 0; yy < 40; yy++) for (int xx = 0; xx < 80; xx++) {
            uint8_t ref = h[256 * level + (size_t)z * fb + (size_t)(y + yy) * stride + x + xx];
            if (r[yy * 80 + xx] != ref) bad++;
        }
        printf("mismatches %d\n", bad);
    }
    return 0;
}
