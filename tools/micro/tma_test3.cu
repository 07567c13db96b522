// variants of the TMA tile load: argv[1] = 0: u8 2-D libcu++, 1: u8 3-D libcu++, 2: u8 3-D raw PTX (as k_fast_nms)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdlib>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int BW = 80, BH = 40;
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int MODE>
__global__ void kernel(const __grid_constant__ CUtensorMap tm, int x, int y, int z, uint8_t* out)
{
    __shared__ alignas(128) uint8_t buf[BH * BW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        if (MODE == 0) cde::cp_async_bulk_tensor_2d_global_to_shared(buf, &tm, x, y, bar);
        if (MODE == 1) cde::cp_async_bulk_tensor_3d_global_to_shared(buf, &tm, x, y, z, bar);
        if (MODE == 2) {
            uint64_t* nb = cuda::device::barrier_native_handle(bar);
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_addr(buf)), "l"(&tm), "r"(x), "r"(y), "r"(z), "r"(smem_addr(nb)) : "memory");
        }
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(buf));
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < BH * BW; i += blockDim.x) out[i] = buf[i];
}
typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv)
{
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int stride = 784, rows = 512, frames = 3; const size_t fb = 1 << 20;
    std::vector<uint8_t> h(fb * frames);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 7 + (i >> 9));
    uint8_t *d, *o; cudaMalloc(&d, h.size()); cudaMalloc(&o, BH * BW); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap tm{};
    cuuint64_t dims[3] = { (cuuint64_t)stride, (cuuint64_t)rows, (cuuint64_t)frames };
    cuuint64_t strides[2] = { (cuuint64_t)stride, (cuuint64_t)fb };
    cuuint32_t es[3] = { 1, 1, 1 }, box[3] = { BW, BH, 1 };
    CUresult r = ((PFN)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, mode == 0 ? 2 : 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int x = 24, y = 12, z = mode == 0 ? 0 : 2;
    if (mode == 0) kernel<0><<<1, 128>>>(tm, x, y, z, o);
    if (mode == 1) kernel<1><<<1, 128>>>(tm, x, y, z, o);
    if (mode == 2) kernel<2><<<1, 128>>>(tm, x, y, z, o);
    cudaError_t e = cudaDeviceSynchronize();
    printf("mode %d encode %d run: %s\n", mode, (int)r, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<uint8_t> res(BH * BW); cudaMemcpy(res.data(), o, BH * BW, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int yy = 0; yy < BH; yy++) for (int xx = 0; xx < BW; xx++) bad += res[yy * BW + xx] != h[(size_t)z * fb + (size_t)(y + yy) * stride + x + xx];
    printf("mismatches %d\n", bad);
    return 0;
}
