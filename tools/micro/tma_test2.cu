// canonical 2-D TMA example (CUDA programming guide) through libcu++ wrappers
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int GW = 1024, GH = 1024, SW = 32, SH = 32;
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int* out)
{
    __shared__ alignas(128) int smem_buffer[SH][SW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < SH * SW; i += blockDim.x) out[i] = smem_buffer[i / SW][i % SW];
}
typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main()
{
    std::vector<int> h(GW * GH);
    for (size_t i = 0; i < h.size(); i++) h[i] = (int)i;
    int *d, *o; cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, SH * SW * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap tm{};
    cuuint64_t size[2] = { GW, GH }; cuuint64_t stride[1] = { GW * sizeof(int) };
    cuuint32_t box[2] = { SW, SH }, es[2] = { 1, 1 };
    CUresult r = ((PFN)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)r);
    kernel<<<1, 128>>>(tm, 64, 32, o);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run: %s\n", cudaGetErrorString(e));
    if (e == cudaSuccess) { std::vector<int> res(SH * SW); cudaMemcpy(res.data(), o, SH * SW * 4, cudaMemcpyDeviceToHost); printf("first %d expect %d\n", res[0], 32 * GW + 64); }
    return 0;
}
