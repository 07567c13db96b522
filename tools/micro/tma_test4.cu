// u8 2-D TMA tile load, parameters from argv: stride box_w box_h x y
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdlib>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
__global__ void kernel(const __grid_constant__ CUtensorMap tm, int x, int y, uint8_t* out, int bytes)
{
    extern __shared__ __align__(128) uint8_t buf[];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(buf, &tm, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, bytes);
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = buf[i];
}
typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv)
{
    const int stride = atoi(argv[1]), bw = atoi(argv[2]), bh = atoi(argv[3]), x = atoi(argv[4]), y = atoi(argv[5]);
    const int rows = 512;
    std::vector<uint8_t> h((size_t)stride * rows);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 7 + (i >> 9));
    uint8_t *d, *o; cudaMalloc(&d, h.size()); cudaMalloc(&o, bw * bh); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap tm{};
    cuuint64_t dims[2] = { (cuuint64_t)stride, (cuuint64_t)rows };
    cuuint64_t strides[1] = { (cuuint64_t)stride };
    cuuint32_t es[2] = { 1, 1 }, box[2] = { (cuuint32_t)bw, (cuuint32_t)bh };
    CUresult r = ((PFN)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    kernel<<<1, 128, bw * bh>>>(tm, x, y, o, bw * bh);
    cudaError_t e = cudaDeviceSynchronize();
    printf("stride %d box %dx%d at (%d,%d): encode %d run: %s", stride, bw, bh, x, y, (int)r, cudaGetErrorString(e));
    if (e != cudaSuccess) { printf("\n"); return 1; }
    std::vector<uint8_t> res(bh * bw); cudaMemcpy(res.data(), o, bh * bw, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int yy = 0; yy < bh; yy++) for (int xx = 0; xx < bw; xx++) bad += res[yy * bw + xx] != h[(size_t)(y + yy) * stride + x + xx];
    printf("  mismatches %d\n", bad);
    return 0;
}
