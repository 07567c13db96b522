// Pipe micro-benchmark: which issue pipe do VIMNMX3.U16x2 / HMNMX2 / IMAD / PRMT / LOP3 share?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on the GPU box.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#define ITERS 4096
template <int MODE>
__global__ void k(unsigned* out)
{
    unsigned a = threadIdx.x * 2654435761u, b = a ^ 0x12345678u, c = a + 99, d = b * 3, e = a * 7 + 1, f = b + 5;
    __half2 h0 = __halves2half2(__ushort_as_half(0x6400 | (a & 255)), __ushort_as_half(0x6400 | (b & 255)));
    __half2 h1 = __halves2half2(__ushort_as_half(0x6400 | (c & 255)), __ushort_as_half(0x6400 | (d & 255)));
    __half2 h2 = h0, h3 = h1;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (MODE == 0 || MODE == 2 || MODE == 4) { a = __vimin3_u16x2(a, b, c); b = __vimax3_u16x2(b, c, d); c = __vimin3_u16x2(c, d, a); d = __vimax3_u16x2(d, a, b); }
            if (MODE == 1 || MODE == 2) { h0 = __hmin2(h0, h1); h1 = __hmax2(h1, h2); h2 = __hmin2(h2, h3); h3 = __hmax2(h3, h0); }
            if (MODE == 3 || MODE == 4) { e = e * 3 + f; f = f * 5 + e; e = e * 7 + a; f = f * 9 + e; }
            if (MODE == 5) { a = __byte_perm(a, b, 0x4140); b = __byte_perm(b, c, 0x4342); c = __byte_perm(c, d, 0x5140); d = __byte_perm(d, a, 0x4143); }
            if (MODE == 6) { a = __funnelshift_r(a, b, 8); b = __funnelshift_r(b, c, 16); c = __funnelshift_r(c, d, 24); d = __funnelshift_r(d, a, 8); }
            if (MODE == 7) { a = min(a, b); b = max(b, c); c = min(c, d); d = max(d, a); }
        }
    }
    unsigned r = a ^ b ^ c ^ d ^ e ^ f ^ __half_as_ushort(__low2half(h0)) ^ __half_as_ushort(__high2half(h1)) ^ __half_as_ushort(__low2half(h2)) ^ __half_as_ushort(__low2half(h3));
    if (r == 0x31415926u) out[0] = r;
}
template <int MODE>
void run(const char* name, double ops_per_iter)
{
    unsigned* d; cudaMalloc(&d, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(d);
    cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(d); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = 148.0 * 8 * 256 * ITERS * 8 * ops_per_iter;
    printf("%-28s %8.3f ms  %7.1f Ginst/s (thread-level)  = %5.1f lanes/clk/SM @1.965GHz\n", name, ms, inst / ms / 1e6, inst / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(d);
}
int main()
{
    run<0>("VIMNMX3.U16x2", 4); run<1>("HMNMX2", 4); run<2>("VIMNMX3 + HMNMX2", 8); run<3>("IMAD", 4);
    run<4>("VIMNMX3 + IMAD", 8); run<5>("PRMT", 4); run<6>("SHF funnel", 4); run<7>("IMNMX 32-bit", 4);
    return 0;
}
