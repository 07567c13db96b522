// Exhaustive pin of the rotation factors of computeOrbDescriptor (reference src/ORBextractor.cc:159-160):
//     float angle = (float)kpt.angle * factorPI;  float a = (float)cos(angle), b = (float)sin(angle);
// kpt.angle is the output of cv::fastAtan2, a float in [0, 360]: 1 135 869 953 bit patterns.  For EVERY one of them this program compares
//   (1) glibc cosf / sinf                      — what the reference calls (std::cos(float) through `using namespace std`),
//   (2) (float)cos((double)x), (float)sin(..)  — what the CPU oracle pins (oracle/orb_oracle.cpp:257),
//   (3) the device: sincos((double)x) rounded to float — what k_describe computes (csrc/orb_extract.cu),
// through position-weighted 64-bit checksums per 2^20 angles, and lists every angle on which they differ.
//   nvcc -O2 -std=c++17 --fmad=false -gencode arch=compute_100a,code=sm_100a -Xcompiler -pthread -o /tmp/sincos_ex tools/cpp/sincos_exhaustive.cu
//   gpurun -- '/tmp/sincos_ex'      (test infrastructure: host libm only as the checker)
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

static constexpr uint32_t LAST = 0x43B40000u;            // bit pattern of 360.0f
static constexpr uint32_t CHUNK = 1u << 20;
static constexpr uint64_t MULB = 0x100000001B3ull;

__host__ __device__ inline uint64_t term(uint32_t i, uint32_t abits, uint32_t bbits)
{
    const uint64_t w = 2ull * i + 1ull;
    return (uint64_t)abits * w + (uint64_t)bbits * w * MULB;
}

__device__ inline void dev_ab(uint32_t u, float& a, float& b)
{
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    const float arad = __fmul_rn(__uint_as_float(u), factorPI);
    double sd, cd;
    sincos((double)arad, &sd, &cd);
    a = (float)cd; b = (float)sd;
}

__global__ void k_sum(unsigned long long* sums)
{
    const uint32_t chunk = blockIdx.x;
    unsigned long long acc = 0;
    for (uint32_t j = threadIdx.x; j < CHUNK; j += blockDim.x) {
        const uint64_t u = (uint64_t)chunk * CHUNK + j;
        if (u > LAST) break;
        float a, b;
        dev_ab((uint32_t)u, a, b);
        acc += term((uint32_t)u, __float_as_uint(a), __float_as_uint(b));
    }
    atomicAdd(&sums[chunk], acc);
}

__global__ void k_dump(uint32_t chunk, uint32_t* ab)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= CHUNK) return;
    const uint64_t u = (uint64_t)chunk * CHUNK + j;
    float a = 0, b = 0;
    if (u <= LAST) dev_ab((uint32_t)u, a, b);
    ab[2 * j] = __float_as_uint(a); ab[2 * j + 1] = __float_as_uint(b);
}

static inline uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float bitsf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static void host_ab(uint32_t u, int variant, float& a, float& b)
{
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    volatile float arad = bitsf(u) * factorPI;           // one FP32 rounding, as written in the reference
    const float x = arad;
    if (variant == 0) { a = cosf(x); b = sinf(x); }
    else { a = (float)cos((double)x); b = (float)sin((double)x); }
}

int main()
{
    const uint32_t nchunk = LAST / CHUNK + 1;
    unsigned long long* d_sums;
    cudaMalloc(&d_sums, nchunk * sizeof(unsigned long long));
    cudaMemset(d_sums, 0, nchunk * sizeof(unsigned long long));
    k_sum<<<nchunk, 256>>>(d_sums);
    std::vector<unsigned long long> gsum(nchunk);
    if (cudaMemcpy(gsum.data(), d_sums, nchunk * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) { printf("CUDA error\n"); return 2; }

    std::vector<uint64_t> hsum[2] = { std::vector<uint64_t>(nchunk, 0), std::vector<uint64_t>(nchunk, 0) };
    const unsigned nt = std::max(1u, std::thread::hardware_concurrency());
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++)
        th.emplace_back([&, t] {
            for (uint32_t c = t; c < nchunk; c += nt) {
                uint64_t s0 = 0, s1 = 0;
                for (uint32_t j = 0; j < CHUNK; j++) {
                    const uint64_t u = (uint64_t)c * CHUNK + j;
                    if (u > LAST) break;
                    float a, b;
                    host_ab((uint32_t)u, 0, a, b); s0 += term((uint32_t)u, fbits(a), fbits(b));
                    host_ab((uint32_t)u, 1, a, b); s1 += term((uint32_t)u, fbits(a), fbits(b));
                }
                hsum[0][c] = s0; hsum[1][c] = s1;
            }
        });
    for (auto& x : th) x.join();

    uint32_t* d_ab;
    cudaMalloc(&d_ab, 2 * CHUNK * sizeof(uint32_t));
    std::vector<uint32_t> ab(2 * CHUNK);
    long long bad_chunks = 0, diff_cosf = 0, diff_dbl = 0, diff_host = 0;
    for (uint32_t c = 0; c < nchunk; c++) {
        if (gsum[c] == hsum[0][c] && gsum[c] == hsum[1][c]) continue;
        bad_chunks++;
        k_dump<<<CHUNK / 256, 256>>>(c, d_ab);
        cudaMemcpy(ab.data(), d_ab, ab.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost);
        for (uint32_t j = 0; j < CHUNK; j++) {
            const uint64_t u = (uint64_t)c * CHUNK + j;
            if (u > LAST) break;
            float a0, b0, a1, b1;
            host_ab((uint32_t)u, 0, a0, b0); host_ab((uint32_t)u, 1, a1, b1);
            const bool d0 = fbits(a0) != ab[2 * j] || fbits(b0) != ab[2 * j + 1];
            const bool d1 = fbits(a1) != ab[2 * j] || fbits(b1) != ab[2 * j + 1];
            const bool dh = fbits(a0) != fbits(a1) || fbits(b0) != fbits(b1);
            diff_cosf += d0; diff_dbl += d1; diff_host += dh;
            if ((d0 || d1 || dh) && diff_cosf + diff_dbl + diff_host <= 60)
                printf("angle %.9g (0x%08x): device cos %08x sin %08x | cosf/sinf %08x %08x | via double %08x %08x\n", bitsf((uint32_t)u), (uint32_t)u,
                       ab[2 * j], ab[2 * j + 1], fbits(a0), fbits(b0), fbits(a1), fbits(b1));
        }
    }
    printf("angles checked: %llu (every float in [0, 360]); chunks with a checksum difference: %lld\n", (unsigned long long)LAST + 1, bad_chunks);
    printf("device != glibc cosf/sinf: %lld   device != (float)cos((double)x): %lld   cosf/sinf != via double (host only): %lld\n", diff_cosf, diff_dbl, diff_host);
    return (diff_cosf || diff_dbl) ? 1 : 0;
}
