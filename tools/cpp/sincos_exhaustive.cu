// Exhaustive pin of the rotation factors of computeOrbDescriptor (reference src/ORBextractor.cc:159-160):
//     float angle = (float)kpt.angle * factorPI;  float a = (float)cos(angle), b = (float)sin(angle);
// kpt.angle is the output of cv::fastAtan2, a float in [0, 360]: 1 135 869 953 bit patterns.  For EVERY one of them this program compares
//   (1) the host's glibc cosf / sinf           — what the reference calls (std::cos(float) through `using namespace std`),
//   (2) orbtrig::sincosf_glibc on the device   — what k_describe computes (csrc/orb_trig.h, the restatement of glibc's algorithm), and the
//       same without the FMA contractions,
//   (3) (float)cos((double)x) on the host and sincos((double)x) rounded to float on the device — the "correctly rounded through double"
//       pin the oracle and the kernel used before,
// through position-weighted 64-bit checksums per 2^20 angles, and lists angles on which (1) and (3) differ.
//   nvcc -O2 -std=c++17 --fmad=false -gencode arch=compute_100a,code=sm_100a -Xcompiler -pthread -I orbslam_jpminipc_b200/csrc \
//        -o tools/cpp/build/sincos_ex tools/cpp/sincos_exhaustive.cu
//   gpurun -- 'tools/cpp/build/sincos_ex'      (test infrastructure: host libm only as the checker; exit code 1 if (2) != (1) anywhere)
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#include "orb_trig.h"

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

// glibc >= 2.28 sinf / cosf (sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, sincosf.h: the ARM optimized-routines algorithm): double arithmetic,
// reduce_fast by n = round(x * 2/pi), a degree-7 sine or degree-8 cosine polynomial, rounded to float once.  FMA = the contractions GCC makes in
// the ifunc variant built with -mfma (__sinf_fma / __cosf_fma), which is what an FMA-capable x86-64 host runs.
template <bool FMA>
__device__ inline double g_mad(double a, double b, double c) { return FMA ? fma(a, b, c) : __dadd_rn(__dmul_rn(a, b), c); }
template <bool FMA>
__device__ inline float g_poly(double x, double x2, bool neg_cos, int n)
{
    const double c0 = neg_cos ? -0x1p0 : 0x1p0, c1 = neg_cos ? 0x1.ffffffd0c621cp-2 : -0x1.ffffffd0c621cp-2,
                 c2 = neg_cos ? -0x1.55553e1068f19p-5 : 0x1.55553e1068f19p-5, c3 = neg_cos ? 0x1.6c087e89a359dp-10 : -0x1.6c087e89a359dp-10,
                 c4 = neg_cos ? -0x1.99343027bf8c3p-16 : 0x1.99343027bf8c3p-16;
    const double s0 = -0x1.555545995a603p-3, s1c = 0x1.1107605230bc4p-7, s2 = -0x1.994eb3774cf24p-13;
    if ((n & 1) == 0) {
        const double x3 = __dmul_rn(x, x2), s1 = g_mad<FMA>(x2, s2, s1c), x7 = __dmul_rn(x3, x2), s = g_mad<FMA>(x3, s0, x);
        return (float)g_mad<FMA>(x7, s1, s);
    }
    const double x4 = __dmul_rn(x2, x2), q2 = g_mad<FMA>(x2, c4, c3), q1 = g_mad<FMA>(x2, c1, c0), x6 = __dmul_rn(x4, x2), c = g_mad<FMA>(x4, c2, q1);
    return (float)g_mad<FMA>(x6, q2, c);
}
template <bool FMA>
__device__ inline void glibc_ab(uint32_t u, float& a, float& b)
{
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    const float y = __fmul_rn(__uint_as_float(u), factorPI);
    const uint32_t top = (__float_as_uint(y) >> 20) & 0x7ff;
    double x = (double)y;
    if (top < 0x3f4) {                                    // abstop12 (y) < abstop12 (pio4)
        const double x2 = __dmul_rn(x, x);
        if (top < 0x398) { a = 1.0f; b = y; return; }     // |y| < 2^-12
        b = g_poly<FMA>(x, x2, false, 0); a = g_poly<FMA>(x, x2, false, 1);
        return;
    }
    const double r = __dmul_rn(x, 0x1.45F306DC9C883p+23);
    const int n = (__double2int_rz(r) + 0x800000) >> 24;
    x = FMA ? fma(-(double)n, 0x1.921FB54442D18p0, x) : __dsub_rn(x, __dmul_rn((double)n, 0x1.921FB54442D18p0));
    const double sg = ((n + 1) & 2) ? -1.0 : 1.0;         // sign[n & 3] = { 1, -1, -1, 1 }
    const bool neg = (n & 2) != 0;
    const double xs = __dmul_rn(x, sg), x2 = __dmul_rn(x, x);
    b = g_poly<FMA>(xs, x2, neg, n); a = g_poly<FMA>(xs, x2, neg, n ^ 1);
}

template <int V>
__global__ void k_sum_glibc(unsigned long long* sums)
{
    const uint32_t chunk = blockIdx.x;
    unsigned long long acc = 0;
    for (uint32_t j = threadIdx.x; j < CHUNK; j += blockDim.x) {
        const uint64_t u = (uint64_t)chunk * CHUNK + j;
        if (u > LAST) break;
        float a, b;
        if (V == 2) {
            const float factorPI = (float)(3.14159265358979323846 / 180.f);
            orbtrig::sincosf_glibc(__fmul_rn(__uint_as_float((uint32_t)u), factorPI), b, a);
        } else glibc_ab<V == 1>((uint32_t)u, a, b);
        acc += term((uint32_t)u, __float_as_uint(a), __float_as_uint(b));
    }
    atomicAdd(&sums[chunk], acc);
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
                printf("angle %.9g (0x%08x): device sincos through double: cos %08x sin %08x | cosf/sinf %08x %08x | via double %08x %08x\n", bitsf((uint32_t)u), (uint32_t)u,
                       ab[2 * j], ab[2 * j + 1], fbits(a0), fbits(b0), fbits(a1), fbits(b1));
        }
    }
    uint32_t kernel_form_bad = 0;
    for (int v = 0; v < 3; v++) {                       // the glibc algorithm on the device, without / with the FMA contractions, and as k_describe includes it, against the host's cosf / sinf
        cudaMemset(d_sums, 0, nchunk * sizeof(unsigned long long));
        if (v == 2) k_sum_glibc<2><<<nchunk, 256>>>(d_sums); else if (v) k_sum_glibc<1><<<nchunk, 256>>>(d_sums); else k_sum_glibc<0><<<nchunk, 256>>>(d_sums);
        std::vector<unsigned long long> g2(nchunk);
        cudaMemcpy(g2.data(), d_sums, nchunk * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        uint32_t nb = 0;
        for (uint32_t c = 0; c < nchunk; c++) nb += g2[c] != hsum[0][c];
        printf("device restatement of glibc sinf / cosf (%s): chunks of 2^20 angles whose checksum differs from the host's cosf / sinf: %u of %u\n",
               v == 2 ? "orbtrig::sincosf_glibc, the function k_describe calls" : v ? "FMA contractions of the -mfma ifunc variant" : "no contraction", nb, nchunk);
        if (v == 2) kernel_form_bad = nb;
    }
    printf("angles checked: %llu (every float in [0, 360]); chunks with a checksum difference: %lld\n", (unsigned long long)LAST + 1, bad_chunks);
    printf("device sincos through double != glibc cosf/sinf: %lld   device sincos through double != (float)cos((double)x): %lld   cosf/sinf != via double (host only): %lld\n", diff_cosf, diff_dbl, diff_host);
    return kernel_form_bad ? 1 : 0;
}
