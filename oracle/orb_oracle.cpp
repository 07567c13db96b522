/*
 * orb_oracle.cpp — CPU ORACLE for the ORB front end.  TEST INFRASTRUCTURE ONLY.
 *
 * A from-scratch restatement of the reference's algorithm (caomw/ORBSLAM_jpMiniPC,
 * ORB-SLAM v1) for the path the B200 product accelerates.  Every function cites the
 * reference file:line it follows.  The OpenCV primitives the reference calls
 * (resize, copyMakeBorder, FAST, KeyPointsFilter::retainBest, GaussianBlur,
 * fastAtan2, cvRound, gemm) live in an un-vendored third party (OpenCV, unpinned
 * `find_package(OpenCV REQUIRED)`, reference CMakeLists.txt:19); they are restated
 * here to the semantics of OpenCV 4.13.0 as observable through python cv2 in the
 * build container and pinned by tests/test_oracle_vs_cv2.py + tests/golden/.
 *
 * Build: g++ -O3 -march=x86-64-v3 -ffp-contract=off (no FMA contraction: every fused
 * multiply-add that is part of the spec is written as an explicit fmaf()).
 * std::nth_element of this toolchain's libstdc++ (GCC 13) is part of the spec
 * (it defines which tied keypoints survive retainBest and in what order).
 */
#include "orb_oracle.h"

#include <algorithm>
#include <climits>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>
#include <set>

namespace {

const int PATCH_SIZE = 31;       /* src/ORBextractor.cc:75 */
const int HALF_PATCH_SIZE = 15;  /* :76 */
const int EDGE_THRESHOLD = 16;   /* :77 */

const int8_t kPattern[1024] = {
#include "orb_pattern.inc"
};

/* cvRound = round-half-even (SSE cvtsd2si under the default rounding mode) */
inline int cv_round(double v) { return (int)lrint(v); }
inline int cv_roundf(float v) { return (int)lrintf(v); }
inline int cv_floor(double v) { int i = (int)v; return i - (i > v); }
inline int cv_ceil(double v) { int i = (int)v; return i + (i < v); }
inline uint8_t sat_u8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

struct KP { float x, y, size, angle, response; int octave, class_id; };
struct RespGreater { bool operator()(const KP& a, const KP& b) const { return a.response > b.response; } };

/* ---------------------------------------------------------------- resize ---
 * cv::resize(src,dst,sz,0,0,INTER_LINEAR) for CV_8UC1 (reference call site
 * src/ORBextractor.cc:800).  OpenCV: inv_scale = (double)dsize/ssize,
 * scale = 1./inv_scale; per destination column fx=(float)((dx+0.5)*scale-0.5),
 * sx=floor(fx), fx-=sx, clamped at both ends; 11-bit fixed-point coefficients
 * (saturate_cast<short>(f*2048) = round-half-even); horizontal pass in int32,
 * vertical pass (((b0*(S0>>4))>>16)+((b1*(S1>>4))>>16)+2)>>2.  Rows are clipped
 * (not re-weighted) in the vertical direction. */
void resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride,
                      uint8_t* dst, int dw, int dh, int dstride)
{
    const double inv_sx = (double)dw / sw, inv_sy = (double)dh / sh;
    const double scale_x = 1. / inv_sx, scale_y = 1. / inv_sy;
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> alpha(2 * dw), beta(2 * dh);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cv_floor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        alpha[2 * dx] = (short)cv_roundf((1.f - fx) * 2048);
        alpha[2 * dx + 1] = (short)cv_roundf(fx * 2048);
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cv_floor(fy);
        fy -= sy;
        yofs[dy] = sy;
        beta[2 * dy] = (short)cv_roundf((1.f - fy) * 2048);
        beta[2 * dy + 1] = (short)cv_roundf(fy * 2048);
    }
    std::vector<int> r0(dw), r1(dw);
    for (int dy = 0; dy < dh; dy++) {
        int sy0 = std::min(std::max(yofs[dy], 0), sh - 1);
        int sy1 = std::min(std::max(yofs[dy] + 1, 0), sh - 1);
        const uint8_t* S0 = src + (size_t)sy0 * sstride;
        const uint8_t* S1 = src + (size_t)sy1 * sstride;
        for (int dx = 0; dx < dw; dx++) {
            int sx = xofs[dx];
            int sx1 = std::min(sx + 1, sw - 1);
            int a0 = alpha[2 * dx], a1 = alpha[2 * dx + 1];
            r0[dx] = S0[sx] * a0 + S0[sx1] * a1;
            r1[dx] = S1[sx] * a0 + S1[sx1] * a1;
        }
        int b0 = beta[2 * dy], b1 = beta[2 * dy + 1];
        uint8_t* D = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; dx++)
            D[dx] = sat_u8((((b0 * (r0[dx] >> 4)) >> 16) + ((b1 * (r1[dx] >> 4)) >> 16) + 2) >> 2);
    }
}

/* copyMakeBorder(..., BORDER_REFLECT_101) in place on a padded plane whose
 * interior is already filled (reference src/ORBextractor.cc:806,814). */
inline int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) { if (p < 0) p = -p; else p = 2 * len - 2 - p; }
    return p;
}
void border_reflect101(uint8_t* plane, int w, int h, int stride, int b)
{
    for (int y = 0; y < h; y++) {
        uint8_t* row = plane + (size_t)(y + b) * stride + b;
        for (int x = -b; x < 0; x++) row[x] = row[reflect101(x, w)];
        for (int x = w; x < w + b; x++) row[x] = row[reflect101(x, w)];
    }
    for (int y = -b; y < h + b; y++) {
        if (y >= 0 && y < h) continue;
        int sy = reflect101(y, h);
        memcpy(plane + (size_t)(y + b) * stride, plane + (size_t)(sy + b) * stride, w + 2 * b);
    }
}

/* ------------------------------------------------------------------ FAST ---
 * cv::FAST(img, kps, th, nonmaxSuppression=true), TYPE_9_16 (reference call
 * sites src/ORBextractor.cc:607,613).  Ring order = OpenCV's 16-pixel Bresenham
 * circle.  corner <=> some 9 contiguous ring pixels all darker than p-th or all
 * brighter than p+th.  Score (OpenCV cornerScore<16>) = max over the 16 arcs of
 * min(|diff| over the arc, signed) - 1, independent of th for corners.  NMS:
 * keep iff score > all 8 neighbours' scores (non-corners and the 3-px margin
 * count 0).  Output in raster order, coordinates local to img. */
const int kRing[16][2] = { {0,3},{1,3},{2,2},{3,1},{3,0},{3,-1},{2,-2},{1,-3},
                           {0,-3},{-1,-3},{-2,-2},{-3,-1},{-3,0},{-3,1},{-2,2},{-1,3} };

/* One image row of corner strengths, written so that the compiler vectorises over x (uint8 min/max):
 *   bright = max_k min(ring[k..k+8]) - v,  dark = v - min_k max(ring[k..k+8]),  strength = max(bright, dark)
 * which equals OpenCV's cornerScore<16> + 1 for corners (strength > th  <=>  corner at threshold th). */
static void fast_strength_row(const uint8_t* row, int stride, int x0, int x1, int th, uint8_t* out /* [x0,x1) */)
{
    enum { MAXW = 4096 };
    alignas(64) uint8_t a[16][MAXW], bmin[MAXW], bmax[MAXW];
    const int n = x1 - x0;
    const uint8_t* r[16];
    for (int k = 0; k < 16; k++) r[k] = row + kRing[k][1] * stride + kRing[k][0] + x0;
    for (int k = 0; k < 16; k++) {
        const uint8_t* __restrict__ p0 = r[k]; const uint8_t* __restrict__ p1 = r[(k + 1) & 15]; const uint8_t* __restrict__ p2 = r[(k + 2) & 15];
        uint8_t* __restrict__ o = a[k];
        for (int x = 0; x < n; x++) { uint8_t m = p0[x] < p1[x] ? p0[x] : p1[x]; o[x] = m < p2[x] ? m : p2[x]; }
    }
    for (int x = 0; x < n; x++) bmin[x] = 0;
    for (int k = 0; k < 16; k++) {
        const uint8_t* __restrict__ p0 = a[k]; const uint8_t* __restrict__ p1 = a[(k + 3) & 15]; const uint8_t* __restrict__ p2 = a[(k + 6) & 15];
        uint8_t* __restrict__ o = bmin;
        for (int x = 0; x < n; x++) { uint8_t m = p0[x] < p1[x] ? p0[x] : p1[x]; m = m < p2[x] ? m : p2[x]; o[x] = o[x] > m ? o[x] : m; }
    }
    for (int k = 0; k < 16; k++) {
        const uint8_t* __restrict__ p0 = r[k]; const uint8_t* __restrict__ p1 = r[(k + 1) & 15]; const uint8_t* __restrict__ p2 = r[(k + 2) & 15];
        uint8_t* __restrict__ o = a[k];
        for (int x = 0; x < n; x++) { uint8_t m = p0[x] > p1[x] ? p0[x] : p1[x]; o[x] = m > p2[x] ? m : p2[x]; }
    }
    for (int x = 0; x < n; x++) bmax[x] = 255;
    for (int k = 0; k < 16; k++) {
        const uint8_t* __restrict__ p0 = a[k]; const uint8_t* __restrict__ p1 = a[(k + 3) & 15]; const uint8_t* __restrict__ p2 = a[(k + 6) & 15];
        uint8_t* __restrict__ o = bmax;
        for (int x = 0; x < n; x++) { uint8_t m = p0[x] > p1[x] ? p0[x] : p1[x]; m = m > p2[x] ? m : p2[x]; o[x] = o[x] < m ? o[x] : m; }
    }
    const uint8_t* v = row + x0;
    for (int x = 0; x < n; x++) {
        const int s = std::max((int)bmin[x] - (int)v[x], (int)v[x] - (int)bmax[x]);
        out[x] = (uint8_t)(s > th ? s - 1 : 0);
    }
}

int fast9_nms(const uint8_t* img, int w, int h, int stride, int th, std::vector<KP>& out)
{
    out.clear();
    if (w < 7 || h < 7 || w > 4096) return 0;
    std::vector<uint8_t> sc((size_t)w * h, 0);
    for (int y = 3; y < h - 3; y++)
        fast_strength_row(img + (size_t)y * stride, stride, 3, w - 3, th, &sc[(size_t)y * w + 3]);
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            int s = sc[(size_t)y * w + x];
            if (!s) continue;
            const uint8_t* r = &sc[(size_t)y * w + x];
            if (s > r[-1] && s > r[1] && s > r[-w - 1] && s > r[-w] && s > r[-w + 1] &&
                s > r[w - 1] && s > r[w] && s > r[w + 1]) {
                KP k; k.x = (float)x; k.y = (float)y; k.size = 7.f; k.angle = -1.f;
                k.response = (float)s; k.octave = 0; k.class_id = -1;
                out.push_back(k);
            }
        }
    return (int)out.size();
}

/* KeyPointsFilter::retainBest(v, n) of OpenCV 4.x followed by the reference's
 * `if(size>n) resize(n)` (src/ORBextractor.cc:683-685, :699-700): the survivors are
 * the first n elements after std::nth_element(begin, begin+n-1, end, response>). */
void retain_best(std::vector<KP>& v, int n)
{
    if (n >= 0 && v.size() > (size_t)n) {
        if (n == 0) { v.clear(); return; }
        std::nth_element(v.begin(), v.begin() + n - 1, v.end(), RespGreater());
        float amb = v[n - 1].response;
        std::vector<KP>::iterator new_end =
            std::partition(v.begin() + n, v.end(), [amb](const KP& k) { return k.response >= amb; });
        v.resize(new_end - v.begin());
    }
    if ((int)v.size() > n) v.resize(n);
}

/* cv::fastAtan2(y,x) (degrees, OpenCV 4.x scalar atan_f32), called from IC_Angle
 * src/ORBextractor.cc:150.  Every operation rounded to FP32, no FMA. */
float fast_atan2(float y, float x)
{
    const float scale = (float)(180 / M_PI);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* IC_Angle, src/ORBextractor.cc:124-151 */
float ic_angle(const uint8_t* center, int step, const int* umax)
{
    int m_01 = 0, m_10 = 0;
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
        int v_sum = 0, d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int val_plus = center[u + v * step], val_minus = center[u - v * step];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return fast_atan2((float)m_01, (float)m_10);
}

/* cosf / sinf as the reference calls them at src/ORBextractor.cc:160 (`cos(angle)` on a float with `using namespace std`).  The algorithm
 * is glibc's (libm >= 2.28, the image has 2.39; not part of /root/reference): sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, sincosf.h — the
 * argument in double, n = round(x * 2/pi) by a scaled truncation, x - n * pi/2, a degree-7 sine or degree-8 cosine polynomial, one rounding
 * to float.  Restated for 0 <= y < 120 in plain double operations; orc_trig_mismatches() compares it with the host's own cosf / sinf
 * (tests/test_oracle_golden.py, strided over every binade of [0, 360] degrees; tools/cpp/sincos_exhaustive.cu does all 1 135 869 953 angles). */
static float trig_poly(double x, double x2, bool neg_cos, int n)
{
    if ((n & 1) == 0) {
        const double s0 = -0x1.555545995a603p-3, s1 = 0x1.1107605230bc4p-7, s2 = -0x1.994eb3774cf24p-13;
        double x3 = x * x2, t1 = s1 + x2 * s2, x7 = x3 * x2, s = x + x3 * s0;
        return (float)(s + x7 * t1);
    }
    const double sg = neg_cos ? -1.0 : 1.0;
    const double c0 = sg, c1 = sg * -0x1.ffffffd0c621cp-2, c2 = sg * 0x1.55553e1068f19p-5, c3 = sg * -0x1.6c087e89a359dp-10, c4 = sg * 0x1.99343027bf8c3p-16;
    double x4 = x2 * x2, q2 = c3 + x2 * c4, q1 = c0 + x2 * c1, x6 = x4 * x2, c = q1 + x4 * c2;
    return (float)(c + x6 * q2);
}
static void ref_sincosf(float y, float* sn, float* cs)
{
    uint32_t bits; memcpy(&bits, &y, 4);
    const uint32_t top = (bits >> 20) & 0x7ff;
    double x = (double)y;
    if (top < 0x3f4) {
        if (top < 0x398) { *cs = 1.0f; *sn = y; return; }
        double x2 = x * x;
        *sn = trig_poly(x, x2, false, 0); *cs = trig_poly(x, x2, false, 1);
        return;
    }
    double r = x * 0x1.45F306DC9C883p+23;
    int n = ((int32_t)r + 0x800000) >> 24;
    x = x - n * 0x1.921FB54442D18p0;
    const bool neg = (n & 2) != 0;
    double xs = ((n + 1) & 2) ? -x : x, x2 = x * x;
    *sn = trig_poly(xs, x2, neg, n); *cs = trig_poly(xs, x2, neg, n ^ 1);
}

/* computeOrbDescriptor, src/ORBextractor.cc:155-194. */
void rbrief(const uint8_t* center, int step, float angle_deg, uint8_t* desc, bool fma_form = false)
{
    const float factorPI = (float)(M_PI / 180.f);        /* :154 */
    float angle = angle_deg * factorPI;                   /* :159 */
    float a, b;
    ref_sincosf(angle, &b, &a);                            /* :160 */
    const int8_t* pat = kPattern;
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int k = 0; k < 8; k++) {
            float x0 = pat[4 * k], y0 = pat[4 * k + 1], x1 = pat[4 * k + 2], y1 = pat[4 * k + 3];
            int t0, t1;
            if (!fma_form) {
                t0 = center[cv_roundf(x0 * b + y0 * a) * step + cv_roundf(x0 * a - y0 * b)];
                t1 = center[cv_roundf(x1 * b + y1 * a) * step + cv_roundf(x1 * a - y1 * b)];
            } else {     /* GCC's contraction under the reference's own -O3 -march=native: vfmadd231ss / vfmsub132ss */
                t0 = center[cv_roundf(fmaf(x0, b, y0 * a)) * step + cv_roundf(fmaf(x0, a, -(y0 * b)))];
                t1 = center[cv_roundf(fmaf(x1, b, y1 * a)) * step + cv_roundf(fmaf(x1, a, -(y1 * b)))];
            }
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

/* GaussianBlur(m, m, Size(7,7), 2, 2, BORDER_REFLECT_101), src/ORBextractor.cc:760,
 * applied to the level ROI of a padded plane: reads the real border (== reflect-101),
 * writes only the ROI.  src points at the ROI origin inside the padded plane. */
const uint32_t kGaussBits[7] = { 0x3d8fafb1u, 0x3e06387eu, 0x3e434a39u, 0x3e5d4ae0u,
                                 0x3e434a39u, 0x3e06387eu, 0x3d8fafb1u };
void gaussian_blur7(const uint8_t* src, int w, int h, int stride, uint8_t* dst, int dstride, int variant)
{
    if (variant == ORC_BLUR_F32_SEPFILTER) {
        float k[7];
        memcpy(k, kGaussBits, sizeof k);
        /* loops run over x innermost so the compiler vectorises them; every pixel still sees the taps in
         * the order i = 0..6 (row pass) and centre, +-1, +-2, +-3 (column pass), each step one fused multiply-add */
        std::vector<float> R((size_t)(h + 6) * w);
        for (int y = -3; y < h + 3; y++) {
            const uint8_t* __restrict__ S = src + (ptrdiff_t)y * stride;
            float* __restrict__ r = &R[(size_t)(y + 3) * w];
            for (int x = 0; x < w; x++) r[x] = (float)S[x - 3] * k[0];
            for (int i = 1; i < 7; i++) {
                const float ki = k[i];
                for (int x = 0; x < w; x++) r[x] = fmaf((float)S[x + i - 3], ki, r[x]);
            }
        }
        std::vector<float> acc(w);
        for (int y = 0; y < h; y++) {
            uint8_t* __restrict__ D = dst + (size_t)y * dstride;
            const float* __restrict__ c = &R[(size_t)(y + 3) * w];
            float* __restrict__ s = acc.data();
            for (int x = 0; x < w; x++) s[x] = k[3] * c[x];
            for (int d = 1; d <= 3; d++) {
                const float* __restrict__ p = c + (ptrdiff_t)d * w;
                const float* __restrict__ m = c - (ptrdiff_t)d * w;
                const float kd = k[3 + d];
                for (int x = 0; x < w; x++) s[x] = fmaf(p[x] + m[x], kd, s[x]);
            }
            for (int x = 0; x < w; x++) {
                float v = nearbyintf(s[x]);
                v = v < 0.f ? 0.f : (v > 255.f ? 255.f : v);
                D[x] = (uint8_t)(int)v;
            }
        }
        return;
    }
    static const int t256[7] = { 18, 34, 48, 56, 48, 34, 18 };
    static const int t257[7] = { 18, 34, 49, 55, 49, 34, 18 };
    const int* t = variant == ORC_BLUR_FIXED_256 ? t256 : t257;
    std::vector<int> R((size_t)(h + 6) * w);
    for (int y = -3; y < h + 3; y++) {
        const uint8_t* S = src + (ptrdiff_t)y * stride;
        for (int x = 0; x < w; x++) {
            int s = 0;
            for (int i = 0; i < 7; i++) s += S[x + i - 3] * t[i];
            R[(size_t)(y + 3) * w + x] = s;
        }
    }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int s = 0;
            for (int i = 0; i < 7; i++) s += R[(size_t)(y + i) * w + x] * t[i];
            dst[(size_t)y * dstride + x] = sat_u8((s + (1 << 15)) >> 16);
        }
}

} // namespace

/* ================================================================ extractor */
struct orc_extractor {
    int nfeatures; double scaleFactor; int nlevels, scoreType, fastTh, blurVariant; bool descFma = false;
    std::vector<float> mvScaleFactor, mvInvScaleFactor;
    std::vector<int> mnFeaturesPerLevel, umax;
    struct Level {
        int w = 0, h = 0, stride = 0, nDesired = 0, cols = 0, rows = 0, cellW = 0, cellH = 0, nfCell = 0, nKept = 0;
        std::vector<uint8_t> plane, blurred;
        std::vector<int> candCell, candX, candY, candScore, nTotal, nRetain;
    };
    std::vector<Level> lv;
};

extern "C" {

/* The restatement of cosf / sinf above against the HOST's libm on the angles with bit patterns lo, lo + step, .. <= hi (degrees, as
 * kpt.angle): the number of angles on which cos or sin differ.  Test infrastructure for tests/test_oracle_golden.py. */
long long orc_trig_mismatches(uint32_t lo, uint32_t hi, uint32_t step)
{
    const float factorPI = (float)(M_PI / 180.f);
    long long bad = 0;
    for (uint64_t u = lo; u <= hi; u += step) {
        uint32_t b32 = (uint32_t)u; float deg; memcpy(&deg, &b32, 4);
        volatile float rad = deg * factorPI;
        float sn, cs;
        ref_sincosf(rad, &sn, &cs);
        const float hc = cosf(rad), hs = sinf(rad);
        bad += memcmp(&cs, &hc, 4) != 0 || memcmp(&sn, &hs, 4) != 0;
    }
    return bad;
}

orc_extractor* orc_extractor_create(int nfeatures, float scale_factor, int nlevels,
                                    int score_type, int fast_th, int blur_variant)
{
    /* src/ORBextractor.cc:457-511 */
    orc_extractor* e = new orc_extractor;
    e->nfeatures = nfeatures; e->scaleFactor = scale_factor; e->nlevels = nlevels;
    e->scoreType = score_type; e->fastTh = fast_th; e->blurVariant = blur_variant;
    e->mvScaleFactor.resize(nlevels); e->mvInvScaleFactor.resize(nlevels);
    e->mvScaleFactor[0] = 1;
    for (int i = 1; i < nlevels; i++) e->mvScaleFactor[i] = (float)(e->mvScaleFactor[i - 1] * e->scaleFactor);
    float invScaleFactor = (float)(1.0f / e->scaleFactor);
    e->mvInvScaleFactor[0] = 1;
    for (int i = 1; i < nlevels; i++) e->mvInvScaleFactor[i] = e->mvInvScaleFactor[i - 1] * invScaleFactor;

    e->mnFeaturesPerLevel.resize(nlevels);
    float factor = (float)(1.0 / e->scaleFactor);
    float nDesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int level = 0; level < nlevels - 1; level++) {
        e->mnFeaturesPerLevel[level] = cv_roundf(nDesired);
        sum += e->mnFeaturesPerLevel[level];
        nDesired *= factor;
    }
    e->mnFeaturesPerLevel[nlevels - 1] = std::max(nfeatures - sum, 0);

    e->umax.assign(HALF_PATCH_SIZE + 1, 0);
    int v, v0, vmax = cv_floor(HALF_PATCH_SIZE * sqrtf(2.f) / 2 + 1);
    int vmin = cv_ceil(HALF_PATCH_SIZE * sqrtf(2.f) / 2);
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (v = 0; v <= vmax; ++v) e->umax[v] = cv_round(sqrt(hp2 - v * v));
    for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
        while (e->umax[v0] == e->umax[v0 + 1]) ++v0;
        e->umax[v] = v0;
        ++v0;
    }
    e->lv.resize(nlevels);
    return e;
}

void orc_extractor_destroy(orc_extractor* e) { delete e; }
void orc_extractor_set_descriptor_fma(orc_extractor* e, int on) { e->descFma = on != 0; }

/* HarrisResponses(img, pts, 7, HARRIS_K), src/ORBextractor.cc:79-120, for one keypoint at (x, y) of the image whose origin is img */
static float harris_response(const uint8_t* img, int step, int x, int y)
{
    const int blockSize = 7, r = blockSize / 2;
    const float harris_k = 0.04f;                                   /* HARRIS_K, :73 */
    float scale = (1 << 2) * blockSize * 255.0f;
    scale = 1.0f / scale;
    const float scale_sq_sq = scale * scale * scale * scale;
    const uint8_t* ptr0 = img + (ptrdiff_t)(y - r) * step + (x - r);
    int a = 0, b = 0, c = 0;
    for (int i = 0; i < blockSize; i++)
        for (int j = 0; j < blockSize; j++) {
            const uint8_t* ptr = ptr0 + (ptrdiff_t)i * step + j;
            const int Ix = (ptr[1] - ptr[-1]) * 2 + (ptr[-step + 1] - ptr[-step - 1]) + (ptr[step + 1] - ptr[step - 1]);
            const int Iy = (ptr[step] - ptr[-step]) * 2 + (ptr[step - 1] - ptr[-step - 1]) + (ptr[step + 1] - ptr[-step + 1]);
            a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
        }
    return ((float)a * b - (float)c * c - harris_k * ((float)a + b) * ((float)a + b)) * scale_sq_sq;
}
float orc_harris_response(const uint8_t* img, int stride, int x, int y) { return harris_response(img, stride, x, y); }

int orc_extract(orc_extractor* e, const uint8_t* img, int w, int h, int stride,
                orc_keypoint* kps, uint8_t* desc, int cap, int* n)
{
    *n = 0;
    if (!img || w <= 0 || h <= 0) return 0;             /* empty image: silent return, :721-722 */
    const int nlevels = e->nlevels, B = EDGE_THRESHOLD;

    /* ---- ComputePyramid, :781-822 (mask pyramid is never consumed, see :601-607) ---- */
    for (int level = 0; level < nlevels; ++level) {
        orc_extractor::Level& L = e->lv[level];
        float scale = e->mvInvScaleFactor[level];
        L.w = cv_roundf((float)w * scale); L.h = cv_roundf((float)h * scale);
        if (L.w < 1 || L.h < 1) return -2;
        L.stride = L.w + 2 * B;
        L.plane.assign((size_t)L.stride * (L.h + 2 * B), 0);
        uint8_t* roi = &L.plane[(size_t)B * L.stride + B];
        if (level != 0) {
            orc_extractor::Level& P = e->lv[level - 1];
            resize_linear_u8(&P.plane[(size_t)B * P.stride + B], P.w, P.h, P.stride, roi, L.w, L.h, L.stride);
        } else {
            for (int y = 0; y < h; y++) memcpy(roi + (size_t)y * L.stride, img + (size_t)y * stride, w);
        }
        border_reflect101(L.plane.data(), L.w, L.h, L.stride, B);
    }

    /* ---- ComputeKeyPoints, :522-707 ---- */
    std::vector<std::vector<KP> > allKeypoints(nlevels);
    float imageRatio = (float)e->lv[0].w / e->lv[0].h;
    for (int level = 0; level < nlevels; ++level) {
        orc_extractor::Level& L = e->lv[level];
        const int nDesiredFeatures = e->mnFeaturesPerLevel[level];
        const int levelCols = (int)sqrtf((float)nDesiredFeatures / (5 * imageRatio));
        const int levelRows = (int)(imageRatio * levelCols);
        L.nDesired = nDesiredFeatures; L.cols = levelCols; L.rows = levelRows;
        L.candCell.clear(); L.candX.clear(); L.candY.clear(); L.candScore.clear();
        if (levelCols < 1 || levelRows < 1) return -2;   /* reference divides by zero here */

        const int minBorderX = EDGE_THRESHOLD, minBorderY = minBorderX;
        const int maxBorderX = L.w - EDGE_THRESHOLD, maxBorderY = L.h - EDGE_THRESHOLD;
        const int W = maxBorderX - minBorderX, H = maxBorderY - minBorderY;
        const int cellW = (int)ceilf((float)W / levelCols);
        const int cellH = (int)ceilf((float)H / levelRows);
        const int nCells = levelRows * levelCols;
        const int nfeaturesCell = (int)ceilf((float)nDesiredFeatures / nCells);
        L.cellW = cellW; L.cellH = cellH; L.nfCell = nfeaturesCell;

        std::vector<std::vector<std::vector<KP> > > cellKeyPoints(levelRows, std::vector<std::vector<KP> >(levelCols));
        std::vector<std::vector<int> > nToRetain(levelRows, std::vector<int>(levelCols, 0));
        std::vector<std::vector<int> > nTotal(levelRows, std::vector<int>(levelCols, 0));
        std::vector<std::vector<char> > bNoMore(levelRows, std::vector<char>(levelCols, 0));
        std::vector<int> iniXCol(levelCols, 0), iniYRow(levelRows, 0);
        int nNoMore = 0, nToDistribute = 0;

        const uint8_t* roi = &L.plane[(size_t)B * L.stride + B];
        float hY = (float)(cellH + 6);
        for (int i = 0; i < levelRows; i++) {
            const float iniY = (float)(minBorderY + i * cellH - 3);
            iniYRow[i] = (int)iniY;
            if (i == levelRows - 1) {
                hY = maxBorderY + 3 - iniY;
                if (hY <= 0) continue;
            }
            float hX = (float)(cellW + 6);
            for (int j = 0; j < levelCols; j++) {
                float iniX;
                if (i == 0) { iniX = (float)(minBorderX + j * cellW - 3); iniXCol[j] = (int)iniX; }
                else iniX = (float)iniXCol[j];
                if (j == levelCols - 1) {
                    hX = maxBorderX + 3 - iniX;
                    if (hX <= 0) continue;
                }
                /* Mat::rowRange/colRange on the ROI: the reference throws if the cell leaves it */
                int x0 = (int)iniX, x1 = (int)(iniX + hX), y0 = (int)iniY, y1 = (int)(iniY + hY);
                if (x0 < 0 || y0 < 0 || x1 > L.w || y1 > L.h || x1 < x0 || y1 < y0) return -2;
                const uint8_t* cellImage = roi + (size_t)y0 * L.stride + x0;
                std::vector<KP>& kc = cellKeyPoints[i][j];
                fast9_nms(cellImage, x1 - x0, y1 - y0, L.stride, e->fastTh, kc);      /* :607 */
                if (kc.size() <= 3) { kc.clear(); fast9_nms(cellImage, x1 - x0, y1 - y0, L.stride, 7, kc); } /* :609-614 */
                if (e->scoreType == 0)                                                   /* HARRIS_SCORE, :616-620 */
                    for (KP& k : kc) k.response = harris_response(cellImage, L.stride, (int)k.x, (int)k.y);
                for (size_t k = 0; k < kc.size(); k++) {
                    L.candCell.push_back(i * levelCols + j); L.candX.push_back((int)kc[k].x);
                    L.candY.push_back((int)kc[k].y); L.candScore.push_back((int)kc[k].response);
                }
                const int nKeys = (int)kc.size();
                nTotal[i][j] = nKeys;
                if (nKeys > nfeaturesCell) { nToRetain[i][j] = nfeaturesCell; bNoMore[i][j] = 0; }
                else { nToRetain[i][j] = nKeys; nToDistribute += nfeaturesCell - nKeys; bNoMore[i][j] = 1; nNoMore++; }
            }
        }
        /* quota redistribution, :644-670 */
        while (nToDistribute > 0 && nNoMore < nCells) {
            int nNewFeaturesCell = nfeaturesCell + (int)ceilf((float)nToDistribute / (nCells - nNoMore));
            nToDistribute = 0;
            for (int i = 0; i < levelRows; i++)
                for (int j = 0; j < levelCols; j++)
                    if (!bNoMore[i][j]) {
                        if (nTotal[i][j] > nNewFeaturesCell) { nToRetain[i][j] = nNewFeaturesCell; bNoMore[i][j] = 0; }
                        else {
                            nToRetain[i][j] = nTotal[i][j];
                            nToDistribute += nNewFeaturesCell - nTotal[i][j];
                            bNoMore[i][j] = 1; nNoMore++;
                        }
                    }
        }
        L.nTotal.clear(); L.nRetain.clear();
        for (int i = 0; i < levelRows; i++) for (int j = 0; j < levelCols; j++) {
            L.nTotal.push_back(nTotal[i][j]); L.nRetain.push_back(nToRetain[i][j]);
        }

        std::vector<KP>& keypoints = allKeypoints[level];
        const int scaledPatchSize = (int)(PATCH_SIZE * e->mvScaleFactor[level]);       /* :675 */
        for (int i = 0; i < levelRows; i++)
            for (int j = 0; j < levelCols; j++) {
                std::vector<KP>& keysCell = cellKeyPoints[i][j];
                retain_best(keysCell, nToRetain[i][j]);                                 /* :683-685 */
                for (size_t k = 0; k < keysCell.size(); k++) {
                    keysCell[k].x += iniXCol[j]; keysCell[k].y += iniYRow[i];
                    keysCell[k].octave = level; keysCell[k].size = (float)scaledPatchSize;
                    keypoints.push_back(keysCell[k]);
                }
            }
        if ((int)keypoints.size() > nDesiredFeatures) retain_best(keypoints, nDesiredFeatures); /* :697-701 */
        L.nKept = (int)keypoints.size();
    }
    /* orientations on the un-blurred planes, :705-706 */
    for (int level = 0; level < nlevels; ++level) {
        orc_extractor::Level& L = e->lv[level];
        const uint8_t* roi = &L.plane[(size_t)B * L.stride + B];
        for (KP& k : allKeypoints[level])
            k.angle = ic_angle(roi + (size_t)cv_roundf(k.y) * L.stride + cv_roundf(k.x), L.stride, e->umax.data());
    }

    /* ---- operator() tail, :733-778 ---- */
    int nkeypoints = 0;
    for (int level = 0; level < nlevels; ++level) nkeypoints += (int)allKeypoints[level].size();
    if (nkeypoints > cap) { *n = nkeypoints; return -3; }
    int offset = 0;
    for (int level = 0; level < nlevels; ++level) {
        orc_extractor::Level& L = e->lv[level];
        std::vector<KP>& keypoints = allKeypoints[level];
        L.blurred = L.plane;
        if (keypoints.empty()) continue;
        gaussian_blur7(&L.plane[(size_t)B * L.stride + B], L.w, L.h, L.stride,
                       &L.blurred[(size_t)B * L.stride + B], L.stride, e->blurVariant);
        const uint8_t* roi = &L.blurred[(size_t)B * L.stride + B];
        for (size_t i = 0; i < keypoints.size(); i++) {
            const KP& k = keypoints[i];
            rbrief(roi + (size_t)cv_roundf(k.y) * L.stride + cv_roundf(k.x), L.stride, k.angle,
                   desc + (size_t)(offset + i) * 32, e->descFma);
        }
        if (level != 0) {
            float scale = e->mvScaleFactor[level];
            for (KP& k : keypoints) { k.x *= scale; k.y *= scale; }
        }
        for (size_t i = 0; i < keypoints.size(); i++) memcpy(&kps[offset + i], &keypoints[i], sizeof(KP));
        offset += (int)keypoints.size();
    }
    *n = nkeypoints;
    return 0;
}

int orc_nlevels(const orc_extractor* e) { return e->nlevels; }
float orc_scale_factor(const orc_extractor* e, int l) { return e->mvScaleFactor[l]; }
float orc_inv_scale_factor(const orc_extractor* e, int l) { return e->mvInvScaleFactor[l]; }
int orc_features_per_level(const orc_extractor* e, int l) { return e->mnFeaturesPerLevel[l]; }
const int* orc_umax(const orc_extractor* e) { return e->umax.data(); }
int orc_level_info(const orc_extractor* e, int l, int* info)
{
    const orc_extractor::Level& L = e->lv[l];
    int v[10] = { L.w, L.h, L.stride, L.nDesired, L.cols, L.rows, L.cellW, L.cellH, L.nfCell, L.nKept };
    memcpy(info, v, sizeof v);
    return 0;
}
const uint8_t* orc_level_plane(const orc_extractor* e, int l, int blurred)
{
    return blurred ? e->lv[l].blurred.data() : e->lv[l].plane.data();
}
int orc_level_candidates(const orc_extractor* e, int l, int cap, int* cell, int* x, int* y, int* score)
{
    const orc_extractor::Level& L = e->lv[l];
    int n = (int)L.candCell.size();
    if (n > cap) return -n;
    for (int i = 0; i < n; i++) { cell[i] = L.candCell[i]; x[i] = L.candX[i]; y[i] = L.candY[i]; score[i] = L.candScore[i]; }
    return n;
}
int orc_level_quota(const orc_extractor* e, int l, int* ntotal, int* nretain)
{
    const orc_extractor::Level& L = e->lv[l];
    for (size_t i = 0; i < L.nTotal.size(); i++) { ntotal[i] = L.nTotal[i]; nretain[i] = L.nRetain[i]; }
    return (int)L.nTotal.size();
}

/* ---- primitives ---- */
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride)
{ resize_linear_u8(src, sw, sh, sstride, dst, dw, dh, dstride); }
void orc_border_reflect101(uint8_t* plane, int w, int h, int stride, int border)
{ border_reflect101(plane, w, h, stride, border); }
int orc_fast9_nms(const uint8_t* img, int w, int h, int stride, int th, int cap, int* x, int* y, int* score)
{
    std::vector<KP> v;
    int n = fast9_nms(img, w, h, stride, th, v);
    if (n > cap) return -n;
    for (int i = 0; i < n; i++) { x[i] = (int)v[i].x; y[i] = (int)v[i].y; score[i] = (int)v[i].response; }
    return n;
}
float orc_fast_atan2(float y, float x) { return fast_atan2(y, x); }
void orc_gaussian_blur7(const uint8_t* src, int w, int h, int stride, uint8_t* dst, int dstride, int variant)
{ gaussian_blur7(src, w, h, stride, dst, dstride, variant); }
void orc_nth_element_desc(float* resp, int32_t* idx, int n, int nth)
{
    std::vector<KP> v(n);
    for (int i = 0; i < n; i++) { v[i].response = resp[i]; v[i].class_id = idx[i]; }
    std::nth_element(v.begin(), v.begin() + nth, v.end(), RespGreater());
    for (int i = 0; i < n; i++) { resp[i] = v[i].response; idx[i] = v[i].class_id; }
}
int orc_retain_best(float* resp, int32_t* idx, int n, int npoints)
{
    std::vector<KP> v(n);
    for (int i = 0; i < n; i++) { v[i].response = resp[i]; v[i].class_id = idx[i]; }
    retain_best(v, npoints);
    for (size_t i = 0; i < v.size(); i++) { resp[i] = v[i].response; idx[i] = v[i].class_id; }
    return (int)v.size();
}
float orc_ic_angle(const uint8_t* center, int stride)
{
    static const int um[16] = { 15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3 };
    return ic_angle(center, stride, um);
}
void orc_rbrief(const uint8_t* center, int stride, float angle_deg, uint8_t* desc32)
{ rbrief(center, stride, angle_deg, desc32); }

/* ================================================================== matcher */
/* ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:1794-1810 (bit-hack popcount) */
int orc_descriptor_distance(const uint8_t* a, const uint8_t* b)
{
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t x, y;
        memcpy(&x, a + 4 * i, 4); memcpy(&y, b + 4 * i, 4);
        unsigned int v = x ^ y;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}
static inline int dist_popcnt(const uint8_t* a, const uint8_t* b)
{
    uint64_t x[4], y[4];
    memcpy(x, a, 32); memcpy(y, b, 32);
    return __builtin_popcountll(x[0] ^ y[0]) + __builtin_popcountll(x[1] ^ y[1]) +
           __builtin_popcountll(x[2] ^ y[2]) + __builtin_popcountll(x[3] ^ y[3]);
}

/* best / second-best sequential scan with strict '<' (src/ORBmatcher.cc:197-222) */
void orc_knn2(const uint8_t* q, int nq, const uint8_t* db, int64_t ndb,
              int32_t* idx1, int32_t* d1, int32_t* d2, int use_popcnt)
{
    for (int i = 0; i < nq; i++) {
        int best1 = INT_MAX, best2 = INT_MAX, bi = -1;
        const uint8_t* qi = q + (size_t)i * 32;
        for (int64_t j = 0; j < ndb; j++) {
            int dist = use_popcnt ? dist_popcnt(qi, db + (size_t)j * 32) : orc_descriptor_distance(qi, db + (size_t)j * 32);
            if (dist < best1) { best2 = best1; best1 = dist; bi = (int)j; }
            else if (dist < best2) best2 = dist;
        }
        idx1[i] = bi; d1[i] = best1; d2[i] = best2;
    }
}

/* acceptance of src/ORBmatcher.cc:224-226: best<=th && (float)best < nnratio*(float)second */
int orc_match_ratio(const int32_t* idx1, const int32_t* d1, const int32_t* d2, int nq,
                    float nnratio, int th, int32_t* match)
{
    int n = 0;
    for (int i = 0; i < nq; i++) {
        match[i] = -1;
        if (idx1[i] >= 0 && d1[i] <= th && (float)d1[i] < nnratio * (float)d2[i]) { match[i] = idx1[i]; n++; }
    }
    return n;
}

/* grid fill src/Frame.cc:109-123, PosInGrid :267-277 (round = half away from zero) */
void orc_frame_grid(const orc_keypoint* kps, int n, int min_x, int max_x, int min_y, int max_y,
                    int32_t* cell_start, int32_t* cell_items)
{
    const int GC = 64, GR = 48;
    const float invW = (float)GC / (float)(max_x - min_x), invH = (float)GR / (float)(max_y - min_y);
    std::vector<std::vector<int> > g(GC * GR);
    for (int i = 0; i < n; i++) {
        int px = (int)roundf((kps[i].x - min_x) * invW);
        int py = (int)roundf((kps[i].y - min_y) * invH);
        if (px < 0 || px >= GC || py < 0 || py >= GR) continue;
        g[px * GR + py].push_back(i);
    }
    int o = 0;
    for (int c = 0; c < GC * GR; c++) {
        cell_start[c] = o;
        for (int v : g[c]) cell_items[o++] = v;
    }
    cell_start[GC * GR] = o;
}

/* Frame::GetFeaturesInArea, src/Frame.cc:200-265 */
int orc_features_in_area(const orc_frame* f, float x, float y, float r, int minLevel, int maxLevel,
                         int32_t* out, int cap)
{
    const int GC = 64, GR = 48;
    const float invW = (float)GC / (float)(f->max_x - f->min_x), invH = (float)GR / (float)(f->max_y - f->min_y);
    int n = 0;
    int nMinCellX = (int)floorf((x - f->min_x - r) * invW);
    nMinCellX = std::max(0, nMinCellX);
    if (nMinCellX >= GC) return 0;
    int nMaxCellX = (int)ceilf((x - f->min_x + r) * invW);
    nMaxCellX = std::min(GC - 1, nMaxCellX);
    if (nMaxCellX < 0) return 0;
    int nMinCellY = (int)floorf((y - f->min_y - r) * invH);
    nMinCellY = std::max(0, nMinCellY);
    if (nMinCellY >= GR) return 0;
    int nMaxCellY = (int)ceilf((y - f->min_y + r) * invH);
    nMaxCellY = std::min(GR - 1, nMaxCellY);
    if (nMaxCellY < 0) return 0;
    bool bCheckLevels = true, bSameLevel = false;
    if (minLevel == -1 && maxLevel == -1) bCheckLevels = false;
    else if (minLevel == maxLevel) bSameLevel = true;
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
            int c = ix * GR + iy;
            for (int j = f->cell_start[c]; j < f->cell_start[c + 1]; j++) {
                int id = f->cell_items[j];
                const orc_keypoint& kp = f->kps[id];
                if (bCheckLevels && !bSameLevel) { if (kp.octave < minLevel || kp.octave > maxLevel) continue; }
                else if (bSameLevel) { if (kp.octave != minLevel) continue; }
                if (std::fabs(kp.x - x) > r || std::fabs(kp.y - y) > r) continue;
                if (n < cap) out[n] = id;
                n++;
            }
        }
    return n;
}

/* ORBmatcher::ComputeThreeMaxima, src/ORBmatcher.cc:1748-1789 */
void orc_three_maxima(const int* hs, int L, int* i1, int* i2, int* i3)
{
    int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
    for (int i = 0; i < L; i++) {
        const int s = hs[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) ind3 = -1;
    *i1 = ind1; *i2 = ind2; *i3 = ind3;
}

static int rot_bin(float a_from, float a_to)
{
    /* src/ORBmatcher.cc:1583-1588 / :234-239; factor = 1/HISTO_LENGTH (sic) */
    const float factor = 1.0f / 30;
    float rot = a_from - a_to;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)roundf(rot * factor);
    if (bin == 30) bin = 0;
    return bin;
}

/* ORBmatcher::SearchByProjection(Frame&, const Frame&, float), src/ORBmatcher.cc:1507-1620.
 * Rcw*x3Dw+tcw is one cv::gemm on CV_32F 3x3 * 3x1 (+C): OpenCV's small-matrix branch sums the
 * three products in FP32 and adds the translation in double (pinned via cv2.gemm, see tests). */
int orc_search_by_projection(const orc_frame* cur, const orc_frame* last, const uint8_t* last_has_mp,
                             const uint8_t* last_outlier, const float* last_xyz, const float* T,
                             float th, int check_ori, int32_t* match_cur)
{
    const int HISTO_LENGTH = 30, TH_HIGH = 100;
    int nmatches = 0;
    std::vector<int> rotHist[30];
    std::vector<float> sf(cur->nlevels);
    sf[0] = 1.0f;
    for (int i = 1; i < cur->nlevels; i++) sf[i] = sf[i - 1] * cur->scale_factor;  /* src/Frame.cc:95-103 */
    std::vector<int32_t> cand(cur->n > 0 ? cur->n : 1);
    for (int i = 0; i < last->n; i++) {
        if (!last_has_mp[i] || last_outlier[i]) continue;
        const float X = last_xyz[3 * i], Y = last_xyz[3 * i + 1], Z = last_xyz[3 * i + 2];
        float c[3];
        for (int r = 0; r < 3; r++) {
            float t0 = T[4 * r + 0] * X + T[4 * r + 1] * Y + T[4 * r + 2] * Z;
            c[r] = (float)((double)t0 * 1.0 + (double)T[4 * r + 3] * 1.0);
        }
        const float xc = c[0], yc = c[1];
        const float invzc = (float)(1.0 / c[2]);
        float u = cur->fx * xc * invzc + cur->cx;
        float v = cur->fy * yc * invzc + cur->cy;
        if (u < cur->min_x || u > cur->max_x) continue;
        if (v < cur->min_y || v > cur->max_y) continue;
        int oct = last->kps[i].octave;
        float radius = th * sf[oct];
        int nc = orc_features_in_area(cur, u, v, radius, oct - 1, oct + 1, cand.data(), cur->n);
        if (nc == 0) continue;
        const uint8_t* dMP = last->desc + (size_t)i * 32;
        int bestDist = INT_MAX, bestIdx2 = -1;
        for (int k = 0; k < nc; k++) {
            int i2 = cand[k];
            if (match_cur[i2] >= 0) continue;
            int dist = orc_descriptor_distance(dMP, cur->desc + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= TH_HIGH) {
            match_cur[bestIdx2] = i;
            nmatches++;
            if (check_ori) rotHist[rot_bin(last->kps[i].angle, cur->kps[bestIdx2].angle)].push_back(bestIdx2);
        }
    }
    if (check_ori) {
        int hs[30], i1, i2, i3;
        for (int i = 0; i < HISTO_LENGTH; i++) hs[i] = (int)rotHist[i].size();
        orc_three_maxima(hs, HISTO_LENGTH, &i1, &i2, &i3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != i1 && i != i2 && i != i3)
                for (int id : rotHist[i]) { match_cur[id] = -1; nmatches--; }
    }
    return nmatches;
}

/* ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...), src/ORBmatcher.cc:155-284 */
int orc_search_by_bow(const orc_featvec* kfv, const uint8_t* kf_desc, const orc_keypoint* kf_kps,
                      const uint8_t* kf_mp_valid, int n_kf,
                      const orc_featvec* ffv, const uint8_t* f_desc, const orc_keypoint* f_kps, int n_f,
                      float nnratio, int check_ori, int32_t* match_f)
{
    (void)n_kf;
    const int HISTO_LENGTH = 30, TH_LOW = 50;
    for (int i = 0; i < n_f; i++) match_f[i] = -1;
    int nmatches = 0;
    std::vector<int> rotHist[30];
    int a = 0, b = 0;
    while (a < kfv->nnodes && b < ffv->nnodes) {
        if (kfv->node_id[a] == ffv->node_id[b]) {
            for (int ik = kfv->start[a]; ik < kfv->start[a + 1]; ik++) {
                const int realIdxKF = kfv->items[ik];
                if (!kf_mp_valid[realIdxKF]) continue;
                const uint8_t* dKF = kf_desc + (size_t)realIdxKF * 32;
                int bestDist1 = INT_MAX, bestIdxF = -1, bestDist2 = INT_MAX;
                for (int jf = ffv->start[b]; jf < ffv->start[b + 1]; jf++) {
                    const int realIdxF = ffv->items[jf];
                    if (match_f[realIdxF] >= 0) continue;
                    const int dist = orc_descriptor_distance(dKF, f_desc + (size_t)realIdxF * 32);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = realIdxF; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 <= TH_LOW && (float)bestDist1 < nnratio * (float)bestDist2) {
                    match_f[bestIdxF] = realIdxKF;
                    if (check_ori) rotHist[rot_bin(kf_kps[realIdxKF].angle, f_kps[bestIdxF].angle)].push_back(bestIdxF);
                    nmatches++;
                }
            }
            a++; b++;
        } else if (kfv->node_id[a] < ffv->node_id[b]) {
            while (a < kfv->nnodes && kfv->node_id[a] < ffv->node_id[b]) a++;     /* lower_bound */
        } else {
            while (b < ffv->nnodes && ffv->node_id[b] < kfv->node_id[a]) b++;
        }
    }
    if (check_ori) {
        int hs[30], i1, i2, i3;
        for (int i = 0; i < HISTO_LENGTH; i++) hs[i] = (int)rotHist[i].size();
        orc_three_maxima(hs, HISTO_LENGTH, &i1, &i2, &i3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == i1 || i == i2 || i == i3) continue;
            for (int id : rotHist[i]) { match_f[id] = -1; nmatches--; }
        }
    }
    return nmatches;
}


/* ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th), src/ORBmatcher.cc:49-125 */
int orc_search_by_projection_mappoints(const orc_frame* f, int nmp, const uint8_t* in_view, const float* proj_x, const float* proj_y,
                                       const int32_t* level, const float* view_cos, const uint8_t* mp_desc, float th, float nnratio,
                                       int32_t* match_f)
{
    const int TH_HIGH = 100;
    int nmatches = 0;
    const bool bFactor = th != 1.0;
    std::vector<float> sf(f->nlevels);
    sf[0] = 1.0f;
    for (int i = 1; i < f->nlevels; i++) sf[i] = sf[i - 1] * f->scale_factor;
    std::vector<int32_t> cand(f->n > 0 ? f->n : 1);
    for (int iMP = 0; iMP < nmp; iMP++) {
        if (!in_view[iMP]) continue;
        const int nPredictedLevel = level[iMP];
        float r = view_cos[iMP] > 0.998 ? 2.5f : 4.0f;                 /* RadiusByViewingCos :127-133 */
        if (bFactor) r *= th;
        int nc = orc_features_in_area(f, proj_x[iMP], proj_y[iMP], r * sf[nPredictedLevel], nPredictedLevel - 1, nPredictedLevel,
                                      cand.data(), f->n);
        if (nc == 0) continue;
        const uint8_t* d = mp_desc + (size_t)iMP * 32;
        int bestDist = INT_MAX, bestLevel = -1, bestDist2 = INT_MAX, bestLevel2 = -1, bestIdx = -1;
        for (int k = 0; k < nc; k++) {
            const int idx = cand[k];
            if (match_f[idx] >= 0) continue;
            const int dist = orc_descriptor_distance(d, f->desc + (size_t)idx * 32);
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = f->kps[idx].octave; bestIdx = idx; }
            else if (dist < bestDist2) { bestLevel2 = f->kps[idx].octave; bestDist2 = dist; }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
            match_f[bestIdx] = iMP;
            nmatches++;
        }
    }
    return nmatches;
}

/* ORBmatcher::WindowSearch, src/ORBmatcher.cc:409-516 */
int orc_window_search(const orc_frame* f1, const orc_frame* f2, const uint8_t* f1_has_mp, int windowSize, int minScaleLevel, int maxScaleLevel,
                      float nnratio, int check_ori, int32_t* match2)
{
    const int HISTO_LENGTH = 30, TH_HIGH = 100;
    int nmatches = 0;
    for (int i = 0; i < f2->n; i++) match2[i] = -1;
    std::vector<int> rotHist[30];
    const bool bMinLevel = minScaleLevel > 0, bMaxLevel = maxScaleLevel < INT_MAX;
    std::vector<int32_t> cand(f2->n > 0 ? f2->n : 1);
    for (int i1 = 0; i1 < f1->n; i1++) {
        if (!f1_has_mp[i1]) continue;
        const orc_keypoint& kp1 = f1->kps[i1];
        const int level1 = kp1.octave;
        if (bMinLevel && level1 < minScaleLevel) continue;
        if (bMaxLevel && level1 > maxScaleLevel) continue;
        int nc = orc_features_in_area(f2, kp1.x, kp1.y, (float)windowSize, level1, level1, cand.data(), f2->n);
        if (nc == 0) continue;
        const uint8_t* d1 = f1->desc + (size_t)i1 * 32;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int k = 0; k < nc; k++) {
            const int i2 = cand[k];
            if (match2[i2] >= 0) continue;
            const int dist = orc_descriptor_distance(d1, f2->desc + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= bestDist2 * nnratio && bestDist <= TH_HIGH) {
            match2[bestIdx2] = i1;
            nmatches++;
            rotHist[rot_bin(f1->kps[i1].angle, f2->kps[bestIdx2].angle)].push_back(bestIdx2);
        }
    }
    if (check_ori) {
        int hs[30], i1, i2, i3;
        for (int i = 0; i < HISTO_LENGTH; i++) hs[i] = (int)rotHist[i].size();
        orc_three_maxima(hs, HISTO_LENGTH, &i1, &i2, &i3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != i1 && i != i2 && i != i3)
                for (int id : rotHist[i]) { match2[id] = -1; nmatches--; }
    }
    return nmatches;
}

/* ORBmatcher::SearchByProjection(Frame &F1, Frame &F2, int windowSize, ...), src/ORBmatcher.cc:519-594 */
int orc_search_by_projection_window(const orc_frame* f1, const orc_frame* f2, const uint8_t* f1_active, const float* f1_xyz,
                                    const float* T, int windowSize, float nnratio, int32_t* match2)
{
    const int TH_HIGH = 100;
    int nmatches = 0;
    std::vector<int32_t> cand(f2->n > 0 ? f2->n : 1);
    for (int i1 = 0; i1 < f1->n; i1++) {
        if (!f1_active[i1]) continue;
        const int level1 = f1->kps[i1].octave;
        const float X = f1_xyz[3 * i1], Y = f1_xyz[3 * i1 + 1], Z = f1_xyz[3 * i1 + 2];
        float c[3];
        for (int r = 0; r < 3; r++) {
            float t0 = T[4 * r + 0] * X + T[4 * r + 1] * Y + T[4 * r + 2] * Z;
            c[r] = (float)((double)t0 * 1.0 + (double)T[4 * r + 3] * 1.0);
        }
        const float invzc2 = (float)(1.0 / c[2]);
        float u2 = f2->fx * c[0] * invzc2 + f2->cx;
        float v2 = f2->fy * c[1] * invzc2 + f2->cy;
        int nc = orc_features_in_area(f2, u2, v2, (float)windowSize, level1, level1, cand.data(), f2->n);
        if (nc == 0) continue;
        const uint8_t* d1 = f1->desc + (size_t)i1 * 32;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int k = 0; k < nc; k++) {
            const int i2 = cand[k];
            if (match2[i2] >= 0) continue;
            const int dist = orc_descriptor_distance(d1, f2->desc + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if ((float)bestDist <= (float)bestDist2 * nnratio && bestDist <= TH_HIGH) {
            match2[bestIdx2] = i1;
            nmatches++;
        }
    }
    return nmatches;
}

/* ORBmatcher::SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist), src/ORBmatcher.cc:1622-1746 */
int orc_search_by_projection_kf(const orc_frame* cur, int nmp, const uint8_t* active, const float* xyz, const float* T,
                                const int32_t* pred_level, const uint8_t* mp_desc, const float* kf_angle, float th, int orb_dist,
                                int check_ori, int32_t* match_cur)
{
    const int HISTO_LENGTH = 30;
    int nmatches = 0;
    std::vector<int> rotHist[30];
    std::vector<float> sf(cur->nlevels);
    sf[0] = 1.0f;
    for (int i = 1; i < cur->nlevels; i++) sf[i] = sf[i - 1] * cur->scale_factor;
    std::vector<int32_t> cand(cur->n > 0 ? cur->n : 1);
    for (int i = 0; i < nmp; i++) {
        if (!active[i]) continue;
        const float X = xyz[3 * i], Y = xyz[3 * i + 1], Z = xyz[3 * i + 2];
        float c[3];
        for (int r = 0; r < 3; r++) {
            float t0 = T[4 * r + 0] * X + T[4 * r + 1] * Y + T[4 * r + 2] * Z;
            c[r] = (float)((double)t0 * 1.0 + (double)T[4 * r + 3] * 1.0);
        }
        const float invzc = (float)(1.0 / c[2]);
        float u = cur->fx * c[0] * invzc + cur->cx;
        float v = cur->fy * c[1] * invzc + cur->cy;
        if (u < cur->min_x || u > cur->max_x) continue;
        if (v < cur->min_y || v > cur->max_y) continue;
        const int nPredictedLevel = pred_level[i];
        float radius = th * sf[nPredictedLevel];
        int nc = orc_features_in_area(cur, u, v, radius, nPredictedLevel - 1, nPredictedLevel + 1, cand.data(), cur->n);
        if (nc == 0) continue;
        const uint8_t* dMP = mp_desc + (size_t)i * 32;
        int bestDist = INT_MAX, bestIdx2 = -1;
        for (int k = 0; k < nc; k++) {
            int i2 = cand[k];
            if (match_cur[i2] >= 0) continue;
            int dist = orc_descriptor_distance(dMP, cur->desc + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= orb_dist) {
            match_cur[bestIdx2] = i;
            nmatches++;
            if (check_ori) rotHist[rot_bin(kf_angle[i], cur->kps[bestIdx2].angle)].push_back(bestIdx2);
        }
    }
    if (check_ori) {
        int hs[30], i1, i2, i3;
        for (int i = 0; i < HISTO_LENGTH; i++) hs[i] = (int)rotHist[i].size();
        orc_three_maxima(hs, HISTO_LENGTH, &i1, &i2, &i3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != i1 && i != i2 && i != i3)
                for (int id : rotHist[i]) { match_cur[id] = -1; nmatches--; }
    }
    return nmatches;
}

/* ORBmatcher::SearchForInitialization, src/ORBmatcher.cc:598-713 */
int orc_search_for_initialization(const orc_frame* f1, const orc_frame* f2, float* prev, int windowSize, float nnratio,
                                  int check_ori, int32_t* vnMatches12)
{
    const int HISTO_LENGTH = 30, TH_LOW = 50;
    int nmatches = 0;
    for (int i = 0; i < f1->n; i++) vnMatches12[i] = -1;
    std::vector<int> rotHist[30];
    std::vector<int> vMatchedDistance(f2->n > 0 ? f2->n : 1, INT_MAX), vnMatches21(f2->n > 0 ? f2->n : 1, -1);
    std::vector<int32_t> cand(f2->n > 0 ? f2->n : 1);
    for (int i1 = 0; i1 < f1->n; i1++) {
        const int level1 = f1->kps[i1].octave;
        if (level1 > 0) continue;
        int nc = orc_features_in_area(f2, prev[2 * i1], prev[2 * i1 + 1], (float)windowSize, level1, level1, cand.data(), f2->n);
        if (nc == 0) continue;
        const uint8_t* d1 = f1->desc + (size_t)i1 * 32;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int k = 0; k < nc; k++) {
            const int i2 = cand[k];
            const int dist = orc_descriptor_distance(d1, f2->desc + (size_t)i2 * 32);
            if (vMatchedDistance[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= TH_LOW) {
            if (bestDist < (float)bestDist2 * nnratio) {
                if (vnMatches21[bestIdx2] >= 0) { vnMatches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
                vnMatches12[i1] = bestIdx2;
                vnMatches21[bestIdx2] = i1;
                vMatchedDistance[bestIdx2] = bestDist;
                nmatches++;
                if (check_ori) rotHist[rot_bin(f1->kps[i1].angle, f2->kps[bestIdx2].angle)].push_back(i1);
            }
        }
    }
    if (check_ori) {
        int hs[30], a, b, c;
        for (int i = 0; i < HISTO_LENGTH; i++) hs[i] = (int)rotHist[i].size();
        orc_three_maxima(hs, HISTO_LENGTH, &a, &b, &c);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == a || i == b || i == c) continue;
            for (int idx1 : rotHist[i])
                if (vnMatches12[idx1] >= 0) { vnMatches12[idx1] = -1; nmatches--; }
        }
    }
    for (int i1 = 0; i1 < f1->n; i1++)
        if (vnMatches12[i1] >= 0) { prev[2 * i1] = f2->kps[vnMatches12[i1]].x; prev[2 * i1 + 1] = f2->kps[vnMatches12[i1]].y; }
    return nmatches;
}

/* src/ORBmatcher.cc:350-404 */
int orc_search_by_projection_sim3(const orc_frame* kf, int nmp, const uint8_t* active, const float* u, const float* v,
                                  const int32_t* pred_level, const uint8_t* mp_desc, int th, int32_t* matched)
{
    const int TH_LOW = 50;
    std::vector<float> sf(kf->nlevels > 0 ? kf->nlevels : 1, 1.0f);
    for (int i = 1; i < kf->nlevels; i++) sf[i] = sf[i - 1] * kf->scale_factor;
    std::vector<int32_t> cand(kf->n > 0 ? kf->n : 1);
    int nmatches = 0;
    for (int i = 0; i < nmp; i++) {
        if (!active[i]) continue;
        const int nPredictedLevel = pred_level[i];
        const float radius = th * sf[nPredictedLevel];
        const int nc = orc_features_in_area(kf, u[i], v[i], radius, -1, -1, cand.data(), kf->n);
        if (nc == 0) continue;
        int bestDist = INT_MAX, bestIdx = -1;
        for (int k = 0; k < nc; k++) {
            const int idx = cand[k];
            if (matched[idx] >= 0) continue;
            const int kpLevel = kf->kps[idx].octave;
            if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
            const int dist = orc_descriptor_distance(mp_desc + (size_t)i * 32, kf->desc + (size_t)idx * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        if (bestDist <= TH_LOW) { matched[bestIdx] = i; nmatches++; }
    }
    return nmatches;
}

void orc_window_best(const orc_frame* f, int nq, const uint8_t* active, const float* u, const float* v, const float* radius,
                     const int32_t* pred_level, const uint8_t* desc, int32_t* best_idx, int32_t* best_dist)
{
    std::vector<int32_t> cand(f->n > 0 ? f->n : 1);
    for (int i = 0; i < nq; i++) {
        best_idx[i] = -1; best_dist[i] = INT_MAX;
        if (!active[i]) continue;
        const int nc = orc_features_in_area(f, u[i], v[i], radius[i], -1, -1, cand.data(), f->n);
        for (int k = 0; k < nc; k++) {
            const int idx = cand[k], lv = f->kps[idx].octave;
            if (lv < pred_level[i] - 1 || lv > pred_level[i]) continue;
            const int dist = orc_descriptor_distance(desc + (size_t)i * 32, f->desc + (size_t)idx * 32);
            if (dist < best_dist[i]) { best_dist[i] = dist; best_idx[i] = idx; }
        }
    }
}

/* ORBmatcher::CheckDistEpipolarLine, src/ORBmatcher.cc:136-153 */
static bool check_dist_epipolar_line(const orc_keypoint& kp1, const orc_keypoint& kp2, const float* F12, const float* sigma2)
{
    const float a = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
    const float b = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
    const float c = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
    const float num = a * kp2.x + b * kp2.y + c;
    const float den = a * a + b * b;
    if (den == 0) return false;
    const float dsqr = num * num / den;
    return dsqr < 3.84 * sigma2[kp2.octave];
}

/* src/ORBmatcher.cc:852-1014 */
int orc_search_for_triangulation(const orc_featvec* fv1, const uint8_t* desc1, const orc_keypoint* kps1, const uint8_t* has_mp1, int n1,
                                 const orc_featvec* fv2, const uint8_t* desc2, const orc_keypoint* kps2, const uint8_t* has_mp2, int n2,
                                 const float* F12, const float* level_sigma2, int check_ori, int32_t* vMatches12)
{
    const int HISTO_LENGTH = 30, TH_LOW = 50;
    int nmatches = 0;
    std::vector<char> vbMatched2(n2 > 0 ? n2 : 1, 0);
    for (int i = 0; i < n1; i++) vMatches12[i] = -1;
    std::vector<int> rotHist[30];
    int a = 0, b = 0;
    while (a < fv1->nnodes && b < fv2->nnodes) {
        if (fv1->node_id[a] == fv2->node_id[b]) {
            for (int i1 = fv1->start[a]; i1 < fv1->start[a + 1]; i1++) {
                const int idx1 = fv1->items[i1];
                if (has_mp1[idx1]) continue;
                std::vector<std::pair<int, size_t> > vDistIndex;
                for (int i2 = fv2->start[b]; i2 < fv2->start[b + 1]; i2++) {
                    const int idx2 = fv2->items[i2];
                    if (vbMatched2[idx2] || has_mp2[idx2]) continue;
                    const int dist = orc_descriptor_distance(desc1 + (size_t)idx1 * 32, desc2 + (size_t)idx2 * 32);
                    if (dist > TH_LOW) continue;
                    vDistIndex.push_back(std::make_pair(dist, (size_t)idx2));
                }
                if (vDistIndex.empty()) continue;
                std::sort(vDistIndex.begin(), vDistIndex.end());
                const int BestDist = vDistIndex.front().first;
                const int DistTh = (int)round(2 * BestDist);
                for (size_t id = 0; id < vDistIndex.size(); id++) {
                    if (vDistIndex[id].first > DistTh) break;
                    const int currentIdx2 = (int)vDistIndex[id].second;
                    if (check_dist_epipolar_line(kps1[idx1], kps2[currentIdx2], F12, level_sigma2)) {
                        vbMatched2[currentIdx2] = 1;
                        vMatches12[idx1] = currentIdx2;
                        nmatches++;
                        if (check_ori) rotHist[rot_bin(kps1[idx1].angle, kps2[currentIdx2].angle)].push_back(idx1);
                        break;
                    }
                }
            }
            a++; b++;
        } else if (fv1->node_id[a] < fv2->node_id[b]) {
            while (a < fv1->nnodes && fv1->node_id[a] < fv2->node_id[b]) a++;
        } else {
            while (b < fv2->nnodes && fv2->node_id[b] < fv1->node_id[a]) b++;
        }
    }
    if (check_ori) {
        int hs[30], i1, i2, i3;
        for (int i = 0; i < HISTO_LENGTH; i++) hs[i] = (int)rotHist[i].size();
        orc_three_maxima(hs, HISTO_LENGTH, &i1, &i2, &i3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == i1 || i == i2 || i == i3) continue;
            for (int id : rotHist[i]) { vMatches12[id] = -1; nmatches--; }
        }
    }
    return nmatches;
}

/* ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, ...), src/ORBmatcher.cc:715-850 */
int orc_search_by_bow_kf(const orc_featvec* fv1, const uint8_t* desc1, const orc_keypoint* kps1, const uint8_t* valid1, int n1,
                         const orc_featvec* fv2, const uint8_t* desc2, const orc_keypoint* kps2, const uint8_t* valid2, int n2,
                         float nnratio, int check_ori, int32_t* match12)
{
    const int HISTO_LENGTH = 30, TH_LOW = 50;
    for (int i = 0; i < n1; i++) match12[i] = -1;
    std::vector<char> vbMatched2(n2 > 0 ? n2 : 1, 0);
    std::vector<int> rotHist[30];
    int nmatches = 0, a = 0, b = 0;
    while (a < fv1->nnodes && b < fv2->nnodes) {
        if (fv1->node_id[a] == fv2->node_id[b]) {
            for (int i1 = fv1->start[a]; i1 < fv1->start[a + 1]; i1++) {
                const int idx1 = fv1->items[i1];
                if (!valid1[idx1]) continue;
                const uint8_t* d1 = desc1 + (size_t)idx1 * 32;
                int bestDist1 = INT_MAX, bestIdx2 = -1, bestDist2 = INT_MAX;
                for (int i2 = fv2->start[b]; i2 < fv2->start[b + 1]; i2++) {
                    const int idx2 = fv2->items[i2];
                    if (vbMatched2[idx2] || !valid2[idx2]) continue;
                    const int dist = orc_descriptor_distance(d1, desc2 + (size_t)idx2 * 32);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 < TH_LOW && (float)bestDist1 < nnratio * (float)bestDist2) {
                    match12[idx1] = bestIdx2;
                    vbMatched2[bestIdx2] = 1;
                    if (check_ori) rotHist[rot_bin(kps1[idx1].angle, kps2[bestIdx2].angle)].push_back(idx1);
                    nmatches++;
                }
            }
            a++; b++;
        } else if (fv1->node_id[a] < fv2->node_id[b]) {
            while (a < fv1->nnodes && fv1->node_id[a] < fv2->node_id[b]) a++;
        } else {
            while (b < fv2->nnodes && fv2->node_id[b] < fv1->node_id[a]) b++;
        }
    }
    if (check_ori) {
        int hs[30], i1, i2, i3;
        for (int i = 0; i < HISTO_LENGTH; i++) hs[i] = (int)rotHist[i].size();
        orc_three_maxima(hs, HISTO_LENGTH, &i1, &i2, &i3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == i1 || i == i2 || i == i3) continue;
            for (int id : rotHist[i]) { match12[id] = -1; nmatches--; }
        }
    }
    return nmatches;
}


/* ------------------------------------------------------------------------------------------------------------------
 * DBoW2 vocabulary: Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h, BowVector.cpp, FeatureVector.cpp, ScoringObject.cpp.
 * The vocabulary file of the reference (Data/ORBvoc.txt, k=10 L=6 L1_NORM TF_IDF) is not in the repository; tests use
 * synthetic trees in the same text format.
 * ------------------------------------------------------------------------------------------------------------------ */
} // extern "C"
#include <map>
#include <cstdio>
#include <cstdlib>
struct orc_vocab {
    int k, L, scoring, weighting;
    std::vector<std::vector<int>> children;      /* m_nodes[i].children */
    std::vector<uint8_t> desc;                   /* nnodes x 32 */
    std::vector<double> weight;
    std::vector<int> word_id;                    /* -1 for inner nodes */
    int nwords;
};
static void vocab_finish(orc_vocab* v)
{
    const int n = (int)v->children.size();
    v->word_id.assign(n, -1);
    v->nwords = 0;
    for (int i = 1; i < n; i++)                  /* :1408-1414: leaves get word ids in node order */
        if (v->children[i].empty()) v->word_id[i] = v->nwords++;
}
extern "C" {

orc_vocab* orc_vocab_create(int k, int L, int scoring, int weighting, int nnodes, const int32_t* parent, const uint8_t* desc,
                            const double* weight)
{
    if (nnodes < 1) return nullptr;
    for (int i = 1; i < nnodes; i++) if (parent[i] < 0 || parent[i] >= i) return nullptr;   /* a parent precedes its children in the file */
    orc_vocab* v = new orc_vocab;
    v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting;
    v->children.resize(nnodes);
    v->desc.assign(desc, desc + (size_t)nnodes * 32);
    v->weight.assign(weight, weight + nnodes);
    for (int i = 1; i < nnodes; i++) v->children[parent[i]].push_back(i);
    vocab_finish(v);
    return v;
}

orc_vocab* orc_vocab_load_text(const char* path)
{
    FILE* f = fopen(path, "r");              /* C stdio: this library may carry a static libstdc++, keep iostreams out of it */
    if (!f) return nullptr;
    std::vector<char> line(1 << 16);
    if (!fgets(line.data(), (int)line.size(), f)) { fclose(f); return nullptr; }
    int k = 0, L = 0, n1 = -1, n2 = -1;
    sscanf(line.data(), "%d %d %d %d", &k, &L, &n1, &n2);
    if (k < 0 || k > 20 || L < 1 || L > 10 || n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3) { fclose(f); return nullptr; }      /* :1358-1362 */
    std::vector<int32_t> parent(1, 0);
    std::vector<uint8_t> desc(32, 0);
    std::vector<double> weight(1, 0.0);
    while (fgets(line.data(), (int)line.size(), f)) {
        const char* p = line.data();
        char* end = nullptr;
        const long pid = strtol(p, &end, 10);
        if (end == p) continue;              /* blank trailing line (the reference would read garbage here) */
        p = end;
        strtol(p, &end, 10);                 /* nIsLeaf */
        p = end;
        parent.push_back((int32_t)pid);
        for (int i = 0; i < 32; i++) { const long b = strtol(p, &end, 10); desc.push_back((uint8_t)b); p = end; }
        weight.push_back(strtod(p, &end));
    }
    fclose(f);
    return orc_vocab_create(k, L, n1, n2, (int)parent.size(), parent.data(), desc.data(), weight.data());
}

void orc_vocab_destroy(orc_vocab* v) { delete v; }
int orc_vocab_nnodes(const orc_vocab* v) { return (int)v->children.size(); }
int orc_vocab_nwords(const orc_vocab* v) { return v->nwords; }

/* :1218-1260.  Where the reference leaves *nid unset (a leaf above level L-levelsup) the leaf itself is reported. */
void orc_vocab_transform_feature(const orc_vocab* v, const uint8_t* d, int levelsup, int32_t* word, double* weight, int32_t* node)
{
    const int nid_level = v->L - levelsup;
    int nid = -1;
    if (nid_level <= 0) nid = 0;
    int final_id = 0, current_level = 0;
    if (v->children[0].empty()) { *word = -1; *weight = 0; *node = 0; return; }
    do {
        ++current_level;
        const std::vector<int>& nodes = v->children[final_id];
        final_id = nodes[0];
        double best_d = orc_descriptor_distance(d, &v->desc[(size_t)final_id * 32]);
        for (size_t c = 1; c < nodes.size(); c++) {
            const double dist = orc_descriptor_distance(d, &v->desc[(size_t)nodes[c] * 32]);
            if (dist < best_d) { best_d = dist; final_id = nodes[c]; }
        }
        if (current_level == nid_level) nid = final_id;
    } while (!v->children[final_id].empty());
    *word = v->word_id[final_id];
    *weight = v->weight[final_id];
    *node = nid >= 0 ? nid : final_id;
}

/* :1127-1193 with BowVector::addWeight / addIfNotExist / normalize (BowVector.cpp:33-95), FeatureVector::addFeature (:31-45) */
void orc_vocab_transform(const orc_vocab* v, const uint8_t* desc, int n, int levelsup, int32_t* bow_word, double* bow_val, int* nbow,
                         int32_t* fv_node, int32_t* fv_start, int32_t* fv_items, int* nfv)
{
    std::map<unsigned, double> bow;
    std::map<unsigned, std::vector<unsigned>> fv;
    *nbow = 0; *nfv = 0; fv_start[0] = 0;
    if (v->children[0].empty()) return;
    const bool l2 = v->scoring == ORC_L2_NORM, must = v->scoring != ORC_DOT_PRODUCT;
    const bool tf = v->weighting == ORC_TF || v->weighting == ORC_TF_IDF;
    for (int i = 0; i < n; i++) {
        int32_t id, nid; double w;
        orc_vocab_transform_feature(v, desc + (size_t)i * 32, levelsup, &id, &w, &nid);
        if (w > 0) {
            auto it = bow.find((unsigned)id);
            if (it == bow.end()) bow[(unsigned)id] = w;
            else if (tf) it->second += w;
            fv[(unsigned)nid].push_back((unsigned)i);
        }
    }
    if (tf && !bow.empty() && !must) {
        const double nd = (double)bow.size();
        for (auto& e : bow) e.second /= nd;
    }
    if (must) {
        double norm = 0.0;
        if (!l2) { for (auto& e : bow) norm += fabs(e.second); }
        else { for (auto& e : bow) norm += e.second * e.second; norm = sqrt(norm); }
        if (norm > 0.0) for (auto& e : bow) e.second /= norm;
    }
    int c = 0;
    for (auto& e : bow) { bow_word[c] = (int32_t)e.first; bow_val[c] = e.second; c++; }
    *nbow = c;
    int nn = 0, pos = 0;
    for (auto& e : fv) {
        fv_node[nn] = (int32_t)e.first;
        for (unsigned idx : e.second) fv_items[pos++] = (int32_t)idx;
        fv_start[++nn] = pos;
    }
    *nfv = nn;
}

/* ScoringObject.cpp:22-64 */
double orc_bow_score_l1(const int32_t* w1, const double* v1, int n1, const int32_t* w2, const double* v2, int n2)
{
    int a = 0, b = 0;
    double score = 0;
    while (a < n1 && b < n2) {
        if (w1[a] == w2[b]) { score += fabs(v1[a] - v2[b]) - fabs(v1[a]) - fabs(v2[b]); a++; b++; }
        else if (w1[a] < w2[b]) a = (int)(std::lower_bound(w1 + a, w1 + n1, w2[b]) - w1);
        else b = (int)(std::lower_bound(w2 + b, w2 + n2, w1[a]) - w2);
    }
    return -score / 2.0;
}

/* src/KeyFrameDatabase.cc:198-252 without the list / covisibility bookkeeping */
void orc_bow_score_db(const int32_t* qw, const double* qv, int nq, int nkf, const int32_t* kf_start, const int32_t* kf_word,
                      const double* kf_val, int32_t* common, float* score, int* max_common)
{
    int mx = 0;
    for (int k = 0; k < nkf; k++) {
        const int32_t* w = kf_word + kf_start[k];
        const int n = kf_start[k + 1] - kf_start[k];
        int c = 0, a = 0, b = 0;
        while (a < nq && b < n) { if (qw[a] == w[b]) { c++; a++; b++; } else if (qw[a] < w[b]) a++; else b++; }
        common[k] = c;
        mx = std::max(mx, c);
    }
    const int min_common = (int)((float)mx * 0.8f);
    for (int k = 0; k < nkf; k++) {
        score[k] = 0.f;
        if (common[k] > 0 && common[k] > min_common)
            score[k] = (float)orc_bow_score_l1(qw, qv, nq, kf_word + kf_start[k], kf_val + kf_start[k], kf_start[k + 1] - kf_start[k]);
    }
    *max_common = mx;
}

/* KeyFrameDatabase::DetectRelocalisationCandidates (src/KeyFrameDatabase.cc:198-308, loop = 0) and DetectLoopCandidates (:75-196,
 * loop = 1) restated on flat arrays; keyframe index = order of KeyFrameDatabase::add, so every inverted-file list (:41-47) is
 * ascending in k.  kf_score = the mRelocScore / mLoopScore members (in / out). */
int orc_bow_detect_candidates(const int32_t* qw, const double* qv, int nq, int nkf, const int32_t* kf_start, const int32_t* kf_word,
                              const double* kf_val, const uint8_t* excluded, int loop, float min_score, const int32_t* cov_start,
                              const int32_t* cov_idx, float* kf_score, int32_t* common, int32_t* cand)
{
    std::map<int, std::vector<int> > inverted;                          /* mvInvertedFile */
    for (int k = 0; k < nkf; k++) for (int j = kf_start[k]; j < kf_start[k + 1]; j++) inverted[kf_word[j]].push_back(k);
    std::vector<int> query(nkf, 0), words(nkf, 0), sharing;             /* mnRelocQuery == F->mnId, mnRelocWords, lKFsSharingWords */
    for (int a = 0; a < nq; a++) {                                      /* :205-222 / :85-104 */
        std::map<int, std::vector<int> >::const_iterator it = inverted.find(qw[a]);
        if (it == inverted.end()) continue;
        for (size_t i = 0; i < it->second.size(); i++) {
            const int k = it->second[i];
            if (!query[k]) {
                words[k] = 0;
                if (!(loop && excluded && excluded[k])) { query[k] = 1; sharing.push_back(k); }
            }
            words[k]++;
        }
    }
    for (int k = 0; k < nkf; k++) common[k] = query[k] ? words[k] : 0;
    if (sharing.empty()) return 0;
    int maxCommonWords = 0;
    for (size_t i = 0; i < sharing.size(); i++) maxCommonWords = std::max(maxCommonWords, words[sharing[i]]);
    const int minCommonWords = (int)((float)maxCommonWords * 0.8f);    /* :233 / :117 */
    std::vector<int> scored;                                            /* lScoreAndMatch */
    for (size_t i = 0; i < sharing.size(); i++) {
        const int k = sharing[i];
        if (words[k] > minCommonWords) {
            const float si = (float)orc_bow_score_l1(qw, qv, nq, kf_word + kf_start[k], kf_val + kf_start[k], kf_start[k + 1] - kf_start[k]);
            kf_score[k] = si;
            if (!loop || si >= min_score) scored.push_back(k);           /* :131-132 */
        }
    }
    if (scored.empty()) return 0;
    std::vector<std::pair<float, int> > acc;                            /* lAccScoreAndMatch */
    float bestAccScore = loop ? min_score : 0.f;                        /* :139 / :260 */
    for (size_t i = 0; i < scored.size(); i++) {
        const int k = scored[i];
        float bestScore = kf_score[k], accScore = kf_score[k];
        int best = k;
        if (cov_start)
            for (int j = cov_start[k]; j < cov_start[k + 1] && j < cov_start[k] + 10; j++) {      /* GetBestCovisibilityKeyFrames(10) */
                const int k2 = cov_idx[j];
                if (!query[k2]) continue;                                /* :274 / :153 */
                if (loop && !(words[k2] > minCommonWords)) continue;     /* :153 */
                accScore += kf_score[k2];
                if (kf_score[k2] > bestScore) { best = k2; bestScore = kf_score[k2]; }
            }
        acc.push_back(std::make_pair(accScore, best));
        if (accScore > bestAccScore) bestAccScore = accScore;
    }
    const float minScoreToRetain = 0.75f * bestAccScore;
    std::set<int> added;
    int n = 0;
    for (size_t i = 0; i < acc.size(); i++)
        if (acc[i].first > minScoreToRetain && !added.count(acc[i].second)) { cand[n++] = acc[i].second; added.insert(acc[i].second); }
    return n;
}


/* ------------------------------------------------------------------------------------------------------------------
 * Frame plumbing: src/Tracking.cc:202-208 (cvtColor), src/Frame.cc:289-349 (undistortPoints).  OpenCV 4.13 arithmetic,
 * checked against cv2 in tests/test_oracle_frame.py.
 * ------------------------------------------------------------------------------------------------------------------ */
void orc_cvt_gray(const uint8_t* src, int w, int h, int stride, int order, uint8_t* dst, int dstride)
{
    const int c0 = order ? 3735 : 9798, c2 = order ? 9798 : 3735;
    for (int y = 0; y < h; y++) {
        const uint8_t* s = src + (size_t)y * stride;
        uint8_t* d = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) d[x] = (uint8_t)((s[3 * x] * c0 + s[3 * x + 1] * 19235 + s[3 * x + 2] * c2 + (1 << 14)) >> 15);
    }
}

void orc_undistort_points(float* xy, int n, float ffx, float ffy, float fcx, float fcy, const float* dist, int ndist)
{
    const double fx = ffx, fy = ffy, cx = fcx, cy = fcy;
    double k[14] = { 0 };
    for (int i = 0; i < ndist && i < 14; i++) k[i] = dist[i];
    const double ifx = 1. / fx, ify = 1. / fy;
    for (int i = 0; i < n; i++) {
        double x = xy[2 * i], y = xy[2 * i + 1];
        const double u = x, v = y;
        x = (x - cx) * ifx; y = (y - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; j++) {                       /* TermCriteria(MAX_ITER, 5, 0.01): count only */
            const double r2 = x * x + y * y;
            const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
            if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
            const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
            const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        const double xx = fx * x + 0.0 * y + cx, yy = 0.0 * x + fy * y + cy, ww = 1. / (0.0 * x + 0.0 * y + 1.0);
        xy[2 * i] = (float)(xx * ww); xy[2 * i + 1] = (float)(yy * ww);
    }
}

void orc_undistort_keypoints(const orc_keypoint* in, int n, float fx, float fy, float cx, float cy, const float* dist, int ndist,
                             orc_keypoint* out)
{
    for (int i = 0; i < n; i++) out[i] = in[i];
    if (ndist == 0 || dist[0] == 0.f) return;               /* src/Frame.cc:291-295 */
    for (int i = 0; i < n; i++) {
        float p[2] = { in[i].x, in[i].y };
        orc_undistort_points(p, 1, fx, fy, cx, cy, dist, ndist);
        out[i].x = p[0]; out[i].y = p[1];
    }
}

void orc_image_bounds(int w, int h, float fx, float fy, float cx, float cy, const float* dist, int ndist, int32_t b[4])
{
    if (ndist == 0 || dist[0] == 0.f) { b[0] = 0; b[1] = w; b[2] = 0; b[3] = h; return; }
    float m[8] = { 0.f, 0.f, (float)w, 0.f, 0.f, (float)h, (float)w, (float)h };
    orc_undistort_points(m, 4, fx, fy, cx, cy, dist, ndist);
    b[0] = (int32_t)std::min(floorf(m[0]), floorf(m[4]));
    b[1] = (int32_t)std::max(ceilf(m[2]), ceilf(m[6]));
    b[2] = (int32_t)std::min(floorf(m[1]), floorf(m[3]));
    b[3] = (int32_t)std::max(ceilf(m[5]), ceilf(m[7]));
}


/* src/MapPoint.cc:214-242 */
void orc_distinctive_descriptors(const uint8_t* desc, const int32_t* start, int npoints, int32_t* best_idx, int32_t* best_median)
{
    for (int p = 0; p < npoints; p++) {
        const int N = start[p + 1] - start[p];
        best_idx[p] = -1; best_median[p] = INT_MAX;
        if (N <= 0) continue;
        const uint8_t* d = desc + (size_t)start[p] * 32;
        std::vector<float> D((size_t)N * N);
        for (int i = 0; i < N; i++) {
            D[(size_t)i * N + i] = 0;
            for (int j = i + 1; j < N; j++) {
                const int dij = orc_descriptor_distance(d + (size_t)i * 32, d + (size_t)j * 32);
                D[(size_t)i * N + j] = (float)dij; D[(size_t)j * N + i] = (float)dij;
            }
        }
        int BestMedian = INT_MAX, BestIdx = 0;
        for (int i = 0; i < N; i++) {
            std::vector<int> v(D.begin() + (size_t)i * N, D.begin() + (size_t)(i + 1) * N);
            std::sort(v.begin(), v.end());
            const int median = v[(size_t)(0.5 * (N - 1))];
            if (median < BestMedian) { BestMedian = median; BestIdx = i; }
        }
        best_idx[p] = BestIdx; best_median[p] = BestMedian;
    }
}

} // extern "C"
