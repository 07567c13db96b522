// orb_frame.cu — the frame plumbing on either side of the extractor:
//   colour -> gray of the incoming image          Tracking::GrabImage, src/Tracking.cc:200-212 (cvtColor RGB2GRAY / BGR2GRAY)
//   Frame::UndistortKeyPoints                     src/Frame.cc:289-320 (cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK))
//   Frame::ComputeImageBounds                     src/Frame.cc:322-349
// OpenCV arithmetic pinned to 4.13 like the rest of the path and checked against cv2 bit for bit (tests/test_oracle_frame.py):
//   gray  = (R*9798 + G*19235 + B*3735 + 2^14) >> 15                      (imgproc color_rgb RGB2Gray<uchar>, 15-bit coefficients)
//   undistortPoints = 5 fixed-point iterations of the inverse Brown model in double, no epsilon test (TermCriteria(MAX_ITER, 5)),
//                     then re-projection with P = K; results rounded to float.
#include "orb_internal.h"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace {

bool on_device(const void* p)
{
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// one thread = 4 pixels: 12 source bytes (three aligned words when the row start is 4-byte aligned) -> one output word
__global__ void __launch_bounds__(256)
k_gray(const uint8_t* __restrict__ src, int w, int h, size_t stride, size_t frame_pitch, int bgr, uint8_t* __restrict__ dst,
       size_t dstride, size_t dpitch)
{
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y, f = blockIdx.z;
    if (x4 >= w) return;
    const uint8_t* s = src + (size_t)f * frame_pitch + (size_t)y * stride + (size_t)x4 * 3;
    uint8_t* d = dst + (size_t)f * dpitch + (size_t)y * dstride + x4;
    const int c0 = bgr ? 3735 : 9798, c2 = bgr ? 9798 : 3735;
    const int n = min(4, w - x4);
    uint32_t out = 0;
    if (n == 4 && (((uintptr_t)s) & 3) == 0) {
        const uint32_t a = __ldg(reinterpret_cast<const uint32_t*>(s)), b = __ldg(reinterpret_cast<const uint32_t*>(s) + 1),
                       c = __ldg(reinterpret_cast<const uint32_t*>(s) + 2);
        const uint32_t px[4][3] = { { a & 255, (a >> 8) & 255, (a >> 16) & 255 }, { a >> 24, b & 255, (b >> 8) & 255 },
                                    { (b >> 16) & 255, b >> 24, c & 255 }, { (c >> 8) & 255, (c >> 16) & 255, c >> 24 } };
#pragma unroll
        for (int k = 0; k < 4; k++) out |= ((px[k][0] * c0 + px[k][1] * 19235 + px[k][2] * c2 + (1u << 14)) >> 15) << (8 * k);
    } else {
        for (int k = 0; k < n; k++) out |= (((uint32_t)s[3 * k] * c0 + (uint32_t)s[3 * k + 1] * 19235 + (uint32_t)s[3 * k + 2] * c2 + (1u << 14)) >> 15) << (8 * k);
    }
    if (n == 4 && (((uintptr_t)d) & 3) == 0) *reinterpret_cast<uint32_t*>(d) = out;
    else for (int k = 0; k < n; k++) d[k] = (uint8_t)(out >> (8 * k));
}

struct Camera { double fx, fy, cx, cy, k[14]; };

// cvUndistortPointsInternal (OpenCV 4.13 calib3d/imgproc undistort.dispatch.cpp) with R = I, P = K, no tilt, 5 iterations.
// Every operation is an individually rounded double operation in the source's order (the library is built without contraction).
__device__ __forceinline__ void undistort_point(const Camera& C, float xin, float yin, float& xo, float& yo)
{
    const double ifx = __ddiv_rn(1.0, C.fx), ify = __ddiv_rn(1.0, C.fy);
    const double u = (double)xin, v = (double)yin;
    double x = __dmul_rn(__dsub_rn(u, C.cx), ifx), y = __dmul_rn(__dsub_rn(v, C.cy), ify);
    const double x0 = x, y0 = y;
    const double* k = C.k;
    for (int j = 0; j < 5; j++) {
        const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
        const double num = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k[7], r2), k[6]), r2), k[5]), r2));
        const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k[4], r2), k[1]), r2), k[0]), r2));
        const double icdist = __ddiv_rn(num, den);
        if (icdist < 0) { x = __dmul_rn(__dsub_rn(u, C.cx), ifx); y = __dmul_rn(__dsub_rn(v, C.cy), ify); break; }
        // deltaX = 2*k[2]*x*y + k[3]*(r2 + 2*x*x) + k[8]*r2 + k[9]*r2*r2   (left to right)
        const double dX = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, k[2]), x), y),
                                                        __dmul_rn(k[3], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x)))),
                                              __dmul_rn(k[8], r2)), __dmul_rn(__dmul_rn(k[9], r2), r2));
        // deltaY = k[2]*(r2 + 2*y*y) + 2*k[3]*x*y + k[10]*r2 + k[11]*r2*r2
        const double dY = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(k[2], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y))),
                                                        __dmul_rn(__dmul_rn(__dmul_rn(2.0, k[3]), x), y)),
                                              __dmul_rn(k[10], r2)), __dmul_rn(__dmul_rn(k[11], r2), r2));
        x = __dmul_rn(__dsub_rn(x0, dX), icdist);
        y = __dmul_rn(__dsub_rn(y0, dY), icdist);
    }
    // xx = RR[0][0]*x + RR[0][1]*y + RR[0][2] with RR = K: the zero products stay in the sum exactly as in the library
    const double xx = __dadd_rn(__dadd_rn(__dmul_rn(C.fx, x), __dmul_rn(0.0, y)), C.cx);
    const double yy = __dadd_rn(__dadd_rn(__dmul_rn(0.0, x), __dmul_rn(C.fy, y)), C.cy);
    const double ww = __ddiv_rn(1.0, __dadd_rn(__dadd_rn(__dmul_rn(0.0, x), __dmul_rn(0.0, y)), 1.0));
    xo = __double2float_rn(__dmul_rn(xx, ww));
    yo = __double2float_rn(__dmul_rn(yy, ww));
}

__global__ void __launch_bounds__(256)
k_undistort_keypoints(Camera C, const orb_keypoint* __restrict__ in, int n, orb_keypoint* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    orb_keypoint kp = in[i];
    undistort_point(C, kp.x, kp.y, kp.x, kp.y);
    out[i] = kp;
}

int make_camera(float fx, float fy, float cx, float cy, const float* dist, int ndist, Camera& C)
{
    if (ndist < 0 || ndist > 14 || (ndist > 0 && !dist)) return ORB_ERR_INVALID;
    if (ndist != 0 && ndist != 4 && ndist != 5 && ndist != 8 && ndist != 12 && ndist != 14) return ORB_ERR_INVALID;   // the sizes cv::undistortPoints accepts
    if (ndist > 12 && (dist[12] != 0.f || dist[13] != 0.f)) return ORB_ERR_UNSUPPORTED;                                  // tilted sensor model
    C.fx = fx; C.fy = fy; C.cx = cx; C.cy = cy;
    for (int i = 0; i < 14; i++) C.k[i] = i < ndist ? (double)dist[i] : 0.0;
    return ORB_OK;
}

} // namespace

extern "C" {

int orb_cvt_gray(orb_ctx* c, const uint8_t* src, int nimg, int w, int h, size_t stride, size_t frame_pitch, int order,
                 uint8_t* dst, size_t dst_stride, size_t dst_pitch)
{
    if (!c || nimg < 0 || w < 0 || h < 0 || (order != ORB_RGB && order != ORB_BGR)) return ORB_ERR_INVALID;
    if (nimg == 0 || w == 0 || h == 0) return ORB_OK;
    if (!src || !dst || stride < (size_t)w * 3 || dst_stride < (size_t)w) return ORB_ERR_INVALID;
    ORB_CUDA(cudaSetDevice(c->device));
    LaneGuard lg(c);
    if (!lg.lane) return ORB_ERR_CUDA;
    cudaStream_t s = lg.lane->stream;
    const bool dev_in = on_device(src), dev_out = on_device(dst);
    const size_t in_bytes = (size_t)(nimg - 1) * frame_pitch + (size_t)(h - 1) * stride + (size_t)w * 3;
    const size_t out_bytes = (size_t)(nimg - 1) * dst_pitch + (size_t)(h - 1) * dst_stride + (size_t)w;
    uint8_t *d_in = nullptr, *d_out = nullptr;
    int rc = ORB_OK;
    auto cleanup = [&]() { if (!dev_in && d_in) cudaFree(d_in); if (!dev_out && d_out) cudaFree(d_out); };
    if (!dev_in) {
        if (cudaMalloc(&d_in, in_bytes) != cudaSuccess) return ORB_ERR_CUDA;
        if (cudaMemcpyAsync(d_in, src, in_bytes, cudaMemcpyHostToDevice, s) != cudaSuccess) { cleanup(); return ORB_ERR_CUDA; }
    }
    if (!dev_out && cudaMalloc(&d_out, out_bytes) != cudaSuccess) { cleanup(); return ORB_ERR_CUDA; }
    k_gray<<<dim3((w + 1023) / 1024, h, nimg), 256, 0, s>>>(dev_in ? src : d_in, w, h, stride, frame_pitch, order == ORB_BGR,
                                                          dev_out ? dst : d_out, dst_stride, dst_pitch);
    if (cudaGetLastError() != cudaSuccess) rc = ORB_ERR_CUDA;
    if (rc == ORB_OK && !dev_out && cudaMemcpyAsync(dst, d_out, out_bytes, cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = ORB_ERR_CUDA;
    if (cudaStreamSynchronize(s) != cudaSuccess) rc = ORB_ERR_CUDA;
    cleanup();
    return rc;
}

int orb_extract_batch_color(orb_ctx* c, const uint8_t* images, int nimg, int w, int h, size_t stride, size_t frame_pitch, int order,
                            orb_keypoint* kps, uint8_t* desc, int cap, int32_t* counts)
{
    if (!c || nimg < 0 || (order != ORB_RGB && order != ORB_BGR)) return ORB_ERR_INVALID;
    if (nimg == 0 || !images || w <= 0 || h <= 0) return orb_extract_batch(c, nullptr, nimg, w, h, w, (size_t)w * h, kps, desc, cap, counts);
    ORB_CUDA(cudaSetDevice(c->device));
    uint8_t* d_gray = nullptr;
    const size_t gpitch = (size_t)w * h;
    ORB_CUDA(cudaMalloc(&d_gray, gpitch * nimg));
    int rc = orb_cvt_gray(c, images, nimg, w, h, stride, frame_pitch, order, d_gray, (size_t)w, gpitch);
    if (rc == ORB_OK) rc = orb_extract_batch(c, d_gray, nimg, w, h, w, gpitch, kps, desc, cap, counts);
    cudaFree(d_gray);
    return rc;
}

int orb_undistort_keypoints(orb_ctx* c, const orb_keypoint* kps, int n, float fx, float fy, float cx, float cy, const float* dist,
                            int ndist, orb_keypoint* kps_un)
{
    if (!c || n < 0) return ORB_ERR_INVALID;
    if (n == 0) return ORB_OK;
    if (!kps || !kps_un) return ORB_ERR_INVALID;
    Camera C;
    int rc = make_camera(fx, fy, cx, cy, dist, ndist, C);
    if (rc) return rc;
    ORB_CUDA(cudaSetDevice(c->device));
    LaneGuard lg(c);
    if (!lg.lane) return ORB_ERR_CUDA;
    cudaStream_t s = lg.lane->stream;
    const bool dev = on_device(kps);
    if (on_device(kps_un) != dev) return ORB_ERR_INVALID;
    const size_t bytes = (size_t)n * sizeof(orb_keypoint);
    if (ndist == 0 || dist[0] == 0.f) {                      // src/Frame.cc:291-295: mvKeysUn = mvKeys
        if (kps_un != kps) ORB_CUDA(cudaMemcpyAsync(kps_un, kps, bytes, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToHost, s));
        ORB_CUDA(cudaStreamSynchronize(s));
        return ORB_OK;
    }
    orb_keypoint *d_in = const_cast<orb_keypoint*>(kps), *d_out = kps_un;
    if (!dev) {
        ORB_CUDA(cudaMalloc(&d_in, bytes));
        d_out = d_in;
        if (cudaMemcpyAsync(d_in, kps, bytes, cudaMemcpyHostToDevice, s) != cudaSuccess) { cudaFree(d_in); return ORB_ERR_CUDA; }
    }
    k_undistort_keypoints<<<(n + 255) / 256, 256, 0, s>>>(C, d_in, n, d_out);
    rc = cudaGetLastError() == cudaSuccess ? ORB_OK : ORB_ERR_CUDA;
    if (rc == ORB_OK && !dev && cudaMemcpyAsync(kps_un, d_out, bytes, cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = ORB_ERR_CUDA;
    if (cudaStreamSynchronize(s) != cudaSuccess) rc = ORB_ERR_CUDA;
    if (!dev) cudaFree(d_in);
    return rc;
}

int orb_image_bounds(orb_ctx* c, int w, int h, float fx, float fy, float cx, float cy, const float* dist, int ndist, int32_t bounds[4])
{
    if (!c || !bounds || w < 0 || h < 0) return ORB_ERR_INVALID;
    if (ndist == 0 || !dist || dist[0] == 0.f) { bounds[0] = 0; bounds[1] = w; bounds[2] = 0; bounds[3] = h; return ORB_OK; }   // :341-347
    orb_keypoint corners[4];
    memset(corners, 0, sizeof corners);
    corners[1].x = (float)w; corners[2].y = (float)h; corners[3].x = (float)w; corners[3].y = (float)h;                    // :326-330
    const int rc = orb_undistort_keypoints(c, corners, 4, fx, fy, cx, cy, dist, ndist, corners);
    if (rc) return rc;
    bounds[0] = (int32_t)std::min(std::floor(corners[0].x), std::floor(corners[2].x));      // mnMinX :337
    bounds[1] = (int32_t)std::max(std::ceil(corners[1].x), std::ceil(corners[3].x));        // mnMaxX
    bounds[2] = (int32_t)std::min(std::floor(corners[0].y), std::floor(corners[1].y));      // mnMinY
    bounds[3] = (int32_t)std::max(std::ceil(corners[2].y), std::ceil(corners[3].y));        // mnMaxY
    return ORB_OK;
}

} // extern "C"
