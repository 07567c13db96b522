/*
 * minicv.hpp — the sliver of the OpenCV C++ API that the reference's ORB front end touches, written from scratch
 * so that the reference's OWN sources (src/ORBextractor.cc, src/ORBmatcher.cc, src/Frame.cc, src/KeyFrame.cc,
 * src/MapPoint.cc, src/Map.cc, src/KeyFrameDatabase.cc, Thirdparty/DBoW2) compile from where they lie under
 * /root/reference into oracle/_ref/ (TEST INFRASTRUCTURE, never product code; see oracle/Makefile).
 *
 * What this is and is not.  The container has no OpenCV C++ SDK, no ROS and no boost, so the reference cannot be
 * built as shipped.  This header gives its code a `cv::` namespace whose containers (Mat, KeyPoint, Point, Rect,
 * InputArray ...) are re-implemented here and whose ARITHMETIC primitives (resize, copyMakeBorder, FAST,
 * GaussianBlur, fastAtan2, undistortPoints, gemm) forward to the oracle's restatements, each of which is pinned
 * bit for bit against the real OpenCV 4.13 through python cv2 (tests/test_oracle_golden.py, test_oracle_frame.py).
 * KeyPointsFilter::retainBest is written here directly over std::nth_element / std::partition (OpenCV-4 form).
 * The result pins the reference's own control flow — cell grid, quota redistribution, threshold fallback, level and
 * keypoint order, descriptor bit order, every ORBmatcher search, the Frame grid, DBoW2's transform and scoring —
 * as executed by the reference's code, against the oracle and the CUDA path.
 */
#ifndef ORB_REFSHIM_MINICV_HPP
#define ORB_REFSHIM_MINICV_HPP
#include <algorithm>
#include <cassert>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <set>
#include <sstream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../orb_oracle.h"

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_PI 3.1415926535897932384626433832795
#define CV_Assert(expr) do { if (!(expr)) throw std::runtime_error(std::string("CV_Assert failed: ") + #expr); } while (0)
#define CV_Error(code, msg) throw std::runtime_error(msg)

/* round half to even under the default rounding mode, like the SSE2 cvtsd2si OpenCV uses */
inline int cvRound(double v) { return (int)lrint(v); }
inline int cvRound(float v) { return (int)lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv {

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename U> Point_(const Point_<U>& p) : x((T)p.x), y((T)p.y) {}
    Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
    Point_ operator-(const Point_& o) const { return Point_(x - o.x, y - o.y); }
    Point_ operator+(const Point_& o) const { return Point_(x + o.x, y + o.y); }
};
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T a, T b, T c) : x(a), y(b), z(c) {}
};
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;
template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;
template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
typedef Rect_<int> Rect;
struct Range { int start, end; Range(int s, int e) : start(s), end(e) {} };
template <typename T, int N> struct Vec { T val[N]; T& operator[](int i) { return val[i]; } const T& operator[](int i) const { return val[i]; } };
typedef Vec<float, 3> Vec3f;

struct KeyPoint {      /* 28 bytes, same member order as cv::KeyPoint */
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
        : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};

template <typename T, size_t fixed = 1032 / sizeof(T) + 8> class AutoBuffer {
    std::vector<T> v;
public:
    AutoBuffer() {}
    explicit AutoBuffer(size_t n) : v(n) {}
    operator T*() { return v.data(); }
    operator const T*() const { return v.data(); }
    T* data() { return v.data(); }
    size_t size() const { return v.size(); }
};

class Mat;
class MatExpr;
/* Mat::zeros / Mat::eye return an initialiser, as in OpenCV: ASSIGNING it to a Mat that already has that size and type fills
 * the existing buffer in place (the reference relies on this at src/ORBextractor.cc:712, where the left-hand side is a
 * row range of the output descriptor matrix) */
struct MatInit { int rows, cols, type; bool eye; };
class _InputArray;
class _OutputArray;
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
typedef const _OutputArray& InputOutputArray;

class Mat {
public:
    enum { AUTO_STEP = 0 };
    int rows, cols;
    uchar* data;
    size_t step;                                   /* bytes per row */
    uchar *datastart, *dataend;                    /* bounds of the allocation this header looks into */

    Mat() : rows(0), cols(0), data(0), step(0), datastart(0), dataend(0), type_(0) {}
    Mat(int r, int c, int type) : Mat() { create(r, c, type); }
    Mat(Size s, int type) : Mat() { create(s.height, s.width, type); }
    Mat(int r, int c, int type, void* d, size_t st = AUTO_STEP) : rows(r), cols(c), data((uchar*)d), type_(type)
    {
        step = st ? st : (size_t)c * elemSize();
        datastart = data;
        dataend = data + (size_t)r * step;
    }
    Mat(const Mat& m, const Rect& roi) : Mat(m) { adjust(roi.y, roi.y + roi.height, roi.x, roi.x + roi.width); }
    Mat(const MatExpr& e);
    Mat& operator=(const MatExpr& e);
    Mat(const MatInit& z) : Mat() { *this = z; }
    Mat& operator=(const MatInit& z)
    {
        create(z.rows, z.cols, z.type);
        for (int y = 0; y < rows; y++) std::memset(data + (size_t)y * step, 0, (size_t)cols * elemSize());
        if (z.eye)
            for (int i = 0; i < std::min(rows, cols); i++) {
                if (depth() == CV_32F) at<float>(i, i) = 1.f; else if (depth() == CV_64F) at<double>(i, i) = 1.0; else at<uchar>(i, i) = 1;
            }
        return *this;
    }

    static int depthSize(int depth) { static const int s[7] = {1, 1, 2, 2, 4, 4, 8}; return s[depth & 7]; }
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize1() const { return (size_t)depthSize(type_); }
    size_t elemSize() const { return elemSize1() * channels(); }
    size_t step1() const { return step / elemSize1(); }
    bool empty() const { return data == 0 || rows == 0 || cols == 0; }
    size_t total() const { return (size_t)rows * cols; }
    Size size() const { return Size(cols, rows); }
    bool isContinuous() const { return rows <= 1 || step == (size_t)cols * elemSize(); }
    bool isSubmatrix() const { return data && (data != datastart || (size_t)(dataend - datastart) != (size_t)rows * step || step != (size_t)cols * elemSize()); }

    void create(int r, int c, int type)
    {
        if (data && r == rows && c == cols && type == type_) return;          /* cv::Mat::create keeps a matching buffer */
        release();
        rows = r; cols = c; type_ = type;
        step = (size_t)c * elemSize();
        const size_t n = std::max<size_t>((size_t)r * step, 1);
        owner = std::shared_ptr<uchar>((uchar*)std::malloc(n), std::free);
        data = datastart = owner.get();
        dataend = data + (size_t)r * step;
    }
    void create(Size s, int type) { create(s.height, s.width, type); }
    void release() { owner.reset(); rows = cols = 0; data = datastart = dataend = 0; step = 0; }
    static MatInit zeros(int r, int c, int type) { MatInit z = {r, c, type, false}; return z; }
    static MatInit eye(int r, int c, int type) { MatInit z = {r, c, type, true}; return z; }
    Mat clone() const { Mat m; copyToMat(m); return m; }
    void copyToMat(Mat& m) const
    {
        m.create(rows, cols, type_);
        for (int y = 0; y < rows; y++) std::memmove(m.data + y * m.step, data + y * step, (size_t)cols * elemSize());
    }
    void copyTo(OutputArray dst) const;

    Mat operator()(const Rect& roi) const { return Mat(*this, roi); }
    Mat rowRange(int a, int b) const { Mat m(*this); m.adjust(a, b, 0, cols); return m; }
    Mat colRange(int a, int b) const { Mat m(*this); m.adjust(0, rows, a, b); return m; }
    Mat row(int y) const { return rowRange(y, y + 1); }
    Mat col(int x) const { return colRange(x, x + 1); }
    Mat reshape(int cn, int = 0) const
    {
        Mat m(*this);
        const int total_cn = cols * channels();
        CV_Assert(isContinuous() || cn == channels());
        CV_Assert(total_cn % cn == 0);
        m.cols = total_cn / cn;
        m.type_ = CV_MAKETYPE(depth(), cn);
        return m;
    }

    template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
    template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }
    uchar* ptr(int y = 0) { return data + (size_t)y * step; }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
    template <typename T> T& at(int y, int x) { return ((T*)(data + (size_t)y * step))[x]; }
    template <typename T> const T& at(int y, int x) const { return ((const T*)(data + (size_t)y * step))[x]; }
    /* single index: element i of a row or column vector */
    template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }

    MatExpr t() const;
    MatExpr inv() const;
    double dot(const Mat& m) const
    {
        CV_Assert(depth() == CV_32F && m.depth() == CV_32F && total() == m.total());
        double s = 0;                                   /* cv::Mat::dot on CV_32F accumulates in double */
        const int n = (int)total();
        for (int i = 0; i < n; i++) s += (double)lin(i) * (double)m.lin(i);
        return s;
    }
    float lin(int i) const { return at<float>(i / cols, i % cols); }

private:
    int type_;
    std::shared_ptr<uchar> owner;
    void adjust(int r0, int r1, int c0, int c1)
    {
        CV_Assert(0 <= r0 && r0 <= r1 && r1 <= rows && 0 <= c0 && c0 <= c1 && c1 <= cols);
        data += (size_t)r0 * step + (size_t)c0 * elemSize();
        rows = r1 - r0; cols = c1 - c0;
    }
};

/* ---- the matrix expressions the reference writes on CV_32F poses.  cv::MatExpr folds  alpha*op(A)*op(B) + beta*C  into one
 *      cv::gemm call; gemm on CV_32F accumulates in double and rounds once (pinned against cv2.gemm, tests/test_oracle_golden.py) ---- */
class MatExpr {
public:
    enum Kind { VALUE, TRANSPOSE, GEMM };
    Kind kind;
    Mat a, b, c;
    double alpha, beta;
    bool ta, tb;
    MatExpr(const Mat& m) : kind(VALUE), a(m), alpha(1), beta(0), ta(false), tb(false) {}
    MatExpr(Kind k, const Mat& a_, double alpha_) : kind(k), a(a_), alpha(alpha_), beta(0), ta(false), tb(false) {}
    Mat eval() const
    {
        if (kind == VALUE && alpha == 1) return a;
        CV_Assert(a.depth() == CV_32F && a.channels() == 1);
        if (kind == VALUE || kind == TRANSPOSE) {
            const bool tr = kind == TRANSPOSE;
            Mat r(tr ? a.cols : a.rows, tr ? a.rows : a.cols, CV_32F);
            for (int i = 0; i < r.rows; i++)
                for (int j = 0; j < r.cols; j++) {
                    const float v = tr ? a.at<float>(j, i) : a.at<float>(i, j);
                    r.at<float>(i, j) = alpha == 1 ? v : (float)(v * alpha);
                }
            return r;
        }
        const int M = ta ? a.cols : a.rows, K = ta ? a.rows : a.cols, N = tb ? b.rows : b.cols;
        CV_Assert((tb ? b.cols : b.rows) == K && b.depth() == CV_32F);
        Mat r(M, N, CV_32F);
        for (int i = 0; i < M; i++)
            for (int j = 0; j < N; j++) {
                double s = 0;
                for (int k = 0; k < K; k++)
                    s += (double)(ta ? a.at<float>(k, i) : a.at<float>(i, k)) * (double)(tb ? b.at<float>(j, k) : b.at<float>(k, j));
                s *= alpha;
                if (!c.empty()) s += beta * (double)c.at<float>(i, j);
                r.at<float>(i, j) = (float)s;
            }
        return r;
    }
    operator Mat() const { return eval(); }
    template <typename T> T at(int i, int j) const { return eval().at<T>(i, j); }
    template <typename T> T at(int i) const { return eval().at<T>(i); }
    Mat row(int y) const { return eval().row(y); }
    Mat col(int x) const { return eval().col(x); }
    Mat rowRange(int a_, int b_) const { return eval().rowRange(a_, b_); }
    Mat colRange(int a_, int b_) const { return eval().colRange(a_, b_); }
    Mat clone() const { return eval().clone(); }
    MatExpr t() const { return MatExpr(TRANSPOSE, eval(), 1); }
    double dot(const Mat& m) const { return eval().dot(m); }
};
inline Mat::Mat(const MatExpr& e) : Mat() { *this = e.eval(); }
inline Mat& Mat::operator=(const MatExpr& e) { *this = e.eval(); return *this; }
inline MatExpr Mat::t() const { return MatExpr(MatExpr::TRANSPOSE, *this, 1); }

inline MatExpr gemm_expr(const MatExpr& x, const MatExpr& y)
{
    MatExpr r(MatExpr::GEMM, Mat(), 1);
    r.alpha = 1;
    if (x.kind == MatExpr::GEMM) r.a = x.eval(); else { r.a = x.a; r.ta = x.kind == MatExpr::TRANSPOSE; r.alpha *= x.alpha; }
    if (y.kind == MatExpr::GEMM) r.b = y.eval(); else { r.b = y.a; r.tb = y.kind == MatExpr::TRANSPOSE; r.alpha *= y.alpha; }
    return r;
}
inline MatExpr operator*(const MatExpr& x, const MatExpr& y) { return gemm_expr(x, y); }
inline MatExpr operator*(const Mat& x, const Mat& y) { return gemm_expr(MatExpr(x), MatExpr(y)); }
inline MatExpr operator*(const MatExpr& x, const Mat& y) { return gemm_expr(x, MatExpr(y)); }
inline MatExpr operator*(const Mat& x, const MatExpr& y) { return gemm_expr(MatExpr(x), y); }
inline MatExpr scaled(const MatExpr& x, double s) { MatExpr r(x); r.alpha *= s; if (r.kind == MatExpr::GEMM) r.beta *= s; return r; }
inline MatExpr operator*(double s, const MatExpr& x) { return scaled(x, s); }
inline MatExpr operator*(const MatExpr& x, double s) { return scaled(x, s); }
inline MatExpr operator*(double s, const Mat& x) { return scaled(MatExpr(x), s); }
inline MatExpr operator*(const Mat& x, double s) { return scaled(MatExpr(x), s); }
inline MatExpr operator-(const MatExpr& x) { return scaled(x, -1); }
inline MatExpr operator-(const Mat& x) { return scaled(MatExpr(x), -1); }
/* element-wise forms: alpha*A + beta*B evaluated in double per element and rounded once (cv::addWeighted / scaleAdd on CV_32F) */
inline Mat lincomb(const Mat& a, double alpha, const Mat& b, double beta)
{
    CV_Assert(a.depth() == CV_32F && b.depth() == CV_32F && a.rows == b.rows && a.cols == b.cols);
    Mat r(a.rows, a.cols, CV_32F);
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < a.cols; j++) {
            if (alpha == 1 && (beta == 1 || beta == -1))          /* cv::add / cv::subtract: one FP32 operation */
                r.at<float>(i, j) = beta == 1 ? a.at<float>(i, j) + b.at<float>(i, j) : a.at<float>(i, j) - b.at<float>(i, j);
            else
                r.at<float>(i, j) = (float)(alpha * (double)a.at<float>(i, j) + beta * (double)b.at<float>(i, j));
        }
    return r;
}
inline MatExpr operator+(const MatExpr& x, const MatExpr& y)
{
    if (x.kind == MatExpr::GEMM && x.c.empty() && y.kind == MatExpr::VALUE) { MatExpr r(x); r.c = y.a; r.beta = y.alpha; return r; }
    if (y.kind == MatExpr::GEMM && y.c.empty() && x.kind == MatExpr::VALUE) { MatExpr r(y); r.c = x.a; r.beta = x.alpha; return r; }
    if (x.kind == MatExpr::VALUE && y.kind == MatExpr::VALUE) return MatExpr(lincomb(x.a, x.alpha, y.a, y.alpha));
    return MatExpr(lincomb(x.eval(), 1, y.eval(), 1));
}
inline MatExpr operator+(const Mat& x, const Mat& y) { return MatExpr(x) + MatExpr(y); }
inline MatExpr operator+(const MatExpr& x, const Mat& y) { return x + MatExpr(y); }
inline MatExpr operator+(const Mat& x, const MatExpr& y) { return MatExpr(x) + y; }
inline MatExpr operator-(const MatExpr& x, const MatExpr& y) { return x + scaled(y, -1); }
inline MatExpr operator-(const Mat& x, const Mat& y) { return MatExpr(x) + scaled(MatExpr(y), -1); }
inline MatExpr operator-(const MatExpr& x, const Mat& y) { return x + scaled(MatExpr(y), -1); }
inline MatExpr operator-(const Mat& x, const MatExpr& y) { return MatExpr(x) + scaled(y, -1); }
inline MatExpr operator/(const MatExpr& x, double s) { return scaled(x, 1.0 / s); }      /* cv::MatExpr: A/s == A*(1/s) */
inline MatExpr operator/(const Mat& x, double s) { return scaled(MatExpr(x), 1.0 / s); }

inline double norm(const Mat& m)
{
    CV_Assert(m.depth() == CV_32F);
    double s = 0;                                       /* NORM_L2 on CV_32F: double accumulator, sqrt in double */
    for (int i = 0; i < m.rows; i++)
        for (int j = 0; j < m.cols * m.channels(); j++) { const double v = m.ptr<float>(i)[j]; s += v * v; }
    return std::sqrt(s);
}
inline double norm(const MatExpr& e) { return norm(e.eval()); }

class _InputArray {
public:
    _InputArray() : m(0) {}
    _InputArray(const Mat& m_) : m(&m_) {}
    _InputArray(const MatExpr& e) : held(e.eval()), m(&held) {}
    Mat getMat() const { return m ? *m : Mat(); }
    bool empty() const { return !m || m->empty(); }
private:
    Mat held;
    const Mat* m;
};
class _OutputArray {
public:
    _OutputArray(Mat& m_) : m(&m_) {}
    _OutputArray(const Mat& m_) : m(const_cast<Mat*>(&m_)) {}       /* a temporary view such as T.rowRange(0,3).col(3) */
    void create(int r, int c, int type) const { m->create(r, c, type); }
    void create(Size s, int type) const { m->create(s.height, s.width, type); }
    void release() const { m->release(); }
    Mat getMat() const { return *m; }
    Mat& getMatRef() const { return *m; }
private:
    Mat* m;
};
inline void Mat::copyTo(OutputArray dst) const
{
    dst.create(rows, cols, type());
    Mat d = dst.getMat();
    for (int y = 0; y < rows; y++) std::memmove(d.data + y * d.step, data + y * step, (size_t)cols * elemSize());
}
inline std::ostream& operator<<(std::ostream& os, const Mat& m)
{
    os << "[";
    for (int i = 0; i < m.rows; i++) {
        for (int j = 0; j < m.cols; j++) os << (m.depth() == CV_32F ? (double)m.at<float>(i, j) : (double)m.at<uchar>(i, j)) << (j + 1 < m.cols ? ", " : "");
        os << (i + 1 < m.rows ? ";\n" : "");
    }
    return os << "]";
}

/* ------------------------------------------------------------------ arithmetic primitives -> the pinned oracle restatements */
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4, BORDER_REFLECT101 = 4,
       BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2 };

inline void resize(InputArray src_, OutputArray dst_, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR)
{
    Mat src = src_.getMat();
    CV_Assert(src.type() == CV_8UC1 && interpolation == INTER_LINEAR && fx == 0 && fy == 0 && dsize.width > 0 && dsize.height > 0);
    dst_.create(dsize, src.type());
    Mat dst = dst_.getMat();
    orc_resize_linear_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
}

inline void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right, int borderType)
{
    Mat src = src_.getMat();
    CV_Assert(src.type() == CV_8UC1 && (borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101 && top == bottom && left == right && top == left);
    /* without BORDER_ISOLATED cv::copyMakeBorder would take real pixels around a sub-matrix: the reference's only non-isolated call
     * passes the caller's whole image (src/ORBextractor.cc:814) */
    CV_Assert((borderType & BORDER_ISOLATED) || !src.isSubmatrix());
    dst_.create(src.rows + top + bottom, src.cols + left + right, src.type());       /* keeps `temp` when it already has this size */
    Mat dst = dst_.getMat();
    uchar* inner = dst.data + (size_t)top * dst.step + left;
    if (inner != src.data)
        for (int y = 0; y < src.rows; y++) std::memmove(inner + (size_t)y * dst.step, src.data + (size_t)y * src.step, (size_t)src.cols);
    orc_border_reflect101(dst.data, src.cols, src.rows, (int)dst.step, top);
}

/* cv::GaussianBlur(7x7, sigma 2) on an 8-bit SUB-matrix that is not isolated: OpenCV 4.13 takes the FP32 separable path and reads the
 * real pixels around the ROI (variant ORC_BLUR_F32_SEPFILTER).  In place, as the reference calls it. */
inline void GaussianBlur(InputArray src_, OutputArray dst_, Size ksize, double sx, double sy = 0, int borderType = BORDER_DEFAULT)
{
    Mat src = src_.getMat();
    CV_Assert(src.type() == CV_8UC1 && ksize.width == 7 && ksize.height == 7 && sx == 2 && sy == 2 && borderType == BORDER_REFLECT_101);
    CV_Assert(src.data - src.datastart >= (ptrdiff_t)(3 * src.step + 3) && src.dataend - (src.data + (size_t)(src.rows - 1) * src.step + src.cols) >= (ptrdiff_t)(3 * src.step + 3));
    dst_.create(src.rows, src.cols, src.type());
    Mat dst = dst_.getMat();
    std::vector<uchar> tmp((size_t)src.rows * src.cols);
    orc_gaussian_blur7(src.data, src.cols, src.rows, (int)src.step, tmp.data(), src.cols, ORC_BLUR_F32_SEPFILTER);
    for (int y = 0; y < src.rows; y++) std::memcpy(dst.data + (size_t)y * dst.step, tmp.data() + (size_t)y * src.cols, (size_t)src.cols);
}

inline float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

/* cv::FAST(image, keypoints, threshold, nonmaxSuppression) — TYPE_9_16 */
inline void FAST(InputArray img_, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true)
{
    Mat img = img_.getMat();
    CV_Assert(img.type() == CV_8UC1 && nonmaxSuppression);
    keypoints.clear();
    if (img.cols < 7 || img.rows < 7) return;
    const int cap = ((img.cols + 1) / 2) * ((img.rows + 1) / 2) + 1;
    std::vector<int> x(cap), y(cap), s(cap);
    const int n = orc_fast9_nms(img.data, img.cols, img.rows, (int)img.step, threshold, cap, x.data(), y.data(), s.data());
    CV_Assert(n >= 0);
    for (int i = 0; i < n; i++) keypoints.push_back(KeyPoint((float)x[i], (float)y[i], 7.f, -1, (float)s[i]));
}

struct KeypointResponseGreater { bool operator()(const KeyPoint& a, const KeyPoint& b) const { return a.response > b.response; } };
struct KeypointResponseGreaterThanOrEqual {
    float value;
    explicit KeypointResponseGreaterThanOrEqual(float v) : value(v) {}
    bool operator()(const KeyPoint& k) const { return k.response >= value; }
};
class KeyPointsFilter {
public:
    /* OpenCV 4.x modules/features2d/src/keypoint.cpp: keep the n_points strongest, and every keypoint tied with the weakest of them */
    static void retainBest(std::vector<KeyPoint>& keypoints, int n_points)
    {
        if (n_points >= 0 && keypoints.size() > (size_t)n_points) {
            if (n_points == 0) { keypoints.clear(); return; }
            std::nth_element(keypoints.begin(), keypoints.begin() + n_points - 1, keypoints.end(), KeypointResponseGreater());
            const float ambiguous = keypoints[n_points - 1].response;
            std::vector<KeyPoint>::const_iterator new_end =
                std::partition(keypoints.begin() + n_points, keypoints.end(), KeypointResponseGreaterThanOrEqual(ambiguous));
            keypoints.resize(new_end - keypoints.begin());
        }
    }
};

/* cv::undistortPoints(src, dst, K, D, R = Mat(), P = K) on N x 1 CV_32FC2 */
inline void undistortPoints(InputArray src_, OutputArray dst_, InputArray K_, InputArray D_, InputArray R_ = _InputArray(), InputArray P_ = _InputArray())
{
    Mat src = src_.getMat(), K = K_.getMat(), D = D_.getMat(), P = P_.getMat();
    CV_Assert(src.type() == CV_32FC2 && src.isContinuous() && R_.empty() && K.type() == CV_32FC1 && D.type() == CV_32FC1);
    CV_Assert(!P.empty() && P.data == K.data);
    const int n = (int)src.total();
    std::vector<float> xy((size_t)n * 2), d(D.total());
    std::memcpy(xy.data(), src.data, xy.size() * 4);
    for (size_t i = 0; i < d.size(); i++) d[i] = D.at<float>((int)i);
    orc_undistort_points(xy.data(), n, K.at<float>(0, 0), K.at<float>(1, 1), K.at<float>(0, 2), K.at<float>(1, 2), d.data(), (int)d.size());
    dst_.create(src.rows, src.cols, src.type());
    std::memcpy(dst_.getMat().data, xy.data(), xy.size() * 4);
}

/* ---- declared because DBoW2's TemplatedVocabulary has virtual save/load(cv::FileStorage); the YAML persistence is outside the hot
 *      path (ORB-SLAM loads the text format, TemplatedVocabulary.h:1338-1425) and is not provided ---- */
class FileNode {
public:
    FileNode operator[](const char*) const { fail(); return FileNode(); }
    FileNode operator[](const std::string&) const { fail(); return FileNode(); }
    FileNode operator[](int) const { fail(); return FileNode(); }
    size_t size() const { fail(); return 0; }
    operator int() const { fail(); return 0; }
    operator double() const { fail(); return 0; }
    operator float() const { fail(); return 0; }
    operator std::string() const { fail(); return std::string(); }
    static void fail() { throw std::runtime_error("cv::FileStorage is not part of the reference shim"); }
};
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const std::string&, int) { FileNode::fail(); }
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const char*) const { FileNode::fail(); return FileNode(); }
    FileNode operator[](const std::string&) const { FileNode::fail(); return FileNode(); }
};
template <typename T> inline FileStorage& operator<<(FileStorage& fs, const T&) { FileNode::fail(); return fs; }

/* the reference names the enum through cv::ORB (src/ORBextractor.cc:616) */
class ORB { public: enum { kBytes = 32, HARRIS_SCORE = 0, FAST_SCORE = 1 }; };

} // namespace cv
#endif
