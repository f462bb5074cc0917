// TEST INFRASTRUCTURE ONLY.  Compiles the reference's lab8 right-hand-side producers UNMODIFIED:
//   labs/lab8/src/OpenCVHW1/hw8_pa.cc:316-498   GradientAt, ZeroGradientAt, MergeImage2<T>, MergeImage<T,channel>,
//                                               MaskImage, EnforceGradientBound          (file scope)
//   labs/lab8/src/OpenCVHW1/hw8_pa.cc:602-676   struct Gradients, both constructors     (local to stitchImages)
// oracle/Makefile cuts those two line ranges out of /root/reference at build time into oracle/_ref/ (a build artefact,
// git-ignored; nothing of the reference is stored in this repository) and this file includes them against a minimal
// stand-in for the few cv:: names they use (cv::Mat::{rows, cols, ptr, at, clone, size}, cv::Vec<T,3>, Scalar).
// OpenCV itself is not available here; the stand-in implements exactly the documented semantics of those members for
// continuous 8UC1 / 8UC3 / 32FC3 images.
//
// Guard rows: several of these functions walk pointers past the end of a row -- and EnforceGradientBound touches rows
// i-1 / i+1 of the first / last image row -- which on a real cv::Mat is undefined behaviour.  Every stand-in Mat
// therefore owns GUARD zero rows before and after its pixels (and one non-zero sentinel byte at the very end, so
// that an unbounded "skip zeros" walk terminates); the exported wrappers copy only the real rows in and out.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <memory>
#include <vector>

typedef unsigned char uchar;
using std::abs;
using std::max;
using std::min;

namespace cv {
template <typename T, int N>
struct Vec {
    T v[N];
    Vec() { for (int i = 0; i < N; ++i) v[i] = T(0); }
    template <typename U>
    Vec(const Vec<U, N> &o) { for (int i = 0; i < N; ++i) v[i] = static_cast<T>(o.v[i]); } // saturate_cast is exact here
    T &operator[](int i) { return v[i]; }
    const T &operator[](int i) const { return v[i]; }
};
template <typename T, int N>
Vec<T, N> operator-(const Vec<T, N> &a, const Vec<T, N> &b) {
    Vec<T, N> r;
    for (int i = 0; i < N; ++i) r.v[i] = a.v[i] - b.v[i];
    return r;
}
typedef Vec<uchar, 3> Vec3b;
typedef Vec<int, 3> Vec3i;
typedef Vec<float, 3> Vec3f;
struct Scalar {
    double v[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : v{a, b, c, d} {}
};
struct Size {
    int width, height;
};
enum { CV_8UC1 = 0, CV_8UC3 = 16, CV_32FC3 = 21 };
static inline int elem_bytes(int type) { return type == CV_8UC1 ? 1 : type == CV_8UC3 ? 3 : 12; }

struct Mat {
    static const int GUARD = 4;
    int rows = 0, cols = 0, type_ = CV_8UC1;
    size_t step = 0;
    std::shared_ptr<std::vector<uchar>> buf; // shared on assignment / copy, as cv::Mat headers are
    uchar *data = nullptr;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, const Scalar &) { create(r, c, type); } // only ever Scalar(0,0,0) in the cut ranges
    Mat(std::initializer_list<int>, std::initializer_list<float>) {}      // the unused 2-tap kernels of Gradients
    void create(int r, int c, int type) {
        rows = r;
        cols = c;
        type_ = type;
        step = (size_t)c * elem_bytes(type);
        buf = std::make_shared<std::vector<uchar>>((size_t)(r + 2 * GUARD) * step + 1, (uchar)0);
        buf->back() = 1;
        data = buf->data() + (size_t)GUARD * step;
    }
    uchar *ptr(int i = 0) { return data + (size_t)i * step; }
    const uchar *ptr(int i = 0) const { return data + (size_t)i * step; }
    template <typename T>
    T &at(int y, int x) { return *reinterpret_cast<T *>(data + (ptrdiff_t)y * (ptrdiff_t)step + (ptrdiff_t)x * (ptrdiff_t)sizeof(T)); }
    template <typename T>
    const T &at(int y, int x) const {
        return *reinterpret_cast<const T *>(data + (ptrdiff_t)y * (ptrdiff_t)step + (ptrdiff_t)x * (ptrdiff_t)sizeof(T));
    }
    Mat clone() const {
        Mat m;
        m.rows = rows;
        m.cols = cols;
        m.type_ = type_;
        m.step = step;
        m.buf = std::make_shared<std::vector<uchar>>(*buf);
        m.data = m.buf->data() + (data - buf->data());
        return m;
    }
    Size size() const { return Size{cols, rows}; }
    int type() const { return type_; }
};
} // namespace cv

// ---- the reference, verbatim, file scope --------------------------------------------------------------------------
#include REF_PANO_FILE_SCOPE

// ---- the reference's local struct Gradients, verbatim, inside a function as upstream -------------------------------
static void ref_gradients(const cv::Mat &m, const cv::Mat *mask, cv::Mat &gx, cv::Mat &gy) {
    using namespace cv;
#include REF_PANO_LOCAL_SCOPE
    if (mask) {
        Gradients g(m, *mask);
        gx = g.x;
        gy = g.y;
    } else {
        Gradients g(m);
        gx = g.x;
        gy = g.y;
    }
}

// ---- flat-buffer wrappers --------------------------------------------------------------------------------------
static cv::Mat wrap(const void *src, int W, int H, int type) {
    cv::Mat m(H, W, type);
    if (src) std::memcpy(m.data, src, (size_t)H * m.step);
    return m;
}
static void unwrap(const cv::Mat &m, void *dst) { std::memcpy(dst, m.data, (size_t)m.rows * m.step); }

extern "C" {
void ref_pano_mask_image(const uchar *src, const uchar *mask, int W, int H, uchar *out) {
    cv::Mat s = wrap(src, W, H, cv::CV_8UC3), m = wrap(mask, W, H, cv::CV_8UC1);
    unwrap(MaskImage(s, m), out);
}
// mask == NULL: first constructor (:604-636); else the mask-driven second one (:638-676).  The reference leaves the
// pixels it does not visit uninitialised in the first constructor; the stand-in's Mat is zero-filled.
void ref_pano_gradients(const uchar *img, const uchar *mask, int W, int H, float *gx, float *gy) {
    cv::Mat m = wrap(img, W, H, cv::CV_8UC3), mk, x, y;
    if (mask) mk = wrap(mask, W, H, cv::CV_8UC1);
    ref_gradients(m, mask ? &mk : nullptr, x, y);
    unwrap(x, gx);
    unwrap(y, gy);
}
void ref_pano_merge2_f32(float *target, const float *src, const uchar *target_mask, const uchar *src_outer_mask,
                         const uchar *src_inner_mask, int W, int H) {
    cv::Mat t = wrap(target, W, H, cv::CV_32FC3), s = wrap(src, W, H, cv::CV_32FC3), tm = wrap(target_mask, W, H, cv::CV_8UC1),
            so = wrap(src_outer_mask, W, H, cv::CV_8UC1), si = wrap(src_inner_mask, W, H, cv::CV_8UC1);
    MergeImage2<float>(t, s, tm, so, si, 1);
    unwrap(t, target);
}
void ref_pano_merge_u8(uchar *target, const uchar *src, const uchar *target_mask, const uchar *src_mask, int channel,
                       double skip_how_many, int W, int H) {
    const int type = channel == 3 ? cv::CV_8UC3 : cv::CV_8UC1;
    cv::Mat t = wrap(target, W, H, type), s = wrap(src, W, H, type), tm = wrap(target_mask, W, H, cv::CV_8UC1),
            sm = wrap(src_mask, W, H, cv::CV_8UC1);
    if (channel == 3)
        MergeImage<uchar, 3>(t, s, tm, sm, skip_how_many);
    else
        MergeImage<uchar, 1>(t, s, tm, sm, skip_how_many);
    unwrap(t, target);
}
void ref_pano_enforce_gradient_bound(float *dx, float *dy, const uchar *src, const uchar *mask, int W, int H) {
    cv::Mat x = wrap(dx, W, H, cv::CV_32FC3), y = wrap(dy, W, H, cv::CV_32FC3);
    EnforceGradientBound(x, y, wrap(src, W, H, cv::CV_8UC3), wrap(mask, W, H, cv::CV_8UC1));
    unwrap(x, dx);
    unwrap(y, dy);
}
}
