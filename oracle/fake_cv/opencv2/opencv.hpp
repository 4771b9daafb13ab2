// TEST INFRASTRUCTURE ONLY (oracle/): a minimal stand-in for <opencv2/opencv.hpp>.
//
// The reference helper class (cpp/include/sister/SisterMultiviewDisparities.hpp) is header-only
// and needs OpenCV core/imgproc, which this image does not ship as a C++ library. This file gives
// the *unmodified* reference header just enough of cv:: to compile and run, so the oracle is the
// reference's own code. Only the image plumbing below is ours; every op was checked against
// Python cv2 4.13 (tests/test_oracle_cvshim.py):
//   cvtColor(BGR2GRAY)   hpp:29-33    OpenCV-4 fixed point (3735 B + 19235 G + 9798 R + 16384) >> 15
//   copyMakeBorder       hpp:35-39    BORDER_REPLICATE
//   flip / transpose     hpp:57-70, 235-236, 250-251
//   convertTo(CV_16UC1)  hpp:111-113  round-half-even + saturate
//   Mat(Rect) * 255      hpp:116-118  saturating u16 multiply of a ROI
// Never include this from the product (sister_b200/); tests/ and oracle/ only.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <vector>

namespace cv {

typedef unsigned char uchar;

enum { CV_8UC1 = 0, CV_16UC1 = 2, CV_32FC1 = 5, CV_8UC3 = 16 };
enum { COLOR_BGR2GRAY = 6 };
enum { BORDER_REPLICATE = 1 };

inline int fakecv_elem_size(int type)
{
    switch (type) {
    case CV_8UC1: return 1;
    case CV_16UC1: return 2;
    case CV_32FC1: return 4;
    case CV_8UC3: return 3;
    }
    throw std::runtime_error("fake cv: unsupported Mat type");
}

struct Rect {
    int x, y, width, height;
    Rect(int x_, int y_, int w_, int h_) : x(x_), y(y_), width(w_), height(h_) {}
};

class Mat;
struct MatExpr;

// cv::MatStep: converts to size_t (bytes between rows), like the real one
struct MatStep {
    size_t v = 0;
    operator size_t() const { return v; }
    size_t operator[](int) const { return v; }
    MatStep &operator=(size_t s) { v = s; return *this; }
};

class Mat {
public:
    int rows = 0, cols = 0;
    uchar *data = nullptr;
    MatStep step;

    Mat() {}
    Mat(int r, int c, int t) { create(r, c, t); }
    Mat(const MatExpr &e);
    Mat &operator=(const MatExpr &e);

    void create(int r, int c, int t)
    {
        rows = r; cols = c; type_ = t;
        step = (size_t)c * fakecv_elem_size(t);
        // The reference's dead "consensus" loop (hpp:91-109) indexes (cols x rows)-shaped Mats with
        // (row<rows', col<cols') of the transposed shape; give every buffer max(r,c)^2 elements of
        // zeroed slack so that stays inside the allocation for portrait frames too.
        size_t m = (size_t)(r > c ? r : c);
        size_t bytes = m * m * fakecv_elem_size(t) + 64;
        void *p = nullptr;
        if (::posix_memalign(&p, 64, bytes) != 0) throw std::bad_alloc();
        std::memset(p, 0, bytes);
        store_ = std::shared_ptr<uchar>((uchar *)p, [](uchar *q) { std::free(q); });
        data = store_.get();
    }
    int type() const { return type_; }
    bool empty() const { return data == nullptr; }
    int channels() const { return type_ == CV_8UC3 ? 3 : 1; }
    bool isContinuous() const { return (size_t)step == (size_t)cols * fakecv_elem_size(type_); }

    template <typename T> T &at(int i, int j) { return *(T *)(data + (size_t)i * step.v + (size_t)j * sizeof(T)); }
    template <typename T> const T &at(int i, int j) const { return *(const T *)(data + (size_t)i * step.v + (size_t)j * sizeof(T)); }

    // ROI view sharing storage (hpp:116-118).
    Mat operator()(const Rect &r) const
    {
        Mat m;
        m.rows = r.height; m.cols = r.width; m.type_ = type_; m.step = step; m.store_ = store_;
        m.data = data + (size_t)r.y * step.v + (size_t)r.x * fakecv_elem_size(type_);
        return m;
    }

    // Only float -> u16 is used (hpp:111-113); in-place allowed.
    void convertTo(Mat &dst, int rtype) const
    {
        if (type_ != CV_32FC1 || rtype != CV_16UC1) throw std::runtime_error("fake cv: convertTo combo");
        Mat out(rows, cols, rtype);
        for (int i = 0; i < rows; i++)
            for (int j = 0; j < cols; j++) {
                float f = at<float>(i, j);
                long v = std::lrintf(f); // round-half-even under the default rounding mode (cvRound)
                if (v < 0) v = 0;
                if (v > 65535) v = 65535;
                out.at<uint16_t>(i, j) = (uint16_t)v;
            }
        dst = out;
    }

private:
    int type_ = 0;
    std::shared_ptr<uchar> store_;
};

struct MatExpr {
    Mat src;
    double scale;
};

inline MatExpr operator*(const Mat &m, double s) { return MatExpr{m, s}; }

inline void fakecv_eval(const MatExpr &e, Mat &dst)
{
    if (e.src.type() != CV_16UC1) throw std::runtime_error("fake cv: MatExpr type");
    Mat out(e.src.rows, e.src.cols, CV_16UC1);
    for (int i = 0; i < e.src.rows; i++)
        for (int j = 0; j < e.src.cols; j++) {
            double v = (double)e.src.at<uint16_t>(i, j) * e.scale;
            long r = std::lrint(v);
            if (r < 0) r = 0;
            if (r > 65535) r = 65535;
            out.at<uint16_t>(i, j) = (uint16_t)r;
        }
    dst = out;
}
inline Mat::Mat(const MatExpr &e) { Mat t; fakecv_eval(e, t); *this = t; }
inline Mat &Mat::operator=(const MatExpr &e) { Mat t; fakecv_eval(e, t); *this = t; return *this; }

inline void cvtColor(const Mat &src, Mat &dst, int code)
{
    if (code != COLOR_BGR2GRAY || src.type() != CV_8UC3) throw std::runtime_error("fake cv: cvtColor combo");
    if (src.empty()) throw std::runtime_error("fake cv: empty input");
    Mat out(src.rows, src.cols, CV_8UC1);
    for (int i = 0; i < src.rows; i++) {
        const uchar *p = src.data + (size_t)i * (size_t)src.step;
        for (int j = 0; j < src.cols; j++) {
            int b = p[3 * j], g = p[3 * j + 1], r = p[3 * j + 2];
            out.at<uchar>(i, j) = (uchar)((3735 * b + 19235 * g + 9798 * r + 16384) >> 15);
        }
    }
    dst = out;
}

inline void copyMakeBorder(const Mat &src, Mat &dst, int top, int bottom, int left, int right, int borderType, int = 0)
{
    if (borderType != BORDER_REPLICATE || src.type() != CV_8UC1) throw std::runtime_error("fake cv: copyMakeBorder combo");
    Mat out(src.rows + top + bottom, src.cols + left + right, CV_8UC1);
    for (int i = 0; i < out.rows; i++) {
        int si = i - top; si = si < 0 ? 0 : (si >= src.rows ? src.rows - 1 : si);
        for (int j = 0; j < out.cols; j++) {
            int sj = j - left; sj = sj < 0 ? 0 : (sj >= src.cols ? src.cols - 1 : sj);
            out.at<uchar>(i, j) = src.at<uchar>(si, sj);
        }
    }
    dst = out;
}

// flipCode 0: around x-axis (rows reversed); >0: around y-axis (cols reversed); <0: both.
inline void flip(const Mat &src, Mat &dst, int flipCode)
{
    if (src.type() != CV_8UC1) throw std::runtime_error("fake cv: flip type");
    Mat out(src.rows, src.cols, CV_8UC1);
    for (int i = 0; i < src.rows; i++)
        for (int j = 0; j < src.cols; j++) {
            int si = (flipCode <= 0) ? src.rows - 1 - i : i;
            int sj = (flipCode != 0) ? src.cols - 1 - j : j;
            out.at<uchar>(i, j) = src.at<uchar>(si, sj);
        }
    dst = out;
}

inline void transpose(const Mat &src, Mat &dst)
{
    if (src.type() != CV_8UC1) throw std::runtime_error("fake cv: transpose type");
    Mat out(src.cols, src.rows, CV_8UC1);
    for (int i = 0; i < src.rows; i++)
        for (int j = 0; j < src.cols; j++)
            out.at<uchar>(j, i) = src.at<uchar>(i, j);
    dst = out;
}

} // namespace cv
