// Drop-in replacement for the reference helper class
//     cpp/include/sister/SisterMultiviewDisparities.hpp:18-26  (CVLAB-Unibo/sister)
// backed by the B200 (sm_100a) path behind the C ABI of sister_b200.h. Same class name, same constructor
// (center, right, top, left, bottom -- hpp:22), same method
//     void compute_disparities(int dispCount, cv::Mat &disp_multiview, cv::Mat &disp_horizontal, cv::Mat &disp_vertical)
// (hpp:26), same outputs: three freshly allocated CV_16UC1 H x W Mats holding saturate_u16(disparity * 255)
// (hpp:111-118). A caller such as cpp/src/compute_disp.cpp:26-35 or the ROS callback of ros/README.md:47-56 switches
// by putting this directory before the reference's on the include path and linking libsister_b200.so instead of
// libsister.a.
//
// Differences, all on the error / side-effect side:
//   * nothing is printed (the reference prints three clock() lines per call, hpp:79,84,89);
//   * violated preconditions (dispCount % 8, padded size % 4 -- sgm.cpp:268, postprocess.cpp:18; the reference
//     segfaults or asserts) and CUDA failures throw std::runtime_error carrying sister_last_error();
//   * dispCount up to 512 and volumes beyond 2^31 cells work (the reference overflows, postprocess.cpp:193, types.h:31-34);
//   * CV_8UC1 inputs are accepted as already-grey views (the reference requires BGR, hpp:29-33);
//   * the device context (streams, scratch volumes) is created on first use and kept for the object's lifetime.
// There is no CPU fallback: without an sm_100 device compute_disparities throws.
#pragma once
#include <memory>
#include <stdexcept>
#include <string>

#include <opencv2/opencv.hpp>

#include "../sister_b200.h"

// The reference header pulls both namespaces into every includer (hpp:15-16) and its callers rely on it
// (compute_disp.cpp:19 uses `string`); a drop-in has to do the same.
using namespace cv;
using namespace std;

class SisterMultiviewDisparities
{
public:
    SisterMultiviewDisparities(cv::Mat center, cv::Mat right, cv::Mat top, cv::Mat left, cv::Mat bottom, int cuda_device = 0)
        : center(center), right(right), top(top), left(left), bottom(bottom), device_(cuda_device)
    {
    }

    void compute_disparities(int dispCount, cv::Mat &disp_multiview, cv::Mat &disp_horizontal, cv::Mat &disp_vertical)
    {
        const cv::Mat *views[5] = {&center, &right, &top, &left, &bottom};
        const int w = center.cols, h = center.rows, type = center.type();
        if (center.empty()) throw std::runtime_error("SisterMultiviewDisparities: empty input view");
        if (type != CV_8UC3 && type != CV_8UC1) throw std::runtime_error("SisterMultiviewDisparities: views must be CV_8UC3 (BGR) or CV_8UC1");
        const uint8_t *ptrs[5];
        for (int k = 0; k < 5; k++) {
            if (views[k]->empty() || views[k]->cols != w || views[k]->rows != h || views[k]->type() != type ||
                (size_t)views[k]->step != (size_t)center.step)
                throw std::runtime_error("SisterMultiviewDisparities: the five views must share size, type and row stride");
            ptrs[k] = views[k]->data;
        }
        ensure_context(w, h, dispCount);
        disp_multiview.create(h, w, CV_16UC1);
        disp_horizontal.create(h, w, CV_16UC1);
        disp_vertical.create(h, w, CV_16UC1);
        uint16_t *out[3] = {(uint16_t *)disp_multiview.data, (uint16_t *)disp_horizontal.data, (uint16_t *)disp_vertical.data};
        const int rc = sister_compute(ctx_.get(), ptrs, w, h, type == CV_8UC3 ? 3 : 1, (size_t)center.step, dispCount,
                                      SISTER_MODE_ALL, out, nullptr);
        if (rc != SISTER_OK)
            throw std::runtime_error(std::string("sister_b200: ") + sister_strerror(rc) + ": " + sister_last_error(ctx_.get()));
    }

protected:
    cv::Mat center, right, top, left, bottom;

private:
    void ensure_context(int w, int h, int d)
    {
        if (ctx_ && w <= cap_w_ && h <= cap_h_ && d <= cap_d_) return;
        ctx_.reset();
        sister_ctx *c = nullptr;
        const int rc = sister_create(&c, device_, w, h, d, 1);
        if (rc != SISTER_OK) throw std::runtime_error(std::string("sister_b200: sister_create: ") + sister_strerror(rc));
        ctx_ = std::shared_ptr<sister_ctx>(c, [](sister_ctx *p) { sister_destroy(p); });
        cap_w_ = w; cap_h_ = h; cap_d_ = d;
    }

    int device_ = 0, cap_w_ = 0, cap_h_ = 0, cap_d_ = 0;
    std::shared_ptr<sister_ctx> ctx_;
};
