// Test driver for include/sister/SisterMultiviewDisparities.hpp (the C++ drop-in of the reference helper class).
// It is written the way the reference's own caller is (cpp/src/compute_disp.cpp:19-35): build five BGR cv::Mat,
// construct the class, call compute_disparities; cv::Mat comes from the test-only stand-in oracle/fake_cv.
//   shim_driver <in.bin> <out.bin> <w> <h> <dispCount>
// in.bin: 5 views, each h*w*3 bytes (BGR). out.bin: 3 maps, each h*w uint16. Exit code 3 + message on exception.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <sister/SisterMultiviewDisparities.hpp>

int main(int argc, char **argv)
{
    if (argc != 6) { std::fprintf(stderr, "usage: shim_driver in.bin out.bin w h dispCount\n"); return 2; }
    const int w = std::atoi(argv[3]), h = std::atoi(argv[4]), D = std::atoi(argv[5]);
    std::vector<cv::Mat> v(5);
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    for (int k = 0; k < 5; k++) {
        v[k].create(h, w, CV_8UC3);
        for (int i = 0; i < h; i++)
            if (std::fread(v[k].data + (size_t)i * (size_t)v[k].step, 1, (size_t)w * 3, f) != (size_t)w * 3) return 2;
    }
    std::fclose(f);
    try {
        SisterMultiviewDisparities multiview(v[0], v[1], v[2], v[3], v[4]);
        cv::Mat result_horizontal, result_vertical, result_multiview;
        multiview.compute_disparities(D, result_multiview, result_horizontal, result_vertical);
        multiview.compute_disparities(D, result_multiview, result_horizontal, result_vertical); // context is reused
        FILE *o = std::fopen(argv[2], "wb");
        if (!o) return 2;
        const cv::Mat *outs[3] = {&result_multiview, &result_horizontal, &result_vertical};
        for (int m = 0; m < 3; m++) {
            if (outs[m]->rows != h || outs[m]->cols != w || outs[m]->type() != CV_16UC1) return 4;
            for (int i = 0; i < h; i++) std::fwrite(outs[m]->data + (size_t)i * (size_t)outs[m]->step, 2, (size_t)w, o);
        }
        std::fclose(o);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 3;
    }
    return 0;
}
