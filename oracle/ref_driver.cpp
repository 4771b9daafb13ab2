// TEST INFRASTRUCTURE ONLY (oracle/). C entry points around the UNMODIFIED reference sources.
//
// Built by oracle/Makefile into oracle/_ref/libsister_ref.so from
//   $(SISTER_REF)/cpp/src/sister/{census,sgm,postprocess}.cpp        (compiled where they lie)
//   $(SISTER_REF)/cpp/include/sister/SisterMultiviewDisparities.hpp  (included below, unmodified)
// against oracle/fake_cv/opencv2/opencv.hpp. No reference source is copied into this repo.
//
// Everything is compiled with -Dposix_memalign=oracle_zero_memalign so that _mm_malloc hands out
// ZERO-FILLED memory: the reference never writes rows h-2,h-1 of a raw cost volume
// (census.cpp:142-145 repeats the first-rows loop), and SURVEY.md §8(c) defines them as 0.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load the resulting library.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <vector>
#include <unistd.h>

#include "opencv2/opencv.hpp"
#include "SisterMultiviewDisparities.hpp" // the reference helper class, as is

#undef posix_memalign
extern "C" int posix_memalign(void **, size_t, size_t) noexcept;

extern "C" int oracle_zero_memalign(void **p, size_t align, size_t size) noexcept
{
    int rc = posix_memalign(p, align, size);
    if (rc == 0) std::memset(*p, 0, size);
    return rc;
}

// census.cpp defines these with external linkage but stereoalgo.h does not declare them.
void censusTransform(uint8 *source, uint64 *dest, uint32 width, uint32 height, uint8 rw, uint8 rh);
void hammingCost(uint64 *im1_data, uint64 *im2_data, int height, int width, int dispCount, uint16 *dsi, sint32 numThreads);

namespace {

struct AlignedBuf {
    void *p = nullptr;
    explicit AlignedBuf(size_t bytes) { if (oracle_zero_memalign(&p, 64, bytes ? bytes : 64) != 0) p = nullptr; }
    ~AlignedBuf() { std::free(p); }
    template <typename T> T *as() { return (T *)p; }
};

// Silence the reference's std::cout timing lines (hpp:79,84,89) while a call runs.
struct StdoutMute {
    int saved = -1;
    explicit StdoutMute(bool on)
    {
        if (!on) return;
        std::fflush(stdout);
        std::cout.flush();
        saved = dup(1);
        FILE *devnull = std::fopen("/dev/null", "w");
        if (devnull) { dup2(fileno(devnull), 1); std::fclose(devnull); }
    }
    ~StdoutMute()
    {
        if (saved < 0) return;
        std::cout.flush();
        std::fflush(stdout);
        dup2(saved, 1);
        close(saved);
    }
};

cv::Mat wrap_bgr(const uint8_t *p, int w, int h)
{
    cv::Mat m(h, w, cv::CV_8UC3);
    for (int i = 0; i < h; i++) std::memcpy(m.data + (size_t)i * (size_t)m.step, p + (size_t)i * w * 3, (size_t)w * 3);
    return m;
}

} // namespace

extern "C" {

// The public API of the reference, end to end: hpp:22 (ctor) + hpp:26 (compute_disparities).
// views: center, right, top, left, bottom; each H x W x 3 (BGR, packed). Outputs H x W uint16.
int ref_compute_disparities(const uint8_t *const *views_bgr, int w, int h, int disp_count,
                            uint16_t *out_mv, uint16_t *out_h, uint16_t *out_v, int quiet)
{
    try {
        StdoutMute mute(quiet != 0);
        SisterMultiviewDisparities s(wrap_bgr(views_bgr[0], w, h), wrap_bgr(views_bgr[1], w, h),
                                     wrap_bgr(views_bgr[2], w, h), wrap_bgr(views_bgr[3], w, h),
                                     wrap_bgr(views_bgr[4], w, h));
        cv::Mat mv, hz, vt;
        s.compute_disparities(disp_count, mv, hz, vt);
        if (mv.rows != h || mv.cols != w) return -2;
        for (int i = 0; i < h; i++) {
            std::memcpy(out_mv + (size_t)i * w, mv.data + (size_t)i * (size_t)mv.step, (size_t)w * 2);
            std::memcpy(out_h + (size_t)i * w, hz.data + (size_t)i * (size_t)hz.step, (size_t)w * 2);
            std::memcpy(out_v + (size_t)i * w, vt.data + (size_t)i * (size_t)vt.step, (size_t)w * 2);
        }
        return 0;
    } catch (...) {
        return -1;
    }
}

// ---- stage taps: the L2 free functions, called directly (stereoalgo.h:4-17, census.cpp) ----

// census.cpp:38-51 with the literals of census.cpp:154 (rw=4, rh=3).
void ref_census(const uint8_t *img, int w, int h, uint64_t *out)
{
    AlignedBuf in((size_t)w * h), o((size_t)w * h * 8);
    std::memcpy(in.p, img, (size_t)w * h);
    censusTransform(in.as<uint8>(), o.as<uint64>(), (uint32)w, (uint32)h, 4, 3);
    std::memcpy(out, o.p, (size_t)w * h * 8);
}

// census.cpp:149-158. dsi: h*w*D uint16 (caller-allocated, any alignment).
void ref_ad_census(const uint8_t *im1, const uint8_t *im2, int h, int w, int D, uint16_t *dsi)
{
    size_t cells = (size_t)w * h * D;
    AlignedBuf a((size_t)w * h), b((size_t)w * h), v(cells * 2);
    std::memcpy(a.p, im1, (size_t)w * h);
    std::memcpy(b.p, im2, (size_t)w * h);
    ad_census(a.as<uint8>(), b.as<uint8>(), h, w, D, v.as<uint16>(), 4);
    std::memcpy(dsi, v.p, cells * 2);
}

// postprocess.cpp:74 / :187 with uniqueness = 1 (hpp:196-197).
void ref_wta(const uint16_t *dsi, int w, int h, int D, float *outL, float *outR)
{
    size_t cells = (size_t)w * h * D;
    AlignedBuf v(cells * 2), l((size_t)w * h * 4), r((size_t)w * h * 4);
    std::memcpy(v.p, dsi, cells * 2);
    uint16 *pv = v.as<uint16>();
    if (outL) { WTALeft_SSE(l.as<float>(), pv, w, h, D - 1, 1); std::memcpy(outL, l.p, (size_t)w * h * 4); }
    if (outR) { WTARight_SSE(r.as<float>(), pv, w, h, D - 1, 1); std::memcpy(outR, r.p, (size_t)w * h * 4); }
}

// postprocess.cpp:15 called in place as at hpp:198.
void ref_median_inplace(float *img, int w, int h)
{
    AlignedBuf m((size_t)w * h * 4);
    std::memcpy(m.p, img, (size_t)w * h * 4);
    median3x3_SSE(m.as<float>(), m.as<float>(), (uint32)w, (uint32)h);
    std::memcpy(img, m.p, (size_t)w * h * 4);
}

// postprocess.cpp:318, threshold as given (hpp:200 uses 5).
void ref_lrcheck(float *L, const float *R, int w, int h, int thr)
{
    std::vector<float> r(R, R + (size_t)w * h);
    doLRCheck(L, r.data(), w, h, thr);
}

// sgm.cpp:457 with the literals of hpp:280 (P1=7, P2min=17, 8 paths).
void ref_sgm(const uint8_t *img, int h, int w, int D, const uint16_t *dsi, uint16_t *sum)
{
    size_t cells = (size_t)w * h * D;
    AlignedBuf c(cells * 2), s(cells * 2), im((size_t)w * (h + 2) + 64);
    std::memcpy(c.p, dsi, cells * 2);
    // sgm.cpp:227-232 reads one pixel before/after the image for adaptP2 (alpha = 0: value unused).
    std::memcpy(im.as<uint8>() + w, img, (size_t)w * h);
    sgm(im.as<uint8>() + w, h, w, D, c.as<uint16>(), s.as<uint16>(), 7, 17, 8);
    std::memcpy(sum, s.p, cells * 2);
}

// One doMultiStereo (hpp:152-295) on already padded grey frames, mirrored over plain arrays so the
// intermediate products can be tapped. The arithmetic is all reference code (the L2 functions);
// only the loops of hpp:201-277 are restated here, and tests/test_oracle_ref.py checks this
// function's disparity against the real header (ref_compute_disparities) on every golden rig.
//   views: center, right, top, left, bottom, each hp x wp uint8 (already grey + padded)
//   masks: 4 x hp x wp uint8 in the image frame, order right(0), left(180), top(90), bottom(270)
//   fused/sum: hp x wp x D uint16 (may be null); disp: hp x wp float
int ref_multistereo_taps(const uint8_t *const *views, int wp, int hp, int D, int mode,
                         uint8_t *masks, uint16_t *fused, uint16_t *sum, float *disp,
                         float *rawL /*4 maps, view frames, after median+LRC; may be null*/)
{
    const int w = wp, h = hp;
    const size_t px = (size_t)w * h, cells = px * D;
    auto idx = [&](int i, int j) { return (size_t)i * w + j; };
    // re-oriented copies, hpp:56-70
    std::vector<uint8_t> C(views[0], views[0] + px), V0(views[1], views[1] + px), C180(px), V180(px), C90(px), V90(px), C270(px), V270(px);
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            C180[idx(i, j)] = views[0][idx(i, w - 1 - j)];
            V180[idx(i, j)] = views[3][idx(i, w - 1 - j)];
        }
    // transpose then flip(-1): T90(r,c) = X(h-1-c, w-1-r); transpose then flip(0): T270(r,c) = X(c, w-1-r)
    for (int r = 0; r < w; r++)
        for (int c = 0; c < h; c++) {
            C90[(size_t)r * h + c] = views[0][idx(h - 1 - c, w - 1 - r)];
            V90[(size_t)r * h + c] = views[2][idx(h - 1 - c, w - 1 - r)];
            C270[(size_t)r * h + c] = views[0][idx(c, w - 1 - r)];
            V270[(size_t)r * h + c] = views[4][idx(c, w - 1 - r)];
        }
    AlignedBuf d0(cells * 2), d90f(cells * 2), d180f(cells * 2), d270f(cells * 2), multi(cells * 2), sumv(cells * 2);
    AlignedBuf bufL(px * 4), bufR(px * 4), im(px + 2 * (size_t)w + 64);
    if (!d0.p || !d90f.p || !d180f.p || !d270f.p || !multi.p || !sumv.p) return -3;
    uint16 *p0 = d0.as<uint16>(), *p90 = d90f.as<uint16>(), *p180 = d180f.as<uint16>(), *p270 = d270f.as<uint16>();
    auto census = [&](std::vector<uint8_t> &a, std::vector<uint8_t> &b, int hh, int ww, uint16 *dst) {
        AlignedBuf x(px), y(px);
        std::memcpy(x.p, a.data(), px);
        std::memcpy(y.p, b.data(), px);
        ad_census(x.as<uint8>(), y.as<uint8>(), hh, ww, D, dst, 4);
    };
    census(C, V0, h, w, p0);       // hpp:181
    census(C90, V90, w, h, p90);   // hpp:182
    census(C180, V180, h, w, p180); // hpp:183
    census(C270, V270, w, h, p270); // hpp:184

    std::vector<uint8_t> m0(px, 0), m180(px, 0), m90(px, 0), m270(px, 0);
    float *L = bufL.as<float>(), *R = bufR.as<float>();
    auto lrc = [&](uint16 *&vol, int ww, int hh) { // hpp:196-200
        WTALeft_SSE(L, vol, ww, hh, D - 1, 1);
        WTARight_SSE(R, vol, ww, hh, D - 1, 1);
        median3x3_SSE(L, L, ww, hh);
        median3x3_SSE(R, R, ww, hh);
        doLRCheck(L, R, ww, hh, 5);
    };
    if (mode != 2) {
        lrc(p0, w, h);
        if (rawL) std::memcpy(rawL + 0 * px, L, px * 4);
        for (int i = 0; i < h; i++)
            for (int j = 0; j < w; j++) m0[idx(i, j)] = !(L[idx(i, j)] <= 0 || j < D); // hpp:201-206
        lrc(p180, w, h);
        if (rawL) std::memcpy(rawL + 1 * px, L, px * 4);
        for (int i = 0; i < h; i++)
            for (int j = 0; j < w; j++) m180[idx(i, w - 1 - j)] = !(L[idx(i, j)] <= 0 || j < D); // hpp:213-218
    }
    if (mode != 1) {
        lrc(p90, h, w);
        if (rawL) std::memcpy(rawL + 2 * px, L, px * 4);
        // hpp:229-236: mask in the (w x h) frame, then transpose + flip(-1) back: M(i,j) = m(w-1-j, h-1-i)
        for (int i = 0; i < h; i++)
            for (int j = 0; j < w; j++) {
                int r = w - 1 - j, c = h - 1 - i;
                m90[idx(i, j)] = !(L[(size_t)r * h + c] <= 0 || c < D);
            }
        lrc(p270, h, w);
        if (rawL) std::memcpy(rawL + 3 * px, L, px * 4);
        // hpp:244-251: transpose + flip(1): M(i,j) = m(w-1-j, i)
        for (int i = 0; i < h; i++)
            for (int j = 0; j < w; j++) {
                int r = w - 1 - j, c = i;
                m270[idx(i, j)] = !(L[(size_t)r * h + c] <= 0 || c < D);
            }
    }
    uint16 *pm = multi.as<uint16>();
    for (int i = 0; i < h; i++) // hpp:255-277
        for (int j = 0; j < w; j++) {
            const uint16 *a0 = p0 + ((size_t)i * w + j) * D;
            const uint16 *a180 = p180 + ((size_t)i * w + (w - 1 - j)) * D;
            const uint16 *a90 = p90 + ((size_t)(w - 1 - j) * h + (h - 1 - i)) * D;
            const uint16 *a270 = p270 + ((size_t)(w - 1 - j) * h + i) * D;
            uint16 *o = pm + ((size_t)i * w + j) * D;
            uint8_t k0 = m0[idx(i, j)], k180 = m180[idx(i, j)], k90 = m90[idx(i, j)], k270 = m270[idx(i, j)];
            for (int d = 0; d < D; d++) {
                switch (mode) {
                case 0: o[d] = (uint16)(k0 * a0[d] + k180 * a180[d] + k90 * a90[d] + k270 * a270[d]); break;
                case 1: o[d] = (uint16)(k0 * a0[d] + k180 * a180[d]); break;
                default: o[d] = (uint16)(k90 * a90[d] + k270 * a270[d]); break;
                }
            }
        }
    std::memcpy(im.as<uint8>() + w, views[0], px);
    uint16 *ps = sumv.as<uint16>();
    sgm(im.as<uint8>() + w, h, w, D, pm, ps, 7, 17, 8); // hpp:280
    WTALeft_SSE(L, ps, w, h, D - 1, 1);                  // hpp:283
    if (disp) std::memcpy(disp, L, px * 4);
    if (masks) {
        std::memcpy(masks + 0 * px, m0.data(), px);
        std::memcpy(masks + 1 * px, m180.data(), px);
        std::memcpy(masks + 2 * px, m90.data(), px);
        std::memcpy(masks + 3 * px, m270.data(), px);
    }
    if (fused) std::memcpy(fused, pm, cells * 2);
    if (sum) std::memcpy(sum, ps, cells * 2);
    return 0;
}

} // extern "C"
