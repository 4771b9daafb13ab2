// compute_disp -- command-line front end of the B200 path, the counterpart of the reference's sample executable
// cpp/src/compute_disp.cpp (CVLAB-Unibo/sister).
//
//   compute_disp <input folder/> <dmax> [output folder/]
//
// Like the reference (compute_disp.cpp:16-23) it reads <folder>center / right / top / left / bottom (the folder string is
// concatenated as is, so it must end in '/', compute_disp.cpp:19), runs compute_disparities (compute_disp.cpp:26-35) and
// derives the same pictures: each map min-max normalised to 8 bits (cv::normalize NORM_MINMAX, compute_disp.cpp:38-40),
// MAGMA colour map (41-43), and the 0.1 / 0.9 blend of the centre view with the coloured multiview map (46-48).
// Where the reference opens windows (51-56) this tool writes files:
//   disp_multiview.pgm  disp_horizontal.pgm  disp_vertical.pgm      16-bit P5, the CV_16UC1 maps (disparity * 255)
//   disp_multiview.ppm  disp_horizontal.ppm  disp_vertical.ppm      the coloured maps;   blended.ppm
// Image files: binary PPM (P6, BGR order is produced internally) or PGM (P5) with the extension .ppm / .pgm; PNG needs a
// codec this image does not have in C++ -- `python -m sister_b200.cli` is the same tool on top of cv2 for PNG folders.
// Host code only; the computation goes through the C ABI (sister_b200.h). Exit code 1 + message on any failure.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/sister_b200.h"
#include "magma_lut.h"

struct Image {
    int w = 0, h = 0, ch = 0;
    std::vector<uint8_t> px; // ch == 3: BGR (cv::imread order)
};

static bool read_token(FILE *f, int &v)
{
    int c = fgetc(f);
    for (;;) {
        while (c == ' ' || c == '\t' || c == '\n' || c == '\r') c = fgetc(f);
        if (c == '#') { while (c != '\n' && c != EOF) c = fgetc(f); continue; }
        break;
    }
    if (c < '0' || c > '9') return false;
    v = 0;
    while (c >= '0' && c <= '9') { v = v * 10 + (c - '0'); c = fgetc(f); }
    return true; // one whitespace byte after the token has been consumed
}

static bool read_pnm(const std::string &path, Image &im)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    char magic[3] = {0, 0, 0};
    int maxv = 0;
    bool ok = fread(magic, 1, 2, f) == 2 && magic[0] == 'P' && (magic[1] == '5' || magic[1] == '6') && read_token(f, im.w) &&
              read_token(f, im.h) && read_token(f, maxv) && maxv == 255 && im.w > 0 && im.h > 0;
    if (ok) {
        im.ch = magic[1] == '6' ? 3 : 1;
        im.px.resize((size_t)im.w * im.h * im.ch);
        ok = fread(im.px.data(), 1, im.px.size(), f) == im.px.size();
        if (ok && im.ch == 3)
            for (size_t k = 0; k < im.px.size(); k += 3) std::swap(im.px[k], im.px[k + 2]); // RGB file -> BGR like cv::imread
    }
    fclose(f);
    return ok;
}

static bool load_view(const std::string &folder, const char *name, Image &im)
{
    return read_pnm(folder + name + ".ppm", im) || read_pnm(folder + name + ".pgm", im);
}

static bool write_pgm16(const std::string &path, const std::vector<uint16_t> &m, int w, int h)
{
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return false;
    fprintf(f, "P5\n%d %d\n65535\n", w, h);
    std::vector<uint8_t> be(m.size() * 2);
    for (size_t k = 0; k < m.size(); k++) { be[2 * k] = (uint8_t)(m[k] >> 8); be[2 * k + 1] = (uint8_t)(m[k] & 0xFF); }
    bool ok = fwrite(be.data(), 1, be.size(), f) == be.size();
    fclose(f);
    return ok;
}

static bool write_ppm_bgr(const std::string &path, const std::vector<uint8_t> &bgr, int w, int h)
{
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return false;
    fprintf(f, "P6\n%d %d\n255\n", w, h);
    std::vector<uint8_t> rgb(bgr.size());
    for (size_t k = 0; k < bgr.size(); k += 3) { rgb[k] = bgr[k + 2]; rgb[k + 1] = bgr[k + 1]; rgb[k + 2] = bgr[k]; }
    bool ok = fwrite(rgb.data(), 1, rgb.size(), f) == rgb.size();
    fclose(f);
    return ok;
}

// cv::normalize(src, dst, 0, 255, NORM_MINMAX, CV_8UC1): scale = 255 / (max - min), dst = saturate_cast<uchar>((src - min) * scale)
static std::vector<uint8_t> normalize_minmax(const std::vector<uint16_t> &m)
{
    uint16_t lo = 65535, hi = 0;
    for (uint16_t v : m) { if (v < lo) lo = v; if (v > hi) hi = v; }
    std::vector<uint8_t> out(m.size(), 0);
    if (hi == lo) return out;
    // OpenCV derives scale and shift in double and converts with a float multiply-add (cvtScale<ushort, uchar, float>)
    const double dscale = 255.0 / ((double)hi - (double)lo), dshift = -(double)lo * dscale;
    const float scale = (float)dscale, shift = (float)dshift;
    for (size_t k = 0; k < m.size(); k++) {
        long r = std::lrintf((float)m[k] * scale + shift); // cvRound: round half to even
        out[k] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
    }
    return out;
}

static std::vector<uint8_t> colorize(const std::vector<uint8_t> &g)
{
    std::vector<uint8_t> out(g.size() * 3);
    for (size_t k = 0; k < g.size(); k++) memcpy(&out[3 * k], kMagmaBGR[g[k]], 3);
    return out;
}

int main(int argc, char **argv)
{
    if (argc != 3 && argc != 4) {
        fprintf(stderr, "expected <input folder> <dmax> [output folder]\n");
        return 1;
    }
    const std::string in = argv[1], out = argc == 4 ? argv[3] : argv[1];
    const int D = atoi(argv[2]);
    const char *names[5] = {"center", "right", "top", "left", "bottom"};
    Image v[5];
    for (int k = 0; k < 5; k++) {
        if (!load_view(in, names[k], v[k])) { fprintf(stderr, "cannot read %s%s.ppm/.pgm\n", in.c_str(), names[k]); return 1; }
        if (v[k].w != v[0].w || v[k].h != v[0].h || v[k].ch != v[0].ch) { fprintf(stderr, "views differ in size or type\n"); return 1; }
    }
    const int w = v[0].w, h = v[0].h, ch = v[0].ch;
    sister_ctx *ctx = nullptr;
    int rc = sister_create(&ctx, 0, w, h, D, 1);
    if (rc != SISTER_OK) { fprintf(stderr, "sister_create: %s\n", sister_strerror(rc)); return 1; }
    const uint8_t *views[5] = {v[0].px.data(), v[1].px.data(), v[2].px.data(), v[3].px.data(), v[4].px.data()};
    std::vector<uint16_t> maps[3];
    uint16_t *outs[3];
    for (int m = 0; m < 3; m++) { maps[m].resize((size_t)w * h); outs[m] = maps[m].data(); }
    rc = sister_compute(ctx, views, w, h, ch, (size_t)w * ch, D, SISTER_MODE_ALL, outs, nullptr);
    if (rc != SISTER_OK) {
        fprintf(stderr, "sister_compute: %s: %s\n", sister_strerror(rc), sister_last_error(ctx));
        sister_destroy(ctx);
        return 1;
    }
    sister_destroy(ctx);
    const char *mnames[3] = {"disp_multiview", "disp_horizontal", "disp_vertical"};
    std::vector<uint8_t> colored[3];
    for (int m = 0; m < 3; m++) {
        colored[m] = colorize(normalize_minmax(maps[m]));
        if (!write_pgm16(out + mnames[m] + ".pgm", maps[m], w, h) || !write_ppm_bgr(out + mnames[m] + ".ppm", colored[m], w, h)) {
            fprintf(stderr, "cannot write to %s\n", out.c_str());
            return 1;
        }
    }
    // cv::addWeighted(center, 0.1, multiview coloured, 0.9, 0.0): saturate_cast<uchar>(a * 0.1 + b * 0.9), centre as BGR
    std::vector<uint8_t> blended((size_t)w * h * 3);
    const float alpha = 0.1f;
    for (size_t p = 0; p < (size_t)w * h; p++)
        for (int c = 0; c < 3; c++) {
            const float a = ch == 3 ? v[0].px[3 * p + c] : v[0].px[p];
            // OpenCV's 8-bit addWeighted works in float: fma(a, alpha, b * beta) on its SIMD path
            long r = std::lrintf(std::fmaf(a, alpha, colored[0][3 * p + c] * (1.0f - alpha)));
            blended[3 * p + c] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
        }
    if (!write_ppm_bgr(out + "blended.ppm", blended, w, h)) { fprintf(stderr, "cannot write to %s\n", out.c_str()); return 1; }
    return 0;
}
