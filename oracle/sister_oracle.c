/*
 * TEST INFRASTRUCTURE ONLY (oracle/). Plain-C, scalar, 64-bit-indexed restatement of the reference's
 * 5-view disparity path (CVLAB-Unibo/sister). It contains no reference code; every function cites the
 * reference file:line it restates. It exists to (a) check the CUDA product bit-for-bit and (b) cover
 * shapes the reference cannot run (D > 264: postprocess.cpp:193 store[288]; >= 2^31 cells: types.h:31-34).
 *
 * PARITY PIN: the reference ships no tests or golden vectors. This file is pinned
 * against the reference ITSELF: oracle/_ref/libsister_ref.so is the unmodified reference compiled in
 * place, and tests/test_oracle_ref.py compares every function below with it on seeded inputs; the
 * committed fixtures in tests/golden/ were produced by that library (scripts/gen_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; the
 * product (sister_b200/) never does.
 *
 * Conventions: all frames are the padded frames (w = W + 2D, h = H + 2D); volumes are [row][col][d],
 * d fastest, uint16 (types.h:31-34).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SO_MAXC 65535u
#define SO_P1 7u      /* hpp:280 */
#define SO_P2 100u    /* sgm.cpp:34-35 with alpha = 0, gamma = 100 >= P2min = 17 (sgm.cpp:15-22) */
#define SO_LRC_THR 5  /* hpp:200 */

/* ------------------------------------------------------------------ staging (hpp:29-70, 111-118) */

/* cvtColor(BGR2GRAY), OpenCV 4.x 8-bit fixed point (hpp:29-33; third-party, see fake_cv shim). */
void so_grey_bgr(const uint8_t *bgr, int w, int h, size_t row_stride, uint8_t *grey)
{
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            const uint8_t *p = bgr + (size_t)i * row_stride + (size_t)j * 3;
            grey[(size_t)i * w + j] = (uint8_t)((3735 * p[0] + 19235 * p[1] + 9798 * p[2] + 16384) >> 15);
        }
}

/* copyMakeBorder(D,D,D,D, BORDER_REPLICATE) (hpp:35-39). out: (h+2D) x (w+2D). */
void so_pad_replicate(const uint8_t *g, int w, int h, int D, uint8_t *out)
{
    int wp = w + 2 * D, hp = h + 2 * D;
    for (int i = 0; i < hp; i++) {
        int si = i - D; si = si < 0 ? 0 : (si >= h ? h - 1 : si);
        for (int j = 0; j < wp; j++) {
            int sj = j - D; sj = sj < 0 ? 0 : (sj >= w ? w - 1 : sj);
            out[(size_t)i * wp + j] = g[(size_t)si * w + sj];
        }
    }
}

/*
 * Re-oriented copies (hpp:56-70). rot 0: identity (h x w); 180: flip(1) (h x w);
 * 90: transpose + flip(-1) -> (w rows x h cols), T(r,c) = X(h-1-c, w-1-r);
 * 270: transpose + flip(0) -> (w rows x h cols), T(r,c) = X(c, w-1-r).
 */
void so_orient(const uint8_t *x, int w, int h, int rot, uint8_t *out)
{
    if (rot == 0) { memcpy(out, x, (size_t)w * h); return; }
    if (rot == 180) {
        for (int i = 0; i < h; i++)
            for (int j = 0; j < w; j++) out[(size_t)i * w + j] = x[(size_t)i * w + (w - 1 - j)];
        return;
    }
    for (int r = 0; r < w; r++)
        for (int c = 0; c < h; c++)
            out[(size_t)r * h + c] = (rot == 90) ? x[(size_t)(h - 1 - c) * w + (w - 1 - r)] : x[(size_t)c * w + (w - 1 - r)];
}

/* convertTo(CV_16UC1) on an integer-valued float map, crop Rect(D,D,W,H), * 255 saturated (hpp:111-118). */
void so_encode_crop(const float *disp, int wp, int hp, int D, uint16_t *out)
{
    int w = wp - 2 * D, h = hp - 2 * D;
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            float f = disp[(size_t)(i + D) * wp + (j + D)];
            long v = (long)f; /* values are integers (postprocess.cpp:141,153) or -10 */
            if (v < 0) v = 0;
            if (v > 65535) v = 65535;
            long m = v * 255;
            out[(size_t)i * w + j] = (uint16_t)(m > 65535 ? 65535 : m);
        }
}

/* ------------------------------------------------------------------ census (census.cpp:30-51) */

/*
 * 9x7 centre-symmetric census, rw = 4, rh = 3 (census.cpp:154). The 64-bit accumulator is declared
 * outside the pixel loops (census.cpp:42) and only shifted, so after 63 shifts the previous pixel's
 * last comparison survives as bit 63 of the next pixel's code ("carry"). Border pixels stay 0.
 */
void so_census(const uint8_t *img, int w, int h, uint64_t *out)
{
    memset(out, 0, (size_t)w * h * sizeof(uint64_t));
    uint64_t v = 0;
    for (int i = 3; i < h - 3; i++)
        for (int j = 4; j < w - 4; j++) {
            for (int y = -3; y <= 3; y++)
                for (int x = -4; x <= 4; x++) {
                    int a = img[(size_t)(i + y) * w + (j + x)], b = img[(size_t)(i - y) * w + (j - x)];
                    v = (v << 1) | (uint64_t)(a - b > 0); /* census.cpp:32-34 */
                }
            out[(size_t)i * w + j] = v;
        }
}

/* ------------------------------------------------------------------ raw cost (census.cpp:54-146) */

static inline uint16_t so_popc64(uint64_t x) { return (uint16_t)__builtin_popcountll(x); }

/*
 * Hamming volume of one (centre, side) pair in the pair's own frame. Rows 0..2 = 255 (census.cpp:95-98);
 * rows 2..h-3 are computed (census.cpp:123-135 cover [2, h-2)) and row 2 is then overwritten with 255 by
 * the mis-indexed "last 3 lines" loop (census.cpp:142-145); rows h-2, h-1 are never written: defined 0.
 */
void so_cost_volume(const uint64_t *c1, const uint64_t *c2, int h, int w, int D, uint16_t *dsi)
{
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            uint16_t *o = dsi + ((size_t)i * w + j) * D;
            if (i < 3) { for (int d = 0; d < D; d++) o[d] = 255; continue; }
            if (i >= h - 2) { for (int d = 0; d < D; d++) o[d] = 0; continue; }
            for (int d = 0; d < D; d++)
                o[d] = (d > j) ? 255 : so_popc64(c1[(size_t)i * w + j] ^ c2[(size_t)i * w + j - d]);
        }
}

/* ------------------------------------------------------------------ WTA (postprocess.cpp:74-315) */

/* uniqueness == 1 makes the second-minimum test a no-op (postprocess.cpp:140,176): first-index argmin. */
void so_wta_left(const uint16_t *vol, int w, int h, int D, float *out)
{
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            const uint16_t *c = vol + ((size_t)i * w + j) * D;
            int end = (D - 1 > j) ? j : D - 1, best = 0;
            uint32_t mn = c[0];
            for (int d = 1; d <= end; d++)
                if (c[d] < mn) { mn = c[d]; best = d; }
            out[(size_t)i * w + j] = (float)best;
        }
}

/* Right-image disparity along the anti-diagonal dsi[i][j+d][d] (postprocess.cpp:215-222, 287-304). */
void so_wta_right(const uint16_t *vol, int w, int h, int D, float *out)
{
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            int end = (D - 1 < w - 1 - j) ? D - 1 : w - 1 - j, best = 0;
            uint32_t mn = vol[((size_t)i * w + j) * D];
            for (int d = 1; d <= end; d++) {
                uint32_t c = vol[((size_t)i * w + j + d) * D + d];
                if (c < mn) { mn = c; best = d; }
            }
            out[(size_t)i * w + j] = (float)best;
        }
}

/* ------------------------------------------------------------------ median (postprocess.cpp:15-71) */

static inline void so_sort2(float *a, float *b)
{
    float lo = *a < *b ? *a : *b, hi = *a < *b ? *b : *a;
    *a = lo; *b = hi;
}

/* The 19-exchange network of postprocess.cpp:52-58; returns the median of 9. */
static float so_median9(float v0, float v1, float v2, float v3, float v4, float v5, float v6, float v7, float v8)
{
    so_sort2(&v1, &v2); so_sort2(&v4, &v5); so_sort2(&v7, &v8);
    so_sort2(&v0, &v1); so_sort2(&v3, &v4); so_sort2(&v6, &v7);
    so_sort2(&v1, &v2); so_sort2(&v4, &v5); so_sort2(&v7, &v8);
    so_sort2(&v0, &v3); so_sort2(&v5, &v8); so_sort2(&v4, &v7);
    so_sort2(&v3, &v6); so_sort2(&v1, &v4); so_sort2(&v2, &v5);
    so_sort2(&v4, &v7); so_sort2(&v4, &v2); so_sort2(&v6, &v4);
    so_sort2(&v4, &v2);
    return v4;
}

/*
 * median3x3_SSE called with source == dest (hpp:198-199). The SSE loop walks the image as one flat
 * array, stores each group of 4 results after loading its inputs, so: the row above is already
 * filtered, the current and next rows are still raw; element [w] receives the zero-initialised
 * "lastMedian" (postprocess.cpp:29,61-63); the first row and the last w+4 elements are untouched
 * (loop bound postprocess.cpp:67; the memcpys at :69-70 are self-copies).
 */
void so_median_inplace(float *m, int w, int h)
{
    size_t N = (size_t)w * h;
    if (N < (size_t)2 * w + 8) { if (N > (size_t)w) m[w] = 0; return; }
    float *raw = (float *)malloc(N * sizeof(float));
    memcpy(raw, m, N * sizeof(float));
    m[w] = 0;
    for (size_t p = (size_t)w + 1; p + w + 5 <= N; p++) /* p <= N - w - 5 */
        m[p] = so_median9(m[p - w - 1], m[p - w], m[p - w + 1], raw[p - 1], raw[p], raw[p + 1], raw[p + w - 1], raw[p + w], raw[p + w + 1]);
    free(raw);
}

/* ------------------------------------------------------------------ LRC (postprocess.cpp:318-341) */

void so_lrcheck(float *L, const float *R, int w, int h, int thr)
{
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            float b = L[(size_t)i * w + j];
            if (b >= 0 && b <= j) {
                float mt = R[(size_t)i * w + (int)(j - b)];
                int diff = (int)(b - mt);
                if (abs(diff) > thr) L[(size_t)i * w + j] = -10;
            } else
                L[(size_t)i * w + j] = -10;
        }
}

/* ------------------------------------------------------------------ SGM (sgm.cpp:26-455) */

static inline uint16_t so_adds(uint32_t a, uint32_t b) { uint32_t s = a + b; return (uint16_t)(s > SO_MAXC ? SO_MAXC : s); }
static inline uint16_t so_subs(uint32_t a, uint32_t b) { return (uint16_t)(a > b ? a - b : 0); }
static inline uint16_t so_min16(uint16_t a, uint16_t b) { return a < b ? a : b; }

/* One path update on a "remaining line" (sgm.cpp:282-305 and the r1/r2/r3 copies :313-383), all in
 * saturating uint16 like PADDUSW/PSUBUSW/PMINUW. Lp has valid entries at [-1..D]. Returns min_d. */
static uint16_t so_path(const uint16_t *Lp, uint16_t mp, const uint16_t *c, int D, uint16_t *Lout)
{
    uint16_t p2 = so_adds(SO_P2, mp), mn = SO_MAXC;
    for (int d = 0; d < D; d++) {
        uint16_t t = so_min16(so_adds(Lp[d - 1], SO_P1), so_adds(Lp[d + 1], SO_P1));
        t = so_min16(t, Lp[d]);
        t = so_min16(t, p2);
        t = so_subs(t, mp);
        Lout[d] = so_adds(c[d], t);
        mn = so_min16(mn, Lout[d]);
    }
    return mn;
}

/*
 * sgm(...,7,17,8) -> accumulateCostsSSE: two passes x four paths. The image argument only feeds
 * adaptP2 with alpha = 0 (sgm.cpp:34,227-232) and is therefore irrelevant.
 * Layout of the line buffers here: (w + 2) columns (index 0 = column -1, index w + 1 = column w),
 * each D + 2 wide (index 0 = disparity -1, index D + 1 = disparity D).
 */
int so_sgm(const uint16_t *C, int h, int w, int D, uint16_t *S)
{
    const size_t colw = (size_t)D + 2, ncol = (size_t)w + 2;
    uint16_t *L1a = malloc(ncol * colw * 2), *L1b = malloc(ncol * colw * 2);
    uint16_t *L2 = malloc(ncol * colw * 2), *L3 = malloc(ncol * colw * 2);
    uint16_t *m1a = malloc(ncol * 2), *m1b = malloc(ncol * 2), *m2 = malloc(ncol * 2), *m3 = malloc(ncol * 2);
    uint16_t *L0a = malloc(colw * 2), *L0b = malloc(colw * 2), *tmp = malloc(colw * 2), *cz = malloc(colw * 2);
    if (!L1a || !L1b || !L2 || !L3 || !m1a || !m1b || !m2 || !m3 || !L0a || !L0b || !tmp || !cz) return -1;
    /* borders: everything MAXC, minima of the off-image columns 0 (sgm.cpp:57-87) */
    memset(L1a, 0xFF, ncol * colw * 2); memset(L1b, 0xFF, ncol * colw * 2);
    memset(L2, 0xFF, ncol * colw * 2); memset(L3, 0xFF, ncol * colw * 2);
    memset(L0a, 0xFF, colw * 2); memset(L0b, 0xFF, colw * 2); memset(tmp, 0xFF, colw * 2);
    memset(m1a, 0, ncol * 2); memset(m1b, 0, ncol * 2); memset(m2, 0, ncol * 2); memset(m3, 0, ncol * 2);
#define COL(buf, j) ((buf) + ((size_t)((j) + 1)) * colw + 1) /* pointer to disparity 0 of column j */
    uint16_t *L1cur = L1a, *L1prev = L1b, *m1cur = m1a + 1, *m1prev = m1b + 1, *mm2 = m2 + 1, *mm3 = m3 + 1;
    uint16_t *L0 = L0a + 1, *L0last = L0b + 1;

    for (int pass = 0; pass < 2; pass++) {
        int i1 = pass ? h - 1 : 0, i2 = pass ? -1 : h, di = pass ? -1 : 1;
        int j1 = pass ? w - 1 : 0, j2 = pass ? -1 : w, dj = di;

        /* ---- first line (sgm.cpp:103-207) ---- */
        uint16_t min0last = SO_MAXC;
        for (int j = j1; j != j2; j += dj) {
            const uint16_t *c = C + ((size_t)i1 * w + j) * D;
            uint16_t *s = S + ((size_t)i1 * w + j) * D;
            uint16_t mc = SO_MAXC, min0 = SO_MAXC;
            for (int d = 0; d < D; d++) {
                uint16_t cost = c[d] == 255 ? 0 : c[d]; /* sgm.cpp:109,123,146 */
                COL(L1prev, j)[d] = cost; COL(L2, j)[d] = cost; COL(L3, j)[d] = cost;
                if (cost < mc) mc = cost;
                if (j == j1) {
                    L0last[d] = cost;
                    if (pass == 0) s[d] = cost; else s[d] = (uint16_t)(s[d] + cost);
                } else {
                    int32_t mp = L0last[d];                                   /* sgm.cpp:159-174, int32 */
                    int32_t a = (int32_t)L0last[d - 1] + (int32_t)SO_P1; if (mp > a) mp = a;
                    int32_t b = (int32_t)L0last[d + 1] + (int32_t)SO_P1; if (mp > b) mp = b;
                    int32_t p2 = (int32_t)min0last + (int32_t)SO_P2; if (mp > p2) mp = p2;
                    mp -= min0last;
                    uint16_t nc = (uint16_t)(uint8_t)(cost + mp);            /* types.h:28 via sgm.cpp:176 */
                    L0[d] = nc;
                    if (min0 > nc) min0 = nc;
                    if (pass == 0) s[d] = nc; else s[d] = (uint16_t)(s[d] + nc);
                }
            }
            m1prev[j] = mc; mm2[j] = mc; mm3[j] = mc;
            if (j == j1) min0last = mc;
            else { uint16_t *t = L0; L0 = L0last; L0last = t; min0last = min0; } /* sgm.cpp:198-199 */
        }

        /* ---- remaining lines (sgm.cpp:213-440) ---- */
        for (int i = i1 + di; i != i2; i += di) {
            memset(L0last, 0, (size_t)D * 2); /* sgm.cpp:215-216; [-1] and [D] stay MAXC */
            min0last = 0;
            for (int j = j1; j != j2; j += dj) {
                const uint16_t *c = C + ((size_t)i * w + j) * D;
                uint16_t *s = S + ((size_t)i * w + j) * D;
                /* r0: in-place over L0last, all reads precede the stores (sgm.cpp:258,282-299) */
                min0last = so_path(L0last, min0last, c, D, tmp + 1);
                memcpy(L0last, tmp + 1, (size_t)D * 2);
                /* r1: predecessor (i-di, j-dj), double buffered (sgm.cpp:308-332) */
                m1cur[j] = so_path(COL(L1prev, j - dj), m1prev[j - dj], c, D, COL(L1cur, j));
                /* r2: predecessor (i-di, j), in place (sgm.cpp:335-358) */
                mm2[j] = so_path(COL(L2, j), mm2[j], c, D, cz + 1);
                memcpy(COL(L2, j), cz + 1, (size_t)D * 2);
                /* r3: predecessor (i-di, j+dj): still the previous line's value (sgm.cpp:361-383) */
                mm3[j] = so_path(COL(L3, j + dj), mm3[j + dj], c, D, cz + 1);
                memcpy(COL(L3, j), cz + 1, (size_t)D * 2);
                const uint16_t *l0 = L0last, *l1 = COL(L1cur, j), *l2 = COL(L2, j), *l3 = COL(L3, j);
                for (int d = 0; d < D; d++) {
                    uint16_t g = so_adds(so_adds(so_adds(l0[d], l1[d]), l2[d]), l3[d]);
                    s[d] = pass == 0 ? g : so_adds(s[d], g); /* sgm.cpp:386-391 */
                }
            }
            { uint16_t *t = L1cur; L1cur = L1prev; L1prev = t; t = m1cur; m1cur = m1prev; m1prev = t; }
        }
    }
#undef COL
    free(L1a); free(L1b); free(L2); free(L3); free(m1a); free(m1b); free(m2); free(m3);
    free(L0a); free(L0b); free(tmp); free(cz);
    return 0;
}

/* ------------------------------------------------------------------ doMultiStereo (hpp:152-295) */

/*
 * views: center, right, top, left, bottom, each hp x wp uint8 (grey, padded). mode 0/1/2 (hpp:262-276).
 * Optional outputs (any may be NULL):
 *   masks  4 x hp x wp uint8, image frame, order right, left, top, bottom (hpp:201-251)
 *   fused  hp x wp x D uint16 (hpp:255-277);  sum  hp x wp x D uint16 (hpp:280)
 *   disp   hp x wp float (hpp:283)
 *   lr     4 x hp x wp float: each view's left map after median + LRC, in the VIEW's frame
 */
int so_multistereo(const uint8_t *const *views, int wp, int hp, int D, int mode,
                   uint8_t *masks, uint16_t *fused, uint16_t *sum, float *disp, float *lr)
{
    const int w = wp, h = hp;
    const size_t px = (size_t)w * h, cells = px * D;
    static const int rot[4] = {0, 180, 90, 270};
    static const int side[4] = {1, 3, 2, 4}; /* right, left, top, bottom within views[] */
    uint8_t *mk = calloc(4 * px, 1), *a = malloc(px), *b = malloc(px);
    uint64_t *ca = malloc(px * 8), *cb = malloc(px * 8);
    uint16_t *vol[4] = {0, 0, 0, 0};
    float *L = malloc(px * 4), *R = malloc(px * 4);
    uint16_t *F = calloc(cells, 2), *Sv = malloc(cells * 2);
    int rc = -1;
    if (!mk || !a || !b || !ca || !cb || !L || !R || !F || !Sv) goto out;
    for (int v = 0; v < 4; v++) {
        int horiz = v < 2;
        if ((mode == 1 && !horiz) || (mode == 2 && horiz)) continue;
        int vw = horiz ? w : h, vh = horiz ? h : w; /* the pair's own frame */
        vol[v] = malloc(cells * 2);
        if (!vol[v]) goto out;
        so_orient(views[0], w, h, rot[v], a);
        so_orient(views[side[v]], w, h, rot[v], b);
        so_census(a, vw, vh, ca);
        so_census(b, vw, vh, cb);
        so_cost_volume(ca, cb, vh, vw, D, vol[v]);       /* hpp:181-184 */
        so_wta_left(vol[v], vw, vh, D, L);               /* hpp:196 ... */
        so_wta_right(vol[v], vw, vh, D, R);
        so_median_inplace(L, vw, vh);
        so_median_inplace(R, vw, vh);
        so_lrcheck(L, R, vw, vh, SO_LRC_THR);
        if (lr) memcpy(lr + (size_t)v * px, L, px * 4);
        for (int i = 0; i < h; i++)
            for (int j = 0; j < w; j++) {
                int r, c; /* this image pixel in the pair's frame */
                switch (rot[v]) {
                case 0: r = i; c = j; break;
                case 180: r = i; c = w - 1 - j; break;
                case 90: r = w - 1 - j; c = h - 1 - i; break;
                default: r = w - 1 - j; c = i; break;
                }
                uint8_t m = !(L[(size_t)r * vw + c] <= 0 || c < D); /* hpp:203 */
                mk[(size_t)v * px + (size_t)i * w + j] = m;
                if (m) {
                    const uint16_t *src = vol[v] + ((size_t)r * vw + c) * D;
                    uint16_t *dst = F + ((size_t)i * w + j) * D;
                    for (int d = 0; d < D; d++) dst[d] = (uint16_t)(dst[d] + src[d]); /* hpp:259-276 */
                }
            }
        free(vol[v]); vol[v] = 0;
    }
    if (so_sgm(F, h, w, D, Sv) != 0) goto out;           /* hpp:280 */
    so_wta_left(Sv, w, h, D, L);                         /* hpp:283 */
    if (masks) memcpy(masks, mk, 4 * px);
    if (fused) memcpy(fused, F, cells * 2);
    if (sum) memcpy(sum, Sv, cells * 2);
    if (disp) memcpy(disp, L, px * 4);
    rc = 0;
out:
    for (int v = 0; v < 4; v++) free(vol[v]);
    free(mk); free(a); free(b); free(ca); free(cb); free(L); free(R); free(F); free(Sv);
    return rc;
}

/*
 * compute_disparities (hpp:26-119). views: center, right, top, left, bottom; channels 3 (BGR) or 1.
 * mode_mask bit0 multiview, bit1 horizontal, bit2 vertical; out[k] H x W uint16 (may be NULL when the
 * bit is clear). raw_disp (optional): 3 x hp x wp float, the un-encoded padded maps.
 */
int so_compute_disparities(const uint8_t *const *views, int w, int h, int channels, size_t row_stride,
                           int D, unsigned mode_mask, uint16_t *const *out, float *raw_disp)
{
    int wp = w + 2 * D, hp = h + 2 * D, rc = -1;
    size_t px = (size_t)wp * hp;
    uint8_t *g = malloc((size_t)w * h), *pad[5] = {0, 0, 0, 0, 0};
    float *disp = malloc(px * 4);
    if (!g || !disp) goto out;
    for (int v = 0; v < 5; v++) {
        pad[v] = malloc(px);
        if (!pad[v]) goto out;
        if (channels == 3) so_grey_bgr(views[v], w, h, row_stride, g);
        else for (int i = 0; i < h; i++) memcpy(g + (size_t)i * w, views[v] + (size_t)i * row_stride, (size_t)w);
        so_pad_replicate(g, w, h, D, pad[v]);
    }
    for (int mode = 0; mode < 3; mode++) {
        if (!(mode_mask & (1u << mode))) continue;
        if (so_multistereo((const uint8_t *const *)pad, wp, hp, D, mode, 0, 0, 0, disp, 0) != 0) goto out;
        if (out && out[mode]) so_encode_crop(disp, wp, hp, D, out[mode]);
        if (raw_disp) memcpy(raw_disp + (size_t)mode * px, disp, px * 4);
    }
    rc = 0;
out:
    for (int v = 0; v < 5; v++) free(pad[v]);
    free(g); free(disp);
    return rc;
}
