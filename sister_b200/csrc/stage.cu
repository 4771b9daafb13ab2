// sister_b200 / stage.cu -- image staging and census kernels (sm_100a).
//
//   k_prep    grey (OpenCV-4 fixed-point BGR2GRAY, hpp:29-33) + replicate pad by D (hpp:35-39) +
//             re-orientation into the 8 view-frame images of hpp:56-70, in one pass over the 5 inputs.
//   k_census  9x7 centre-symmetric census (census.cpp:30-51, rw = 4, rh = 3 from census.cpp:154),
//             including the accumulator carry of census.cpp:42 (bit 63 = last comparison of the pixel
//             scanned just before, in the view frame's raster order).
#include "kernels.cuh"

namespace sister {

__device__ __forceinline__ uint8_t padded_grey(const uint8_t *__restrict__ img, int row_stride, int ch,
                                               int W, int H, int D, int i, int j)
{
    int ii = min(max(i - D, 0), H - 1);
    int jj = min(max(j - D, 0), W - 1);
    const uint8_t *p = img + (size_t)ii * row_stride + (size_t)jj * ch;
    if (ch == 3) return (uint8_t)((3735 * p[0] + 19235 * p[1] + 9798 * p[2] + 16384) >> 15);
    return p[0];
}

// grid (ceil(wv/32), ceil(hv/32), 8), block (32, 8)
__global__ void __launch_bounds__(256) k_prep(const uint8_t *__restrict__ in, size_t view_stride, int row_stride,
                                              int ch, Dims d, uint8_t *__restrict__ oriented)
{
    __shared__ uint8_t tile[32][33];
    const int o = blockIdx.z, v = o >> 1;
    const int src_view = (o & 1) ? (v == 0 ? 1 : v == 1 ? 3 : v == 2 ? 2 : 4) : 0; // center,right,top,left,bottom
    const int hv = view_rows(d, v), wv = view_cols(d, v);
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    if (r0 >= hv || c0 >= wv) return;
    const uint8_t *img = in + (size_t)src_view * view_stride;
    uint8_t *out = oriented + (size_t)o * d.px;
    const int tx = threadIdx.x, ty = threadIdx.y;
    if (v < 2) {
        // rows of the view frame are rows of the image: coalesced both ways
        for (int k = 0; k < 4; k++) {
            int r = r0 + ty + 8 * k, c = c0 + tx;
            if (r < hv && c < wv) {
                int i, j;
                view_to_image(d, v, r, c, i, j);
                out[(size_t)r * wv + c] = padded_grey(img, row_stride, ch, d.W, d.H, d.D, i, j);
            }
        }
    } else {
        // view rows are image columns: read along image rows (tx <-> view row), transpose through smem
        for (int k = 0; k < 4; k++) {
            int cl = ty + 8 * k, rl = tx;
            int r = r0 + rl, c = c0 + cl;
            if (r < hv && c < wv) {
                int i, j;
                view_to_image(d, v, r, c, i, j);
                tile[cl][rl] = padded_grey(img, row_stride, ch, d.W, d.H, d.D, i, j);
            }
        }
        __syncthreads();
        for (int k = 0; k < 4; k++) {
            int rl = ty + 8 * k, cl = tx;
            int r = r0 + rl, c = c0 + cl;
            if (r < hv && c < wv) out[(size_t)r * wv + c] = tile[cl][rl];
        }
    }
}

void launch_prep(const uint8_t *in, size_t view_stride, int row_stride, int channels, const Dims &d,
                 uint8_t *oriented, cudaStream_t st, LaunchCounter &lc)
{
    int m = d.Wp > d.Hp ? d.Wp : d.Hp;
    dim3 grid((m + 31) / 32, (m + 31) / 32, 8), block(32, 8);
    k_prep<<<grid, block, 0, st>>>(in, view_stride, row_stride, channels, d, oriented);
    lc.add();
}

// ---------------------------------------------------------------------------------------------- census

constexpr int kCenPx = 2;                        // horizontally adjacent pixels per thread (they share 8 of 9 window columns)
constexpr int kCenTileW = 32 * kCenPx, kCenTileH = 8;
constexpr int kCenOrg = 8;                       // staged column of the tile's first pixel
constexpr int kCenWords = (kCenOrg + kCenTileW + 8) / 4; // staged columns c0-8 .. c0+71 as 32-bit words (needed: c0-5 .. c0+67)
constexpr int kCenHaloV = 3;                     // rows r-3..r+3
constexpr int kCenSmemH = kCenTileH + 2 * kCenHaloV; // 14

// One census bit per two instructions: t = X(p - o) - X(p + o) is negative exactly when X(p + o) > X(p - o)
// (census.cpp:24-28), and a funnel shift moves its sign bit into a group accumulator, first scanned bit first. The
// offsets o and -o compare the same two pixels, so the window is walked by ROW PAIRS (r-3, r+3), (r-2, r+2), (r-1, r+1),
// (r, r): a pair yields the 9 bits of its upper row (offsets y < 0) and the 9 bits of its lower row (y > 0) from the same
// values, only two rows are live at a time, and a thread's two pixels share the loads (35 byte loads per pixel; the first
// version kept the window in 73 registers and spent 64 loads + 77 ISETP + 65 SEL + 72 LOP3 / IMAD on the 63 bits). The
// tile is staged with aligned 32-bit loads (the padded sizes are multiples of 4, hpp:35-41 + postprocess.cpp:18): a word
// is inside the image or outside it as a whole; pixels that get a code never look outside (census.cpp:33-36), so outside
// words are simply zero.
// grid (ceil(wv/64), ceil(hv/8), 8), block (32, 8)
__global__ void __launch_bounds__(256) k_census(const uint8_t *__restrict__ oriented, Dims d,
                                                unsigned long long *__restrict__ census)
{
    __shared__ __align__(16) uint8_t T[kCenSmemH][kCenWords * 4];
    const int o = blockIdx.z, v = o >> 1;
    const int hv = view_rows(d, v), wv = view_cols(d, v);
    const int r0 = blockIdx.y * kCenTileH, c0 = blockIdx.x * kCenTileW;
    if (r0 >= hv || c0 >= wv) return;
    const uint8_t *img = oriented + (size_t)o * d.px;
    unsigned long long *out = census + (size_t)o * d.px;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int e = tid; e < kCenSmemH * kCenWords; e += 256) {
        const int tr = e / kCenWords, tw = e % kCenWords;
        const int r = r0 - kCenHaloV + tr, c = c0 - kCenOrg + 4 * tw;
        uint32_t w = 0;
        if (r >= 0 && r < hv && c >= 0 && c < wv) w = __ldg(reinterpret_cast<const uint32_t *>(img + (size_t)r * wv + c));
        reinterpret_cast<uint32_t *>(&T[tr][0])[tw] = w;
    }
    __syncthreads();
    const int r = r0 + threadIdx.y, ca = c0 + kCenPx * threadIdx.x;
    if (r >= hv || ca >= wv) return;
    const int tr = threadIdx.y + kCenHaloV, tc = kCenPx * threadIdx.x + kCenOrg;
    const bool row_ok = r >= 3 && r <= hv - 4;
    // 9-bit groups per pixel, raster order: up[0..2] = rows r-3, r-2, r-1, mid = row r, dn[0..2] = rows r+1, r+2, r+3
    unsigned up[kCenPx][3], dn[kCenPx][3], mid[kCenPx];
    constexpr int NC = 9 + kCenPx - 1; // window columns ca-4 .. ca+4+(kCenPx-1)
#pragma unroll
    for (int g = 0; g < 3; g++) {
        int A[NC], B[NC]; // rows r-3+g and r+3-g
#pragma unroll
        for (int x = 0; x < NC; x++) { A[x] = T[tr - 3 + g][tc - 4 + x]; B[x] = T[tr + 3 - g][tc - 4 + x]; }
#pragma unroll
        for (int p = 0; p < kCenPx; p++) {
            unsigned u = 0, l = 0;
#pragma unroll
            for (int x = 0; x < 9; x++) {
                // offset (y, x-4) with y = g-3 < 0: X(r+y, c+x-4) > X(r-y, c-x+4);   offset (-y, x-4): the mirrored pair
                u = __funnelshift_l((unsigned)(B[p + 8 - x] - A[p + x]), u, 1);
                l = __funnelshift_l((unsigned)(A[p + 8 - x] - B[p + x]), l, 1);
            }
            up[p][g] = u; dn[p][2 - g] = l;
        }
    }
    {
        int A[NC];
#pragma unroll
        for (int x = 0; x < NC; x++) A[x] = T[tr][tc - 4 + x];
#pragma unroll
        for (int p = 0; p < kCenPx; p++) {
            unsigned m = 0;
#pragma unroll
            for (int x = 0; x < 9; x++) m = __funnelshift_l((unsigned)(A[p + 8 - x] - A[p + x]), m, 1);
            mid[p] = m;
        }
    }
    unsigned prev_last = 0; // bit 0 of the code of the pixel to the left, when that pixel has one
    bool prev_valid = false;
#pragma unroll
    for (int p = 0; p < kCenPx; p++) {
        const int c = ca + p;
        if (c >= wv) break;
        unsigned long long code = 0;
        const bool valid = row_ok && c >= 4 && c <= wv - 5;
        if (valid) {
            // bit 63: the accumulator is not reset between pixels (census.cpp:30-51), it still holds the last bit of the
            // pixel scanned before
            unsigned carry;
            if (prev_valid) {
                carry = prev_last;
            } else if (c > 4) {
                carry = T[tr + 3][tc + p + 3] > T[tr - 3][tc + p - 5]; // previous pixel (r, c-1): X(r+3, c-1+4) > X(r-3, c-1-4)
            } else if (r > 3) {
                // previous pixel in raster order is (r-1, wv-5): X(r+2, wv-1) > X(r-4, wv-9)
                carry = img[(size_t)(r + 2) * wv + (wv - 1)] > img[(size_t)(r - 4) * wv + (wv - 9)];
            } else {
                carry = 0;
            }
            // bit positions: carry 63, up[0] 62..54, up[1] 53..45, up[2] 44..36, mid 35..27, dn[0] 26..18, dn[1] 17..9, dn[2] 8..0
            const unsigned hi = (carry << 31) | (up[p][0] << 22) | (up[p][1] << 13) | (up[p][2] << 4) | (mid[p] >> 5);
            const unsigned lo = (mid[p] << 27) | (dn[p][0] << 18) | (dn[p][1] << 9) | dn[p][2];
            code = ((unsigned long long)hi << 32) | lo;
            prev_last = lo & 1u;
        }
        prev_valid = valid;
        out[(size_t)r * wv + c] = code;
    }
}

void launch_census(const uint8_t *oriented, const Dims &d, unsigned long long *census, cudaStream_t st, LaunchCounter &lc)
{
    int mw = d.Wp > d.Hp ? d.Wp : d.Hp;
    dim3 grid((mw + kCenTileW - 1) / kCenTileW, (mw + kCenTileH - 1) / kCenTileH, 8), block(32, 8);
    k_census<<<grid, block, 0, st>>>(oriented, d, census);
    lc.add();
}

} // namespace sister
