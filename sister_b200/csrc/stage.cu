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

constexpr int kCenTileW = 32, kCenTileH = 8;
constexpr int kCenHaloL = 5, kCenHaloR = 4, kCenHaloV = 3; // cols c-5..c+4 (the -5 is for the carry), rows r-3..r+3
constexpr int kCenSmemW = kCenTileW + kCenHaloL + kCenHaloR + 3; // 44
constexpr int kCenSmemH = kCenTileH + 2 * kCenHaloV;             // 14

// grid (ceil(wv/32), ceil(hv/8), 8), block (32, 8)
__global__ void __launch_bounds__(256) k_census(const uint8_t *__restrict__ oriented, Dims d,
                                                unsigned long long *__restrict__ census)
{
    __shared__ uint8_t T[kCenSmemH][kCenSmemW];
    const int o = blockIdx.z, v = o >> 1;
    const int hv = view_rows(d, v), wv = view_cols(d, v);
    const int r0 = blockIdx.y * kCenTileH, c0 = blockIdx.x * kCenTileW;
    if (r0 >= hv || c0 >= wv) return;
    const uint8_t *img = oriented + (size_t)o * d.px;
    unsigned long long *out = census + (size_t)o * d.px;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int e = tid; e < kCenSmemH * (kCenTileW + kCenHaloL + kCenHaloR); e += 256) {
        int tr = e / (kCenTileW + kCenHaloL + kCenHaloR), tc = e % (kCenTileW + kCenHaloL + kCenHaloR);
        int r = min(max(r0 - kCenHaloV + tr, 0), hv - 1);
        int c = min(max(c0 - kCenHaloL + tc, 0), wv - 1);
        T[tr][tc] = img[(size_t)r * wv + c];
    }
    __syncthreads();
    const int r = r0 + threadIdx.y, c = c0 + threadIdx.x;
    if (r >= hv || c >= wv) return;
    unsigned long long code = 0;
    if (r >= 3 && r <= hv - 4 && c >= 4 && c <= wv - 5) {
        const int tr = threadIdx.y + kCenHaloV, tc = threadIdx.x + kCenHaloL;
        unsigned carry;
        if (c > 4) {
            carry = T[tr + 3][tc + 3] > T[tr - 3][tc - 5]; // previous pixel (r, c-1): X(r+3, c-1+4) > X(r-3, c-1-4)
        } else if (r > 3) {
            // previous pixel in raster order is (r-1, wv-5): X(r+2, wv-1) > X(r-4, wv-9)
            carry = img[(size_t)(r + 2) * wv + (wv - 1)] > img[(size_t)(r - 4) * wv + (wv - 9)];
        } else {
            carry = 0;
        }
        unsigned hi = carry << 31, lo = 0;
#pragma unroll
        for (int y = -3; y <= 3; y++) {
#pragma unroll
            for (int x = -4; x <= 4; x++) {
                const int k = (y + 3) * 9 + (x + 4); // 0..62, bit position 62 - k
                unsigned bit = T[tr + y][tc + x] > T[tr - y][tc - x];
                if (k < 31) hi |= bit << (30 - k);
                else lo |= bit << (62 - k);
            }
        }
        code = ((unsigned long long)hi << 32) | lo;
    }
    out[(size_t)r * wv + c] = code;
}

void launch_census(const uint8_t *oriented, const Dims &d, unsigned long long *census, cudaStream_t st, LaunchCounter &lc)
{
    int mw = d.Wp > d.Hp ? d.Wp : d.Hp;
    dim3 grid((mw + kCenTileW - 1) / kCenTileW, (mw + kCenTileH - 1) / kCenTileH, 8), block(32, 8);
    k_census<<<grid, block, 0, st>>>(oriented, d, census);
    lc.add();
}

} // namespace sister
