// sister_b200: shared device/host definitions. sm_100a only.
//
// Frames and volumes (all on the PADDED frame, Wp = W + 2D, Hp = H + 2D, like the reference, hpp:35-41):
//   image frame   row i in [0,Hp), col j in [0,Wp)
//   view v        0 right (rot 0), 1 left (rot 180), 2 top (rot 90), 3 bottom (rot 270)  -- hpp:56-70
//   view frame    the frame the reference runs the (centre, side) pair in: Hp x Wp for v < 2, Wp x Hp else
//                   v0: (r,c) = (i, j)           v1: (r,c) = (i, Wp-1-j)
//                   v2: (r,c) = (Wp-1-j, Hp-1-i) v3: (r,c) = (Wp-1-j, i)
//   oriented image o = 2v (centre re-oriented for view v) or 2v+1 (the side view re-oriented)
// The 8 oriented images and their census maps are stored in their view frames, so every matching partner
// of a pixel is a contiguous run census[r][c - d], d = 0..D-1, for all four views (no strided reads for
// the vertical-baseline views).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sister {

constexpr int kViews = 4;
constexpr int kP1 = 7;            // hpp:280
constexpr int kP2 = 100;          // sgm.cpp:34-35 (alpha = 0, gamma = 100)
constexpr int kLrcThreshold = 5;  // hpp:200
constexpr int kInvalidCost = 255; // census.cpp:76
constexpr uint32_t kInf2 = 0x3FFF3FFFu; // "MAX_SGM_COST" stand-in for two packed u16 lanes (never selected, never overflows)

struct Dims {
    int W, H, D;   // input size, dispCount
    int Wp, Hp;    // padded
    long long px;  // Wp * Hp
    long long cells; // px * D
    // Byte order inside a cell (the D bytes of one pixel) of the fused-cost volume and of the 8 path volumes, fixed by how
    // the path kernel (sgm.cu) spreads a chain over lanes: `lpc` lanes per chain, 2 * nr disparities per lane. When the
    // lanes are exactly full (D == 2 * lpc * nr: 32, 64, 96, 128, 192, 256, ... 512) the cell is stored LANE-INTERLEAVED:
    //     32-bit word (t, sl), t < nr / 2, sl < lpc, at byte 4 * (t * lpc + sl), holds the disparities
    //     sl * 2nr + { 2t, 2t + 1, nr + 2t, nr + 2t + 1 }
    // which is the pair of packed 16-bit registers (2t, 2t + 1) of lane sl with the odd one shifted up by a byte: a load
    // or store instruction of a chain covers 4 * lpc contiguous bytes, and bytes <-> registers is two instructions per
    // word. Otherwise (interleaved == 0) the cell is in natural disparity order. Producers (k_fuse), consumers
    // (k_sgm_paths, k_sgm_final) and the test taps (api.cu) all go through cell_disp / cell_pos.
    int lpc, nr, lpc_shift, interleaved;
};

// disparity stored at byte `pos` of a cell, and its inverse
__host__ __device__ inline int cell_disp(const Dims &d, int pos)
{
    if (!d.interleaved) return pos;
    const int word = pos >> 2, b = pos & 3, t = word >> d.lpc_shift, sl = word & (d.lpc - 1);
    return sl * 2 * d.nr + 2 * t + (b & 1) + (b >> 1) * d.nr;
}
__host__ __device__ inline int cell_pos(const Dims &d, int disp)
{
    if (!d.interleaved) return disp;
    const int sl = disp / (2 * d.nr), r = disp % (2 * d.nr), hi = r >= d.nr, k = r - hi * d.nr;
    return 4 * ((k >> 1) * d.lpc + sl) + 2 * hi + (k & 1);
}
// fills lpc, nr, lpc_shift, interleaved from D (sgm.cu)
void set_cell_order(Dims &d);

__host__ __device__ inline int view_rows(const Dims &d, int v) { return v < 2 ? d.Hp : d.Wp; }
__host__ __device__ inline int view_cols(const Dims &d, int v) { return v < 2 ? d.Wp : d.Hp; }
// first row of share k when hv rows are dealt to n shares (share k = rows [share_row0(k), share_row0(k + 1)))
__host__ __device__ inline int share_row0(int hv, int k, int n) { return (int)((long long)hv * k / n); }

// image (i,j) -> view frame (r,c)
__host__ __device__ inline void image_to_view(const Dims &d, int v, int i, int j, int &r, int &c)
{
    switch (v) {
    case 0: r = i; c = j; break;
    case 1: r = i; c = d.Wp - 1 - j; break;
    case 2: r = d.Wp - 1 - j; c = d.Hp - 1 - i; break;
    default: r = d.Wp - 1 - j; c = i; break;
    }
}
// view frame (r,c) -> image (i,j)
__host__ __device__ inline void view_to_image(const Dims &d, int v, int r, int c, int &i, int &j)
{
    switch (v) {
    case 0: i = r; j = c; break;
    case 1: i = r; j = d.Wp - 1 - c; break;
    case 2: i = d.Hp - 1 - c; j = d.Wp - 1 - r; break;
    default: i = c; j = d.Wp - 1 - r; break;
    }
}

// status bits written by kernels into the slot's status word (checked by the host after every frame)
constexpr int kStatusFusedOverflow = 1; // a fused cost exceeded 8 bits (cannot happen)
constexpr int kStatusSpinTimeout = 2;   // a pipelined kernel waited too long for its neighbour

} // namespace sister
