// sister_b200 / sgm.cu -- semi-global aggregation and final selection (sm_100a).
//
// Restates accumulateCostsSSE (sgm.cpp:26-455; P1 = 7, P2 = 100, 2 passes x 4 paths) on the uint8 fused volume.
// One warp works on one pixel: the D disparities are spread over the lanes, DPL = 2*NR consecutive disparities
// per lane, two per 32-bit register as packed u16 (VIMNMX.U16x2 / VIADDMNMX.U16x2 / VIADD.16x2 on sm_100a).
//
// State is kept NORMALISED: A(d) = L(d) - min_d L. With it the reference's update
//     L'(d) = C(d) (+) ( min(L(d), L(d-1) (+) P1, L(d+1) (+) P1, P2 (+) m) (-) m )           sgm.cpp:282-297
// becomes L'(d) = C(d) + min(A(d), A(d-1) + P1, A(d+1) + P1, P2), which needs no saturation because C <= 252
// (match.cu) gives L' <= 352 and the 8-path sum <= 2816. The reference's sentinels map as follows:
//     L(-1) = L(D) = 65535 (sgm.cpp:84-87)                 -> kInf2 in the neighbour slots (never the minimum)
//     off-image predecessor column: L = 65535, m = 0       -> A = kInf2 everywhere  => L' = C + P2  (sgm.cpp:57-81)
//     r0 at the start of a row: L = 0, m = 0               -> A = 0 everywhere      => L' = C       (sgm.cpp:215-216)
//     first line of a pass: L1 = L2 = L3 = C, m = min C    -> A = C - min C, no contribution to the sum (sgm.cpp:103-138)
//     first line, r0: int32 arithmetic + 8-bit truncation  -> k_sgm_first_line (sgm.cpp:141-190, types.h:28)
// The "cost == 255 -> 0" substitution of sgm.cpp:109,123,146 can never fire on C <= 252.
#include "kernels.cuh"

namespace sister {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr uint32_t kP1x2 = (uint32_t)kP1 * 0x10001u;
constexpr uint32_t kP2x2 = (uint32_t)kP2 * 0x10001u;

struct PassGeom {
    int i1, di, j1, dj, jl; // first line, row step, first column, column step, last column in scan order
};
__host__ __device__ inline PassGeom pass_geom(const Dims &d, int pass)
{
    PassGeom g;
    if (pass == 0) { g.i1 = 0; g.di = 1; g.j1 = 0; g.dj = 1; g.jl = d.Wp - 1; }
    else { g.i1 = d.Hp - 1; g.di = -1; g.j1 = d.Wp - 1; g.dj = -1; g.jl = 0; }
    return g;
}

// number of valid packed registers of this lane (disparities lane*2NR + 2k, +1 are valid for k < nvalid)
template <int NR> __device__ __forceinline__ int lane_nvalid(int D, int lane)
{
    int n = D / 2 - lane * NR;
    return n < 0 ? 0 : (n > NR ? NR : n);
}

template <int NR> __device__ __forceinline__ void load_cost(const uint8_t *__restrict__ pix, int lane, int nvalid, uint32_t (&c)[NR])
{
    const uint16_t *p = reinterpret_cast<const uint16_t *>(pix) + lane * NR;
#pragma unroll
    for (int k = 0; k < NR; k++) c[k] = (k < nvalid) ? __byte_perm((uint32_t)p[k], 0u, 0x4140) : 0u;
}

__device__ __forceinline__ unsigned warp_min_u16x2(uint32_t m2)
{
    unsigned m = min(m2 & 0xFFFFu, m2 >> 16);
    return __reduce_min_sync(kFull, m);
}

// One SGM step for one path at one pixel. A: normalised state of the predecessor (pad registers = kInf2).
// Writes L (pad registers = kInf2) and returns min_d L.
template <int NR>
__device__ __forceinline__ unsigned path_step(const uint32_t (&A)[NR], const uint32_t (&c)[NR], int lane, int nvalid, uint32_t (&L)[NR])
{
    uint32_t up = __shfl_up_sync(kFull, A[NR - 1], 1);
    uint32_t dn = __shfl_down_sync(kFull, A[0], 1);
    if (lane == 0) up = kInf2;
    if (lane == 31) dn = kInf2;
    uint32_t E[NR + 1]; // E[k] = (d-1 of the low half, low half) ; E[k+1] = (high half, d+1 of the high half)
    E[0] = __byte_perm(up, A[0], 0x5432);
#pragma unroll
    for (int k = 1; k < NR; k++) E[k] = __byte_perm(A[k - 1], A[k], 0x5432);
    E[NR] = __byte_perm(A[NR - 1], dn, 0x5432);
    uint32_t m2 = kInf2;
#pragma unroll
    for (int k = 0; k < NR; k++) {
        uint32_t x = __viaddmin_u16x2(E[k], kP1x2, A[k]);
        uint32_t y = __viaddmin_u16x2(E[k + 1], kP1x2, kP2x2);
        uint32_t l = __vminu2(x, y) + c[k];
        L[k] = (k < nvalid) ? l : kInf2;
        m2 = __vminu2(m2, L[k]);
    }
    return warp_min_u16x2(m2);
}

template <int NR> __device__ __forceinline__ void normalise(const uint32_t (&L)[NR], unsigned m, int nvalid, uint32_t (&A)[NR])
{
    const uint32_t mm = m * 0x10001u;
#pragma unroll
    for (int k = 0; k < NR; k++) A[k] = (k < nvalid) ? (L[k] - mm) : kInf2;
}

template <int NR> __device__ __forceinline__ void sum_add(uint16_t *__restrict__ spix, int lane, int nvalid, const uint32_t (&L)[NR])
{
    uint32_t *p = reinterpret_cast<uint32_t *>(spix) + lane * NR;
#pragma unroll
    for (int k = 0; k < NR; k++)
        if (k < nvalid) p[k] += L[k];
}

// ------------------------------------------------------------------------------------ first line, path r0
// One warp per launch: the horizontal path on the first line of a pass (sgm.cpp:141-190): plain int arithmetic,
// then saturate_cast<uint16>(uint8) truncation (types.h:28) -- the state carried along the line is the truncated
// value, not normalised.
template <int NR> __global__ void __launch_bounds__(32) k_sgm_first_line(const uint8_t *__restrict__ fused, Dims d, int pass, uint16_t *__restrict__ sum)
{
    const int lane = threadIdx.x;
    const int nvalid = lane_nvalid<NR>(d.D, lane);
    const PassGeom g = pass_geom(d, pass);
    uint32_t Lq[NR], c[NR];
    unsigned m = 0;
    for (int j = g.j1, n = 0; n < d.Wp; j += g.dj, n++) {
        const size_t pix = (size_t)g.i1 * d.Wp + j;
        load_cost<NR>(fused + pix * d.D, lane, nvalid, c);
        uint32_t nw[NR];
        if (n == 0) {
#pragma unroll
            for (int k = 0; k < NR; k++) nw[k] = (k < nvalid) ? c[k] : kInf2;
        } else {
            uint32_t up = __shfl_up_sync(kFull, Lq[NR - 1], 1);
            uint32_t dn = __shfl_down_sync(kFull, Lq[0], 1);
            if (lane == 0) up = kInf2;
            if (lane == 31) dn = kInf2;
            uint32_t E[NR + 1];
            E[0] = __byte_perm(up, Lq[0], 0x5432);
#pragma unroll
            for (int k = 1; k < NR; k++) E[k] = __byte_perm(Lq[k - 1], Lq[k], 0x5432);
            E[NR] = __byte_perm(Lq[NR - 1], dn, 0x5432);
            const uint32_t mm = m * 0x10001u, p2 = mm + kP2x2;
#pragma unroll
            for (int k = 0; k < NR; k++) {
                uint32_t x = __viaddmin_u16x2(E[k], kP1x2, Lq[k]);
                uint32_t y = __viaddmin_u16x2(E[k + 1], kP1x2, p2);
                uint32_t t = __vminu2(x, y) - mm;
                nw[k] = (k < nvalid) ? ((c[k] + t) & 0x00FF00FFu) : kInf2;
            }
        }
        uint32_t m2 = kInf2;
#pragma unroll
        for (int k = 0; k < NR; k++) { Lq[k] = nw[k]; m2 = __vminu2(m2, nw[k]); }
        m = warp_min_u16x2(m2);
        sum_add<NR>(sum + pix * d.D, lane, nvalid, nw);
    }
}

// ------------------------------------------------------------------------------------ path chains
// One warp follows one path chain of one direction through the frame and adds its L into the sum volume.
// (Correctness-first decomposition: every path is an independent kernel; paths commute because nothing saturates.)
//   path 0: r0, predecessor (i, j - dj)       chains = rows other than the first line, start state A = 0
//   path 1: r1, predecessor (i - di, j - dj)  chains start on the first line (state from C) or at column j1 (A = inf)
//   path 2: r2, predecessor (i - di, j)       chains start on the first line
//   path 3: r3, predecessor (i - di, j + dj)  chains start on the first line or at the last column (A = inf)
template <int NR>
__global__ void __launch_bounds__(256) k_sgm_chains(const uint8_t *__restrict__ fused, Dims d, int pass, int path, uint16_t *__restrict__ sum)
{
    const int lane = threadIdx.x & 31;
    const int chain = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const PassGeom g = pass_geom(d, pass);
    const int row_off = (pass == 0) ? 1 : 0; // rows other than the first line: row_off .. row_off + Hp - 2
    int i, j, si, sj, init; // init 0: A = 0; 1: first-line cell; 2: A = inf
    if (path == 0) {
        if (chain >= d.Hp - 1) return;
        i = chain + row_off; j = g.j1; si = 0; sj = g.dj; init = 0;
    } else if (path == 2) {
        if (chain >= d.Wp) return;
        i = g.i1; j = chain; si = g.di; sj = 0; init = 1;
    } else {
        if (chain >= d.Wp + d.Hp - 1) return;
        si = g.di; sj = (path == 1) ? g.dj : -g.dj;
        if (chain < d.Wp) { i = g.i1; j = chain; init = 1; }
        else { i = chain - d.Wp + row_off; j = (path == 1) ? g.j1 : g.jl; init = 2; }
    }
    const int nvalid = lane_nvalid<NR>(d.D, lane);
    uint32_t A[NR], c[NR], L[NR];
#pragma unroll
    for (int k = 0; k < NR; k++) A[k] = (init == 0 && k < nvalid) ? 0u : kInf2;
    bool first_line_cell = (init == 1);
    while (i >= 0 && i < d.Hp && j >= 0 && j < d.Wp) {
        const size_t pix = (size_t)i * d.Wp + j;
        load_cost<NR>(fused + pix * d.D, lane, nvalid, c);
        if (first_line_cell) {
            uint32_t m2 = kInf2;
#pragma unroll
            for (int k = 0; k < NR; k++) { L[k] = (k < nvalid) ? c[k] : kInf2; m2 = __vminu2(m2, L[k]); }
            normalise<NR>(L, warp_min_u16x2(m2), nvalid, A);
            first_line_cell = false;
        } else {
            unsigned m = path_step<NR>(A, c, lane, nvalid, L);
            sum_add<NR>(sum + pix * d.D, lane, nvalid, L);
            normalise<NR>(L, m, nvalid, A);
        }
        i += si; j += sj;
    }
}

template <int NR> static void launch_sgm_nr(const uint8_t *fused, const Dims &d, uint16_t *sum, cudaStream_t st, LaunchCounter &lc)
{
    cudaMemsetAsync(sum, 0, (size_t)d.cells * sizeof(uint16_t), st);
    const int wpb = 8;
    for (int pass = 0; pass < 2; pass++) {
        k_sgm_first_line<NR><<<1, 32, 0, st>>>(fused, d, pass, sum);
        lc.add();
        for (int path = 0; path < 4; path++) {
            int chains = path == 0 ? d.Hp - 1 : path == 2 ? d.Wp : d.Wp + d.Hp - 1;
            k_sgm_chains<NR><<<(chains + wpb - 1) / wpb, wpb * 32, 0, st>>>(fused, d, pass, path, sum);
            lc.add();
        }
    }
}

// disparities per lane = 2 * NR, chosen so that D fits in 32 lanes
static int pick_nr(int D) { return (D + 63) / 64; }

void launch_sgm(const uint8_t *fused, const Dims &d, uint16_t *sum, int *status, cudaStream_t st, LaunchCounter &lc)
{
    (void)status;
    switch (pick_nr(d.D)) {
    case 1: launch_sgm_nr<1>(fused, d, sum, st, lc); break;
    case 2: launch_sgm_nr<2>(fused, d, sum, st, lc); break;
    case 3: launch_sgm_nr<3>(fused, d, sum, st, lc); break;
    case 4: launch_sgm_nr<4>(fused, d, sum, st, lc); break;
    case 5: launch_sgm_nr<5>(fused, d, sum, st, lc); break;
    case 6: launch_sgm_nr<6>(fused, d, sum, st, lc); break;
    case 7: launch_sgm_nr<7>(fused, d, sum, st, lc); break;
    default: launch_sgm_nr<8>(fused, d, sum, st, lc); break;
    }
}

// ------------------------------------------------------------------------------------ final WTA + encode
// WTALeft_SSE with uniqueness 1 (hpp:283): first-index argmin over d <= min(j, D-1); then convertTo(CV_16UC1),
// crop Rect(D, D, W, H) and * 255 with saturation (hpp:111-118). One warp per pixel.
template <int NR>
__global__ void __launch_bounds__(256) k_select(const uint16_t *__restrict__ sum, Dims d, int16_t *__restrict__ raw_disp, uint16_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    const int nvalid = lane_nvalid<NR>(d.D, lane);
    for (long long pix = warp0; pix < d.px; pix += nw) {
        const int i = (int)(pix / d.Wp), j = (int)(pix % d.Wp);
        const int dmax = min(j, d.D - 1);
        const uint32_t *p = reinterpret_cast<const uint32_t *>(sum + pix * d.D) + lane * NR;
        unsigned best = 0xFFFFFFFFu;
#pragma unroll
        for (int k = 0; k < NR; k++) {
            if (k < nvalid) {
                const uint32_t s2 = p[k];
                const int d0 = lane * 2 * NR + 2 * k;
                if (d0 <= dmax) best = min(best, ((s2 & 0xFFFFu) << 16) | (unsigned)d0);
                if (d0 + 1 <= dmax) best = min(best, (s2 & 0xFFFF0000u) | (unsigned)(d0 + 1));
            }
        }
        best = __reduce_min_sync(kFull, best);
        if (lane == 0) {
            const int disp = (int)(best & 0xFFFFu);
            if (raw_disp) raw_disp[pix] = (int16_t)disp;
            const int oi = i - d.D, oj = j - d.D;
            if (out && oi >= 0 && oi < d.H && oj >= 0 && oj < d.W) out[(size_t)oi * d.W + oj] = (uint16_t)min(disp * 255, 65535);
        }
    }
}

void launch_select(const uint16_t *sum, const Dims &d, int16_t *raw_disp, uint16_t *out, cudaStream_t st, LaunchCounter &lc)
{
    const int blocks = 148 * 8;
#define SISTER_SELECT_CASE(N) case N: k_select<N><<<blocks, 256, 0, st>>>(sum, d, raw_disp, out); break;
    switch (pick_nr(d.D)) {
        SISTER_SELECT_CASE(1) SISTER_SELECT_CASE(2) SISTER_SELECT_CASE(3) SISTER_SELECT_CASE(4)
        SISTER_SELECT_CASE(5) SISTER_SELECT_CASE(6) SISTER_SELECT_CASE(7)
    default: k_select<8><<<blocks, 256, 0, st>>>(sum, d, raw_disp, out); break;
    }
#undef SISTER_SELECT_CASE
    lc.add();
}

} // namespace sister
