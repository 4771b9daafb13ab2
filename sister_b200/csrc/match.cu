// sister_b200 / match.cu -- per-view matching, confidence masks and the fused cost volume (sm_100a).
//
//   k_match_wta   Hamming cost popc64(c1[r][a] ^ c2[r][a-d]) (census.cpp:54-89) evaluated ONCE per (a, b = a-d)
//                 pair and reduced both ways: over d for the left map (postprocess.cpp:74-185) and along the
//                 anti-diagonal for the right map (postprocess.cpp:187-315). The raw volume is never stored.
//   k_median      median3x3_SSE called in place (hpp:198-199): recursive flat-array semantics.
//   k_lrc_mask    doLRCheck (postprocess.cpp:318-341) + the mask loops of hpp:201-251.
//   k_fuse        C(i,j,d) = sum_v mask_v(i,j) * cost_v(i,j,d) (hpp:255-277) as uint8.
//
// Why uint8 is exact for C: mask_v = 0 wherever the view-frame column is < D (hpp:203) and wherever the raw
// left map is 0 (rows 0..2 and h-2,h-1 of a raw volume are constant -> argmin 0 -> masked), so the 255 marker
// (census.cpp:76,95-98) never survives the masking; a census code holds 31 antisymmetric bit pairs, one always-0
// centre bit and the carry bit, so two codes differ in at most 63 bits; 4 * 63 = 252. k_fuse still evaluates the
// reference's formula literally and raises kStatusFusedOverflow if a sum ever exceeded 255.
#include "kernels.cuh"

namespace sister {

// ---------------------------------------------------------------------------------------------- WTA L/R

// grid (max(hv), 4), block 256, dynamic smem: 2 * wv u64 + wv u32
__global__ void __launch_bounds__(256) k_match_wta(const unsigned long long *__restrict__ census, Dims d, unsigned view_mask,
                                                   int16_t *__restrict__ wtaL, int16_t *__restrict__ wtaR)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int v = blockIdx.y, r = blockIdx.x;
    if (!((view_mask >> v) & 1u)) return;
    const int hv = view_rows(d, v), wv = view_cols(d, v);
    if (r >= hv) return;
    int16_t *outL = wtaL + (size_t)v * d.px + (size_t)r * wv;
    int16_t *outR = wtaR + (size_t)v * d.px + (size_t)r * wv;
    const int tid = threadIdx.x;
    if (r < 3 || r >= hv - 2) {
        // rows 0..2 are all 255 (census.cpp:95-98,142-145), rows h-2,h-1 are never written (defined 0):
        // constant cost -> first-index argmin 0 for both maps
        for (int c = tid; c < wv; c += blockDim.x) { outL[c] = 0; outR[c] = 0; }
        return;
    }
    unsigned long long *c1 = reinterpret_cast<unsigned long long *>(smem_raw);
    unsigned long long *c2 = c1 + wv;
    unsigned *rkey = reinterpret_cast<unsigned *>(c2 + wv);
    const unsigned long long *g1 = census + (size_t)(2 * v) * d.px + (size_t)r * wv;
    const unsigned long long *g2 = census + (size_t)(2 * v + 1) * d.px + (size_t)r * wv;
    for (int c = tid; c < wv; c += blockDim.x) { c1[c] = g1[c]; c2[c] = g2[c]; rkey[c] = 0xFFFFFFFFu; }
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int D = d.D;
    const int nblk = (wv + 31) / 32;
    for (int blk = warp; blk < nblk; blk += nwarps) {
        const int A = blk * 32, a = A + lane;
        const bool a_ok = a < wv;
        const unsigned long long x1 = a_ok ? c1[a] : 0ull;
        unsigned lkey = 0xFFFFFFFFu;
        const int b_lo = max(0, A - D + 1), b_hi = min(A + 31, wv - 1);
        for (int b = b_lo; b <= b_hi; b++) {
            const unsigned long long x2 = c2[b]; // broadcast
            const int dd = a - b;
            const unsigned cost = __popcll(x1 ^ x2);
            const bool ok = a_ok && dd >= 0 && dd < D;
            const unsigned key = ok ? ((cost << 16) | (unsigned)dd) : 0xFFFFFFFFu;
            lkey = min(lkey, key);
            const unsigned rk = __reduce_min_sync(0xFFFFFFFFu, key);
            if (lane == 0 && rk != 0xFFFFFFFFu) atomicMin(&rkey[b], rk);
        }
        if (a_ok) outL[a] = (int16_t)(lkey & 0xFFFFu);
    }
    __syncthreads();
    for (int c = tid; c < wv; c += blockDim.x) outR[c] = (int16_t)(rkey[c] & 0xFFFFu);
}

void launch_match_wta(const unsigned long long *census, const Dims &d, unsigned view_mask, int16_t *wtaL, int16_t *wtaR,
                      cudaStream_t st, LaunchCounter &lc)
{
    int m = d.Wp > d.Hp ? d.Wp : d.Hp;
    size_t smem = (size_t)m * (8 + 8 + 4);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(k_match_wta, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_done = true;
    }
    dim3 grid(m, 4);
    k_match_wta<<<grid, 256, smem, st>>>(census, d, view_mask, wtaL, wtaR);
    lc.add();
}

// ---------------------------------------------------------------------------------------------- median

__device__ __forceinline__ void sort2(int &a, int &b)
{
    int lo = min(a, b), hi = max(a, b);
    a = lo; b = hi;
}

// the 19-exchange network of postprocess.cpp:52-58
__device__ __forceinline__ int median9(int v0, int v1, int v2, int v3, int v4, int v5, int v6, int v7, int v8)
{
    sort2(v1, v2); sort2(v4, v5); sort2(v7, v8);
    sort2(v0, v1); sort2(v3, v4); sort2(v6, v7);
    sort2(v1, v2); sort2(v4, v5); sort2(v7, v8);
    sort2(v0, v3); sort2(v5, v8); sort2(v4, v7);
    sort2(v3, v6); sort2(v1, v4); sort2(v2, v5);
    sort2(v4, v7); sort2(v4, v2); sort2(v6, v4);
    sort2(v4, v2);
    return v4;
}

// One block per map (8 maps: L and R of 4 views). Flat-array recursion (see oracle/sister_oracle.c
// so_median_inplace): out[p] = med9(out[p-w-1..p-w+1], raw[p-1..p+1], raw[p+w-1..p+w+1]) for
// p in [w+1, N-w-5], out[w] = 0, everything else unchanged. A row depends on the finished row above, and its
// last element on its own first element (flat wrap-around), hence the two phases per row.
// grid 8, block 1024, dynamic smem 3 * wv int16
__global__ void __launch_bounds__(1024) k_median(const int16_t *__restrict__ wtaL, const int16_t *__restrict__ wtaR, Dims d,
                                                 unsigned view_mask, int16_t *__restrict__ medL, int16_t *__restrict__ medR)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int m = blockIdx.x, v = m >> 1;
    if (!((view_mask >> v) & 1u)) return;
    const int hv = view_rows(d, v), wv = view_cols(d, v);
    const int16_t *raw = ((m & 1) ? wtaR : wtaL) + (size_t)v * d.px;
    int16_t *out = ((m & 1) ? medR : medL) + (size_t)v * d.px;
    int16_t *rows = reinterpret_cast<int16_t *>(smem_raw); // 3 rotating rows of filtered output
    const long long N = (long long)hv * wv;
    const long long p_lo = wv + 1, p_hi = N - wv - 5;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int c = tid; c < wv; c += nt) { int16_t x = raw[c]; rows[c] = x; out[c] = x; }
    __syncthreads();
    for (int r = 1; r < hv; r++) {
        int16_t *cur = rows + (size_t)(r % 3) * wv;
        const int16_t *prev = rows + (size_t)((r + 2) % 3) * wv;  // row r-1
        const int16_t *prev2 = rows + (size_t)((r + 1) % 3) * wv; // row r-2 (valid for r >= 2)
        const long long base = (long long)r * wv;
        for (int c = tid; c < wv; c += nt) {
            const long long p = base + c;
            int val;
            if (p == wv) val = 0; // the zero-initialised lastMedian lands here (postprocess.cpp:29,61-63)
            else if (p < p_lo || p > p_hi) val = raw[p];
            else if (c == wv - 1) continue; // phase 2
            else {
                int a0 = (c == 0) ? prev2[wv - 1] : prev[c - 1];
                val = median9(a0, prev[c], prev[c + 1], raw[p - 1], raw[p], raw[p + 1], raw[p + wv - 1], raw[p + wv], raw[p + wv + 1]);
            }
            cur[c] = (int16_t)val;
        }
        __syncthreads();
        if (tid == 0) {
            const long long p = base + wv - 1;
            if (p >= p_lo && p <= p_hi && p != wv)
                cur[wv - 1] = (int16_t)median9(prev[wv - 2], prev[wv - 1], cur[0], raw[p - 1], raw[p], raw[p + 1], raw[p + wv - 1], raw[p + wv], raw[p + wv + 1]);
        }
        __syncthreads();
        for (int c = tid; c < wv; c += nt) out[base + c] = cur[c];
    }
}

// grid (ceil(wv/256), max(hv), 4), block 256
__global__ void __launch_bounds__(256) k_lrc_mask(const int16_t *__restrict__ medL, const int16_t *__restrict__ medR, Dims d,
                                                  unsigned view_mask, int16_t *__restrict__ lr_final, uint8_t *__restrict__ masks)
{
    const int v = blockIdx.z, r = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
    if (!((view_mask >> v) & 1u)) return;
    const int hv = view_rows(d, v), wv = view_cols(d, v);
    if (r >= hv || c >= wv) return;
    const size_t off = (size_t)v * d.px + (size_t)r * wv;
    int b = medL[off + c];
    if (b >= 0 && b <= c) { // postprocess.cpp:327
        int mt = medR[off + c - b];
        int diff = b - mt;
        if (abs(diff) > kLrcThreshold) b = -10;
    } else {
        b = -10;
    }
    lr_final[off + c] = (int16_t)b;
    int i, j;
    view_to_image(d, v, r, c, i, j);
    masks[(size_t)v * d.px + (size_t)i * d.Wp + j] = !(b <= 0 || c < d.D); // hpp:203
}

void launch_median_lrc_mask(const int16_t *wtaL, const int16_t *wtaR, const Dims &d, unsigned view_mask, int16_t *medL,
                            int16_t *medR, int16_t *lr_final, uint8_t *masks, cudaStream_t st, LaunchCounter &lc)
{
    int m = d.Wp > d.Hp ? d.Wp : d.Hp;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(k_median, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_done = true;
    }
    k_median<<<8, 1024, (size_t)3 * m * sizeof(int16_t), st>>>(wtaL, wtaR, d, view_mask, medL, medR);
    lc.add();
    dim3 grid((m + 255) / 256, m, 4);
    k_lrc_mask<<<grid, 256, 0, st>>>(medL, medR, d, view_mask, lr_final, masks);
    lc.add();
}

// ---------------------------------------------------------------------------------------------- fuse

// Block = T x T image tile. For each view the tile's matching partners are T view-frame lines of
// (T + D - 1) consecutive census codes; they are staged in shared memory once and every (pixel, d) cell of the
// tile is evaluated from there. One warp per pixel at a time, lanes over d, byte stores coalesced along d.
// grid (ceil(Wp/T), ceil(Hp/T)), block 256, dynamic smem 4 * T * (T + D) u64
__global__ void __launch_bounds__(256) k_fuse(const unsigned long long *__restrict__ census, const uint8_t *__restrict__ masks,
                                              Dims d, unsigned view_mask, int T, uint8_t *__restrict__ fused, int *__restrict__ status)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *s = reinterpret_cast<unsigned long long *>(smem_raw);
    const int D = d.D, P = T + D; // line pitch (T + D - 1 used)
    const int i0 = blockIdx.y * T, j0 = blockIdx.x * T;
    const int tid = threadIdx.x;
    int cv_lo[4] = {j0, d.Wp - j0 - T, d.Hp - i0 - T, i0};
    for (int v = 0; v < 4; v++) {
        if (!((view_mask >> v) & 1u)) continue;
        const int hv = view_rows(d, v), wv = view_cols(d, v);
        const unsigned long long *c2 = census + (size_t)(2 * v + 1) * d.px;
        unsigned long long *sv = s + (size_t)v * T * P;
        const int n = T * (P - 1);
        for (int e = tid; e < n; e += blockDim.x) {
            const int line = e / (P - 1), x = e % (P - 1);
            const int rv = (v < 2) ? i0 + line : d.Wp - 1 - (j0 + line);
            const int col = cv_lo[v] - (D - 1) + x;
            unsigned long long val = 0;
            if (rv >= 0 && rv < hv && col >= 0 && col < wv) val = c2[(size_t)rv * wv + col];
            sv[line * P + x] = val;
        }
    }
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    bool overflow = false;
    for (int pidx = warp; pidx < T * T; pidx += nwarps) {
        const int li = pidx / T, lj = pidx % T;
        const int i = i0 + li, j = j0 + lj;
        if (i >= d.Hp || j >= d.Wp) continue;
        const size_t pix = (size_t)i * d.Wp + j;
        // per-view setup (warp-uniform)
        unsigned long long c1[4];
        int e0[4], cv[4], kind[4]; // kind 0: skip, 1: popc, 2: all 255, 3: all 0
        const unsigned long long *line_ptr[4];
#pragma unroll
        for (int v = 0; v < 4; v++) {
            kind[v] = 0; c1[v] = 0; e0[v] = 0; cv[v] = 0; line_ptr[v] = s;
            if (!((view_mask >> v) & 1u)) continue;
            if (!masks[(size_t)v * d.px + pix]) continue;
            const int hv = view_rows(d, v), wv = view_cols(d, v);
            int rv, cc;
            image_to_view(d, v, i, j, rv, cc);
            cv[v] = cc;
            if (rv < 3) kind[v] = 2;              // census.cpp:95-98,142-145
            else if (rv >= hv - 2) kind[v] = 3;   // never written by the reference: defined 0
            else {
                kind[v] = 1;
                c1[v] = census[(size_t)(2 * v) * d.px + (size_t)rv * wv + cc];
                e0[v] = cc - cv_lo[v] + D - 1;
                line_ptr[v] = s + (size_t)v * T * P + (size_t)((v < 2) ? li : lj) * P;
            }
        }
        uint8_t *dst = fused + pix * D;
        for (int dd = lane; dd < D; dd += 32) {
            unsigned sum = 0;
#pragma unroll
            for (int v = 0; v < 4; v++) {
                if (kind[v] == 1) sum += (dd > cv[v]) ? (unsigned)kInvalidCost : (unsigned)__popcll(c1[v] ^ line_ptr[v][e0[v] - dd]);
                else if (kind[v] == 2) sum += kInvalidCost;
            }
            if (sum > 255u) { overflow = true; sum = 255u; }
            dst[dd] = (uint8_t)sum;
        }
    }
    if (overflow) atomicOr(status, kStatusFusedOverflow);
}

void launch_fuse(const unsigned long long *census, const uint8_t *masks, const Dims &d, unsigned view_mask, uint8_t *fused,
                 int *status, cudaStream_t st, LaunchCounter &lc)
{
    int T = 16;
    while (T > 4 && (size_t)4 * T * (T + d.D) * 8 > 200 * 1024) T /= 2;
    size_t smem = (size_t)4 * T * (T + d.D) * 8;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(k_fuse, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        attr_done = true;
    }
    dim3 grid((d.Wp + T - 1) / T, (d.Hp + T - 1) / T);
    k_fuse<<<grid, 256, smem, st>>>(census, masks, d, view_mask, T, fused, status);
    lc.add();
}

} // namespace sister
