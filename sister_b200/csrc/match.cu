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

constexpr unsigned kNoKey = 0xFFFFFFFFu;
constexpr int kMatchK = 4; // consecutive view columns per lane

// One pass over the (a, d) pairs of a block of 32 * K view columns a = A + lane * K + k, d = 0 .. D-1.
// Every pair's Hamming cost popc64(c1[a] ^ c2[a - d]) is evaluated ONCE and feeds both reductions as the key
// (cost << 16) | d, whose minimum is the first-index argmin of WTALeft_SSE / WTARight_SSE (postprocess.cpp:74-315,
// uniqueness test dead at 1.0):
//   left map  L[a]: the lane's own running minimum over d;
//   right map R[b], b = a - d: a systolic accumulator that travels with b. At step d a lane's K columns meet the K
//     partners b = a0 - d .. a0 - d + K - 1; one step later that window has slid down by one, so the top partner (its
//     census code and its accumulator) moves to the next lane by one shuffle, lane 0 takes a new partner from shared
//     memory and lane 31 retires one accumulator into rkey[] with a shared-memory atomicMin. The window is a rotating
//     register file (slot (k - s) mod K at sub-step s = d mod K), so nothing else moves.
// CHECK = false when every pair of the block is in range (A >= D - 1 and A + 32K <= wv).
__device__ __forceinline__ void red_min_shared_if(unsigned addr, unsigned val, bool pred)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p red.shared.min.u32 [%0], %1;\n\t}\n" ::"r"(addr), "r"(val), "r"((unsigned)pred) : "memory");
}
// 8-byte asynchronous global -> shared copy; nothing is read and zeros are written when !valid
__device__ __forceinline__ void cp_async8_zfill(unsigned dst_s, const void *src, bool valid)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst_s), "l"(src), "r"(valid ? 8 : 0) : "memory");
}
__device__ __forceinline__ uint2 lds64(unsigned addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];\n" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}

template <int K, bool CHECK>
__device__ __forceinline__ void match_block(unsigned c1_s, unsigned c2_s, unsigned rkey_s, int wv, int D, int A, int lane,
                                            int16_t *__restrict__ outL)
{
    // c1_s, c2_s, rkey_s: shared-memory addresses of the row's census codes (8 bytes each) and of the R keys (4 bytes)
    const int a0 = A + lane * K;
    unsigned x1lo[K], x1hi[K], wlo[K], whi[K], racc[K], lkey[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        const bool in = !CHECK || a0 + k < wv;
        const uint2 x = in ? lds64(c1_s + 8u * (a0 + k)) : make_uint2(0u, 0u), y = in ? lds64(c2_s + 8u * (a0 + k)) : make_uint2(0u, 0u);
        x1lo[k] = x.x; x1hi[k] = x.y;
        wlo[k] = y.x; whi[k] = y.y;
        racc[k] = kNoKey; lkey[k] = kNoKey;
    }
    const bool last_lane = lane == 31, first_lane = lane == 0;
    // step dd: lane 31 retires rkey[a0 + K - dd], lane 0 takes c2[A - dd]; both walk down one element per step
    unsigned top_s = rkey_s + 4u * (a0 + K); // minus 4 * dd
    unsigned new_s = c2_s + 8u * A;          // minus 8 * dd
#pragma unroll 1
    for (int t = 0; t < D / K; t++) {
#pragma unroll
        for (int s = 0; s < K; s++) {
            const int dd = t * K + s;
            const int p = (K - s) % K; // slot of the partner that leaves the window / of the one that enters
            if (s > 0 || t > 0) {
                const int b_top = a0 - dd + K; // partner that was on top at step dd - 1
                red_min_shared_if(top_s - 4u * s, racc[p], last_lane && (!CHECK || (b_top >= 0 && b_top < wv)));
                uint2 y = make_uint2(0u, 0u);
                if (!CHECK || A - dd >= 0) y = lds64(new_s - 8u * s); // warp-uniform address: one broadcast read
                const unsigned slo = __shfl_up_sync(0xFFFFFFFFu, wlo[p], 1);
                const unsigned shi = __shfl_up_sync(0xFFFFFFFFu, whi[p], 1);
                const unsigned sac = __shfl_up_sync(0xFFFFFFFFu, racc[p], 1);
                wlo[p] = first_lane ? y.x : slo;
                whi[p] = first_lane ? y.y : shi;
                racc[p] = first_lane ? kNoKey : sac;
            }
            const unsigned kd = (unsigned)dd;
#pragma unroll
            for (int k = 0; k < K; k++) {
                const int q = (k - s + K) % K;
                unsigned key = (unsigned)__popc(x1lo[k] ^ wlo[q]) * 65536u + kd;
                key = (unsigned)__popc(x1hi[k] ^ whi[q]) * 65536u + key;
                if (CHECK && (a0 + k - dd < 0 || a0 + k >= wv)) key = kNoKey;
                lkey[k] = min(lkey[k], key);
                racc[q] = min(racc[q], key);
            }
        }
        top_s -= 4u * K;
        new_s -= 8u * K;
    }
    // retire what is still in the window: logical position e holds b = a0 - (D - 1) + e, slot (e - (K - 1)) mod K
#pragma unroll
    for (int e = 0; e < K; e++) {
        const int b = a0 - (D - 1) + e;
        const unsigned v = racc[(e - (K - 1) + K) % K];
        red_min_shared_if(rkey_s + 4u * (unsigned)b, v, !CHECK || (b >= 0 && b < wv));
    }
    if (!CHECK) {
        static_assert(K == 4, "the vector store below assumes 4 columns per lane");
        short4 o = make_short4((short)(lkey[0] & 0xFFFFu), (short)(lkey[1] & 0xFFFFu), (short)(lkey[2] & 0xFFFFu), (short)(lkey[3] & 0xFFFFu));
        *reinterpret_cast<short4 *>(outL + a0) = o; // a0 % 4 == 0 and wv % 4 == 0: 8-byte aligned
    } else {
#pragma unroll
        for (int k = 0; k < K; k++)
            if (a0 + k < wv) outL[a0 + k] = (int16_t)(lkey[k] & 0xFFFFu);
    }
}

// grid (max(hv) / shares, 4), block <= 224 threads (see launch_match_wta), dynamic smem: 2 * wv u64 + wv u32; five or more blocks per SM
template <bool WIDE> // WIDE: rows too long for five blocks' worth of shared memory per SM -- three blocks of up to 13 warps
__global__ void __launch_bounds__(WIDE ? 416 : 224, WIDE ? 3 : 5) k_match_wta(const unsigned long long *__restrict__ census, Dims d, unsigned view_mask,
                                                    int16_t *__restrict__ wtaL, int16_t *__restrict__ wtaR, int share, int n_shares)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int v = blockIdx.y;
    if (!((view_mask >> v) & 1u)) return;
    const int hv = view_rows(d, v), wv = view_cols(d, v);
    // (row bands: this launch matches the rows [hv * share / n, hv * (share + 1) / n) of every view)
    const int r = blockIdx.x + share_row0(hv, share, n_shares);
    if (r >= share_row0(hv, share + 1, n_shares)) return;
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
    for (int c = tid; c < wv; c += blockDim.x) { c1[c] = g1[c]; c2[c] = g2[c]; rkey[c] = kNoKey; }
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int D = d.D;
    constexpr int kCols = 32 * kMatchK;
    const int nblk = (wv + kCols - 1) / kCols;
    const unsigned c1_s = (unsigned)__cvta_generic_to_shared(c1), c2_s = (unsigned)__cvta_generic_to_shared(c2);
    const unsigned rkey_s = (unsigned)__cvta_generic_to_shared(rkey);
    for (int blk = warp; blk < nblk; blk += nwarps) {
        const int A = blk * kCols;
        if (A >= D - 1 && A + kCols <= wv) match_block<kMatchK, false>(c1_s, c2_s, rkey_s, wv, D, A, lane, outL);
        else match_block<kMatchK, true>(c1_s, c2_s, rkey_s, wv, D, A, lane, outL);
    }
    __syncthreads();
    for (int c = tid; c < wv; c += blockDim.x) outR[c] = (int16_t)(rkey[c] & 0xFFFFu);
}

void launch_match_wta(const unsigned long long *census, const Dims &d, unsigned view_mask, int16_t *wtaL, int16_t *wtaR,
                      cudaStream_t st, LaunchCounter &lc, int share, int n_shares)
{
    int m = d.Wp > d.Hp ? d.Wp : d.Hp;
    const size_t smem = (size_t)m * (8 + 8 + 4);
    // a row's column blocks (32 * kMatchK columns each) are dealt to at most 7 warps in as few equal rounds as possible: five
    // or more small blocks per SM keep the POPC pipe busier through the blocks' load / store phases than three of 13 warps
    // (0.936 -> 0.906 ms at c2) -- as long as five blocks' rows fit the SM's shared memory (frames up to ~2300 wide)
    const int nblk = (m + 32 * kMatchK - 1) / (32 * kMatchK);
    dim3 grid((m + n_shares - 1) / n_shares, 4);
    if (smem * 5 <= 220 * 1024) {
        const int rounds = (nblk + 6) / 7, warps = (nblk + rounds - 1) / rounds;
        lc.fail(optin_dynamic_smem((const void *)k_match_wta<false>, smem));
        k_match_wta<false><<<grid, 32 * warps, smem, st>>>(census, d, view_mask, wtaL, wtaR, share, n_shares);
    } else {
        const int warps = nblk < 13 ? nblk : 13;
        lc.fail(optin_dynamic_smem((const void *)k_match_wta<true>, smem));
        k_match_wta<true><<<grid, 32 * warps, smem, st>>>(census, d, view_mask, wtaL, wtaR, share, n_shares);
    }
    lc.add();
}

// ---------------------------------------------------------------------------------------------- median

// Median of nine on two pixels at once (packed s16x2). The value of a median does not depend on the network that finds
// it, so instead of the 19 compare-exchanges of postprocess.cpp:52-58 (38 min / max): sort the three columns, then
// med3(max of the minima, med3 of the middles, min of the maxima). With three-input packed min / max (VIMNMX3) and the
// middle of three as a ^ b ^ c ^ min ^ max (the three outputs are a permutation of the three inputs) that is 22
// instructions, and the kernel is bound by exactly these.
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) { return a ^ b ^ c; }
__device__ __forceinline__ void sort3p(uint32_t a, uint32_t b, uint32_t c, uint32_t &lo, uint32_t &mid, uint32_t &hi)
{
    lo = __vimin3_s16x2(a, b, c);
    hi = __vimax3_s16x2(a, b, c);
    mid = xor3(xor3(a, b, c), lo, hi);
}
__device__ __forceinline__ uint32_t med3p(uint32_t a, uint32_t b, uint32_t c)
{
    return xor3(xor3(a, b, c), __vimin3_s16x2(a, b, c), __vimax3_s16x2(a, b, c));
}
// v0 v1 v2 / v3 v4 v5 / v6 v7 v8: rows of the 3 x 3 window; columns are (v0, v3, v6), (v1, v4, v7), (v2, v5, v8)
__device__ __forceinline__ uint32_t median9p(uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3, uint32_t v4, uint32_t v5, uint32_t v6,
                                             uint32_t v7, uint32_t v8)
{
    uint32_t l0, m0, h0, l1, m1, h1, l2, m2, h2;
    sort3p(v0, v3, v6, l0, m0, h0);
    sort3p(v1, v4, v7, l1, m1, h1);
    sort3p(v2, v5, v8, l2, m2, h2);
    return med3p(__vimax3_s16x2(l0, l1, l2), med3p(m0, m1, m2), __vimin3_s16x2(h0, h1, h2));
}

// One block per map (8 maps: L and R of 4 views). Flat-array recursion (see oracle/sister_oracle.c
// so_median_inplace): out[p] = med9(out[p-w-1..p-w+1], raw[p-1..p+1], raw[p+w-1..p+w+1]) for
// p in [w+1, N-w-5], out[w] = 0, everything else unchanged. A row depends on the finished row above and its last
// element on its own first element (flat wrap-around), so the recurrence is one block barrier per row and the work
// between two barriers has to be as short as possible:
//   * a thread filters TWO adjacent pixels at once with packed s16x2 min/max (VIMNMX.S16x2);
//   * the raw rows are streamed through a flat shared-memory ring (kMedRing rows, cp.async kMedAhead rows ahead) and
//     the filtered rows through a flat ring of 4 rows, so that "previous element of column 0" and "next element of
//     column w-1" are plain flat neighbours, exactly as in the reference's flat pointer walk;
//   * one thread does the first pair and then the last pair of the row (which needs the row's first output).
//   * warps are given roles by scheduler: the border warp alone on one, the interior warps on the other three, and the
//     warps without a role leave at once (the barrier counts the ones that stay).
// grid 8, block 1024, dynamic smem (kMedRing + 4) * wv int16
constexpr int kMedRing = 8, kMedAhead = 6;

__global__ void __launch_bounds__(1024) k_median(const int16_t *__restrict__ wtaL, const int16_t *__restrict__ wtaR, Dims d,
                                                 unsigned view_mask, int16_t *__restrict__ medL, int16_t *__restrict__ medR)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int m = blockIdx.x, v = m >> 1;
    if (!((view_mask >> v) & 1u)) return;
    const int hv = view_rows(d, v), wv = view_cols(d, v);
    const int16_t *raw = ((m & 1) ? wtaR : wtaL) + (size_t)v * d.px;
    int16_t *out = ((m & 1) ? medR : medL) + (size_t)v * d.px;
    int16_t *ring = reinterpret_cast<int16_t *>(smem_raw);  // raw rows, flat, modulo kMedRing * wv
    int16_t *filt = ring + (size_t)kMedRing * wv;           // filtered rows, flat, modulo 4 * wv
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
    const int RS = kMedRing * wv, FS = 4 * wv;
    const long long N = (long long)hv * wv;
    const long long p_lo = wv + 1, p_hi = N - wv - 5;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // Roles by scheduler (a warp runs on scheduler warp & 3): warp 3 is the border warp and has scheduler 3 to itself -- the
    // other warps of that scheduler leave at once, so that not even their loop bookkeeping competes with it -- and the 24
    // warps of schedulers 0..2 are the interior threads, numbered ti = 0 .. nt-1.
    // Only as many interior warps stay as the row has pairs of pairs (13 of 24 at 1664 columns): every warp that stays
    // pays the row loop's bookkeeping on its scheduler.
    if ((warp & 3) == 3 && warp != 3) return;
    const bool interior = (warp & 3) != 3;
    const int iw = warp - (warp >> 2);                          // interior warp number 0..23
    int n_iw = ((((wv - 4) >> 1) + 1) / 2 + 31) / 32;           // warps needed for two pairs per thread
    n_iw = n_iw < 1 ? 1 : n_iw > 24 ? 24 : n_iw;
    if (interior && iw >= n_iw) return;
    const int ti = iw * 32 + lane, nt = n_iw * 32;
    const int n_sync = nt + 32;                                 // the interior warps that stay + the border warp
    auto block_sync = [&]() { asm volatile("bar.sync 1, %0;\n" ::"r"(n_sync) : "memory"); };
    const int chunks = wv >> 2; // 8-byte chunks per row (wv % 4 == 0, postprocess.cpp:18)
    auto stage = [&](int row) {
        if (row < hv) {
            const int16_t *src = raw + (size_t)row * wv;
            const unsigned dst = ring_s + 2u * (unsigned)((row % kMedRing) * wv);
            for (int c = interior ? ti : chunks; c < chunks; c += nt)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst + 8u * c), "l"(src + 4 * c) : "memory");
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    for (int row = 0; row < kMedAhead; row++) stage(row);
    // row 0 is copied unchanged
    asm volatile("cp.async.wait_group %0;\n" ::"n"(kMedAhead - 1) : "memory");
    block_sync();
    for (int c = interior ? ti : wv; c < wv; c += nt) { const int16_t x = ring[c]; filt[c] = x; out[c] = x; }
    const uint32_t *ring32 = reinterpret_cast<const uint32_t *>(ring);
    const uint32_t *filt32 = reinterpret_cast<const uint32_t *>(filt);
    for (int r = 1; r < hv; r++) {
        // rows <= r + 2 must have landed: rows 0 .. kMedAhead + r - 2 are committed
        asm volatile("cp.async.wait_group %0;\n" ::"n"(kMedAhead - 4) : "memory");
        block_sync(); // publishes the raw rows and the filtered row r - 1; everyone is done with row r - 1
        stage(kMedAhead + r - 1);
        const long long base = (long long)r * wv;
        const int fprev = ((r - 1) & 3) * wv, fcur = (r & 3) * wv; // flat ring index of column 0 of rows r-1, r
        const int rb = (r % kMedRing) * wv;                         // raw ring index of (r, 0)
        const int rb1 = ((r + 1) % kMedRing) * wv;                  // raw ring index of (r+1, 0)
        const bool plain_row = r >= 2 && r <= hv - 3;               // no position of this row is special
        // ---- interior pairs (c, c+1), c even in [2, wv-4]: all nine neighbours lie inside the rows' own slots ----
        // Plain rows (all but the first and the last two): the warps of three of the SM's four schedulers take two adjacent
        // pairs per thread -- 8-byte loads, and the sorted column between the two pairs is shared -- and leave the fourth
        // scheduler to the border warp below, whose dependent chain (column w-1 waits for column 0) is the row's critical
        // path and otherwise gets one issue slot in eight.
        if (plain_row) {
            if (interior) {
                const int hmax = (wv - 4) >> 1; // pairs 1 .. hmax
                const uint2 *fp2 = reinterpret_cast<const uint2 *>(filt32 + (fprev >> 1));
                const uint2 *rq2 = reinterpret_cast<const uint2 *>(ring32 + (rb >> 1));
                const uint2 *rs2 = reinterpret_cast<const uint2 *>(ring32 + (rb1 >> 1));
                uint32_t *fcw = reinterpret_cast<uint32_t *>(filt) + (fcur >> 1);
                uint32_t *orow = reinterpret_cast<uint32_t *>(out + base);
                for (int t = ti; 1 + 2 * t <= hmax; t += nt) {
                    // words 2t .. 2t+3 of the three window rows: pairs h0 = 2t+1 (a b c) and h1 = 2t+2 (b c d)
                    const uint2 p0 = fp2[t], p1 = fp2[t + 1], q0 = rq2[t], q1 = rq2[t + 1], s0 = rs2[t], s1 = rs2[t + 1];
                    uint32_t l[5], m[5], h[5];
                    sort3p(__byte_perm(p0.x, p0.y, 0x5432), __byte_perm(q0.x, q0.y, 0x5432), __byte_perm(s0.x, s0.y, 0x5432), l[0], m[0], h[0]);
                    sort3p(p0.y, q0.y, s0.y, l[1], m[1], h[1]);
                    sort3p(__byte_perm(p0.y, p1.x, 0x5432), __byte_perm(q0.y, q1.x, 0x5432), __byte_perm(s0.y, s1.x, 0x5432), l[2], m[2], h[2]);
                    sort3p(p1.x, q1.x, s1.x, l[3], m[3], h[3]);
                    sort3p(__byte_perm(p1.x, p1.y, 0x5432), __byte_perm(q1.x, q1.y, 0x5432), __byte_perm(s1.x, s1.y, 0x5432), l[4], m[4], h[4]);
                    const uint32_t v0 = med3p(__vimax3_s16x2(l[0], l[1], l[2]), med3p(m[0], m[1], m[2]), __vimin3_s16x2(h[0], h[1], h[2]));
                    const uint32_t v1 = med3p(__vimax3_s16x2(l[2], l[3], l[4]), med3p(m[2], m[3], m[4]), __vimin3_s16x2(h[2], h[3], h[4]));
                    fcw[2 * t + 1] = v0; orow[2 * t + 1] = v0;
                    if (2 * t + 2 <= hmax) { fcw[2 * t + 2] = v1; orow[2 * t + 2] = v1; }
                }
            }
        } else
        for (int c = interior ? 2 + 2 * ti : wv; c <= wv - 4; c += 2 * nt) {
            const int h = c >> 1;
            const uint32_t pa = filt32[(fprev >> 1) + h - 1], pb = filt32[(fprev >> 1) + h], pc = filt32[(fprev >> 1) + h + 1];
            const uint32_t qa = ring32[(rb >> 1) + h - 1], qb = ring32[(rb >> 1) + h], qc = ring32[(rb >> 1) + h + 1];
            const uint32_t sa = ring32[(rb1 >> 1) + h - 1], sb = ring32[(rb1 >> 1) + h], sc = ring32[(rb1 >> 1) + h + 1];
            uint32_t val;
            if (r < hv - 1)
                val = median9p(__byte_perm(pa, pb, 0x5432), pb, __byte_perm(pb, pc, 0x5432), __byte_perm(qa, qb, 0x5432), qb,
                               __byte_perm(qb, qc, 0x5432), __byte_perm(sa, sb, 0x5432), sb, __byte_perm(sb, sc, 0x5432));
            else
                val = qb;
            if (!plain_row) { // per-pixel exceptions: p < p_lo or p > p_hi keep the raw value
                const long long p = base + c;
                if (p < p_lo || p > p_hi) val = (val & 0xFFFF0000u) | (qb & 0xFFFFu);
                if (p + 1 < p_lo || p + 1 > p_hi) val = (val & 0xFFFFu) | (qb & 0xFFFF0000u);
            }
            reinterpret_cast<uint32_t *>(filt)[(fcur >> 1) + h] = val;
            *reinterpret_cast<uint32_t *>(out + base + c) = val;
        }
        // ---- the two border pairs, literal (flat neighbours wrap through the rings): lanes 0..3 of the last warp take
        // columns 0, 1, w-2, w-1. Column w-1 needs this row's column-0 output as ONE of its nine inputs: the median of
        // nine with one unknown x is clamp(x, k3, k4) with k3, k4 the 4th and 5th smallest of the other eight, and
        // those are the network's outputs for x = -inf and x = +inf -- evaluated together as the two packed halves, in
        // parallel with column 0, so the row's critical path is one network, not two. ----
        if (warp == 3) {
            auto F = [&](int idx) -> int { if (idx < 0) idx += FS; if (idx >= FS) idx -= FS; return filt[idx]; };
            auto R = [&](int idx) -> int { if (idx < 0) idx += RS; if (idx >= RS) idx -= RS; return ring[idx]; };
            auto dup = [](int x) -> uint32_t { return ((uint32_t)x & 0xFFFFu) * 0x10001u; };
            const int c = lane == 0 ? 0 : lane == 1 ? 1 : lane == 2 ? wv - 2 : wv - 1;
            const long long p = base + c;
            const bool is_median = lane < 4 && p != wv && p >= p_lo && p <= p_hi;
            uint32_t packed = 0;
            int val = 0;
            if (lane < 4) {
                if (is_median) {
                    const uint32_t a2 = (lane == 3) ? 0x7FFF8000u /* (lo: -32768, hi: +32767) */ : dup(F(fprev + c + 1));
                    packed = median9p(dup(F(fprev + c - 1)), dup(F(fprev + c)), a2, dup(R(rb + c - 1)), dup(R(rb + c)), dup(R(rb + c + 1)),
                                      dup(R(rb + c + wv - 1)), dup(R(rb + c + wv)), dup(R(rb + c + wv + 1)));
                    val = (int)(int16_t)(packed & 0xFFFFu);
                } else if (p != wv) {
                    val = R(rb + c); // p < p_lo or p > p_hi keeps the raw value; p == w is the zero of postprocess.cpp:29,61-63
                }
            }
            const int first = __shfl_sync(0xFFFFFFFFu, val, 0); // this row's column-0 output
            if (lane == 3 && is_median) {
                const int k3 = (int)(int16_t)(packed & 0xFFFFu), k4 = (int)(int16_t)(packed >> 16);
                val = min(max(first, k3), k4);
            }
            if (lane < 4) {
                filt[fcur + c] = (int16_t)val;
                out[base + c] = (int16_t)val;
            }
        }
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

// ---- the same recurrence with the warps running free behind each other (padded frames of 256 .. 5120 columns). In flat
// terms out[p] depends on out[p - w - 1 .. p - w + 1]: a column chunk of a row needs the same chunk of the row above and one
// pixel of either neighbouring chunk -- not the whole row. So warp c owns the 64 K columns [64 K c, 64 K (c + 1)) (K packed
// words = 2 K pixels per lane; K = 2 up to 2048 columns, K = 4 above), walks down the rows, and waits only for its two
// neighbours to have finished the row above (the first chunk's left neighbour is the last chunk two rows up, the last
// chunk's right neighbour is the first chunk of the SAME row: the flat wrap-around of postprocess.cpp:31-67). No block
// barrier: a row costs one neighbour hand-over (an mbarrier phase per warp and row, two barriers per warp alternating with
// the row parity so that a waiter can never be two phases behind). The filtered row above stays in the lanes' registers;
// neighbouring lanes exchange their edge pixels with shuffles, neighbouring warps through one boundary word per side and
// row in shared memory (ring of 4 rows). Raw rows are streamed per warp with cp.async (kMed2Depth rows ahead) into the
// warp's own ring of [32 x K words | left halo word | right halo word].
// What bounds it (measured with clock64 per phase, c2): a warp's row is ONE serial instruction stream -- about 120
// instructions at 4-7 cycles each when a warp runs alone on its scheduler -- and the rows are serial; the hand-over itself
// (arrive -> the neighbour's try_wait returns) is the smaller part. Hence four pixels per lane rather than two (26 warps
// were slower than 13: the per-row overhead is per warp) and everything besides the median itself kept to a handful of
// instructions: shared memory is addressed in its own window with offsets carried from row to row, the boundary loads and
// stores are one instruction for the whole warp with per-lane addresses (no divergent branch), the two special rows are a
// uniform branch.
constexpr int kMed2Depth = 8, kMed2MaxWarps = 20;
__host__ __device__ constexpr int med2_slot_bytes(int K) { return 32 * 4 * K + 16; } // one raw row of a warp: words, 2 halo words, pad

__device__ __forceinline__ void med_mbar_init(unsigned a, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(c) : "memory"); }
__device__ __forceinline__ void med_mbar_arrive(unsigned a) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(a) : "memory"); }
// bounded by time: a broken pipeline must never hang the device (returns false after 4 s)
__device__ __forceinline__ bool med_mbar_wait(unsigned a, unsigned parity)
{
    unsigned ok;
    unsigned long long t0 = 0;
    for (unsigned n = 0;; n++) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return true;
        if ((n & 255u) != 255u) continue;
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
        if (t0 == 0) t0 = t;
        else if (t - t0 > 4000000000ull) return false;
    }
}

__device__ __forceinline__ uint32_t med_lds32(unsigned a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void med_sts32(unsigned a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(a), "r"(v) : "memory"); }
template <int K> __device__ __forceinline__ void med_lds_words(unsigned a, uint32_t (&w)[K])
{
    if constexpr (K == 2) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];\n" : "=r"(w[0]), "=r"(w[1]) : "r"(a) : "memory");
    else asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(a) : "memory");
}
template <int K> __device__ __forceinline__ void med_stg_words(int16_t *g, const uint32_t (&w)[K])
{
    if constexpr (K == 2) *reinterpret_cast<uint2 *>(g) = make_uint2(w[0], w[1]);
    else *reinterpret_cast<uint4 *>(g) = make_uint4(w[0], w[1], w[2], w[3]);
}

// grid 8 (one block per map), block 32 * ceil(max(wv) / (64 K)), dynamic smem: barriers, boundary words, the warps' rings
template <int K>
__global__ void __launch_bounds__(K == 2 ? 512 : 32 * kMed2MaxWarps) k_median_chunks(const int16_t *__restrict__ wtaL, const int16_t *__restrict__ wtaR,
                                                                      Dims d, unsigned view_mask, int16_t *__restrict__ medL,
                                                                      int16_t *__restrict__ medR, int *__restrict__ status)
{
    constexpr int kSlot = med2_slot_bytes(K), kCols = 64 * K, kWin = 2 * K + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // [warp][row parity] barriers | [row & 3][warp][first word, last word] | rings
    const unsigned bars_s = (unsigned)__cvta_generic_to_shared(smem_raw);
    const unsigned bnd_s = bars_s + 16u * kMed2MaxWarps;
    constexpr unsigned kBndRow = 8u * kMed2MaxWarps;
    const int m = blockIdx.x, v = m >> 1;
    if (!((view_mask >> v) & 1u)) return;
    const int hv = view_rows(d, v), wv = view_cols(d, v);
    const int lane = threadIdx.x & 31, c = threadIdx.x >> 5, nch = (wv + kCols - 1) / kCols, L = nch - 1; // (the block is sized for the wider orientation)
    if (threadIdx.x < 2 * nch) med_mbar_init(bars_s + 8u * threadIdx.x, 32);
    __syncthreads();
    if (c >= nch) return;
    const int N = hv * wv, p_lo = wv + 1, p_hi = N - wv - 5;
    const int x = kCols * c + 2 * K * lane;                               // the lane's columns x .. x + 2K - 1
    const int nl = min(32, (wv - kCols * c) / (2 * K)), last = nl - 1;    // lanes of this chunk that hold pixels (>= 2)
    const bool on = lane < nl, first_lane = lane == 0, last_lane = lane == last;
    const unsigned ring_s = bnd_s + 4u * kBndRow + (unsigned)(c * (kMed2Depth * kSlot));
    constexpr unsigned kRingBytes = kMed2Depth * kSlot;
    // ---- staging of the raw rows. Row j -> slot j % depth: 32 x K words, then the word before the chunk and the word after
    // it (flat neighbours, whatever row they belong to; nothing outside [0, N) is touched: the word before row 0 of chunk 0
    // and the word after the last row of the last chunk do not exist)
    const char *gsrc = reinterpret_cast<const char *>(((m & 1) ? wtaR : wtaL) + (size_t)v * d.px + x);
    const int halo_src = first_lane ? -4 : 4 * K;
    const unsigned own_dst = 4u * K * lane, halo_dst = 128u * K + (first_lane ? 0u : 4u);
    const unsigned halo_rd = (first_lane || last_lane) ? halo_dst : own_dst; // (the other lanes do not use what they read here)
    const int halo_j0 = first_lane ? (c == 0 ? 1 : 0) : (last_lane ? 0 : hv);           // rows [halo_j0, halo_j1] have the halo word
    const int halo_j1 = (last_lane && !first_lane && c == L) ? hv - 2 : hv - 1;
    const size_t row_bytes = (size_t)wv * 2;
    unsigned st_off = 0; // slot of the next row to stage
    int st_j = 0;
    auto stage = [&]() {
        if (st_j < hv) {
            if (on) {
                if constexpr (K == 2) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(ring_s + st_off + own_dst), "l"(gsrc) : "memory");
                else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(ring_s + st_off + own_dst), "l"(gsrc) : "memory");
            }
            if (st_j >= halo_j0 && st_j <= halo_j1)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(ring_s + st_off + halo_dst), "l"(gsrc + halo_src) : "memory");
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        gsrc += row_bytes;
        st_j++;
        st_off += kSlot;
        if (st_off == kRingBytes) st_off = 0;
    };
    // the window words of the next raw row for this lane: (x-1 x) (x x+1) (x+1 x+2) ... (x+2K-1 x+2K)
    unsigned rd_off = 0;
    auto raw_row = [&](uint32_t (&w)[kWin]) {
        uint32_t own[K];
        med_lds_words<K>(ring_s + rd_off + own_dst, own);
        const uint32_t halo = med_lds32(ring_s + rd_off + halo_rd);
        uint32_t left = __shfl_up_sync(0xFFFFFFFFu, own[K - 1], 1), right = __shfl_down_sync(0xFFFFFFFFu, own[0], 1);
        if (first_lane) left = halo;
        if (last_lane) right = halo;
        w[0] = __byte_perm(left, own[0], 0x5432);
#pragma unroll
        for (int i = 0; i < K; i++) {
            w[2 * i + 1] = own[i];
            w[2 * i + 2] = __byte_perm(own[i], i + 1 < K ? own[i + 1 < K ? i + 1 : i] : right, 0x5432);
        }
        rd_off += kSlot;
        if (rd_off == kRingBytes) rd_off = 0;
    };
    // ---- hand-over with the neighbouring warps. Left: warp c - 1 at row r - 1 (chunk 0: the last chunk at row r - 2);
    // right: warp c + 1 at row r - 1 (the last chunk: chunk 0 at row r). Lane 0 takes the left boundary word, every other
    // lane the right one (only the chunk's last lane uses it).
    const int lw = c > 0 ? c - 1 : L, ld = c > 0 ? 1 : 2, rw = c < L ? c + 1 : 0, rdl = c < L ? 1 : 0;
    const int b_delta = first_lane ? ld : rdl;
    const unsigned b_base = bnd_s + (first_lane ? 8u * lw + 4u : 8u * rw);
    const unsigned my_bnd = bnd_s + 8u * c + (first_lane ? 0u : 4u);
    const bool bnd_writer = first_lane || last_lane;
    int16_t *gout = ((m & 1) ? medR : medL) + (size_t)v * d.px + x;

    for (int j = 0; j < kMed2Depth - 1; j++) stage();
    bool broken = false;
    uint32_t q[kWin], s[kWin], f[K]; // raw row r, raw row r + 1, this lane's part of the finished row above
#pragma unroll
    for (int k = 0; k < kWin; k++) s[k] = 0u;
    // row 0 is copied unchanged
    asm volatile("cp.async.wait_group %0;\n" ::"n"(kMed2Depth - 2) : "memory");
    raw_row(q);
#pragma unroll
    for (int i = 0; i < K; i++) f[i] = q[2 * i + 1];
    if (on) med_stg_words<K>(gout, f);
    if (bnd_writer) med_sts32(my_bnd, first_lane ? f[0] : f[K - 1]);
    med_mbar_arrive(bars_s + 8u * (2 * c));
    stage();
    asm volatile("cp.async.wait_group %0;\n" ::"n"(kMed2Depth - 2) : "memory");
    raw_row(q);
    for (int r = 1; r < hv; r++) {
        stage();
        gout += wv;
        asm volatile("cp.async.wait_group %0;\n" ::"n"(kMed2Depth - 2) : "memory"); // rows <= r + 1 have landed
        if (r + 1 < hv) raw_row(s);
        uint32_t fl = __shfl_up_sync(0xFFFFFFFFu, f[K - 1], 1), fr = __shfl_down_sync(0xFFFFFFFFu, f[0], 1);
        uint32_t val[K];
#pragma unroll
        for (int i = 0; i < K; i++) val[i] = q[2 * i + 1];
        if (r < hv - 1) {
            // the row above: this chunk is in the registers; one pixel of either neighbouring chunk
            if (!broken) {
                bool ok = true;
                const int rl = r - ld, rr = r - rdl;
                if (rl >= 0) ok = med_mbar_wait(bars_s + 8u * (2 * lw + (rl & 1)), (unsigned)((rl >> 1) & 1));
                if (ok) ok = med_mbar_wait(bars_s + 8u * (2 * rw + (rr & 1)), (unsigned)((rr >> 1) & 1));
                if (!ok) {
                    broken = true;
                    if (lane == 0) atomicOr(status, kStatusSpinTimeout);
                }
            }
            const uint32_t edge = med_lds32(b_base + kBndRow * ((unsigned)(r - b_delta) & 3u)); // (row 1, chunk 0, lane 0: unused, out[w] = 0)
            if (first_lane) fl = edge;
            if (last_lane) fr = edge;
            uint32_t l[kWin], md[kWin], h[kWin];
            sort3p(__byte_perm(fl, f[0], 0x5432), q[0], s[0], l[0], md[0], h[0]);
#pragma unroll
            for (int i = 0; i < K; i++) {
                sort3p(f[i], q[2 * i + 1], s[2 * i + 1], l[2 * i + 1], md[2 * i + 1], h[2 * i + 1]);
                sort3p(__byte_perm(f[i], i + 1 < K ? f[i + 1 < K ? i + 1 : i] : fr, 0x5432), q[2 * i + 2], s[2 * i + 2], l[2 * i + 2], md[2 * i + 2], h[2 * i + 2]);
            }
#pragma unroll
            for (int i = 0; i < K; i++)
                val[i] = med3p(__vimax3_s16x2(l[2 * i], l[2 * i + 1], l[2 * i + 2]), med3p(md[2 * i], md[2 * i + 1], md[2 * i + 2]),
                               __vimin3_s16x2(h[2 * i], h[2 * i + 1], h[2 * i + 2]));
            if (r == 1 || r == hv - 2) { // the first and the last filtered positions: postprocess.cpp:29,61-63
                const int base = r * wv + x;
#pragma unroll
                for (int i = 0; i < K; i++) {
                    const int p = base + 2 * i;
                    const uint32_t rawv = q[2 * i + 1];
                    if (p == wv) val[i] &= 0xFFFF0000u;                                   // out[w] = 0
                    else if (p < p_lo || p > p_hi) val[i] = (val[i] & 0xFFFF0000u) | (rawv & 0xFFFFu);
                    if (p + 1 < p_lo || p + 1 > p_hi) val[i] = (val[i] & 0xFFFFu) | (rawv & 0xFFFF0000u);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < K; i++) f[i] = val[i];
        if (bnd_writer) med_sts32(my_bnd + kBndRow * ((unsigned)r & 3u), first_lane ? f[0] : f[K - 1]);
        // (the last row has no dependencies and nobody waits for it: arriving for it could put this warp two phases ahead
        // of a neighbour that still waits for row hv - 3 on the same barrier)
        if (r < hv - 1) med_mbar_arrive(bars_s + 8u * (2 * c + (r & 1)));
        if (on) med_stg_words<K>(gout, f);
#pragma unroll
        for (int k = 0; k < kWin; k++) q[k] = s[k];
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

template <int K>
static void launch_median_chunks(const int16_t *wtaL, const int16_t *wtaR, const Dims &d, unsigned view_mask, int16_t *medL, int16_t *medR,
                                 int *status, int m, cudaStream_t st, LaunchCounter &lc)
{
    const int warps = (m + 64 * K - 1) / (64 * K);
    const size_t smem = (size_t)16 * kMed2MaxWarps + (size_t)4 * 8 * kMed2MaxWarps + (size_t)warps * kMed2Depth * med2_slot_bytes(K);
    if (smem > 48 * 1024) lc.fail(optin_dynamic_smem((const void *)k_median_chunks<K>, smem));
    k_median_chunks<K><<<8, 32 * warps, smem, st>>>(wtaL, wtaR, d, view_mask, medL, medR, status);
    lc.add();
}

// grid (ceil(max(wv) / 32), ceil(max(hv) / 32), 4), block (32, 8): a 32 x 32 tile of the view frame. The masks live in the
// image frame (hpp:203-252 un-flips / un-transposes them): for the two transposed views a tile's bytes go through shared
// memory so that they, too, leave as runs of 32 consecutive bytes instead of one byte per image row.
__global__ void __launch_bounds__(256) k_lrc_mask(const int16_t *__restrict__ medL, const int16_t *__restrict__ medR, Dims d,
                                                  unsigned view_mask, int16_t *__restrict__ lr_final, uint8_t *__restrict__ masks)
{
    __shared__ uint8_t tile[32][36];
    const int v = blockIdx.z, tx = threadIdx.x, ty = threadIdx.y;
    if (!((view_mask >> v) & 1u)) return;
    const int hv = view_rows(d, v), wv = view_cols(d, v);
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    if (r0 >= hv || c0 >= wv) return;
    const int c = c0 + tx;
#pragma unroll
    for (int rr = ty; rr < 32; rr += 8) {
        const int r = r0 + rr;
        uint8_t mk = 0;
        if (r < hv && c < wv) {
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
            mk = !(b <= 0 || c < d.D); // hpp:203
            if (v < 2) {
                int i, j;
                view_to_image(d, v, r, c, i, j);
                masks[(size_t)v * d.px + (size_t)i * d.Wp + j] = mk;
            }
        }
        tile[rr][tx] = mk;
    }
    if (v < 2) return;
    __syncthreads();
    // one view column = one image row; the lanes take the tile's view rows = 32 consecutive image columns
    const int r = r0 + tx;
#pragma unroll
    for (int cc = ty; cc < 32; cc += 8) {
        const int cv = c0 + cc;
        if (r < hv && cv < wv) {
            int i, j;
            view_to_image(d, v, r, cv, i, j);
            masks[(size_t)v * d.px + (size_t)i * d.Wp + j] = tile[tx][cc];
        }
    }
}

void launch_median_lrc_mask(const int16_t *wtaL, const int16_t *wtaR, const Dims &d, unsigned view_mask, int16_t *medL,
                            int16_t *medR, int16_t *lr_final, uint8_t *masks, int *status, cudaStream_t st, LaunchCounter &lc)
{
    int m = d.Wp > d.Hp ? d.Wp : d.Hp;
    const int lo = d.Wp < d.Hp ? d.Wp : d.Hp;
    // warps free-running behind each other, one per 128 (256) columns; the block is sized for the wider of the two frame
    // orientations and the warps a narrower map does not need leave at once
    if (lo >= 256 && m <= 2048 && d.Wp % 8 == 0 && d.Hp % 8 == 0) {
        launch_median_chunks<2>(wtaL, wtaR, d, view_mask, medL, medR, status, m, st, lc);
    } else if (lo >= 512 && m <= 256 * kMed2MaxWarps && d.Wp % 16 == 0 && d.Hp % 16 == 0) {
        launch_median_chunks<4>(wtaL, wtaR, d, view_mask, medL, medR, status, m, st, lc);
    } else {
        const size_t med_smem = (size_t)(kMedRing + 4) * m * sizeof(int16_t);
        lc.fail(optin_dynamic_smem((const void *)k_median, med_smem));
        k_median<<<8, 1024, med_smem, st>>>(wtaL, wtaR, d, view_mask, medL, medR);
        lc.add();
    }
    dim3 grid((m + 31) / 32, (m + 31) / 32, 4);
    k_lrc_mask<<<grid, dim3(32, 8), 0, st>>>(medL, medR, d, view_mask, lr_final, masks);
    lc.add();
}

// ---------------------------------------------------------------------------------------------- fuse

// Block = T x T image tile, T warps; warp w owns tile row w and walks its T pixels. For each view the tile's matching
// partners are T view-frame lines of (T + D - 1) consecutive census codes; they are staged in shared memory once
// (views whose mask is 0 on the whole tile are skipped) and every (pixel, d) cell of the tile is evaluated from
// there: lanes over d (d = lane + 32k), consecutive lanes read consecutive 8-byte codes (conflict-free), the D cost
// bytes of a pixel leave as full 32-byte sectors.
// Fast path (the only one the masks ever select, see the header comment): the view row is an ordinary census row and
// the view column is >= D - 1, so every d has a partner and the cost is a plain popcount. Anything else takes the
// literal per-cell formula of census.cpp:63-88,95-98 / hpp:264-276 below.
// grid (ceil(Wp/T), ceil(Hp/T)), block 32 * T, dynamic smem 4 * T * (T + D) u64 + 4 * T * T u64 + T * T bytes
// ORDER: 0 = the cell's byte order is read from Dims at run time; 1 = natural; 2 = lane-interleaved with lpc 16 and
// nr = NK (D = 32 * nr: the per-k disparity offsets are then compile-time constants and fold into the load addresses)
// TRIPLE: one pass produces the fused volumes of all three modes (hpp:262-276): the horizontal pair (right + left) into
// fused_h, the vertical pair (top + bottom) into fused_v and their sum, the multiview volume, into fused -- every Hamming
// distance is evaluated once instead of once per mode that contains its view. The two partial sums ride in the two halves
// of one accumulator (<= 2 * 255 each).
template <int T, int NK, int ORDER, bool TRIPLE> // NK = D / 32 when D is a multiple of 32 (fully unrolled), 0 = any D
__global__ void __launch_bounds__(32 * T) k_fuse(const unsigned long long *__restrict__ census, const uint8_t *__restrict__ masks,
                                                 Dims d, unsigned view_mask, uint8_t *__restrict__ fused, uint8_t *__restrict__ fused_h,
                                                 uint8_t *__restrict__ fused_v, int *__restrict__ status, int row_lo, int row_hi)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned s_any;
    unsigned long long *s = reinterpret_cast<unsigned long long *>(smem_raw);
    const int D = d.D, P = T + D; // line pitch (T + D - 1 used)
    unsigned long long *sc1 = s + (size_t)4 * T * P;        // the tile's own (centre) codes, per view: [v][li][lj]
    uint8_t *smask = reinterpret_cast<uint8_t *>(sc1 + 4 * T * T);
    const int i0 = row_lo + blockIdx.y * T, j0 = blockIdx.x * T;
    const int tid = threadIdx.x, nthr = 32 * T;
    if (tid == 0) s_any = 0;
    __syncthreads();
    {
        unsigned mine = 0;
        for (int e = tid; e < T * T; e += nthr) {
            const int i = i0 + e / T, j = j0 + e % T;
            unsigned m = 0;
            if (i < d.Hp && j < d.Wp) {
                const size_t pix = (size_t)i * d.Wp + j;
#pragma unroll
                for (int v = 0; v < 4; v++)
                    if (((view_mask >> v) & 1u) && masks[(size_t)v * d.px + pix]) m |= 1u << v;
            }
            smask[e] = (uint8_t)m;
            mine |= m;
        }
        if (mine) atomicOr(&s_any, mine);
    }
    __syncthreads();
    const unsigned any = s_any;
    const int cv_lo[4] = {j0, d.Wp - j0 - T, d.Hp - i0 - T, i0};
#pragma unroll
    for (int v = 0; v < 4; v++) {
        if (!((any >> v) & 1u)) continue;
        const int hv = view_rows(d, v), wv = view_cols(d, v);
        const unsigned long long *c2 = census + (size_t)(2 * v + 1) * d.px;
        unsigned long long *sv = s + (size_t)v * T * P;
        // warp w stages line w: consecutive lanes, consecutive codes
        const int line = tid >> 5;
        const int rv = (v < 2) ? i0 + line : d.Wp - 1 - (j0 + line);
        const bool row_ok = rv >= 0 && rv < hv;
        const unsigned long long *src = c2 + (size_t)(row_ok ? rv : 0) * wv;
        const int col0 = cv_lo[v] - (D - 1);
        // asynchronous copies (zero-filled outside the view): the 4 x 7 loads of a thread are all in flight at once instead of
        // each waiting for its shared-memory store
        for (int x = tid & 31; x < P - 1; x += 32) {
            const int col = col0 + x;
            const bool in = row_ok && col >= 0 && col < wv;
            cp_async8_zfill((unsigned)__cvta_generic_to_shared(sv + line * P + x), src + (in ? col : 0), in);
        }
        // the centre codes of the tile's pixels (row `line` of the tile)
        if ((tid & 31) < T) {
            const int i = i0 + line, j = j0 + (tid & 31);
            unsigned long long c = 0;
            if (i < d.Hp && j < d.Wp) {
                int rv2, cc2;
                image_to_view(d, v, i, j, rv2, cc2);
                c = __ldg(census + (size_t)(2 * v) * d.px + (size_t)rv2 * wv + cc2);
            }
            sc1[(v * T + line) * T + (tid & 31)] = c;
        }
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    __syncthreads();
    const int lane = tid & 31, li = tid >> 5;
    const int i = i0 + li;
    if (i >= row_hi) return; // (row_hi <= Hp; a band's volumes end with its last row)
    constexpr int NKC = NK > 0 ? NK : 16;
    bool overflow = false;
    // The lane writes bytes lane + 32k of the cell; which disparity that is depends on the cell's byte order (Dims,
    // common.cuh). It separates into a lane part and a per-k part (lpc is 8 or 16, so a block of 32 bytes = 8 words never
    // straddles a row of the word matrix): d = dl + dku(k). Within a k the 32 lanes still read 32 distinct 8-byte census
    // codes whose indices cover every residue mod 16 twice (once per half warp): conflict-free like the natural order.
    const int dl = d.interleaved ? 2 * d.nr * (lane >> 2) + (lane & 1) + d.nr * ((lane >> 1) & 1) : lane;
    auto dku = [&](int k) {
        if constexpr (ORDER == 1) return 32 * k;
        else if constexpr (ORDER == 2) return 2 * NK * ((8 * k) & 15) + 2 * ((8 * k) >> 4);
        else return d.interleaved ? 2 * d.nr * ((8 * k) & (d.lpc - 1)) + 2 * ((8 * k) >> d.lpc_shift) : 32 * k;
    };
    // Interior tiles (all but a frame of width ~D): every cell of every active view is a plain popcount, so the row loop
    // needs no geometry at all -- one running shared-memory pointer per view (+-1 code per pixel for the horizontal views,
    // one staged line per pixel for the vertical ones), a zero cell leaves as 16-byte stores, the 8-bit range is checked
    // once per pixel. 244 -> ~95 instructions per pixel and warp; the popcounts (XU pipe) are then what is left.
    if constexpr (NK > 0) {
        auto plain_at = [&](int v, int ii, int jj) {
            int rv, cc;
            image_to_view(d, v, ii, jj, rv, cc);
            return rv >= 3 && rv < view_rows(d, v) - 2 && cc >= D - 1;
        };
        bool interior = i0 + T <= d.Hp && j0 + T <= d.Wp;
#pragma unroll
        for (int v = 0; v < 4; v++)
            if ((any >> v) & 1u)
                interior = interior && plain_at(v, i0, j0) && plain_at(v, i0, j0 + T - 1) && plain_at(v, i0 + T - 1, j0) && plain_at(v, i0 + T - 1, j0 + T - 1);
        if (interior) { // block-uniform
            const unsigned s_base = (unsigned)__cvta_generic_to_shared(s), c1_base = (unsigned)__cvta_generic_to_shared(sc1);
            // code index of the d = 0 partner of pixel (li, lj = 0) in view v's staged lines, minus the lane's disparity part
            unsigned pv[4];
            pv[0] = s_base + 8u * (unsigned)((0 * T + li) * P + (D - 1) - dl);
            pv[1] = s_base + 8u * (unsigned)((1 * T + li) * P + (T - 1) + (D - 1) - dl);
            pv[2] = s_base + 8u * (unsigned)((2 * T + 0) * P + (T - 1 - li) + (D - 1) - dl);
            pv[3] = s_base + 8u * (unsigned)((3 * T + 0) * P + li + (D - 1) - dl);
            const int pstep[4] = {8, -8, 8 * P, 8 * P};
            unsigned koff[NKC];
#pragma unroll
            for (int k = 0; k < NKC; k++) koff[k] = 8u * (unsigned)dku(k);
            unsigned c1a = c1_base + 8u * (unsigned)(li * T); // + 8 * (v * T * T + lj)
            uint8_t *dst = fused + ((size_t)i * d.Wp + j0) * D;
            const ptrdiff_t to_h = TRIPLE ? fused_h - fused : 0, to_v = TRIPLE ? fused_v - fused : 0;
#pragma unroll 1
            for (int lj = 0; lj < T; lj++, dst += D, c1a += 8u) {
                const unsigned m = smask[li * T + lj];
                if (m == 0) { // warp-uniform; D is a multiple of 32 here: the cell is D / 16 aligned 16-byte stores
#pragma unroll
                    for (int vol = 0; vol < (TRIPLE ? 3 : 1); vol++) {
                        uint8_t *z = dst + (vol == 1 ? to_h : vol == 2 ? to_v : 0);
                        if (lane < D / 16) *reinterpret_cast<uint4 *>(z + 16 * lane) = make_uint4(0u, 0u, 0u, 0u);
                        if (NK > 2 && lane + 32 < D / 16) *reinterpret_cast<uint4 *>(z + 16 * (lane + 32)) = make_uint4(0u, 0u, 0u, 0u);
                    }
                } else {
                    unsigned acc[NKC];
                    if (m == 0xFu) { // all four views (the common case of mode 0): straight-line, no per-view branches
                        uint2 c1[4];
#pragma unroll
                        for (int v = 0; v < 4; v++) c1[v] = lds64(c1a + 8u * (unsigned)(v * T * T));
#pragma unroll
                        for (int k = 0; k < NKC; k++) {
                            unsigned a = 0, b = 0;
#pragma unroll
                            for (int v = 0; v < 4; v++) {
                                const uint2 x = lds64(pv[v] - koff[k]);
                                const unsigned cst = __popc(x.x ^ c1[v].x) + __popc(x.y ^ c1[v].y);
                                if (TRIPLE && v >= 2) b += cst;
                                else a += cst;
                            }
                            acc[k] = TRIPLE ? b * 65536u + a : a;
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < NKC; k++) acc[k] = 0;
#pragma unroll
                        for (int v = 0; v < 4; v++) {
                            if (!((m >> v) & 1u)) continue; // warp-uniform
                            const uint2 c1 = lds64(c1a + 8u * (unsigned)(v * T * T));
#pragma unroll
                            for (int k = 0; k < NKC; k++) {
                                const uint2 x = lds64(pv[v] - koff[k]);
                                acc[k] += (__popc(x.x ^ c1.x) + __popc(x.y ^ c1.y)) << ((TRIPLE && v >= 2) ? 16 : 0);
                            }
                        }
                    }
                    if constexpr (TRIPLE) {
                        unsigned all = 0;
#pragma unroll
                        for (int k = 0; k < NKC; k++) all |= (acc[k] & 0xFFFFu) + (acc[k] >> 16);
                        if (all > 255u) overflow = true; // cannot happen: flagged, bytes saturate below
#pragma unroll
                        for (int k = 0; k < NKC; k++) {
                            const unsigned h = acc[k] & 0xFFFFu, v2 = acc[k] >> 16;
                            dst[lane + 32 * k] = (uint8_t)min(h + v2, 255u);
                            dst[to_h + lane + 32 * k] = (uint8_t)min(h, 255u);
                            dst[to_v + lane + 32 * k] = (uint8_t)min(v2, 255u);
                        }
                    } else {
                        unsigned all = acc[0];
#pragma unroll
                        for (int k = 1; k < NKC; k++) all |= acc[k];
                        if (all > 255u) { // cannot happen; kept as the literal formula's saturation + flag
                            overflow = true;
#pragma unroll
                            for (int k = 0; k < NKC; k++) acc[k] = min(acc[k], 255u);
                        }
#pragma unroll
                        for (int k = 0; k < NKC; k++) dst[lane + 32 * k] = (uint8_t)acc[k];
                    }
                }
#pragma unroll
                for (int v = 0; v < 4; v++) pv[v] += (unsigned)pstep[v];
            }
            if (overflow) atomicOr(status, kStatusFusedOverflow);
            return;
        }
    }
#pragma unroll 1
    for (int lj = 0; lj < T; lj++) {
        const int j = j0 + lj;
        if (j >= d.Wp) break;
        const size_t pix = (size_t)i * d.Wp + j;
        const unsigned m = smask[li * T + lj];
        unsigned acc[NKC];
#pragma unroll
        for (int k = 0; k < NKC; k++) acc[k] = 0;
#pragma unroll
        for (int v = 0; v < 4; v++) {
            if (!((m >> v) & 1u)) continue; // warp-uniform
            const int hv = view_rows(d, v), wv = view_cols(d, v);
            int rv, cc;
            image_to_view(d, v, i, j, rv, cc);
            const unsigned long long *line = s + (size_t)v * T * P + (size_t)((v < 2) ? li : lj) * P + (cc - cv_lo[v] + D - 1);
            if (rv >= 3 && rv < hv - 2 && cc >= D - 1) {
                const unsigned long long c1 = sc1[(v * T + li) * T + lj];
                const unsigned c1lo = (unsigned)c1, c1hi = (unsigned)(c1 >> 32);
                const unsigned long long *p = line - dl;
#pragma unroll
                for (int k = 0; k < NKC; k++) {
                    if (NK > 0 || lane + 32 * k < D) {
                        const uint2 x = *reinterpret_cast<const uint2 *>(p - dku(k));
                        acc[k] += (__popc(x.x ^ c1lo) + __popc(x.y ^ c1hi)) << ((TRIPLE && v >= 2) ? 16 : 0);
                    }
                }
            } else {
                // literal formula: rows 0..2 are 255 (census.cpp:95-98,142-145), rows h-2,h-1 are never written
                // (defined 0), d beyond the view column is 255 (census.cpp:76)
                unsigned long long c1 = 0;
                const bool popc_row = rv >= 3 && rv < hv - 2;
                if (popc_row) c1 = sc1[(v * T + li) * T + lj];
#pragma unroll 1
                for (int k = 0; k < NKC; k++) {
                    if (lane + 32 * k >= D) break;
                    const int dd = dl + dku(k);
                    unsigned cst;
                    if (rv < 3) cst = kInvalidCost;
                    else if (!popc_row) cst = 0;
                    else cst = (dd > cc) ? (unsigned)kInvalidCost : (unsigned)__popcll(c1 ^ line[-dd]);
                    acc[k] += cst << ((TRIPLE && v >= 2) ? 16 : 0);
                }
            }
        }
        uint8_t *dst = fused + pix * D + lane;
#pragma unroll
        for (int k = 0; k < NKC; k++) {
            if (NK > 0 || lane + 32 * k < D) {
                if constexpr (TRIPLE) {
                    const unsigned h = acc[k] & 0xFFFFu, v2 = acc[k] >> 16;
                    if (h + v2 > 255u) overflow = true;
                    dst[32 * k] = (uint8_t)min(h + v2, 255u);
                    fused_h[pix * D + lane + 32 * k] = (uint8_t)min(h, 255u);
                    fused_v[pix * D + lane + 32 * k] = (uint8_t)min(v2, 255u);
                } else {
                    unsigned sum = acc[k];
                    if (sum > 255u) { overflow = true; sum = 255u; }
                    dst[32 * k] = (uint8_t)sum;
                }
            }
        }
    }
    if (overflow) atomicOr(status, kStatusFusedOverflow);
}

template <int T, int NK, int ORDER>
static void launch_fuse_o(const unsigned long long *census, const uint8_t *masks, const Dims &d, unsigned view_mask, uint8_t *fused,
                          uint8_t *fused_h, uint8_t *fused_v, int *status, cudaStream_t st, int row_lo, int row_hi, LaunchCounter &lc)
{
    const size_t smem = (size_t)4 * T * (T + d.D) * 8 + (size_t)4 * T * T * 8 + T * T;
    dim3 grid((d.Wp + T - 1) / T, (row_hi - row_lo + T - 1) / T);
    if (fused_h && fused_v) {
        lc.fail(optin_dynamic_smem((const void *)k_fuse<T, NK, ORDER, true>, smem));
        k_fuse<T, NK, ORDER, true><<<grid, 32 * T, smem, st>>>(census, masks, d, view_mask, fused, fused_h, fused_v, status, row_lo, row_hi);
        return;
    }
    lc.fail(optin_dynamic_smem((const void *)k_fuse<T, NK, ORDER, false>, smem));
    k_fuse<T, NK, ORDER, false><<<grid, 32 * T, smem, st>>>(census, masks, d, view_mask, fused, nullptr, nullptr, status, row_lo, row_hi);
}

template <int T, int NK>
static void launch_fuse_t(const unsigned long long *census, const uint8_t *masks, const Dims &d, unsigned view_mask, uint8_t *fused,
                          uint8_t *fused_h, uint8_t *fused_v, int *status, cudaStream_t st, int row_lo, int row_hi, LaunchCounter &lc)
{
    if (NK > 0 && d.interleaved && d.lpc == 16 && d.nr == NK) launch_fuse_o<T, NK, (NK > 0 ? 2 : 0)>(census, masks, d, view_mask, fused, fused_h, fused_v, status, st, row_lo, row_hi, lc);
    else if (NK > 0 && !d.interleaved) launch_fuse_o<T, NK, (NK > 0 ? 1 : 0)>(census, masks, d, view_mask, fused, fused_h, fused_v, status, st, row_lo, row_hi, lc);
    else launch_fuse_o<T, NK, 0>(census, masks, d, view_mask, fused, fused_h, fused_v, status, st, row_lo, row_hi, lc);
}

void launch_fuse(const unsigned long long *census, const uint8_t *masks, const Dims &d, unsigned view_mask, uint8_t *fused,
                 int *status, cudaStream_t st, LaunchCounter &lc, int row_lo, int row_hi, uint8_t *fused_h, uint8_t *fused_v)
{
    if (row_hi < 0) row_hi = d.Hp;
    if (row_hi <= row_lo) return;
    // 16 x 16 tiles while two blocks fit an SM (D <= 200), else 8 x 8
    if (d.D <= 200) {
        switch (d.D) {
        case 64: launch_fuse_t<16, 2>(census, masks, d, view_mask, fused, fused_h, fused_v, status, st, row_lo, row_hi, lc); break;
        case 128: launch_fuse_t<16, 4>(census, masks, d, view_mask, fused, fused_h, fused_v, status, st, row_lo, row_hi, lc); break;
        case 192: launch_fuse_t<16, 6>(census, masks, d, view_mask, fused, fused_h, fused_v, status, st, row_lo, row_hi, lc); break;
        default: launch_fuse_t<16, 0>(census, masks, d, view_mask, fused, fused_h, fused_v, status, st, row_lo, row_hi, lc); break;
        }
    } else {
        switch (d.D) {
        case 256: launch_fuse_t<8, 8>(census, masks, d, view_mask, fused, fused_h, fused_v, status, st, row_lo, row_hi, lc); break;
        case 384: launch_fuse_t<8, 12>(census, masks, d, view_mask, fused, fused_h, fused_v, status, st, row_lo, row_hi, lc); break;
        case 512: launch_fuse_t<8, 16>(census, masks, d, view_mask, fused, fused_h, fused_v, status, st, row_lo, row_hi, lc); break;
        default: launch_fuse_t<8, 0>(census, masks, d, view_mask, fused, fused_h, fused_v, status, st, row_lo, row_hi, lc); break;
        }
    }
    lc.add();
}

} // namespace sister
