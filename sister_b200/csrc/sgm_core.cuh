// sister_b200 / sgm_core.cuh -- the per-step arithmetic of the semi-global aggregation, of the chain kernel (sgm.cu). See the header comment of sgm.cu for the formulation
// (normalised clamped state a, penalty byte Q = L' - C) and its mapping to sgm.cpp:26-455.
#pragma once
#include "kernels.cuh"

namespace sister {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr uint32_t kP1x2 = (uint32_t)kP1 * 0x10001u;
constexpr uint32_t kP2x2 = (uint32_t)kP2 * 0x10001u;
// make a value opaque to the optimiser (no rematerialisation from its definition)
__device__ __forceinline__ void opaque(unsigned &x) { asm volatile("mov.u32 %0, %0;\n" : "+r"(x)); }
template <class T> __device__ __forceinline__ void opaque_ptr(T *&p)
{
    unsigned long long v = reinterpret_cast<unsigned long long>(p);
    asm volatile("mov.u64 %0, %0;\n" : "+l"(v));
    p = reinterpret_cast<T *>(v);
}

// x + y issued as IMAD (x * one + y with `one` an opaque register holding 1): the chain kernel is bound by the ALU pipe
// (VIMNMX / VIADDMNMX / PRMT live there) while the FMA pipe idles, and ptxas puts a plain add on either.
#ifndef SISTER_SGM_FMA_ADDS
#define SISTER_SGM_FMA_ADDS 1
#endif
__device__ __forceinline__ uint32_t add_fma(uint32_t x, uint32_t y, uint32_t one)
{
#if SISTER_SGM_FMA_ADDS
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;\n" : "=r"(r) : "r"(x), "r"(one), "r"(y));
    return r;
#else
    (void)one;
    return x + y;
#endif
}

// min over the chain's lanes of both halves of m2, returned in both halves: (m, m). Xor-butterfly inside the chain's LPC
// lanes: log2(LPC) shuffles serve every chain of the warp at once. (One CREDUX.MIN per chain with the other chains
// masked out has a shorter dependent latency but costs more issue slots; measured no faster on B200.)
#ifndef SISTER_SGM_REDUX
#define SISTER_SGM_REDUX 1
#endif
template <int LPC> __device__ __forceinline__ uint32_t chain_min2(uint32_t m2)
{
    uint32_t v = __vmins2(m2, __byte_perm(m2, m2, 0x1032));
    if constexpr (LPC == 32) {
        return __reduce_min_sync(kFull, v);
    } else if constexpr ((LPC == 16 && SISTER_SGM_REDUX) || (LPC == 8 && SISTER_SGM_REDUX >= 2)) {
        // 32 / LPC chains per warp: one warp-wide reduction per chain (the other chains' lanes pass infinity), independent of
        // each other so that their latencies overlap, instead of log2(LPC) dependent shuffle rounds -- the step's critical
        // path is what bounds the kernel whenever an SM holds few warps (the tail of the grid). v holds the same value in
        // both halves, so the unsigned 32-bit order is the 16-bit order.
        const int cid = (threadIdx.x & 31) / LPC;
        uint32_t r = 0;
#pragma unroll
        for (int k = 0; k < 32 / LPC; k++) {
            const uint32_t m = __reduce_min_sync(kFull, cid == k ? v : 0x7FFF7FFFu);
            r = cid == k ? m : r;
        }
        return r;
    } else {
#pragma unroll
        for (int o = 1; o < LPC; o <<= 1) v = __vmins2(v, __shfl_xor_sync(kFull, v, o));
        return v;
    }
}

// Register layout of a chain's state ("split lanes"). Lane sl of a chain owns the 2 * NR consecutive disparities
// d0 = sl * 2NR ...; register k holds the pair (d0 + k, d0 + NR + k) in its (low, high) 16-bit halves. The d-1 / d+1
// neighbours of register k are then simply registers k-1 / k+1 -- whole registers, both halves at once -- and only the
// two ends need a byte permute with a value from the adjacent lane:
//     left of register 0       = (previous lane's high half of register NR-1, own low half of register NR-1)
//     right of register NR-1   = (own high half of register 0, next lane's low half of register 0)
// With b = a + P1 kept next to a, the penalty is ONE three-input packed minimum per register:
//     Q[k] = min(a[k], b[k-1], b[k+1])                      (VIMNMX3.S16x2; reference: sgm.cpp:282-297)
// At the chain's first / last lane the missing neighbour is kInf2 (L(-1) = L(D) = 65535, sgm.cpp:84-87). Disparities
// >= D (only when D < LPC * 2NR, FULL = false) carry kInfHalf in a, b, L so that they never win a minimum.
constexpr uint32_t kInfLo = 0x00003FFFu, kInfHi = 0x3FFF0000u;

template <int NR, int LPC, bool FULL> struct LaneInfo {
    int sl;                          // lane within the chain
    uint32_t one;                    // 1, opaque to the optimiser (add_fma)
    uint32_t up_mask, dn_mask;       // kInf2 where the lane below / above holds no neighbour (chain ends, disparities >= D)
    uint32_t pad[FULL ? 1 : NR];     // !FULL: kInfLo / kInfHi where the register's half is a disparity >= D
    __device__ __forceinline__ void init(int lane, int D)
    {
        sl = lane % LPC;
        one = 1u;
        up_mask = sl == 0 ? kInf2 : 0u;
        dn_mask = (sl == LPC - 1 || (!FULL && (sl + 1) * 2 * NR >= D)) ? kInf2 : 0u;
        if constexpr (!FULL) {
#pragma unroll
            for (int k = 0; k < NR; k++)
                pad[k] = ((sl * 2 * NR + k >= D) ? kInfLo : 0u) | ((sl * 2 * NR + NR + k >= D) ? kInfHi : 0u);
        } else {
            pad[0] = 0u;
        }
    }
    __device__ __forceinline__ uint32_t padded(uint32_t v, int k) const
    {
        if constexpr (FULL) return v;
        else return v | pad[k];
    }
    // where the lane's bytes start inside a cell: lane-interleaved words when every lane is full (Dims, common.cuh), else the
    // lane's 2NR consecutive bytes of the natural order
    template <bool IL> __device__ __forceinline__ int cell_offset() const { return IL ? sl * 4 : sl * 2 * NR; }
    // bytes of this lane's 2NR that lie inside the cell (a multiple of 4)
    __device__ __forceinline__ int valid_bytes(int D) const
    {
        const int n = D - sl * 2 * NR;
        return n < 0 ? 0 : (n > 2 * NR ? 2 * NR : n);
    }
};

// the two end neighbours of b (see above)
template <int NR> __device__ __forceinline__ void end_neighbours(const uint32_t (&b)[NR], uint32_t up_mask, uint32_t dn_mask, uint32_t &left, uint32_t &right)
{
    const uint32_t up = __shfl_up_sync(kFull, b[NR - 1], 1) | up_mask;
    const uint32_t dn = __shfl_down_sync(kFull, b[0], 1) | dn_mask;
    left = __byte_perm(up, b[NR - 1], 0x5432);
    right = __byte_perm(b[0], dn, 0x5432);
}

// State of a chain between steps: a = min(L - m, P2) (clamped normalised costs), b = a + P1, and the two end
// neighbours of b (left of register 0, right of register NR-1) already assembled for the next step.
template <int NR> struct ChainState {
    uint32_t a[NR], b[NR], left, right;
};

template <int NR> __device__ __forceinline__ uint32_t lane_min(const uint32_t (&L)[NR])
{
    uint32_t m2 = __vmins2(L[0], L[1]);
#pragma unroll
    for (int k = 2; k < NR; k += 2) m2 = __vimin3_s16x2(m2, L[k], L[k + 1]);
    return m2;
}

#ifndef SISTER_SGM_LATE_ENDS
#define SISTER_SGM_LATE_ENDS 0
#endif
// normalise and clamp L into the next state. The end registers of L travel to the adjacent lanes BEFORE the minimum is
// known (the shuffles overlap the reduction) and are normalised by the receiver: a step waits for one MIO round trip.
// off_next: the NEXT cell of the chain follows a border crossing, its predecessor lies outside the frame: L_prev = 65535,
// min = 0 in the reference (sgm.cpp:57-81), i.e. the state it must see is a = P2 everywhere. Adding a large positive
// constant instead of -m makes every clamp below return exactly that, for one select per step.
template <int NR, int LPC, bool FULL>
__device__ __forceinline__ void renormalise(const uint32_t (&L)[NR], const LaneInfo<NR, LPC, FULL> &li, ChainState<NR> &st, bool off_next = false)
{
#if !SISTER_SGM_LATE_ENDS
    const uint32_t up = __shfl_up_sync(kFull, L[NR - 1], 1);
    const uint32_t dn = __shfl_down_sync(kFull, L[0], 1);
#endif
    const uint32_t mm = chain_min2<LPC>(lane_min<NR>(L));
    // -m as a 16-bit two's complement value in both halves (L <= 0x3FFF, so L + 0x3000 stays positive)
    const uint32_t neg2 = off_next ? 0x30003000u : __byte_perm(0u - mm, 0u, 0x1010);
#pragma unroll
    for (int k = 0; k < NR; k++) {
        st.a[k] = li.padded(__viaddmin_s16x2(L[k], neg2, kP2x2), k);
        st.b[k] = add_fma(st.a[k], kP1x2, li.one);
    }
#if SISTER_SGM_LATE_ENDS
    // the end registers of b travel after the normalisation: two instructions of the (binding) ALU pipe fewer than
    // normalising the neighbours' raw values a second time, for one more shuffle latency per step
    end_neighbours<NR>(st.b, li.up_mask, li.dn_mask, st.left, st.right);
#else
    // a missing neighbour (mask = kInf2) normalises to P2 + P1, which no a <= P2 ever loses to: as good as infinity
    const uint32_t bu = add_fma(__viaddmin_s16x2(up | li.up_mask, neg2, kP2x2), kP1x2, li.one);
    const uint32_t bd = add_fma(__viaddmin_s16x2(dn | li.dn_mask, neg2, kP2x2), kP1x2, li.one);
    st.left = __byte_perm(bu, st.b[NR - 1], 0x5432);
    st.right = __byte_perm(st.b[0], bd, 0x5432);
#endif
}

// One SGM step of one chain. Writes q = L' - C (in [0, P2]) and the new state.
template <int NR, int LPC, bool FULL>
__device__ __forceinline__ void chain_step(ChainState<NR> &st, const uint32_t (&c)[NR], const LaneInfo<NR, LPC, FULL> &li, uint32_t (&q)[NR],
                                           bool off_next = false)
{
    uint32_t L[NR];
#pragma unroll
    for (int k = 0; k < NR; k++) {
        q[k] = __vimin3_s16x2(st.a[k], k == 0 ? st.left : st.b[k - 1], k == NR - 1 ? st.right : st.b[k + 1]);
        L[k] = li.padded(q[k] + c[k], k);
    }
    renormalise<NR, LPC, FULL>(L, li, st, off_next);
}

// constant state (a = v in every valid disparity): v = 0 at the start of a row for r0 (sgm.cpp:215-216), v = P2 behind
// an off-image predecessor column (sgm.cpp:57-81)
template <int NR, int LPC, bool FULL>
__device__ __forceinline__ void chain_set(ChainState<NR> &st, uint32_t v2, const LaneInfo<NR, LPC, FULL> &li)
{
#pragma unroll
    for (int k = 0; k < NR; k++) {
        st.a[k] = li.padded(v2, k);
        st.b[k] = st.a[k] + kP1x2;
    }
    st.left = __byte_perm((v2 + kP1x2) | li.up_mask, st.b[NR - 1], 0x5432);
    st.right = __byte_perm(st.b[0], (v2 + kP1x2) | li.dn_mask, 0x5432);
}

// continue a chain from a stored state vector a (a band border, sgm.cu Band)
template <int NR, int LPC, bool FULL>
__device__ __forceinline__ void chain_resume(ChainState<NR> &st, const uint32_t (&a)[NR], const LaneInfo<NR, LPC, FULL> &li)
{
#pragma unroll
    for (int k = 0; k < NR; k++) {
        st.a[k] = li.padded(a[k], k);
        st.b[k] = st.a[k] + kP1x2;
    }
    end_neighbours<NR>(st.b, li.up_mask, li.dn_mask, st.left, st.right);
}

// The horizontal path on the first line of a pass (sgm.cpp:141-190): plain int arithmetic on the un-normalised
// values, then saturate_cast<uint16>(uint8) truncation (types.h:28). The state carried along the line is the truncated
// value Lq (b = Lq + P1) and its minimum (mm, both halves); the byte written to the path volume is the truncated
// value itself.
template <int NR, int LPC, bool FULL>
__device__ __forceinline__ void first_line_step(uint32_t (&Lq)[NR], uint32_t (&b)[NR], uint32_t &mm, const uint32_t (&c)[NR],
                                                const LaneInfo<NR, LPC, FULL> &li, bool first_column, uint32_t (&q)[NR])
{
    if (first_column) {
#pragma unroll
        for (int k = 0; k < NR; k++) q[k] = c[k];
    } else {
        uint32_t left, right;
        end_neighbours<NR>(b, li.up_mask, li.dn_mask, left, right);
        const uint32_t p2 = mm + kP2x2;
#pragma unroll
        for (int k = 0; k < NR; k++) {
            const uint32_t x = __vimin3_s16x2(Lq[k], k == 0 ? left : b[k - 1], k == NR - 1 ? right : b[k + 1]);
            q[k] = (c[k] + (__vmins2(x, p2) - mm)) & 0x00FF00FFu;
        }
    }
#pragma unroll
    for (int k = 0; k < NR; k++) {
        Lq[k] = li.padded(q[k], k);
        b[k] = Lq[k] + kP1x2;
    }
    mm = chain_min2<LPC>(lane_min<NR>(Lq));
}

// ---- byte <-> register conversion.
// IL (lane-interleaved cell, common.cuh): the lane's word t IS the pair word X[t] = (lo[2t], lo[2t+1], hi[2t], hi[2t+1]);
// registers 2t, 2t+1 are its even / odd bytes: two instructions per word each way.
// !IL (natural order): a lane's 2NR cost bytes are contiguous in the cell; 16-bit unit u (u < NR) holds disparities
// d0 + 2u, d0 + 2u + 1 and X[t] = (unit t, unit NR/2 + t) takes one more byte permute per word.

template <int NR, bool IL> __device__ __forceinline__ void unpack_cost(const uint32_t (&w)[NR / 2], uint32_t (&c)[NR])
{
    static_assert(NR % 2 == 0, "an even number of packed registers per lane");
#pragma unroll
    for (int t = 0; t < NR / 2; t++) {
        uint32_t X;
        if constexpr (IL) {
            X = w[t];
        } else {
            constexpr int H = NR / 2;
            const int ua = t, ub = H + t; // units
            const uint32_t sel = ((ub & 1) ? 0x7600u : 0x5400u) | ((ua & 1) ? 0x32u : 0x10u);
            X = __byte_perm(w[ua >> 1], w[ub >> 1], sel);
        }
        c[2 * t + 1] = __byte_perm(X, 0u, 0x4341);
        c[2 * t] = X - (c[2 * t + 1] << 8); // == X & 0x00FF00FF, written so that it can issue as an IMAD on the FMA pipe
    }
}

template <int NR, bool IL> __device__ __forceinline__ void pack_q(const uint32_t (&q)[NR], uint32_t (&w)[NR / 2])
{
    uint32_t X[NR / 2];
#pragma unroll
    for (int t = 0; t < NR / 2; t++) X[t] = q[2 * t + 1] * 256u + q[2 * t];
#pragma unroll
    for (int j = 0; j < NR / 2; j++) {
        if constexpr (IL) {
            w[j] = X[j];
        } else {
            constexpr int H = NR / 2;
            const int u0 = 2 * j, u1 = 2 * j + 1;
            const uint32_t sel = ((u1 >= H) ? 0x7600u : 0x5400u) | ((u0 >= H) ? 0x32u : 0x10u);
            w[j] = __byte_perm(X[u0 % H], X[u1 % H], sel);
        }
    }
}

// store the step's penalty bytes (q <= 255 in both halves) at dst = the cell + the lane's cell_offset()
__device__ __forceinline__ void stg32(uint8_t *p, uint32_t x) { asm volatile("st.global.u32 [%0], %1;\n" ::"l"(p), "r"(x) : "memory"); }
__device__ __forceinline__ void stg64(uint8_t *p, uint32_t x, uint32_t y) { asm volatile("st.global.v2.u32 [%0], {%1, %2};\n" ::"l"(p), "r"(x), "r"(y) : "memory"); }
__device__ __forceinline__ void stg128(uint8_t *p, uint32_t x, uint32_t y, uint32_t z, uint32_t w)
{
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
template <int NR, int LPC, bool FULL, bool IL> __device__ __forceinline__ void store_q(uint8_t *dst, const uint32_t (&q)[NR], int valid_bytes)
{
    static_assert(FULL || !IL, "only full lanes interleave");
#ifdef SISTER_SGM_NOSTORE // measurement aid: the chain arithmetic without its stores (results are then missing)
    if (q[0] != 0xDEADBEEFu) return;
#endif
    uint32_t w[NR / 2];
    pack_q<NR, IL>(q, w);
    if constexpr (IL) {
#pragma unroll
        for (int k = 0; k < NR / 2; k++) stg32(dst + 4 * LPC * k, w[k]); // 4 * LPC contiguous bytes per chain and instruction
    } else if constexpr (FULL && NR % 8 == 0) {
#pragma unroll
        for (int k = 0; k < NR / 8; k++) stg128(dst + 16 * k, w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
    } else if constexpr (FULL && NR % 4 == 0) {
#pragma unroll
        for (int k = 0; k < NR / 4; k++) stg64(dst + 8 * k, w[2 * k], w[2 * k + 1]);
    } else {
#pragma unroll
        for (int k = 0; k < NR / 2; k++)
            if (FULL || 4 * k < valid_bytes) stg32(dst + 4 * k, w[k]);
    }
}

// ---- cost stream. The fused-cost bytes a chain will need are known in advance: each lane loads its own 2NR bytes of
// the cell kAhead steps early straight into registers (LDG, L1-cached: the few instructions that cover one cell hit
// the same lines) and asks L2 for the cell kAhead + kFar steps early (PREFETCH.L2), so no step waits on DRAM or L2.
#ifndef SISTER_SGM_AHEAD
#define SISTER_SGM_AHEAD 3
#endif
#ifndef SISTER_SGM_FAR
#define SISTER_SGM_FAR 8
#endif
constexpr int kAhead = SISTER_SGM_AHEAD;   // register lookahead, steps (loop unrolled by kAhead, buffers rotate at compile time)
constexpr int kFar = SISTER_SGM_FAR;       // L2 prefetch distance beyond the register lookahead, steps

#ifndef SISTER_SGM_PF_L1
#define SISTER_SGM_PF_L1 0
#endif
__device__ __forceinline__ void prefetch_l2(const uint8_t *p)
{
#if SISTER_SGM_PF_L1
    asm volatile("prefetch.global.L1 [%0];\n" ::"l"(p));
#else
    asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p));
#endif
}

template <int NR, int LPC, bool FULL, bool IL> __device__ __forceinline__ void load_cost(const uint8_t *src, int valid_bytes, uint32_t (&w)[NR / 2])
{
    if constexpr (IL) {
#pragma unroll
        for (int k = 0; k < NR / 2; k++) w[k] = __ldg(reinterpret_cast<const uint32_t *>(src) + LPC * k);
    } else if constexpr (FULL && NR % 8 == 0) {
#pragma unroll
        for (int k = 0; k < NR / 8; k++) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src) + k);
            w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
        }
    } else if constexpr (FULL && NR % 4 == 0) {
#pragma unroll
        for (int k = 0; k < NR / 4; k++) {
            const uint2 v = __ldg(reinterpret_cast<const uint2 *>(src) + k);
            w[2 * k] = v.x; w[2 * k + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < NR / 2; k++) w[k] = (FULL || 4 * k < valid_bytes) ? __ldg(reinterpret_cast<const uint32_t *>(src) + k) : 0u;
    }
}

} // namespace sister
