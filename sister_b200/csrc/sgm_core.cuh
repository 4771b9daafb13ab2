// sister_b200 / sgm_core.cuh -- the per-step arithmetic of the semi-global aggregation, shared by the chain kernel
// (sgm.cu) and the lock-step pair sweeps (sweep.cu). See the header comment of sgm.cu for the formulation
// (normalised clamped state a, penalty byte Q = L' - C) and its mapping to sgm.cpp:26-455.
#pragma once
#include "kernels.cuh"

namespace sister {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr uint32_t kP1x2 = (uint32_t)kP1 * 0x10001u;
constexpr uint32_t kP2x2 = (uint32_t)kP2 * 0x10001u;
constexpr int kRing = 12;         // steps of fused cost in flight per chain

// Lane mapping. A chain occupies LPC lanes of a warp (LPC = 16 for D <= 256: two chains per warp, so the per-step
// fixed work -- shuffles, border selects, the min reduction, loop and cursor arithmetic -- is paid once for two
// chains; LPC = 32 above). Lane sl of a chain owns the 2 * NR consecutive disparities sl * 2NR ..., two per register.

// ---------------------------------------------------------------------------------------------- small helpers

__device__ __forceinline__ void cp_async8(unsigned smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16(unsigned smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
// the same, issued only by lanes whose `on` register is non-zero (a register predicate keeps ptxas from
// re-deriving the lane test from %tid in every step)
__device__ __forceinline__ void cp_async16_if(unsigned smem_dst, const void *gmem_src, unsigned on)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}\n" ::"r"(smem_dst), "l"(gmem_src), "r"(on) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// number of valid packed registers of this lane (disparities sl*2NR + 2k, +1 are valid for k < nvalid)
template <int NR> __device__ __forceinline__ int lane_nvalid(int D, int sl)
{
    int n = D / 2 - sl * NR;
    return n < 0 ? 0 : (n > NR ? NR : n);
}

// min over the chain's lanes of both halves of m2, returned in both halves: (m, m). All values are in [0, 0x3FFF], so
// with equal halves the unsigned 32-bit order is the 16-bit order and one CREDUX.MIN per chain does it.
template <int LPC> __device__ __forceinline__ uint32_t chain_min2(uint32_t m2, const uint32_t (&others)[32 / LPC])
{
    (void)others;
    uint32_t v = __vmins2(m2, __byte_perm(m2, m2, 0x1032));
    if constexpr (LPC == 32) {
        return __reduce_min_sync(kFull, v);
    } else {
        // xor-butterfly inside the chain's LPC lanes: log2(LPC) shuffles serve every chain of the warp at once and the
        // dependent latency is a few shuffles instead of one CREDUX per chain
#pragma unroll
        for (int o = 1; o < LPC; o <<= 1) v = __vmins2(v, __shfl_xor_sync(kFull, v, o));
        return v;
    }
}

// Neighbour registers of a packed state vector: E[k] = (d-1 of the low half, low half), E[k+1] = (high half, d+1 of
// the high half). The two values that live in the adjacent lanes come by shuffle; at the chain's first / last lane
// they are kInf2 (L(-1) = L(D) = 65535 in the reference, sgm.cpp:84-87).
template <int NR> __device__ __forceinline__ void neighbours(const uint32_t (&a)[NR], uint32_t up_mask, uint32_t dn_mask, uint32_t (&E)[NR + 1])
{
    // up_mask / dn_mask = kInf2 at the chain's first / last lane, 0 elsewhere: x | kInf2 >= kInf2 never wins a minimum
    const uint32_t up = __shfl_up_sync(kFull, a[NR - 1], 1) | up_mask;
    const uint32_t dn = __shfl_down_sync(kFull, a[0], 1) | dn_mask;
    E[0] = __byte_perm(up, a[0], 0x5432);
#pragma unroll
    for (int k = 1; k < NR; k++) E[k] = __byte_perm(a[k - 1], a[k], 0x5432);
    E[NR] = __byte_perm(a[NR - 1], dn, 0x5432);
}

template <int LPC> struct LaneInfo {
    int sl, nvalid;
    uint32_t up_mask, dn_mask;       // kInf2 at the chain's first / last lane
    uint32_t others[32 / LPC];       // 0x7FFF7FFF for the chains of the warp this lane does not belong to
};

// One SGM step of one chain: a = clamped normalised state of the predecessor (pad registers = kInf2).
// Writes q = L' - C (in [0, P2]) and the new state.
template <int NR, int LPC, bool FULL>
__device__ __forceinline__ void chain_step(uint32_t (&a)[NR], const uint32_t (&c)[NR], const LaneInfo<LPC> &li, uint32_t (&q)[NR])
{
    uint32_t E[NR + 1], L[NR];
    neighbours<NR>(a, li.up_mask, li.dn_mask, E);
    uint32_t m2 = kInf2;
#pragma unroll
    for (int k = 0; k < NR; k++) {
        const uint32_t x = __viaddmin_s16x2(E[k], kP1x2, a[k]);
        q[k] = __viaddmin_s16x2(E[k + 1], kP1x2, x);
        L[k] = q[k] + c[k];
        if (FULL || k < li.nvalid) m2 = __vmins2(m2, L[k]);
    }
    const uint32_t mm = chain_min2<LPC>(m2, li.others);
    const uint32_t cap = mm + kP2x2;
#pragma unroll
    for (int k = 0; k < NR; k++) {
        const uint32_t n = __vmins2(L[k], cap) - mm; // min(L - m, P2); L >= m in both halves, no borrow
        a[k] = (FULL || k < li.nvalid) ? n : kInf2;
    }
}

// The first cell of a column / diagonal chain lies on the first line of the pass: L = C (sgm.cpp:103-138).
template <int NR, int LPC, bool FULL>
__device__ __forceinline__ void chain_first_cell(uint32_t (&a)[NR], const uint32_t (&c)[NR], const LaneInfo<LPC> &li, uint32_t (&q)[NR])
{
    uint32_t m2 = kInf2;
#pragma unroll
    for (int k = 0; k < NR; k++) {
        q[k] = 0u;
        if (FULL || k < li.nvalid) m2 = __vmins2(m2, c[k]);
    }
    const uint32_t mm = chain_min2<LPC>(m2, li.others);
    const uint32_t cap = mm + kP2x2;
#pragma unroll
    for (int k = 0; k < NR; k++) {
        const uint32_t n = __vmins2(c[k], cap) - mm;
        a[k] = (FULL || k < li.nvalid) ? n : kInf2;
    }
}

// The horizontal path on the first line of a pass (sgm.cpp:141-190): plain int arithmetic on the un-normalised
// values, then saturate_cast<uint16>(uint8) truncation (types.h:28). The state carried along the line is the truncated
// value Lq and its minimum (mm, both halves); the byte written to the path volume is the truncated value itself.
template <int NR, int LPC, bool FULL>
__device__ __forceinline__ void first_line_step(uint32_t (&Lq)[NR], uint32_t &mm, const uint32_t (&c)[NR], const LaneInfo<LPC> &li,
                                                bool first_column, uint32_t (&q)[NR])
{
    if (first_column) {
#pragma unroll
        for (int k = 0; k < NR; k++) q[k] = c[k];
    } else {
        uint32_t E[NR + 1];
        neighbours<NR>(Lq, li.up_mask, li.dn_mask, E);
        const uint32_t p2 = mm + kP2x2;
#pragma unroll
        for (int k = 0; k < NR; k++) {
            const uint32_t x = __viaddmin_s16x2(E[k], kP1x2, Lq[k]);
            const uint32_t y = __viaddmin_s16x2(E[k + 1], kP1x2, p2);
            q[k] = (c[k] + (__vmins2(x, y) - mm)) & 0x00FF00FFu;
        }
    }
    uint32_t m2 = kInf2;
#pragma unroll
    for (int k = 0; k < NR; k++) {
        Lq[k] = (FULL || k < li.nvalid) ? q[k] : kInf2;
        m2 = __vmins2(m2, Lq[k]);
    }
    mm = chain_min2<LPC>(m2, li.others);
}

// store the step's penalty bytes (q <= 255 in both halves): 2 * NR bytes per lane
template <int NR, bool FULL> __device__ __forceinline__ void store_q(uint8_t *dst, const uint32_t (&q)[NR], int nvalid)
{
    if constexpr (NR % 2 == 0) {
        // sl * 2NR is a multiple of 4 (of 8 when NR % 4 == 0) and cell * D a multiple of 8: the stores are aligned
        if (FULL || nvalid == NR) {
            uint32_t w[NR / 2];
#pragma unroll
            for (int k = 0; k < NR / 2; k++) w[k] = __byte_perm(q[2 * k], q[2 * k + 1], 0x6420);
            if constexpr (FULL && NR % 8 == 0) {
#pragma unroll
                for (int k = 0; k < NR / 8; k++) reinterpret_cast<uint4 *>(dst)[k] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
            } else if constexpr (NR % 4 == 0) {
#pragma unroll
                for (int k = 0; k < NR / 4; k++) reinterpret_cast<uint2 *>(dst)[k] = make_uint2(w[2 * k], w[2 * k + 1]);
            } else {
#pragma unroll
                for (int k = 0; k < NR / 2; k++) reinterpret_cast<uint32_t *>(dst)[k] = w[k];
            }
        } else {
#pragma unroll
            for (int k = 0; k < NR; k++)
                if (k < nvalid) reinterpret_cast<uint16_t *>(dst)[k] = (uint16_t)__byte_perm(q[k], 0u, 0x4420);
        }
    } else {
#pragma unroll
        for (int k = 0; k < NR; k++)
            if (FULL || k < nvalid) reinterpret_cast<uint16_t *>(dst)[k] = (uint16_t)__byte_perm(q[k], 0u, 0x4420);
    }
}

// Per-lane view of one chain: cursors are 32-bit offsets in units of 8 bytes (D % 8 == 0) from the volume base, turned
// into addresses with one IMAD.WIDE; the cost ring is kRing slots of LPC * 2NR bytes per chain.
template <int NR, int LPC, bool FULL> struct ChainRun {
    static constexpr int kSlotBytes = LPC * 2 * NR;
    static constexpr unsigned kRingBytes = kRing * kSlotBytes;
    // the two chains of a warp read their rings in the same instruction: offset the second ring by 16 banks
    static constexpr unsigned kChainPitch = kRingBytes + ((kRingBytes % 128 == 0 && LPC < 32) ? 64 : 0);
    // FULL (D == LPC * 2NR, a multiple of 16): 16-byte copies, LPC * NR / 8 of them per cell; otherwise 8-byte copies
    static constexpr int kRounds = (NR + 3) / 4;    // 8-byte cp.async rounds: LPC lanes fetch LPC * 8 bytes per round
    static constexpr int kRounds16 = (NR + 7) / 8;  // 16-byte rounds
    const uint8_t *fused_lane; // fused + sl * 8 (or sl * 16)
    uint8_t *q_lane;           // path volume + sl * 2NR
    unsigned ring_ld;          // shared address of the chain's ring + sl * 2NR (reads)
    unsigned ring_st;          // shared address of the chain's ring + sl * 8 (or sl * 16): cp.async destination
    int D, sl;
    unsigned on16[kRounds16];  // FULL: does this lane copy in round r
    unsigned rd_off = 0;                           // ring slot of the step being consumed
    unsigned wr_off = (kRing - 1) * kSlotBytes;    // free slot: the one consumed in the previous step

    __device__ __forceinline__ void issue(unsigned slot_off, int off8) const
    {
        const uint8_t *src = fused_lane + (long long)off8 * 8;
        if constexpr (FULL) {
#pragma unroll
            for (int r = 0; r < kRounds16; r++) {
                if ((r + 1) * LPC * 16 <= LPC * 2 * NR) cp_async16(ring_st + slot_off + r * LPC * 16, src + r * LPC * 16);
                else cp_async16_if(ring_st + slot_off + r * LPC * 16, src + r * LPC * 16, on16[r]);
            }
        } else {
#pragma unroll
            for (int r = 0; r < kRounds; r++)
                if ((sl + r * LPC) * 8 < D) cp_async8(ring_st + slot_off + r * LPC * 8, src + r * LPC * 8);
        }
    }
    __device__ __forceinline__ void consume(uint32_t (&c)[NR]) const
    {
        cp_async_wait<kRing - 2>();
        __syncwarp();
        const unsigned src = ring_ld + rd_off;
        if constexpr (NR % 4 == 0) {
#pragma unroll
            for (int k = 0; k < NR / 4; k++) {
                uint32_t v0, v1;
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];\n" : "=r"(v0), "=r"(v1) : "r"(src + 8 * k) : "memory");
                c[4 * k] = __byte_perm(v0, 0u, 0x4140);
                c[4 * k + 1] = __byte_perm(v0, 0u, 0x4342);
                c[4 * k + 2] = __byte_perm(v1, 0u, 0x4140);
                c[4 * k + 3] = __byte_perm(v1, 0u, 0x4342);
            }
        } else if constexpr (NR % 2 == 0) {
#pragma unroll
            for (int k = 0; k < NR / 2; k++) {
                uint32_t v;
                asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(src + 4 * k) : "memory");
                c[2 * k] = __byte_perm(v, 0u, 0x4140);
                c[2 * k + 1] = __byte_perm(v, 0u, 0x4342);
            }
        } else {
#pragma unroll
            for (int k = 0; k < NR; k++) {
                unsigned short v;
                asm volatile("ld.shared.u16 %0, [%1];\n" : "=h"(v) : "r"(src + 2 * k) : "memory");
                c[k] = __byte_perm((uint32_t)v, 0u, 0x4140);
            }
        }
    }
    // returns the free slot (consumed one step ago, every lane is past its reads: a __syncwarp lies in between),
    // frees the slot consumed in this step for the next one and moves on
    __device__ __forceinline__ unsigned advance_ring()
    {
        const unsigned free_slot = wr_off;
        wr_off = rd_off;
        rd_off = (rd_off + kSlotBytes == kRingBytes) ? 0u : rd_off + kSlotBytes;
        return free_slot;
    }
};

} // namespace sister
