// sister_b200 / sgm.cu -- semi-global aggregation and final selection (sm_100a).
//
// Restates accumulateCostsSSE (sgm.cpp:26-455; P1 = 7, P2 = 100, 2 passes x 4 paths) on the uint8 fused volume C.
//
// Decomposition (proven equal to the reference recurrence on the CPU by tests/sgm_spec.py):
//   * Each of the 8 paths is a set of INDEPENDENT chains (rows for r0, columns for r2, wrapped diagonals for r1/r3).
//     A chain occupies 8 or 16 lanes of a warp (4 or 2 chains per warp); the D disparities are spread over those
//     lanes, 2*NR consecutive disparities per lane, two per 32-bit register as packed 16-bit lanes in the "split"
//     layout of sgm_core.cuh (VIMNMX3.S16x2 / VIADDMNMX.S16x2 on sm_100a: 4 instructions per register and step).
//   * The state a chain carries is normalised and clamped:  a(d) = min(L(d) - min_d L, P2).  With it the reference update
//         L'(d) = C(d) (+) ( min(L(d), L(d-1) (+) P1, L(d+1) (+) P1, P2 (+) m) (-) m )                 sgm.cpp:282-297
//     becomes  Q(d) = min(a(d), a(d-1) + P1, a(d+1) + P1),  L'(d) = C(d) + Q(d),  with no saturation anywhere because
//     C <= 252 (match.cu) bounds L' by 352 and the 8-path sum by 2816. The reference's sentinels map as follows:
//         L(-1) = L(D) = 65535 (sgm.cpp:84-87)                 -> kInf2 in the neighbour slots (never the minimum)
//         off-image predecessor column: L = 65535, m = 0       -> a = P2 everywhere  => Q = P2   (sgm.cpp:57-81)
//         r0 at the start of a row: L = 0, m = 0               -> a = 0 everywhere   => Q = 0    (sgm.cpp:215-216)
//         first line of a pass: L1 = L2 = L3 = C, m = min C    -> a = min(C - min C, P2), nothing added to the sum
//         first line, r0: int32 arithmetic + 8-bit truncation  -> first_line_step (sgm.cpp:141-190, types.h:28)
//     The "cost == 255 -> 0" substitution of sgm.cpp:109,123,146 can never fire on a fused C <= 252; it is applied all
//     the same (first lines only), because the two-view path (sister_stereo, hpp:122-150) aggregates a raw volume.
//   * A chain does not touch the sum volume. It emits the penalty term Q(d) = L'(d) - C(d) in [0, P2] as ONE BYTE per
//     cell into the path's own volume (a diagonal chain that leaves the frame re-enters at the opposite border with
//     a = P2, so every column/diagonal chain has exactly Hp steps and every cell of a path volume is written once).
//   * k_sgm_final adds up  S = nC * C + sum of the 8 path bytes  (nC = 8, or 4 on the first line of either pass where
//     r1..r3 contribute nothing and r0's byte is the whole truncated value) and does the final WTA (hpp:283) and the
//     output encoding (hpp:111-118) in the same sweep; S itself is only written when a test asks for it.
//
// Traffic per padded cell: 8 x (1 B read C + 1 B write Q) + (1 + 8) B read = 25 B, no read-modify-write, against
// 40 B for the path-by-path accumulation into a uint16 sum volume this replaces (crop only: 14 B per padded cell).
//
// How the path kernel runs (with the measurements behind each choice):
//   * cells are stored in the CELL ORDER of common.cuh (lane-interleaved words when the chain's lanes are exactly full), so
//     a chain's load / store instruction covers 4 * lpc contiguous bytes and bytes <-> packed registers is two
//     instructions per word;
//   * full-lane chains stage the cost stream through a per-warp cp.async ring in shared memory (run_chain_ring: lookahead
//     of kRing - 1 steps, no lookahead registers), the other sizes keep the register lookahead of run_chain;
//   * between border crossings, away from the chain's end and on one side of a region edge the step loop runs a body
//     without wrap tests, guards or border selects ("fast phases");
//   * warps are handed out from an interleaved table (warp_order_for) so that every SM gets the same mix of sections.
// The kernel is bound by issue slots while all its blocks are resident and by the step's dependent latency in the thinly
// occupied second wave (5568 warps on 148 x 24 slots at config 2); it is not bound by DRAM (no stores: -0.12 ms, no
// loads: -0.09 ms of 0.74 ms).
#include <algorithm>
#include <cstdlib>
#include <initializer_list>
#include <map>
#include <mutex>
#include <type_traits>
#include <vector>

#include "kernels.cuh"
#include "sgm_core.cuh"

namespace sister {

#ifndef SISTER_SGM_CHAIN_WARPS
#define SISTER_SGM_CHAIN_WARPS 8
#endif
constexpr int kChainWarps = SISTER_SGM_CHAIN_WARPS;    // warps per block (they share nothing)
// 24 resident warps per SM (80 registers) up to this many packed registers per lane when every lane is full, else 16
#ifndef SISTER_SGM_FULL_NR20
#define SISTER_SGM_FULL_NR20 0
#endif
#ifndef SISTER_SGM_FULL_NR32
#define SISTER_SGM_FULL_NR32 0
#endif
#ifndef SISTER_SGM_NR24
#define SISTER_SGM_NR24 8
#endif
#ifndef SISTER_SGM_FULL_NR24
#define SISTER_SGM_FULL_NR24 12
#endif
constexpr int resident_warps(int NR, bool FULL)
{
    return (FULL && NR <= SISTER_SGM_FULL_NR32) ? 32 : (NR <= SISTER_SGM_NR24 || (FULL && NR <= SISTER_SGM_FULL_NR24)) ? 24 : (FULL && NR <= SISTER_SGM_FULL_NR20) ? 20 : 16;
}

// ---------------------------------------------------------------------------------------------- chain geometry

struct Chain {
    int i, j;       // first cell
    int si, sj;     // step
    int enter;      // column a diagonal chain re-enters at after leaving the frame
    int vol;        // path volume index: 4 * pass + path
};

// Region of interest: the cells whose aggregated cost is consumed. The caller only ever sees the crop
// Rect(D, D, W, H) of the disparity map (hpp:116-118), so outside of test / raw-disparity runs the path bytes are
// needed there and nowhere else. SGM state still flows in from the borders of the padded frame, so a chain starts
// where it always did, but
//   * a chain that never touches the region is not run (rows above / below it, columns left / right of it);
//   * a chain stops once it has left the region for good (its remaining cells feed nothing);
//   * path bytes are stored, and summed by k_sgm_final, inside the region only.
// The full frame (r0 = c0 = 0, r1 = Hp, c1 = Wp) reproduces the reference's whole aggregated volume.
struct Roi {
    int r0, r1, c0, c1;
};
__host__ __device__ inline bool roi_is_full(const Roi &r, const Dims &d) { return r.r0 == 0 && r.c0 == 0 && r.r1 == d.Hp && r.c1 == d.Wp; }

// Row band [b0, b1) of the padded frame: what one GPU owns when a large frame is split over several (SURVEY section
// 8(e)). Row chains of the band's rows are local. A column / diagonal chain is cut at the band's borders: it picks
// up its state -- the clamped normalised vector a(d), one byte per disparity, D bytes per chain -- where the
// neighbouring band left it (in[pass]) and leaves it for the next band (out[pass]). Pass 0 flows down (from the band
// above, to the band below), pass 1 up. State layout: [path r1, r2, r3][first-line column of the chain][D].
// The whole frame is the band b0 = 0, b1 = Hp with no state pointers.
struct Band {
    int b0, b1;
    const uint8_t *in[2];
    uint8_t *out[2];
};

// Chain numbering: nine sections, each padded to a multiple of `cpw` (chains per warp) so that a warp never mixes
// sections; a padding slot repeats the section's last chain (it recomputes and rewrites the same bytes):
//   0        kind 1   r0 on the first line of pass 0 / pass 1 (rows 0 and Hp-1; only when the region contains them)
//   1, 2     kind 0   r0 of pass 0 / pass 1 on the other rows of the region
//   3 .. 8   kind 2   [pass][path r1, r2, r3][column]: the chains of a warp sit on adjacent columns of the same row, so
//                     a step reads and writes one contiguous run of cells; the three paths of a pass start together
//                     and advance at the same rate, so every row of C is read three times within a short window and
//                     two of the three reads hit L2. r2 runs on the region's columns only, the diagonals on all.
struct Sections {
    int n[9], lo[9];   // chains in the section, first row / column
    long long o[10];   // first chain index (padded)
};
inline Sections chain_sections(const Dims &d, const Roi &r, const Band &bd, int cpw)
{
    Sections s;
    const bool first = r.r0 == 0 && bd.b0 == 0, last = r.r1 == d.Hp && bd.b1 == d.Hp;
    const int lo = r.r0 > bd.b0 ? r.r0 : bd.b0, hi = r.r1 < bd.b1 ? r.r1 : bd.b1; // rows of the region inside the band
    s.n[0] = (first ? 1 : 0) + (last ? 1 : 0);
    s.lo[0] = first ? 0 : 1; // pass of the section's first chain
    s.lo[1] = lo > 1 ? lo : 1;                            // pass 0: row 0 is the first line
    s.n[1] = hi - s.lo[1];
    s.lo[2] = lo;                                         // pass 1: row Hp-1 is the first line
    s.n[2] = (hi < d.Hp - 1 ? hi : d.Hp - 1) - lo;
    for (int p = 0; p < 2; p++)
        for (int t = 0; t < 3; t++) {
            // pass 0 walks rows [b0, min(b1, r1)), pass 1 rows [max(b0, r0), b1) downwards: none if that is empty
            const bool live = p == 0 ? bd.b0 < (bd.b1 < r.r1 ? bd.b1 : r.r1) : (bd.b0 > r.r0 ? bd.b0 : r.r0) < bd.b1;
            s.lo[3 + 3 * p + t] = t == 1 ? r.c0 : 0;
            s.n[3 + 3 * p + t] = !live ? 0 : t == 1 ? r.c1 - r.c0 : d.Wp;
        }
    s.o[0] = 0;
    for (int k = 0; k < 9; k++) {
        if (s.n[k] < 0) s.n[k] = 0;
        s.o[k + 1] = s.o[k] + ((long long)s.n[k] + cpw - 1) / cpw * cpw;
    }
    return s;
}

// returns the kind (0..2) or -1 when g is past the end
__device__ __forceinline__ int chain_decode(const Dims &d, const Roi &r, const Band &bd, const Sections &sec, long long g, Chain &ch, int &nsteps,
                                            int &section, long long &state_off, bool &imports, bool &exports)
{
    imports = exports = false;
    state_off = 0;
    if (g >= sec.o[9]) return -1;
    int k = 0;
#pragma unroll
    for (int q = 1; q < 9; q++) k += g >= sec.o[q];
    section = k;
    const int idx = (int)min(g - sec.o[k], (long long)sec.n[k] - 1);
    if (k == 0) {
        const int p = sec.lo[0] + idx;
        ch.i = p ? d.Hp - 1 : 0; ch.j = p ? d.Wp - 1 : 0; ch.si = 0; ch.sj = p ? -1 : 1; ch.enter = 0;
        ch.vol = 4 * p; nsteps = p ? d.Wp - r.c0 : r.c1;
        return 1;
    }
    if (k < 3) {
        const int p = k - 1;
        ch.i = sec.lo[k] + idx; ch.j = p ? d.Wp - 1 : 0; ch.si = 0; ch.sj = p ? -1 : 1; ch.enter = 0;
        ch.vol = 4 * p; nsteps = p ? d.Wp - r.c0 : r.c1;
        return 0;
    }
    const int p = (k - 3) / 3, type = (k - 3) % 3; // 0: r1, 1: r2, 2: r3
    const int dj = p ? -1 : 1, j1 = p ? d.Wp - 1 : 0, jl = p ? 0 : d.Wp - 1;
    const int c = sec.lo[k] + idx;                 // the chain's column on the first line of the pass
    ch.si = dj;
    ch.sj = type == 0 ? dj : type == 1 ? 0 : -dj;
    ch.enter = type == 0 ? j1 : jl; // unused by r2 (sj = 0 never leaves the frame)
    ch.vol = 4 * p + 1 + type;
    // rows of this band the chain walks, and how many steps lie behind it when it enters the band
    int done;
    if (p == 0) {
        const int end = bd.b1 < r.r1 ? bd.b1 : r.r1;
        ch.i = bd.b0; nsteps = end - bd.b0; done = bd.b0;
        imports = bd.b0 > 0; exports = end < r.r1;
    } else {
        const int end = bd.b0 > r.r0 ? bd.b0 : r.r0;
        ch.i = bd.b1 - 1; nsteps = bd.b1 - end; done = d.Hp - bd.b1;
        imports = bd.b1 < d.Hp; exports = end > r.r0;
    }
    int j = (c + (int)(((long long)ch.sj * done) % d.Wp)) % d.Wp; // a wrapped diagonal advances modulo Wp
    if (j < 0) j += d.Wp;
    ch.j = j;
    state_off = ((long long)type * d.Wp + c) * d.D;
    return 2;
}

// ---------------------------------------------------------------------------------------------- the path kernel

// Cursor of a chain: 32-bit offset in units of 8 bytes (D % 8 == 0) from the volume base plus the cell's row and
// column: a diagonal chain needs the column to notice that it stepped over a side border (it then re-enters at the
// opposite border of the same row, a fixed correction of one row of cells), the store cursor needs both to know
// whether the cell lies in the region of interest.
struct Cursor {
    int off8, i, j;
};
struct Walk {
    int stride8, wrapfix8, si, sj, enter, Wp;
    int r0, c0;            // region of interest ...
    unsigned nr, nc;       // ... and its height / width
};
template <bool DIAG> __device__ __forceinline__ bool advance(Cursor &c, const Walk &w)
{
    c.off8 += w.stride8;
    c.j += w.sj;
    if constexpr (DIAG) {
        c.i += w.si;
        if ((unsigned)c.j >= (unsigned)w.Wp) { c.j = w.enter; c.off8 += w.wrapfix8; return true; }
    }
    return false;
}
template <bool DIAG> __device__ __forceinline__ bool in_roi(const Cursor &c, const Walk &w)
{
    const bool col = (unsigned)(c.j - w.c0) < w.nc;
    if constexpr (DIAG) return col && (unsigned)(c.i - w.r0) < w.nr;
    else return col; // a row chain only runs on rows of the region
}

// KIND 0: r0 on an ordinary row; 1: r0 on the first line of a pass; 2: r1 / r2 / r3 (columns ride along in the
// wrapped-diagonal loop, sj = 0 never wraps).
template <int NR, int LPC, bool FULL, bool IL, int KIND>
__device__ __forceinline__ void run_chain(const uint8_t *__restrict__ fused_lane, const uint8_t *__restrict__ fused_pfl, uint8_t *__restrict__ q_lane, const LaneInfo<NR, LPC, FULL> &li,
                                          const Walk &wk, Cursor first, int nsteps, int valid_bytes,
                                          const uint8_t *state_in = nullptr, uint8_t *state_out = nullptr, bool starts_on_first_line = true)
{
    constexpr bool DIAG = KIND == 2;
    uint32_t buf[kAhead][NR / 2];
    Cursor ld = first, pf, st = first;
    // prologue: kAhead cells in registers, kFar more requested from L2 (nsteps >= 12 > kAhead + kFar is not required:
    // every request is guarded by the step count)
#pragma unroll
    for (int t = 0; t < kAhead; t++) {
        load_cost<NR, LPC, FULL, IL>(fused_lane + (long long)ld.off8 * 8, valid_bytes, buf[t]); // nsteps >= 12 (check_shape, sister_test_sgm)
        advance<DIAG>(ld, wk); // nsteps > kAhead: ld now points at step kAhead
    }
    // the first line of a pass reads an invalid cost (255, census.cpp:76) as 0 (sgm.cpp:109,123,146); fused volumes
    // never hold 255 (match.cu), the raw two-view volume of sister_stereo does
    if (KIND == 2 && starts_on_first_line) {
#pragma unroll
        for (int k = 0; k < NR / 2; k++) buf[0][k] &= ~__vcmpeq4(buf[0][k], 0xFFFFFFFFu);
    }
    pf = ld;
#pragma unroll 1
    for (int t = 0; t < kFar && kAhead + t < nsteps; t++) {
        if (valid_bytes > 0) prefetch_l2(fused_pfl + (long long)pf.off8 * 8);
        advance<DIAG>(pf, wk);
    }
    ChainState<NR> cs;
    uint32_t mm = 0;          // KIND 1: minimum of the truncated state
    // KIND 0: a = 0 at the start of a row (sgm.cpp:215-216). KIND 2: the first cell lies on the first line of the pass
    // where L = C and nothing is added to the sum (sgm.cpp:103-138) -- exactly what a step from a = 0 produces
    // (q = 0, L = C). KIND 1 takes L = C in its first column and ignores the state.
    chain_set<NR, LPC, FULL>(cs, 0u, li);
    if (KIND == 2 && state_in) { // warp-uniform: the chain continues from the band before (Band)
        uint32_t w[NR / 2], a[NR];
        load_cost<NR, LPC, FULL, IL>(state_in, valid_bytes, w);
        unpack_cost<NR, IL>(w, a);
        chain_resume<NR, LPC, FULL>(cs, a, li);
    }
    // one step: consume buffer u (step s), refill it with step s + kAhead (past the end of the chain the last cell is
    // simply loaded again: an unconditional load keeps the buffer in place, a predicated one costs a copy per register)
    auto step = [&](const int u, const int s) {
        uint32_t c[NR], q[NR];
        if constexpr (KIND == 1) { // every cell of this chain lies on the first line
#pragma unroll
            for (int k = 0; k < NR / 2; k++) buf[u][k] &= ~__vcmpeq4(buf[u][k], 0xFFFFFFFFu);
        }
        unpack_cost<NR, IL>(buf[u], c);
        load_cost<NR, LPC, FULL, IL>(fused_lane + (long long)ld.off8 * 8, valid_bytes, buf[u]);
        if (s + kAhead + 1 < nsteps) advance<DIAG>(ld, wk);
        if (s + kAhead + kFar < nsteps) {
            if (FULL || valid_bytes > 0) prefetch_l2(fused_pfl + (long long)pf.off8 * 8);
            advance<DIAG>(pf, wk);
        }
        uint8_t *dst = q_lane + (long long)st.off8 * 8;
        const bool wanted = in_roi<DIAG>(st, wk);
        const bool off_next = advance<DIAG>(st, wk); // KIND 2: the next cell follows a border crossing
        if constexpr (KIND == 1) {
            first_line_step<NR, LPC, FULL>(cs.a, cs.b, mm, c, li, s == 0, q);
        } else if constexpr (KIND == 0) {
            chain_step<NR, LPC, FULL>(cs, c, li, q);
        } else {
            chain_step<NR, LPC, FULL>(cs, c, li, q, off_next);
        }
        if (wanted) store_q<NR, LPC, FULL, IL>(dst, q, valid_bytes);
    };
    // The same step between border crossings and away from the chain's end, where all three cursors are plain arithmetic
    // progressions: the load / prefetch addresses are fixed offsets from the store cursor, no wrap test, no end-of-chain
    // guard, no border select in the renormalisation. Whether a step's cell can lie in the region of interest is
    // warp-uniform (a section's chains sit on the same row -- KIND 2 -- or column -- KIND 0 -- at the same step): steps
    // [w_lo, w_hi). A fast run never straddles one of those edges, so it either stores nothing (STORE = false) or
    // stores under the per-lane column test only (KIND 2) / unconditionally (KIND 0). The kernel is bound by the ALU
    // pipe; the general step costs about 20 more instructions of that pipe.
    int w_lo, w_hi;
    if constexpr (DIAG) {
        w_lo = wk.si > 0 ? wk.r0 - first.i : first.i - (wk.r0 + (int)wk.nr - 1);
        w_hi = wk.si > 0 ? wk.r0 + (int)wk.nr - first.i : first.i - wk.r0 + 1;
    } else {
        w_lo = wk.sj > 0 ? wk.c0 - first.j : first.j - (wk.c0 + (int)wk.nc - 1);
        w_hi = wk.sj > 0 ? wk.c0 + (int)wk.nc - first.j : first.j - wk.c0 + 1;
    }
    const int ld_ahead = kAhead * wk.stride8, pf_ahead = (kAhead + kFar) * wk.stride8;
    auto fast_step = [&](const int u, auto store_tag) {
        constexpr bool STORE = decltype(store_tag)::value;
        uint32_t c[NR], q[NR];
        unpack_cost<NR, IL>(buf[u], c);
        load_cost<NR, LPC, FULL, IL>(fused_lane + (long long)(st.off8 + ld_ahead) * 8, valid_bytes, buf[u]);
        if (FULL || valid_bytes > 0) prefetch_l2(fused_pfl + (long long)(st.off8 + pf_ahead) * 8);
        chain_step<NR, LPC, FULL>(cs, c, li, q);
        if constexpr (STORE) {
            uint8_t *dst = q_lane + (long long)st.off8 * 8;
            if constexpr (DIAG) {
                if ((unsigned)(st.j - wk.c0) < wk.nc) store_q<NR, LPC, FULL, IL>(dst, q, valid_bytes);
                st.j += wk.sj;
            } else {
                store_q<NR, LPC, FULL, IL>(dst, q, valid_bytes);
            }
        }
        st.off8 += wk.stride8;
    };
    int s0 = 0;
#pragma unroll 1
    while (s0 + kAhead <= nsteps) {
        int nfast = 0;
        bool inside = false;
        if constexpr (KIND != 1) {
            // advances the store cursor can make before the one that crosses a border, the fewest over the warp's chains;
            // the prefetch cursor runs kAhead + kFar steps ahead of it
            int room = nsteps - s0 - (kAhead + kFar) - 1;
            if constexpr (DIAG) {
                const int tw = wk.sj > 0 ? wk.Wp - 1 - st.j : wk.sj < 0 ? st.j : 0x3FFFFFFF;
                room = min(room, __reduce_min_sync(kFull, tw) - (kAhead + kFar));
            }
            inside = s0 >= w_lo && s0 < w_hi;
            room = min(room, (s0 < w_lo ? w_lo : s0 < w_hi ? w_hi : 0x3FFFFFFF) - s0);
            nfast = room >= 2 * kAhead ? room / kAhead : 0; // groups of kAhead steps
        }
        if (nfast > 0) {
            if (inside) {
#pragma unroll 1
                for (int g = 0; g < nfast; g++) {
#pragma unroll
                    for (int u = 0; u < kAhead; u++) fast_step(u, std::true_type{});
                }
                if constexpr (!DIAG) st.j += nfast * kAhead * wk.sj;
            } else {
#pragma unroll 1
                for (int g = 0; g < nfast; g++) {
#pragma unroll
                    for (int u = 0; u < kAhead; u++) fast_step(u, std::false_type{});
                }
                st.j += nfast * kAhead * wk.sj;
            }
            st.i += nfast * kAhead * wk.si;
            s0 += nfast * kAhead;
            // the general step's cursors again: load at step s0 + kAhead, prefetch at s0 + kAhead + kFar (no crossing in between)
            ld = st; ld.off8 += ld_ahead; ld.j += kAhead * wk.sj; ld.i += kAhead * wk.si;
            pf = st; pf.off8 += pf_ahead; pf.j += (kAhead + kFar) * wk.sj; pf.i += (kAhead + kFar) * wk.si;
        } else {
#pragma unroll
            for (int u = 0; u < kAhead; u++) step(u, s0 + u);
            s0 += kAhead;
        }
    }
#pragma unroll
    for (int u = 0; u < kAhead - 1; u++)
        if (s0 + u < nsteps) step(u, s0 + u); // warp-uniform
    if (KIND == 2 && state_out) store_q<NR, LPC, FULL, IL>(state_out, cs.a, valid_bytes); // a <= P2 fits a byte
}

// ---- the same chain with the fused costs staged through a per-warp shared-memory ring (full lanes: D == 2 * NR * LPC).
// The kernel is bound by the latency of its cost loads (stall sampling: long scoreboard on the first use of a loaded
// cell), registers cap the lookahead of run_chain at 3 steps and more resident warps help more than more lookahead. Here
// every step the warp copies the 32 / LPC cells of the step kRing - 1 ahead with 16-byte cp.async (cells are 16-byte
// aligned, a warp step is 4 * NR chunks) into the ring slot it consumed one step ago and reads its own words of the
// current slot back: no lookahead registers, a lookahead of kRing - 1 steps.
#ifndef SISTER_SGM_RING
#define SISTER_SGM_RING 8
#endif
constexpr int kRing = SISTER_SGM_RING;

__device__ __forceinline__ void cp_async16(unsigned dst_s, const uint8_t *src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst_s), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint32_t lds32(unsigned a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint2 lds64v(unsigned a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];\n" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint4 lds128v(unsigned a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory"); return v; }

// the lane's NR / 2 words of its chain's cell in a ring slot (a = slot + the lane's first byte)
template <int NR, int LPC, bool IL> __device__ __forceinline__ void ring_read(unsigned a, uint32_t (&w)[NR / 2])
{
    if constexpr (IL) {
#pragma unroll
        for (int k = 0; k < NR / 2; k++) w[k] = lds32(a + 4 * LPC * k);
    } else if constexpr (NR % 8 == 0) {
#pragma unroll
        for (int k = 0; k < NR / 8; k++) { const uint4 v = lds128v(a + 16 * k); w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w; }
    } else if constexpr (NR % 4 == 0) {
#pragma unroll
        for (int k = 0; k < NR / 4; k++) { const uint2 v = lds64v(a + 8 * k); w[2 * k] = v.x; w[2 * k + 1] = v.y; }
    } else {
#pragma unroll
        for (int k = 0; k < NR / 2; k++) w[k] = lds32(a + 4 * k);
    }
}

template <int NR, int LPC, bool IL, int KIND>
__device__ __forceinline__ void run_chain_ring(const uint8_t *__restrict__ fused, uint8_t *__restrict__ q_lane, const LaneInfo<NR, LPC, true> &li, const Walk &wk,
                                               Cursor first, int nsteps, unsigned ring_s, int lane, const uint8_t *state_in = nullptr,
                                               uint8_t *state_out = nullptr, bool starts_on_first_line = true)
{
    constexpr bool DIAG = KIND == 2;
    constexpr int D = 2 * NR * LPC;    // bytes per cell
    constexpr int CH = D / 16;         // 16-byte chunks per cell
    constexpr int NCH = 4 * NR;        // chunks per warp and step: 32 / LPC cells
    constexpr int NCP = (NCH + 31) / 32;
    constexpr int SS = (32 / LPC) * D; // bytes per ring slot
    constexpr int R = kRing, A = R - 1;
    constexpr int U = R / 2;           // steps per trip of the main loops: half a ring, so that slot offsets are immediates
    static_assert(R == 2 * U, "a trip of the step loops walks half of the ring");
    // copy cursors: chunk g = lane + 32 n of the warp step belongs to the cell of chain g / CH
    Cursor cp[NCP];
    const uint8_t *cp_src[NCP];
    unsigned cp_dst[NCP];
    int cp_stride8[NCP];
    bool cp_on[NCP];
#pragma unroll
    for (int n = 0; n < NCP; n++) {
        const int g = lane + 32 * n;
        cp_on[n] = g < NCH;
        const int c = cp_on[n] ? g / CH : 0;
        cp[n].off8 = __shfl_sync(kFull, first.off8, c * LPC);
        cp[n].i = __shfl_sync(kFull, first.i, c * LPC);
        cp[n].j = __shfl_sync(kFull, first.j, c * LPC);
        cp_src[n] = fused + (g - c * CH) * 16;
        cp_dst[n] = ring_s + g * 16;
        // the two first-line chains of a warp (KIND 1) run in opposite directions: the stride is that of the chain copied for
        cp_stride8[n] = __shfl_sync(kFull, wk.stride8, c * LPC);
    }
    // copy the cells of step t (where the copy cursors stand) into the slot at byte offset slot_off, move on
    auto copy_step = [&](const int t, const unsigned slot_off) {
        if (t < nsteps) {
#pragma unroll
            for (int n = 0; n < NCP; n++)
                if (cp_on[n]) cp_async16(cp_dst[n] + slot_off, cp_src[n] + (long long)cp[n].off8 * 8);
            if (t + 1 < nsteps) {
#pragma unroll
                for (int n = 0; n < NCP; n++) {
                    if constexpr (DIAG) advance<DIAG>(cp[n], wk); // one section, one walk
                    else cp[n].off8 += cp_stride8[n];
                }
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int t = 0; t < A; t++) copy_step(t, t * SS);
    const unsigned rd_lane = ring_s + (lane / LPC) * D + li.template cell_offset<IL>();
    // Step u of a trip reads slot half / SS + u and refills the slot consumed one step earlier with the step A ahead; the
    // trips alternate between the ring's two halves (half <-> other), so inside a trip every slot offset is an immediate.
    unsigned half = 0, other = U * SS;
    Cursor st = first;
    ChainState<NR> cs;
    uint32_t mm = 0;          // KIND 1: minimum of the truncated state
    // initial state as in run_chain
    chain_set<NR, LPC, true>(cs, 0u, li);
    if (KIND == 2 && state_in) { // warp-uniform: the chain continues from the band before (Band)
        uint32_t w[NR / 2], a[NR];
        load_cost<NR, LPC, true, IL>(state_in, 2 * NR, w);
        unpack_cost<NR, IL>(w, a);
        chain_resume<NR, LPC, true>(cs, a, li);
    }
    // the slot of the current step has landed (every lane waits for its own copies, then the warp meets); read it
    auto fetch = [&](const int u, uint32_t (&c)[NR], const bool zero_invalid) {
        cp_async_wait<A - 1>();
        __syncwarp();
        uint32_t w[NR / 2];
        ring_read<NR, LPC, IL>(rd_lane + half + u * SS, w);
        // the first line of a pass reads an invalid cost (255, census.cpp:76) as 0 (sgm.cpp:109,123,146); fused volumes
        // never hold 255 (match.cu), the raw two-view volume of sister_stereo does
        if (zero_invalid) {
#pragma unroll
            for (int k = 0; k < NR / 2; k++) w[k] &= ~__vcmpeq4(w[k], 0xFFFFFFFFu);
        }
        unpack_cost<NR, IL>(w, c);
    };
    auto wr_off = [&](const int u) { return u == 0 ? other + (U - 1) * SS : half + (u - 1) * SS; };
    auto next_trip = [&]() { const unsigned t = half; half = other; other = t; };
    auto step = [&](const int s, const int u) {
        uint32_t c[NR], q[NR];
        fetch(u, c, KIND == 1 || (KIND == 2 && starts_on_first_line && s == 0));
        copy_step(s + A, wr_off(u));
        uint8_t *dst = q_lane + (long long)st.off8 * 8;
        const bool wanted = in_roi<DIAG>(st, wk);
        const bool off_next = advance<DIAG>(st, wk); // KIND 2: the next cell follows a border crossing
        if constexpr (KIND == 1) {
            first_line_step<NR, LPC, true>(cs.a, cs.b, mm, c, li, s == 0, q);
        } else if constexpr (KIND == 0) {
            chain_step<NR, LPC, true>(cs, c, li, q);
        } else {
            chain_step<NR, LPC, true>(cs, c, li, q, off_next);
        }
        if (wanted) store_q<NR, LPC, true, IL>(dst, q, 2 * NR);
    };
    // The same step between border crossings and away from the chain's end, where every cursor is a plain arithmetic
    // progression: no wrap test, no end-of-chain guard, no border select in the renormalisation. Whether a step's cell can
    // lie in the region of interest is warp-uniform (a section's chains sit on the same row -- KIND 2 -- or column --
    // KIND 0 -- at the same step): steps [w_lo, w_hi). A fast run never straddles one of those edges, so it either stores
    // nothing or stores under the per-lane column test only (KIND 2) / unconditionally (KIND 0).
    int w_lo, w_hi;
    if constexpr (DIAG) {
        w_lo = wk.si > 0 ? wk.r0 - first.i : first.i - (wk.r0 + (int)wk.nr - 1);
        w_hi = wk.si > 0 ? wk.r0 + (int)wk.nr - first.i : first.i - wk.r0 + 1;
    } else {
        w_lo = wk.sj > 0 ? wk.c0 - first.j : first.j - (wk.c0 + (int)wk.nc - 1);
        w_hi = wk.sj > 0 ? wk.c0 + (int)wk.nc - first.j : first.j - wk.c0 + 1;
    }
    const uint8_t *cp_ptr[NCP]; // fast runs: the copy cursors as running pointers
    auto fast_step = [&](const int u, auto store_tag) {
        constexpr bool STORE = decltype(store_tag)::value;
        uint32_t c[NR], q[NR];
        fetch(u, c, false);
#pragma unroll
        for (int n = 0; n < NCP; n++) {
            if (cp_on[n]) cp_async16(cp_dst[n] + wr_off(u), cp_ptr[n]);
            cp_ptr[n] += (long long)cp_stride8[n] * 8;
        }
        cp_async_commit();
        chain_step<NR, LPC, true>(cs, c, li, q);
        if constexpr (STORE) {
            uint8_t *dst = q_lane + (long long)st.off8 * 8;
            if constexpr (DIAG) {
                if ((unsigned)(st.j - wk.c0) < wk.nc) store_q<NR, LPC, true, IL>(dst, q, 2 * NR);
                st.j += wk.sj;
            } else {
                store_q<NR, LPC, true, IL>(dst, q, 2 * NR);
            }
        }
        st.off8 += wk.stride8;
    };
    int s0 = 0;
#pragma unroll 1
    while (s0 + U <= nsteps) {
        int nfast = 0;
        bool inside = false;
        if constexpr (KIND != 1) {
            // advances the store cursor can make before the one that crosses a border, the fewest over the warp's chains;
            // the copy cursors run A steps ahead of it and must not reach the chain's last cell either
            int room = nsteps - s0 - A - 1;
            if constexpr (DIAG) {
                const int tw = wk.sj > 0 ? wk.Wp - 1 - st.j : wk.sj < 0 ? st.j : 0x3FFFFFFF;
                room = min(room, __reduce_min_sync(kFull, tw) - A);
            }
            inside = s0 >= w_lo && s0 < w_hi;
            room = min(room, (s0 < w_lo ? w_lo : s0 < w_hi ? w_hi : 0x3FFFFFFF) - s0);
            nfast = (room >= 2 * U && s0 > 0) ? room / U : 0; // trips of U steps; step 0 is special (first line)
        }
        if (nfast > 0) {
#pragma unroll
            for (int n = 0; n < NCP; n++) cp_ptr[n] = cp_src[n] + (long long)cp[n].off8 * 8;
            if (inside) {
#pragma unroll 1
                for (int g = 0; g < nfast; g++) {
#pragma unroll
                    for (int u = 0; u < U; u++) fast_step(u, std::true_type{});
                    next_trip();
                }
                if constexpr (!DIAG) st.j += nfast * U * wk.sj;
            } else {
#pragma unroll 1
                for (int g = 0; g < nfast; g++) {
#pragma unroll
                    for (int u = 0; u < U; u++) fast_step(u, std::false_type{});
                    next_trip();
                }
                st.j += nfast * U * wk.sj;
            }
            st.i += nfast * U * wk.si;
#pragma unroll
            for (int n = 0; n < NCP; n++) {
                cp[n].off8 += nfast * U * cp_stride8[n];
                cp[n].j += nfast * U * wk.sj; cp[n].i += nfast * U * wk.si;
            }
            s0 += nfast * U;
        } else {
#pragma unroll
            for (int u = 0; u < U; u++) step(s0 + u, u);
            next_trip();
            s0 += U;
        }
    }
#pragma unroll
    for (int u = 0; u < U - 1; u++)
        if (s0 + u < nsteps) step(s0 + u, u); // warp-uniform
    cp_async_wait<0>();
    if (KIND == 2 && state_out) store_q<NR, LPC, true, IL>(state_out, cs.a, 2 * NR); // a <= P2 fits a byte
}

// grid ceil(chains / (kChainWarps * 32 / LPC)), block kChainWarps * 32, dynamic shared memory: the warps' cost rings (full lanes)
template <int NR, int LPC, bool FULL, bool IL>
__global__ void __launch_bounds__(kChainWarps * 32, resident_warps(NR, FULL) / kChainWarps) k_sgm_paths(const uint8_t *__restrict__ fused, Dims d, Roi roi, Band band, Sections sec, uint8_t *__restrict__ qvol, unsigned section_mask,
                                                                                                             const int *__restrict__ warp_order, int n_warps, unsigned one)
{
    constexpr int CPW = 32 / LPC;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    LaneInfo<NR, LPC, FULL> li;
    li.init(lane, d.D);
    Chain ch;
    int nsteps = 0, section = 0;
    long long state_off = 0;
    bool imports, exports;
    const int w = blockIdx.x * kChainWarps + warp;
    if (w >= n_warps) return;
    // warp_order[w]: which CPW consecutive chains (of the padded numbering of Sections) this warp runs
    const int kind = chain_decode(d, roi, band, sec, (long long)__ldg(warp_order + w) * CPW + lane / LPC, ch, nsteps, section, state_off, imports,
                                  exports);
    if (kind < 0) return; // cannot happen: the table only holds warps of the requested sections
    const int D = d.D, Wp = d.Wp, D8 = D >> 3;
    const uint8_t *fused_lane = fused + li.template cell_offset<IL>();
    uint8_t *q_lane = qvol + (size_t)ch.vol * (size_t)d.cells + li.template cell_offset<IL>();
    const int valid_bytes = li.valid_bytes(D); // bytes of this lane inside the cell
    // per-lane constants derived from %tid: pin them in registers, otherwise ptxas re-derives them inside every step
    opaque(li.up_mask); opaque(li.dn_mask);
    li.one = one; // a kernel argument: the only 1 neither nvvm nor ptxas can fold (add_fma)
    // the prefetch requests of a chain's lanes cover the whole cell whatever its byte order
    const uint8_t *fused_pfl = fused + li.sl * 2 * NR;
    opaque_ptr(fused_lane); opaque_ptr(q_lane);
    if constexpr (IL) opaque_ptr(fused_pfl);
    else fused_pfl = fused_lane;
    Walk wk;
    wk.stride8 = (ch.si * Wp + ch.sj) * D8;
    wk.wrapfix8 = -ch.sj * Wp * D8;
    if (section_mask >> 31) wk.stride8 = wk.wrapfix8 = 0; // measurement aid (SISTER_DEBUG_PATH_KINDS bit 3): every step on the chain's first cell
    wk.si = ch.si; wk.sj = ch.sj; wk.enter = ch.enter; wk.Wp = Wp;
    wk.r0 = roi.r0; wk.c0 = roi.c0; wk.nr = (unsigned)(roi.r1 - roi.r0); wk.nc = (unsigned)(roi.c1 - roi.c0);
    Cursor first;
    first.off8 = (ch.i * Wp + ch.j) * D8;
    first.i = ch.i;
    first.j = ch.j;
    const int p = kind == 2 ? (section - 3) / 3 : 0;
    const uint8_t *sin = (kind == 2 && imports && band.in[p]) ? band.in[p] + state_off + li.template cell_offset<IL>() : nullptr;
    uint8_t *sout = (kind == 2 && exports && band.out[p]) ? band.out[p] + state_off + li.template cell_offset<IL>() : nullptr;
    if constexpr (FULL) {
        extern __shared__ __align__(16) unsigned char ring_raw[];
        const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring_raw) + warp * (kRing * (32 / LPC) * 2 * NR * LPC);
        if (kind == 1) run_chain_ring<NR, LPC, IL, 1>(fused, q_lane, li, wk, first, nsteps, ring_s, lane);
        else if (kind == 0) run_chain_ring<NR, LPC, IL, 0>(fused, q_lane, li, wk, first, nsteps, ring_s, lane);
        else run_chain_ring<NR, LPC, IL, 2>(fused, q_lane, li, wk, first, nsteps, ring_s, lane, sin, sout, !imports);
    } else {
        if (kind == 1) run_chain<NR, LPC, FULL, IL, 1>(fused_lane, fused_pfl, q_lane, li, wk, first, nsteps, valid_bytes);
        else if (kind == 0) run_chain<NR, LPC, FULL, IL, 0>(fused_lane, fused_pfl, q_lane, li, wk, first, nsteps, valid_bytes);
        else run_chain<NR, LPC, FULL, IL, 2>(fused_lane, fused_pfl, q_lane, li, wk, first, nsteps, valid_bytes, sin, sout, !imports);
    }
}

// ---------------------------------------------------------------------------------------------- final sum + WTA + encode

__device__ __forceinline__ uint2 ldg8(const uint8_t *p) { return __ldg(reinterpret_cast<const uint2 *>(p)); }

// S = nC * C + sum_v Q_v; WTALeft_SSE with uniqueness 1 (hpp:283): first-index argmin over d <= min(j, D-1); then
// convertTo(CV_16UC1), crop Rect(D, D, W, H) and * 255 with saturation (hpp:111-118).
// Eight lanes per pixel, each lane owns 8-byte chunks sub, sub + 8, ... of the pixel's D bytes in all nine volumes.
// grid-stride over groups of 4 pixels per warp.
template <bool IL>
__global__ void __launch_bounds__(256, 6) k_sgm_final(const uint8_t *__restrict__ fused, const uint8_t *__restrict__ qvol, Dims d, Roi roi,
                                                   uint16_t *__restrict__ sum, int16_t *__restrict__ raw_disp, uint16_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31, sub = lane & 7, grp = lane >> 3;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const int D = d.D, nchunk = D >> 3;
    const size_t cells = (size_t)d.cells;
    const int wroi = roi.c1 - roi.c0;
    const long long npx = (long long)(roi.r1 - roi.r0) * wroi; // pixels of the region of interest, row-major
    for (long long base = warp0 * 4; base < npx; base += nwarps * 4) {
        const long long t = base + grp;
        const bool live = t < npx;
        const int i = live ? roi.r0 + (int)(t / wroi) : 0, j = live ? roi.c0 + (int)(t % wroi) : 0;
        const long long pix = (long long)i * d.Wp + j;
        const int dmax = min(j, D - 1);
        const int sh = (i == 0 || i == d.Hp - 1) ? 2 : 3; // nC = 4 on the first line of either pass, else 8
        unsigned best = 0xFFFFFFFFu;
        if (live) {
            const size_t off0 = (size_t)pix * D;
            for (int chunk = sub; chunk < nchunk; chunk += 8) {
                const size_t off = off0 + chunk * 8;
                // disparities of the chunk's four byte pairs: db + (k & 1) * s1 + (k >> 1) * s2 and the one after it
                // (natural order: consecutive; lane-interleaved cell, common.cuh: words (t, sl), (t, sl + 1))
                int db = chunk * 8, s1 = 2, s2 = 4;
                if constexpr (IL) {
                    const int word = chunk * 2;
                    db = (word & (d.lpc - 1)) * 2 * d.nr + 2 * (word >> d.lpc_shift);
                    s1 = d.nr; s2 = 2 * d.nr;
                }
                const uint2 cc = ldg8(fused + off);
                uint2 qq[8];
#pragma unroll
                for (int v = 0; v < 8; v++) qq[v] = ldg8(qvol + (size_t)v * cells + off);
                // byte-wise pair sums stay below 256 except on first lines (r0's byte may be up to 255 there, its
                // partner r1 writes 0), so a plain 32-bit add is a 4-lane byte add
                uint32_t w[2][4];
#pragma unroll
                for (int k = 0; k < 4; k++) { w[0][k] = qq[2 * k].x + qq[2 * k + 1].x; w[1][k] = qq[2 * k].y + qq[2 * k + 1].y; }
                uint32_t S[4]; // 8 cells as packed u16: S[k] = bytes 2k, 2k + 1 of the chunk
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t cw = h ? cc.y : cc.x;
                    uint32_t lo = __byte_perm(cw, 0u, 0x4140) << sh, hi = __byte_perm(cw, 0u, 0x4342) << sh;
#pragma unroll
                    for (int k = 0; k < 4; k++) { lo += __byte_perm(w[h][k], 0u, 0x4140); hi += __byte_perm(w[h][k], 0u, 0x4342); }
                    S[2 * h] = lo; S[2 * h + 1] = hi;
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int da = db + (k & 1) * s1 + (k >> 1) * s2;
                    if (sum) *reinterpret_cast<uint32_t *>(sum + off0 + da) = S[k]; // the test tap is in natural order
                    if (da <= dmax) best = min(best, ((S[k] & 0xFFFFu) << 16) | (unsigned)da);
                    if (da + 1 <= dmax) best = min(best, (S[k] & 0xFFFF0000u) | (unsigned)(da + 1));
                }
            }
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) best = min(best, __shfl_xor_sync(kFull, best, o));
        if (live && sub == 0) {
            const int disp = (int)(best & 0xFFFFu);
            if (raw_disp) raw_disp[pix] = (int16_t)disp;
            const int oi = i - d.D, oj = j - d.D;
            if (out && oi >= 0 && oi < d.H && oj >= 0 && oj < d.W) out[(size_t)oi * d.W + oj] = (uint16_t)min(disp * 255, 65535);
        }
    }
}

// ---------------------------------------------------------------------------------------------- WTA-right on S (two-view path)

// WTARight_SSE on the aggregated volume (hpp:138, postprocess.cpp:187-315, uniqueness 1): R(i, j) = first-index argmin
// over d <= min(w-1-j, D-1) of S[i][j+d][d]. One warp per pixel, lanes over d; a row of S stays in L2 while its
// pixels are visited.
__global__ void __launch_bounds__(256) k_wta_right_sum(const uint16_t *__restrict__ sum, Dims d, int16_t *__restrict__ outR)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long pix = warp0; pix < d.px; pix += nwarps) {
        const int j = (int)(pix % d.Wp);
        const int dmax = min(d.Wp - 1 - j, d.D - 1);
        unsigned best = 0xFFFFFFFFu;
        for (int dd = lane; dd <= dmax; dd += 32) best = min(best, ((unsigned)sum[(size_t)(pix + dd) * d.D + dd] << 16) | (unsigned)dd);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(kFull, best, o));
        if (lane == 0) outR[pix] = (int16_t)(best & 0xFFFFu);
    }
}

void launch_wta_right_sum(const uint16_t *sum, const Dims &d, int16_t *outR, cudaStream_t st, LaunchCounter &lc)
{
    k_wta_right_sum<<<148 * 8, 256, 0, st>>>(sum, d, outR);
    lc.add();
}

// ---------------------------------------------------------------------------------------------- launch

// How a chain is spread over lanes, and with it the byte order of a cell (common.cuh): disparities per lane = 2 * nr, nr
// even, chosen so that D fits in lpc lanes. Crop-only the kernel is bound by the ALU pipe with few chains left:
// measured on B200 at D = 192, two chains per warp (twice the warps) beat four chains per warp by 20 %.
void set_cell_order(Dims &d)
{
    // measurement aid, only in builds with -DSISTER_DEBUG_HOOKS: SISTER_DEBUG_LPC8_MAXD=<largest D that runs four chains per warp>
    static int lpc8_max = -1;
    if (lpc8_max < 0) {
        lpc8_max = 128;
#ifdef SISTER_DEBUG_HOOKS
        if (const char *e = getenv("SISTER_DEBUG_LPC8_MAXD")) lpc8_max = atoi(e);
#endif
    }
    d.lpc = (d.D <= lpc8_max && d.D <= 192) ? 8 : 16;
    d.lpc_shift = d.lpc == 8 ? 3 : 4;
    d.nr = 2 * ((d.D + 4 * d.lpc - 1) / (4 * d.lpc));
    // interleave when the lanes are exactly full and k_fuse's shared-memory reads stay conflict-free in that order: the
    // codes a half warp reads are 2nr apart in groups of four lanes, distinct modulo 16 iff nr = 2 (mod 4): D = 32, 96
    // (lpc 8), 64, 192, 320, 448 (lpc 16)
    d.interleaved = (d.D == 2 * d.lpc * d.nr && d.nr % 4 == 2) ? 1 : 0;
}

// measurement aid, only in builds with -DSISTER_DEBUG_HOOKS (the shipped library never reads the environment):
// SISTER_DEBUG_PATH_KINDS=<bit mask of chain kinds to run> (results are then incomplete)
static unsigned debug_section_mask()
{
#ifndef SISTER_DEBUG_HOOKS
    return 0x1FFu;
#endif
    static int m = -1;
    if (m < 0) {
        const char *e = getenv("SISTER_DEBUG_PATH_KINDS");
        const int kinds = e ? atoi(e) & 7 : 7;
        m = ((kinds & 2) ? 0x001 : 0) | ((kinds & 1) ? 0x006 : 0) | ((kinds & 4) ? 0x1F8 : 0) | ((e && (atoi(e) & 8)) ? (int)0x80000000u : 0);
    }
    return (unsigned)m;
}

// The order warps are handed out in (blocks take kChainWarps consecutive warps, the grid is dispatched in block order).
// Sections differ in cost per step and in length, and the three column / diagonal paths of a pass re-read each row of C
// within a window of steps only if they advance at the same rate on every SM. So the warps of the sections that run
// together are INTERLEAVED in proportion to the sections' sizes -- every block, hence every SM, gets the same mix -- in
// two groups: first the row chains (the longest) with pass 0, then pass 1, which mostly forms the second wave of blocks.
// Built once per launch geometry and kept on the device.
struct WarpOrder {
    int *dev = nullptr;
    int n = 0;
};
static const WarpOrder &warp_order_for(const Sections &sec, unsigned section_mask, int cpw, cudaStream_t st, LaunchCounter &lc)
{
    struct Key {
        long long o[10];
        unsigned mask;
        int cpw, device;
        bool operator<(const Key &b) const
        {
            for (int k = 0; k < 10; k++) if (o[k] != b.o[k]) return o[k] < b.o[k];
            if (mask != b.mask) return mask < b.mask;
            if (cpw != b.cpw) return cpw < b.cpw;
            return device < b.device;
        }
    };
    static std::map<Key, WarpOrder> cache;
    static std::mutex mu;
    Key key;
    for (int k = 0; k < 10; k++) key.o[k] = sec.o[k];
    key.mask = section_mask & 0x1FFu; key.cpw = cpw;
    cudaGetDevice(&key.device);
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    std::vector<int> order;
    auto merge = [&](std::initializer_list<int> group) {
        struct Item { double pos; int section, slot; };
        std::vector<Item> items;
        for (int k : group) {
            if (!((section_mask >> k) & 1u)) continue;
            const int nw = (int)((sec.o[k + 1] - sec.o[k]) / cpw);
            for (int w = 0; w < nw; w++) items.push_back({(w + 0.5) / nw, k, (int)(sec.o[k] / cpw) + w});
        }
        std::stable_sort(items.begin(), items.end(), [](const Item &a, const Item &b) { return a.pos < b.pos; });
        for (const Item &i : items) order.push_back(i.slot);
    };
    merge({0});
    merge({1, 2, 3, 4, 5});
    merge({6, 7, 8});
    WarpOrder wo;
    wo.n = (int)order.size();
    if (wo.n > 0) {
        // built once per (geometry, device); ordered before the first kernel that reads it by the stream it is copied on
        cudaError_t e = cudaMalloc((void **)&wo.dev, order.size() * sizeof(int));
        if (e == cudaSuccess) e = cudaMemcpyAsync(wo.dev, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st); // `order` dies with this call; other streams may use the table next
        if (e != cudaSuccess) {
            lc.fail(e);
            if (wo.dev) cudaFree(wo.dev);
            static const WarpOrder none;
            return none; // not cached: the next launch tries again
        }
    }
    return cache.emplace(key, wo).first->second;
}

template <int NR, int LPC, bool FULL, bool IL>
static void launch_paths(const uint8_t *fused, const Dims &d, const Roi &roi, const Band &band, unsigned section_mask, uint8_t *qvol, cudaStream_t st, LaunchCounter &lc)
{
    constexpr int CPW = 32 / LPC;
    const Sections sec = chain_sections(d, roi, band, CPW);
    const WarpOrder &wo = warp_order_for(sec, section_mask, CPW, st, lc);
    if (wo.n == 0) return;
    const size_t smem = FULL ? (size_t)kChainWarps * kRing * CPW * d.D : 0; // the cost ring of run_chain_ring
    if (smem > 48 * 1024) lc.fail(optin_dynamic_smem((const void *)k_sgm_paths<NR, LPC, FULL, IL>, smem));
    k_sgm_paths<NR, LPC, FULL, IL><<<(unsigned)((wo.n + kChainWarps - 1) / kChainWarps), kChainWarps * 32, smem, st>>>(fused, d, roi, band, sec, qvol,
                                                                                                            section_mask, wo.dev, wo.n, 1u);
}

template <int LPC, int NRMAX>
static void launch_paths_lpc(const uint8_t *fused, const Dims &d, const Roi &roi, const Band &band, unsigned section_mask, uint8_t *qvol, cudaStream_t st, LaunchCounter &lc)
{
    const int nr = d.nr;
    const bool full = d.D == 2 * LPC * nr;
#define SISTER_PATHS_CASE(N)                                                                    \
    case N:                                                                                     \
        if constexpr (N <= NRMAX) {                                                             \
            if constexpr (N % 4 == 2) {                                                         \
                if (d.interleaved) { launch_paths<N, LPC, true, true>(fused, d, roi, band, section_mask, qvol, st, lc); break; } \
            }                                                                                   \
            if (full) launch_paths<N, LPC, true, false>(fused, d, roi, band, section_mask, qvol, st, lc);                    \
            else launch_paths<N, LPC, false, false>(fused, d, roi, band, section_mask, qvol, st, lc);                        \
        }                                                                                       \
        break;
    switch (nr) {
        SISTER_PATHS_CASE(2) SISTER_PATHS_CASE(4) SISTER_PATHS_CASE(6) SISTER_PATHS_CASE(8)
        SISTER_PATHS_CASE(10) SISTER_PATHS_CASE(12) SISTER_PATHS_CASE(14) SISTER_PATHS_CASE(16)
    }
#undef SISTER_PATHS_CASE
}

static Roi make_roi(const Dims &d, bool full_frame)
{
    Roi roi;
    if (full_frame) { roi.r0 = 0; roi.r1 = d.Hp; roi.c0 = 0; roi.c1 = d.Wp; }
    else { roi.r0 = d.D; roi.r1 = d.D + d.H; roi.c0 = d.D; roi.c1 = d.D + d.W; } // Rect(D, D, W, H), hpp:116-118
    return roi;
}

static void launch_paths_any(const uint8_t *fused, const Dims &d, const Roi &roi, const Band &band, unsigned section_mask, uint8_t *qvol,
                             cudaStream_t st, LaunchCounter &lc)
{
    if (d.lpc == 8) launch_paths_lpc<8, 12>(fused, d, roi, band, section_mask, qvol, st, lc); // four chains per warp
    else launch_paths_lpc<16, 16>(fused, d, roi, band, section_mask, qvol, st, lc);           // two (D <= 512, check_shape)
}

static void launch_final(const uint8_t *fused, const uint8_t *qvol, const Dims &d, const Roi &roi, uint16_t *sum, int16_t *raw_disp, uint16_t *out,
                         cudaStream_t st)
{
    const long long groups = ((long long)(roi.r1 - roi.r0) * (roi.c1 - roi.c0) + 3) / 4;
    if (groups <= 0) return;
    long long blocks = (groups + 7) / 8;
    if (blocks > 148LL * 64) blocks = 148LL * 64;
    if (d.interleaved) k_sgm_final<true><<<(unsigned)blocks, 256, 0, st>>>(fused, qvol, d, roi, sum, raw_disp, out);
    else k_sgm_final<false><<<(unsigned)blocks, 256, 0, st>>>(fused, qvol, d, roi, sum, raw_disp, out);
}

void launch_sgm(const uint8_t *fused, const Dims &d, bool full_frame, uint8_t *qvol, uint16_t *sum, int16_t *raw_disp, uint16_t *out,
                int *status, cudaStream_t st, LaunchCounter &lc)
{
    (void)status;
    const Roi roi = make_roi(d, full_frame);
    Band band;
    band.b0 = 0; band.b1 = d.Hp;
    band.in[0] = band.in[1] = nullptr; band.out[0] = band.out[1] = nullptr;
    launch_paths_any(fused, d, roi, band, debug_section_mask(), qvol, st, lc);
    lc.add();
    launch_final(fused, qvol, d, roi, sum, raw_disp, out, st);
    lc.add();
}

// ---- row bands (one band per GPU; crop-only aggregation). what: 0 = the band's row chains (r0 of both passes),
// 1 = the column / diagonal chains of pass 0 (state_in from the band above, state_out for the band below),
// 2 = those of pass 1 (state_in from the band below, state_out for the band above), 3 = final sum / WTA / encode of the
// band's rows of the crop.
void launch_sgm_band(int what, const uint8_t *fused, const Dims &d, int band_r0, int band_r1, const uint8_t *state_in, uint8_t *state_out,
                     uint8_t *qvol, int16_t *raw_disp, uint16_t *out, cudaStream_t st, LaunchCounter &lc)
{
    Roi roi = make_roi(d, false);
    Band band;
    band.b0 = band_r0; band.b1 = band_r1;
    band.in[0] = band.in[1] = nullptr; band.out[0] = band.out[1] = nullptr;
    if (what == 3) {
        roi.r0 = roi.r0 > band_r0 ? roi.r0 : band_r0;
        roi.r1 = roi.r1 < band_r1 ? roi.r1 : band_r1;
        launch_final(fused, qvol, d, roi, nullptr, raw_disp, out, st);
    } else {
        if (what == 1) { band.in[0] = state_in; band.out[0] = state_out; }
        if (what == 2) { band.in[1] = state_in; band.out[1] = state_out; }
        launch_paths_any(fused, d, roi, band, what == 0 ? 0x006u : what == 1 ? 0x038u : 0x1C0u, qvol, st, lc);
    }
    lc.add();
}

} // namespace sister
