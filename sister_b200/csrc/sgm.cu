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
// 40 B for the path-by-path accumulation into a uint16 sum volume this replaces. The fused-cost cells a chain will
// need are known in advance, so each lane loads its bytes kAhead steps early into registers and asks L2 for the cell
// kFar steps beyond that. On B200 the path kernel is bound by DRAM (about 5 TB/s of mixed reads and writes in
// scattered runs of 192-byte cells), not by issue slots: halving the instruction count, doubling the warps or
// deepening the prefetch each moved it by less than 5 % (profiles/r01_ncu_v16_paths.txt).
#include <cstdlib>

#include "kernels.cuh"
#include "sgm_core.cuh"

namespace sister {

constexpr int kChainWarps = 8;    // warps per block (they share nothing)

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
template <int NR, int LPC, bool FULL, int KIND>
__device__ __forceinline__ void run_chain(const uint8_t *__restrict__ fused_lane, uint8_t *__restrict__ q_lane, const LaneInfo<NR, LPC, FULL> &li,
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
        load_cost<NR, FULL>(fused_lane + (long long)ld.off8 * 8, valid_bytes, buf[t]); // nsteps >= 12 (check_shape, sister_test_sgm)
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
        if (valid_bytes > 0) prefetch_l2(fused_lane + (long long)pf.off8 * 8);
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
        load_cost<NR, FULL>(state_in, valid_bytes, w);
        unpack_cost<NR>(w, a);
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
        unpack_cost<NR>(buf[u], c);
        load_cost<NR, FULL>(fused_lane + (long long)ld.off8 * 8, valid_bytes, buf[u]);
        if (s + kAhead + 1 < nsteps) advance<DIAG>(ld, wk);
        if (s + kAhead + kFar < nsteps) {
            if (FULL || valid_bytes > 0) prefetch_l2(fused_lane + (long long)pf.off8 * 8);
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
        if (wanted) store_q<NR, FULL>(dst, q, valid_bytes);
    };
    int s0 = 0;
#pragma unroll 1
    for (; s0 + kAhead <= nsteps; s0 += kAhead) {
#pragma unroll
        for (int u = 0; u < kAhead; u++) step(u, s0 + u);
    }
#pragma unroll
    for (int u = 0; u < kAhead - 1; u++)
        if (s0 + u < nsteps) step(u, s0 + u); // warp-uniform
    if (KIND == 2 && state_out) store_q<NR, FULL>(state_out, cs.a, valid_bytes); // a <= P2 fits a byte
}

// grid ceil(chains / (kChainWarps * 32 / LPC)), block kChainWarps * 32, no shared memory
template <int NR, int LPC, bool FULL>
__global__ void __launch_bounds__(kChainWarps * 32, ((NR <= 8 || (FULL && NR <= 12)) ? 24 : 16) / kChainWarps) k_sgm_paths(const uint8_t *__restrict__ fused, Dims d, Roi roi, Band band, Sections sec, uint8_t *__restrict__ qvol, unsigned section_mask,
                                                                                                             long long first_block)
{
    constexpr int CPW = 32 / LPC;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    LaneInfo<NR, LPC, FULL> li;
    li.init(lane, d.D);
    Chain ch;
    int nsteps = 0, section = 0;
    long long state_off = 0;
    bool imports, exports;
    const int kind = chain_decode(d, roi, band, sec, ((first_block + blockIdx.x) * kChainWarps + warp) * CPW + lane / LPC, ch, nsteps, section,
                                  state_off, imports, exports);
    if (kind < 0) return; // warp-uniform: sections are padded to whole warps
    if (!((section_mask >> section) & 1u)) return; // band pipelines run the sections in separate launches
    const int D = d.D, Wp = d.Wp, D8 = D >> 3;
    const uint8_t *fused_lane = fused + li.sl * 2 * NR;
    uint8_t *q_lane = qvol + (size_t)ch.vol * (size_t)d.cells + li.sl * 2 * NR;
    const int valid_bytes = li.valid_bytes(D); // bytes of this lane inside the cell
    // per-lane constants derived from %tid: pin them in registers, otherwise ptxas re-derives them inside every step
    opaque(li.up_mask); opaque(li.dn_mask);
    opaque_ptr(fused_lane); opaque_ptr(q_lane);
    Walk wk;
    wk.stride8 = (ch.si * Wp + ch.sj) * D8;
    wk.wrapfix8 = -ch.sj * Wp * D8;
    wk.si = ch.si; wk.sj = ch.sj; wk.enter = ch.enter; wk.Wp = Wp;
    wk.r0 = roi.r0; wk.c0 = roi.c0; wk.nr = (unsigned)(roi.r1 - roi.r0); wk.nc = (unsigned)(roi.c1 - roi.c0);
    Cursor first;
    first.off8 = (ch.i * Wp + ch.j) * D8;
    first.i = ch.i;
    first.j = ch.j;
    if (kind == 1) run_chain<NR, LPC, FULL, 1>(fused_lane, q_lane, li, wk, first, nsteps, valid_bytes);
    else if (kind == 0) run_chain<NR, LPC, FULL, 0>(fused_lane, q_lane, li, wk, first, nsteps, valid_bytes);
    else {
        const int p = (section - 3) / 3;
        const uint8_t *sin = (imports && band.in[p]) ? band.in[p] + state_off + li.sl * 2 * NR : nullptr;
        uint8_t *sout = (exports && band.out[p]) ? band.out[p] + state_off + li.sl * 2 * NR : nullptr;
        run_chain<NR, LPC, FULL, 2>(fused_lane, q_lane, li, wk, first, nsteps, valid_bytes, sin, sout, !imports);
    }
}

// ---------------------------------------------------------------------------------------------- final sum + WTA + encode

__device__ __forceinline__ uint2 ldg8(const uint8_t *p) { return __ldg(reinterpret_cast<const uint2 *>(p)); }

// S = nC * C + sum_v Q_v; WTALeft_SSE with uniqueness 1 (hpp:283): first-index argmin over d <= min(j, D-1); then
// convertTo(CV_16UC1), crop Rect(D, D, W, H) and * 255 with saturation (hpp:111-118).
// Eight lanes per pixel, each lane owns 8-byte chunks sub, sub + 8, ... of the pixel's D bytes in all nine volumes.
// grid-stride over groups of 4 pixels per warp.
__global__ void __launch_bounds__(256) k_sgm_final(const uint8_t *__restrict__ fused, const uint8_t *__restrict__ qvol, Dims d, Roi roi,
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
                const int d0 = chunk * 8;
                const size_t off = off0 + d0;
                const uint2 cc = ldg8(fused + off);
                uint2 qq[8];
#pragma unroll
                for (int v = 0; v < 8; v++) qq[v] = ldg8(qvol + (size_t)v * cells + off);
                // byte-wise pair sums stay below 256 except on first lines (r0's byte may be up to 255 there, its
                // partner r1 writes 0), so a plain 32-bit add is a 4-lane byte add
                uint32_t w[2][4];
#pragma unroll
                for (int k = 0; k < 4; k++) { w[0][k] = qq[2 * k].x + qq[2 * k + 1].x; w[1][k] = qq[2 * k].y + qq[2 * k + 1].y; }
                uint32_t S[4]; // 8 cells as packed u16: S[0] = d0, d0+1; S[1] = d0+2, d0+3; ...
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t cw = h ? cc.y : cc.x;
                    uint32_t lo = __byte_perm(cw, 0u, 0x4140) << sh, hi = __byte_perm(cw, 0u, 0x4342) << sh;
#pragma unroll
                    for (int k = 0; k < 4; k++) { lo += __byte_perm(w[h][k], 0u, 0x4140); hi += __byte_perm(w[h][k], 0u, 0x4342); }
                    S[2 * h] = lo; S[2 * h + 1] = hi;
                }
                if (sum) *reinterpret_cast<uint4 *>(sum + off) = make_uint4(S[0], S[1], S[2], S[3]);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int da = d0 + 2 * k;
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

// measurement aid: SISTER_DEBUG_PATH_KINDS=<bit mask of chain kinds to run> (results are then incomplete)
static unsigned debug_section_mask()
{
    static int m = -1;
    if (m < 0) {
        const char *e = getenv("SISTER_DEBUG_PATH_KINDS");
        const int kinds = e ? atoi(e) & 7 : 7;
        m = ((kinds & 2) ? 0x001 : 0) | ((kinds & 1) ? 0x006 : 0) | ((kinds & 4) ? 0x1F8 : 0);
    }
    return (unsigned)m;
}

template <int NR, int LPC, bool FULL>
static void launch_paths(const uint8_t *fused, const Dims &d, const Roi &roi, const Band &band, unsigned section_mask, uint8_t *qvol, cudaStream_t st)
{
    constexpr int CPW = 32 / LPC;
    const Sections sec = chain_sections(d, roi, band, CPW);
    // launch only the span of chain indices the requested sections cover
    int k0 = 0, k1 = 9;
    while (k0 < 9 && (!((section_mask >> k0) & 1u) || sec.n[k0] == 0)) k0++;
    while (k1 > k0 && (!((section_mask >> (k1 - 1)) & 1u) || sec.n[k1 - 1] == 0)) k1--;
    if (k0 >= k1) return;
    const long long per_block = (long long)kChainWarps * CPW;
    const long long first_block = sec.o[k0] / per_block, n = sec.o[k1] - first_block * per_block;
    k_sgm_paths<NR, LPC, FULL><<<(unsigned)((n + per_block - 1) / per_block), kChainWarps * 32, 0, st>>>(fused, d, roi, band, sec, qvol, section_mask,
                                                                                                       first_block);
}

template <int LPC, int NRMAX>
static void launch_paths_lpc(const uint8_t *fused, const Dims &d, const Roi &roi, const Band &band, unsigned section_mask, uint8_t *qvol, cudaStream_t st)
{
    // disparities per lane = 2 * NR, NR even, chosen so that D fits in LPC lanes
    const int nr = 2 * ((d.D + 4 * LPC - 1) / (4 * LPC));
    const bool full = d.D == 2 * LPC * nr;
#define SISTER_PATHS_CASE(N)                                                                    \
    case N:                                                                                     \
        if constexpr (N <= NRMAX) {                                                             \
            if (full) launch_paths<N, LPC, true>(fused, d, roi, band, section_mask, qvol, st);                           \
            else launch_paths<N, LPC, false>(fused, d, roi, band, section_mask, qvol, st);                               \
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
                             cudaStream_t st)
{
    // Whole-frame the kernel is bound by DRAM, crop-only by the ALU pipe with few chains left: measured on B200 at D = 192
    // two chains per warp (twice the warps) beat four chains per warp by 4 % / 20 %.
    static int lpc8_max = -1; // measurement aid: SISTER_DEBUG_LPC8_MAXD=<largest D that runs four chains per warp>
    if (lpc8_max < 0) { const char *e = getenv("SISTER_DEBUG_LPC8_MAXD"); lpc8_max = e ? atoi(e) : 128; }
    if (d.D <= lpc8_max && d.D <= 192) launch_paths_lpc<8, 12>(fused, d, roi, band, section_mask, qvol, st); // four chains per warp
    else launch_paths_lpc<16, 16>(fused, d, roi, band, section_mask, qvol, st);                             // two (D <= 512, check_shape)
}

static void launch_final(const uint8_t *fused, const uint8_t *qvol, const Dims &d, const Roi &roi, uint16_t *sum, int16_t *raw_disp, uint16_t *out,
                         cudaStream_t st)
{
    const long long groups = ((long long)(roi.r1 - roi.r0) * (roi.c1 - roi.c0) + 3) / 4;
    if (groups <= 0) return;
    long long blocks = (groups + 7) / 8;
    if (blocks > 148LL * 64) blocks = 148LL * 64;
    k_sgm_final<<<(unsigned)blocks, 256, 0, st>>>(fused, qvol, d, roi, sum, raw_disp, out);
}

void launch_sgm(const uint8_t *fused, const Dims &d, bool full_frame, uint8_t *qvol, uint16_t *sum, int16_t *raw_disp, uint16_t *out,
                int *status, cudaStream_t st, LaunchCounter &lc)
{
    (void)status;
    const Roi roi = make_roi(d, full_frame);
    Band band;
    band.b0 = 0; band.b1 = d.Hp;
    band.in[0] = band.in[1] = nullptr; band.out[0] = band.out[1] = nullptr;
    launch_paths_any(fused, d, roi, band, debug_section_mask(), qvol, st);
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
        launch_paths_any(fused, d, roi, band, what == 0 ? 0x006u : what == 1 ? 0x038u : 0x1C0u, qvol, st);
    }
    lc.add();
}

} // namespace sister
