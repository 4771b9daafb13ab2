// sister_b200 / sgm.cu -- semi-global aggregation and final selection (sm_100a).
//
// Restates accumulateCostsSSE (sgm.cpp:26-455; P1 = 7, P2 = 100, 2 passes x 4 paths) on the uint8 fused volume C.
//
// Formulation (proven equal to the reference recurrence on the CPU by tests/sgm_spec.py):
//   * The state a path carries is normalised and clamped:  a(d) = min(L(d) - min_d L, P2).  With it the reference update
//         L'(d) = C(d) (+) ( min(L(d), L(d-1) (+) P1, L(d+1) (+) P1, P2 (+) m) (-) m )                 sgm.cpp:282-297
//     becomes  Q(d) = min(a(d), a(d-1) + P1, a(d+1) + P1),  L'(d) = C(d) + Q(d),  with no saturation anywhere because
//     C <= 252 (match.cu) bounds L' by 352 and the 8-path sum by 2816. The reference's sentinels map as follows:
//         L(-1) = L(D) = 65535 (sgm.cpp:84-87)                 -> kInf2 in the neighbour slots (never the minimum)
//         off-image predecessor column: L = 65535, m = 0       -> a = P2 everywhere  => Q = P2   (sgm.cpp:57-81)
//         r0 at the start of a row: L = 0, m = 0               -> a = 0 everywhere   => Q = 0    (sgm.cpp:215-216)
//         first line of a pass: L1 = L2 = L3 = C, m = min C    -> a step from a = 0 (Q = 0, a' = min(C - min C, P2)),
//                                                                 with the invalid cost 255 read as 0 (sgm.cpp:109,123,146)
//         first line, r0: int32 arithmetic + 8-bit truncation  -> first_line_step (sgm.cpp:141-190, types.h:28)
//   * PAIRED SWEEPS. The 8 paths are aggregated by four sweeps of two paths each; a sweep writes ONE byte per cell, the
//     sum of its two penalty terms Q (<= 2 * P2; on a first line the r0 byte is the whole truncated value and its partner
//     adds 0), into its own volume:
//         sweep 0 (pass 0)  carrier r0 along a row, left to right      + rider r1 (predecessor (i-1, j-1))
//         sweep 1 (pass 0)  carrier r2 down a column                    + rider r3 (predecessor (i-1, j+1))
//         sweep 2 (pass 1)  carrier r0 along a row, right to left      + rider r1 (predecessor (i+1, j+1))
//         sweep 3 (pass 1)  carrier r2 up a column                      + rider r3 (predecessor (i+1, j-1))
//     A sweep is a set of chains n (rows or columns) that all take step t (a column or a row) together. The carrier's
//     state stays in its chain's registers. The rider's state is handed from chain n-1 to chain n between steps: what chain
//     n-1 left after step t-1 IS the diagonal predecessor of chain n at step t. The hand-over is one-directional, so the
//     blocks of a sweep form a pipeline (block b consumes what block b-1 published), never a two-sided wavefront:
//         inside a block    through shared memory, double-buffered, one block barrier per step;
//         between blocks    through a full-length mailbox in global memory (one entry per step; every 32-bit word carries
//                           the launch's 4-bit epoch tag in the top bits of its bytes -- states are <= P2 < 128 -- so a word
//                           is valid or not by itself: no flags, no fences, no ring, no back-pressure, and a block that is
//                           not resident yet cannot stall the one that feeds it). A helper warp per block polls / publishes
//                           ahead of the compute warps.
//   * k_sgm_final adds up  S = nC * C + the 4 pair bytes  (nC = 8, or 4 on the first line of either pass) and does the final
//     WTA (hpp:283) and the output encoding (hpp:111-118) in the same sweep; S itself is only written when a test asks.
//
// Traffic per cell of the region: 4 x (1 B read C + 1 B write) + (1 + 4) B read = 13 B (round 1: 25 B with eight
// independent-chain path volumes).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "kernels.cuh"
#include "sgm_core.cuh"

namespace sister {

// compute warps per block (+ 1 mailbox warp): up to 18 while a lane holds few registers of state, fewer for the long
// disparity ranges, whose state needs the registers a smaller block leaves per thread
__host__ __device__ constexpr int sweep_warps_max(int NR) { return NR <= 8 ? 20 : NR <= 12 ? 17 : 8; }
constexpr int kSweepWarpsMin = 6;  // the mailbox allocation of a context is sized for this many (sgm_mailbox_bytes)
constexpr int kRing = 8;           // cost ring slots per warp
constexpr unsigned long long kWaitNs = 4000000000ull; // a hand-over that takes 4 s is a broken pipeline, not a slow one

// Region of interest: the cells whose aggregated cost is consumed. The caller only ever sees the crop
// Rect(D, D, W, H) of the disparity map (hpp:116-118), so outside of test / raw-disparity runs the pair bytes are
// needed there and nowhere else. SGM state still flows in from the borders of the padded frame, so a sweep starts
// where it always did, but rows / columns behind the region are not run, rows / columns in front of it run their rider
// (the diagonal path that will enter the region) only, and bytes are stored inside the region only.
struct Roi {
    int r0, r1, c0, c1;
};

// One sweep (mirrors tests/sgm_spec.py Sweep). Offsets are in units of 8 bytes (D % 8 == 0) and fit 32 bits (check_shape).
struct SweepGeo {
    int n0, n1;        // chains [n0, n1): row sweeps: rows counted from the pass's first line; column sweeps: columns counted
                       // from the border the diagonal enters at
    int t0, t1;        // steps [t0, t1): the other coordinate, counted from where the pass starts
    int car0, car1;    // chains that run their carrier (rows / columns of the region of interest)
    int ts0, ts1;      // steps whose cells lie in the region (bytes are stored for carrier chains at these steps)
    int base8, sn8, st8; // cell of (chain 0, step 0), increment per chain, increment per step
    int row;           // 1: row sweep (first line: chain 0; rider off the image: step 0); 0: column sweep (the reverse)
    int lead;          // dead slots in front of chain n0 (a row sweep that starts on the first line keeps that chain alone in its warp)
    int nw;            // compute warps per block of this sweep
    int nblk;          // blocks
    int vol;           // pair volume
    long long mb_off;  // first mailbox entry; block b publishes entries [mb_off + b * (t1 - t0), + (t1 - t0))
    // row bands: row sweeps read / leave one rider state per step, column sweeps [carrier states | rider states] per chain
    const uint8_t *band_in;
    uint8_t *band_out;
    // != 0: the row sweep's band entries are a mailbox between two GPUs that run at the same time -- every word carries this
    // tag (like the block mailbox), the reader polls for it, and both sides use system-scope accesses (peer memory)
    unsigned band_tag;
};
struct SweepPlan {
    SweepGeo g[4];
    int lvl_vid[5], lvl_pos[4], lvl_n[4], lvl_sweep[4][4]; // block id -> (sweep, block of the sweep): round-robin over the sweeps
    int nw_max;        // the largest of the sweeps' compute warps per block: the mailbox warp is warp nw_max
    int total_blocks;
    long long vol_stride; // bytes between two pair volumes
    unsigned tagword;  // the launch's epoch in bit 7 of each byte
    long long entries; // mailbox entries used
};

__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#ifdef SISTER_DEBUG_HOOKS
// measurement aid: per block {start ns, end ns, SM, sweep * 65536 + block of the sweep} of the last sweep launch
__device__ unsigned long long g_sweep_times[2048][4];
__device__ __forceinline__ unsigned sm_id() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
extern "C" int sister_debug_sweep_times(void *host, size_t bytes)
{
    return (int)cudaMemcpyFromSymbol(host, g_sweep_times, bytes < sizeof(g_sweep_times) ? bytes : sizeof(g_sweep_times));
}
#endif

// ---------------------------------------------------------------------------------------------- small device helpers

__device__ __forceinline__ void cp_async16(unsigned dst_s, const uint8_t *src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst_s), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async8(unsigned dst_s, const uint8_t *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst_s), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint32_t lds32(unsigned a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint2 lds64v(unsigned a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];\n" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint4 lds128v(unsigned a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts64(unsigned a, uint32_t x, uint32_t y) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};\n" ::"r"(a), "r"(x), "r"(y) : "memory"); }
// mailbox words: relaxed, never cached in L1. System scope because the same code hands the row sweeps' states to the GPU of the
// neighbouring row band (peer memory); on local memory it costs the same as .gpu (measured), and one form keeps the mailbox
// warp's step free of branches (a run-time choice between the two cost 3 % of the kernel)
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint8_t *p) { uint32_t v; asm volatile("ld.relaxed.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_relaxed_sys(uint8_t *p, uint32_t v) { asm volatile("st.relaxed.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory"); }
// The step barrier of a block (compute warps + mailbox warp; warps without a role have left). Warps of different roles reach
// it from different places of the code, which bar.sync permits (every warp executes it as a whole) but compute-sanitizer's
// synccheck reports as "divergent threads in block"; -DSISTER_ONE_BARRIER_INSTRUCTION routes every role through one
// out-of-line instance for that tool (profiles/r02_sanitize_synccheck.txt) -- at the cost of a call per step, so not by default.
#ifdef SISTER_ONE_BARRIER_INSTRUCTION
__device__ __noinline__ void block_sync(int n_threads) { asm volatile("bar.sync 1, %0;\n" ::"r"(n_threads) : "memory"); }
#else
__device__ __forceinline__ void block_sync(int n_threads) { asm volatile("bar.sync 1, %0;\n" ::"r"(n_threads) : "memory"); }
#endif

// the lane's NR / 2 words of its chain's cell in a ring slot (a = slot + the chain's cell + the lane's first byte)
template <int NR, int LPC, bool FULL, bool IL> __device__ __forceinline__ void ring_read(unsigned a, int valid_bytes, uint32_t (&w)[NR / 2])
{
    if constexpr (IL) {
#pragma unroll
        for (int k = 0; k < NR / 2; k++) w[k] = lds32(a + 4 * LPC * k);
    } else if constexpr (FULL && NR % 8 == 0) {
#pragma unroll
        for (int k = 0; k < NR / 8; k++) { const uint4 v = lds128v(a + 16 * k); w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w; }
    } else if constexpr (FULL && NR % 4 == 0) {
#pragma unroll
        for (int k = 0; k < NR / 4; k++) { const uint2 v = lds64v(a + 8 * k); w[2 * k] = v.x; w[2 * k + 1] = v.y; }
    } else {
#pragma unroll
        for (int k = 0; k < NR / 2; k++) w[k] = (FULL || 4 * k < valid_bytes) ? lds32(a + 4 * k) : 0u;
    }
}

// A chain state (NR packed registers per lane) as it travels: in shared memory as the registers themselves, 8-byte unit k
// of lane sl at (k * LPC + sl) * 8; in a mailbox / band entry (2 * NR * LPC bytes) as NR / 2 "pair words" per lane (bytes:
// a[2k].lo, a[2k+1].lo, a[2k].hi, a[2k+1].hi), word k of lane sl at (k * LPC + sl) * 4, bit 7 of every byte free for the tag.
template <int NR, int LPC> __device__ __forceinline__ void xch_read(unsigned entry_s, int sl, uint32_t (&a)[NR])
{
#pragma unroll
    for (int k = 0; k < NR / 2; k++) { const uint2 v = lds64v(entry_s + 8u * (unsigned)(k * LPC + sl)); a[2 * k] = v.x; a[2 * k + 1] = v.y; }
}
template <int NR, int LPC> __device__ __forceinline__ void xch_write(unsigned entry_s, int sl, const uint32_t (&a)[NR])
{
#pragma unroll
    for (int k = 0; k < NR / 2; k++) sts64(entry_s + 8u * (unsigned)(k * LPC + sl), a[2 * k], a[2 * k + 1]);
}
template <int NR> __device__ __forceinline__ void state_to_words(const uint32_t (&a)[NR], uint32_t tag, uint32_t (&w)[NR / 2])
{
#pragma unroll
    for (int k = 0; k < NR / 2; k++) w[k] = ((a[2 * k + 1] * 256u + a[2 * k]) & 0x7F7F7F7Fu) | tag; // pads (disparities >= D) are dropped
}
template <int NR, int LPC, bool FULL> __device__ __forceinline__ void words_to_state(const uint32_t (&w)[NR / 2], const LaneInfo<NR, LPC, FULL> &li, uint32_t (&a)[NR])
{
#pragma unroll
    for (int k = 0; k < NR / 2; k++) {
        const uint32_t x = w[k] & 0x7F7F7F7Fu;
        a[2 * k + 1] = li.padded(__byte_perm(x, 0u, 0x4341), 2 * k + 1);
        a[2 * k] = li.padded(x & 0x00FF00FFu, 2 * k);
    }
}
// plain (untagged) entry in global memory: a band state
template <int NR, int LPC, bool FULL> __device__ __forceinline__ void entry_load(const uint8_t *entry, const LaneInfo<NR, LPC, FULL> &li, uint32_t (&a)[NR])
{
    uint32_t w[NR / 2];
#pragma unroll
    for (int k = 0; k < NR / 2; k++) w[k] = __ldg(reinterpret_cast<const uint32_t *>(entry) + k * LPC + li.sl);
    words_to_state<NR, LPC, FULL>(w, li, a);
}
template <int NR, int LPC, bool FULL> __device__ __forceinline__ void entry_store(uint8_t *entry, const LaneInfo<NR, LPC, FULL> &li, const uint32_t (&a)[NR])
{
    uint32_t w[NR / 2];
    state_to_words<NR>(a, 0u, w);
#pragma unroll
    for (int k = 0; k < NR / 2; k++) reinterpret_cast<uint32_t *>(entry)[k * LPC + li.sl] = w[k];
}

// One step of the rider: the state arrives as `a` only (from the neighbouring chain), b and the end neighbours are formed
// here, and the new state leaves as `a` only.
template <int NR, int LPC, bool FULL>
__device__ __forceinline__ void rider_step(const uint32_t (&a)[NR], const uint32_t (&c)[NR], const LaneInfo<NR, LPC, FULL> &li, uint32_t (&q)[NR], uint32_t (&out)[NR])
{
    uint32_t b[NR], left, right, L[NR];
#pragma unroll
    for (int k = 0; k < NR; k++) b[k] = add_fma(a[k], kP1x2, li.one);
    end_neighbours<NR>(b, li.up_mask, li.dn_mask, left, right);
#pragma unroll
    for (int k = 0; k < NR; k++) {
        q[k] = __vimin3_s16x2(a[k], k == 0 ? left : b[k - 1], k == NR - 1 ? right : b[k + 1]);
        L[k] = li.padded(q[k] + c[k], k);
    }
    const uint32_t mm = chain_min2<LPC>(lane_min<NR>(L));
    const uint32_t neg2 = __byte_perm(0u - mm, 0u, 0x1010);
#pragma unroll
    for (int k = 0; k < NR; k++) out[k] = li.padded(__viaddmin_s16x2(L[k], neg2, kP2x2), k);
}

// ---------------------------------------------------------------------------------------------- the sweep kernel

// MODE 0: carrier + rider; 1: rider only (chains in front of the region); 2: the warp that holds the first line of a row
// sweep (carrier = the literal first-line arithmetic, rider restarts from the zero state at every step)
template <int NR, int LPC, bool FULL, bool IL, int MODE>
__device__ __forceinline__ void sweep_warp(const uint8_t *__restrict__ fused, uint8_t *__restrict__ vol, const LaneInfo<NR, LPC, FULL> &li, const SweepGeo &g, int D,
                                           int n, bool alive, bool car, int cl, unsigned ring_s, unsigned x_s, int x_stride, int lane, int n_sync)
{
    constexpr int CPW = 32 / LPC;
    constexpr int DS = 2 * NR * LPC;          // bytes of a cell in a ring slot (>= D), of a mailbox / band entry
    constexpr int CB = FULL ? 16 : 8;         // copy granule
    constexpr int NCP = (CPW * DS / CB + 31) / 32;
    constexpr int SS = CPW * DS;              // bytes per ring slot
    constexpr int R = kRing, A = R - 1, U = R / 2;
    constexpr int EX = NR * LPC * 4;          // bytes of an exchange entry
    const int T = g.t1 - g.t0;
    const int sub = lane / LPC;
    const int valid_bytes = li.valid_bytes(D);
    // ---- cost ring: chunk gch = lane + 32 m of a warp step belongs to the cell of sub-chain gch / chunks-per-cell
    const int cpc = FULL ? DS / CB : D / CB;
    const uint8_t *cp_src[NCP];
    unsigned cp_dst[NCP];
    bool cp_on[NCP];
    const int my_off8 = g.base8 + n * g.sn8 + g.t0 * g.st8; // the lane's chain at the first step
#pragma unroll
    for (int m = 0; m < NCP; m++) {
        const int gch = lane + 32 * m;
        cp_on[m] = gch < CPW * cpc;
        const int cs = cp_on[m] ? gch / cpc : 0;
        const int off8 = __shfl_sync(kFull, my_off8, cs * LPC);
        cp_src[m] = fused + (long long)off8 * 8 + (gch - cs * cpc) * CB;
        cp_dst[m] = ring_s + cs * DS + (gch - cs * cpc) * CB;
    }
    const long long step_bytes = (long long)g.st8 * 8;
    auto copy_step = [&](const bool go, const unsigned slot_off) { // the cells where the copy cursors stand -> the slot, move on
        if (go) {
#pragma unroll
            for (int m = 0; m < NCP; m++) {
                if (cp_on[m]) {
                    if constexpr (FULL) cp_async16(cp_dst[m] + slot_off, cp_src[m]);
                    else cp_async8(cp_dst[m] + slot_off, cp_src[m]);
                }
                cp_src[m] += step_bytes;
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < A; s++) copy_step(s < T, s * SS);
    const unsigned rd_lane = ring_s + sub * DS + li.template cell_offset<IL>();
    unsigned half = 0, other = U * SS;
    auto wr_off = [&](const int u) { return u == 0 ? other + (U - 1) * SS : half + (u - 1) * SS; };
    auto next_trip = [&]() { const unsigned t = half; half = other; other = t; };
    // ---- state
    ChainState<NR> cs;
    uint32_t mm = 0; // MODE 2: minimum of the truncated first-line state
    chain_set<NR, LPC, FULL>(cs, 0u, li);
    uint32_t rider_out[NR];
#pragma unroll
    for (int k = 0; k < NR; k++) rider_out[k] = 0u;
    const unsigned x_in = x_s + (unsigned)cl * EX, x_out = x_in + EX; // + parity * x_stride
    const bool frame_start = g.t0 == 0;
    const int ts_lo = g.ts0 - g.t0, ts_n = g.ts1 - g.ts0; // steps (relative to the first) whose cells are stored
    const long long band_chain = (long long)(n - g.n0) * DS, band_riders = (long long)(g.n1 - g.n0) * DS;
    if (!g.row && g.band_in) { // a column sweep continued from the band before: [carrier states | rider states] per chain
        uint32_t a[NR];
        entry_load<NR, LPC, FULL>(g.band_in + band_chain, li, a);
        chain_resume<NR, LPC, FULL>(cs, a, li);
        entry_load<NR, LPC, FULL>(g.band_in + band_riders + band_chain, li, rider_out);
        xch_write<NR, LPC>(x_out, li.sl, rider_out); // what this chain left after the step before the band: parity 0 = first step of the band
    }
    uint8_t *dst = vol + (long long)my_off8 * 8 + li.template cell_offset<IL>();
    block_sync(n_sync); // the exchange entries of the first step (helper: predecessor block / constants; above: band states) are in place
    // GEN: the general step (first trip, trips that straddle an edge of the stored range, the last trips); otherwise STORE
    // says whether the whole trip lies inside the stored range, and nothing is tested
    auto step = [&](const int s, const int u, auto gen_tag, auto store_tag) {
        constexpr bool GEN = decltype(gen_tag)::value, STORE = decltype(store_tag)::value;
        uint32_t c[NR], q1[NR], ra[NR];
        // ---- the rider's state: what the neighbouring chain left after the previous step
        if constexpr (MODE == 2) {
#pragma unroll
            for (int k = 0; k < NR; k++) ra[k] = li.padded(0u, k);
        } else {
            xch_read<NR, LPC>(x_in + ((u & 1) ? (unsigned)x_stride : 0u), li.sl, ra);
            if (GEN && frame_start && s == 0) { // row sweep: the predecessor column is off the image; column sweep: first line
                const uint32_t v = g.row ? kP2x2 : 0u;
#pragma unroll
                for (int k = 0; k < NR; k++) ra[k] = li.padded(v, k);
            }
        }
        {
            cp_async_wait<A - 1>();
            __syncwarp();
            uint32_t w[NR / 2];
            ring_read<NR, LPC, FULL, IL>(rd_lane + half + u * SS, valid_bytes, w);
            // a first-line cell reads the invalid cost 255 (census.cpp:76) as 0 (sgm.cpp:109,123,146); fused volumes never
            // hold 255 (match.cu), the raw two-view volume of sister_stereo does
            if (MODE == 2 || (GEN && !g.row && frame_start && s == 0)) {
#pragma unroll
                for (int k = 0; k < NR / 2; k++) w[k] &= ~__vcmpeq4(w[k], 0xFFFFFFFFu);
            }
            unpack_cost<NR, IL>(w, c);
        }
        copy_step(!GEN || s + A < T, wr_off(u));
        // ---- rider
        rider_step<NR, LPC, FULL>(ra, c, li, q1, rider_out);
        xch_write<NR, LPC>(x_out + ((u & 1) ? 0u : (unsigned)x_stride), li.sl, rider_out);
        // ---- carrier, sum, store
        if constexpr (MODE != 1) {
            uint32_t q0[NR];
            if constexpr (MODE == 2) first_line_step<NR, LPC, FULL>(cs.a, cs.b, mm, c, li, s == 0, q0);
            else chain_step<NR, LPC, FULL>(cs, c, li, q0);
            if (GEN ? (car && (unsigned)(s - ts_lo) < (unsigned)ts_n) : STORE) {
#pragma unroll
                for (int k = 0; k < NR; k++) q0[k] += q1[k];
                store_q<NR, LPC, FULL, IL>(dst, q0, valid_bytes);
            }
            dst += step_bytes;
        }
        block_sync(n_sync);
    };
    const bool all_car = __all_sync(kFull, car);
    int s0 = 0;
#pragma unroll 1
    while (s0 + U <= T) {
        const bool fast = s0 > 0 && s0 + U + A <= T;
        const bool inside = s0 >= ts_lo && s0 + U <= ts_lo + ts_n, outside = s0 + U <= ts_lo || s0 >= ts_lo + ts_n;
        if (fast && MODE == 0 && inside && all_car) {
#pragma unroll
            for (int u = 0; u < U; u++) step(s0 + u, u, std::false_type{}, std::true_type{});
        } else if (fast && (MODE == 1 || (MODE == 0 && outside))) {
#pragma unroll
            for (int u = 0; u < U; u++) step(s0 + u, u, std::false_type{}, std::false_type{});
        } else {
#pragma unroll
            for (int u = 0; u < U; u++) step(s0 + u, u, std::true_type{}, std::false_type{});
        }
        next_trip();
        s0 += U;
    }
#pragma unroll
    for (int u = 0; u < U - 1; u++)
        if (s0 + u < T) step(s0 + u, u, std::true_type{}, std::false_type{}); // warp-uniform
    cp_async_wait<0>();
    if (!g.row && g.band_out && alive) { // leave the column sweep's states for the next band
        entry_store<NR, LPC, FULL>(g.band_out + band_chain, li, cs.a);
        entry_store<NR, LPC, FULL>(g.band_out + band_riders + band_chain, li, rider_out);
    }
}

// The mailbox warp of a block: lanes [0, LPC) bring the predecessor block's rider state for the NEXT step into exchange
// entry 0, lanes [LPC, 2 LPC) publish what the block's last chain left after the PREVIOUS step; both meet the compute warps
// at the step's barrier. Entries are read two steps ahead so that the L2 round trip is off the step's critical path.
template <int NR, int LPC, bool FULL>
__device__ __forceinline__ void mailbox_warp(const LaneInfo<NR, LPC, FULL> &li, const SweepGeo &g, int b, int ch, int last_e, uint8_t *__restrict__ mailbox, unsigned tagword,
                                             unsigned x_s, int x_stride, int lane, int n_sync, int *__restrict__ status, const int role)
{
    constexpr int NH = NR / 2;
    constexpr int EX = NR * LPC * 4;
    constexpr long long EB = 2 * NR * LPC; // bytes of a mailbox / band entry
    const int T = g.t1 - g.t0;
    // role 0: the importing warp, role 1: the exporting warp (one warp doing both spends a step on each in turn, and the mailbox
    // warp's step is as long as a compute warp's)
    const bool imp = role == 0 && lane < LPC, exp = role == 1 && lane >= LPC && lane < 2 * LPC;
    // where the predecessor's states come from, where the last chain's states go
    const uint8_t *src = nullptr;
    unsigned src_tag = 0u, src_mask = 0u;
    if (b > 0) { src = mailbox + (g.mb_off + (long long)(b - 1) * T) * EB; src_tag = tagword; src_mask = 0x80808080u; }
    else if (g.row && g.band_in) { // the band before leaves one state per step: all of them already there (untagged), or
        src = g.band_in;           // arriving while this band runs (tagged, written by the neighbouring GPU)
        if (g.band_tag) { src_tag = g.band_tag; src_mask = 0x80808080u; }
    }
    uint8_t *dst = nullptr;
    unsigned dst_tag = 0u;
    if (b < g.nblk - 1) { dst = mailbox + (g.mb_off + (long long)b * T) * EB; dst_tag = tagword; }
    else if (g.row && g.band_out) { dst = g.band_out; dst_tag = g.band_tag; }
    if (role == 0) dst = nullptr;
    else src = nullptr;
    const long long lane_off = (long long)li.sl * 4;
    uint32_t a[NR];
    // entry 0 at the first step. A column sweep continued from the band before takes the predecessor of the block's first
    // chain from the band state. With no predecessor at all (column sweep: the border column, rider = P2 for good; row sweep:
    // the first line, whose warp ignores the entry) the constant goes into both parities once.
    if (imp) {
        if (!g.row && g.band_in && b > 0) {
            const long long pred = (long long)b * ch - 1; // chain before the block's first (column sweeps have no lead)
            entry_load<NR, LPC, FULL>(g.band_in + ((long long)(g.n1 - g.n0) + pred) * EB, li, a);
            xch_write<NR, LPC>(x_s, li.sl, a);
        } else if (!src) {
#pragma unroll
            for (int k = 0; k < NR; k++) a[k] = li.padded(kP2x2, k);
            xch_write<NR, LPC>(x_s, li.sl, a);
            xch_write<NR, LPC>(x_s + x_stride, li.sl, a);
        }
    }
    uint32_t pf0[NH], pf1[NH];
    auto fetch = [&](const int s, uint32_t (&w)[NH]) { // entry s of the source: the predecessor's state after step s
        if (imp && src && s < T - 1) {
            const uint8_t *e = src + (long long)s * EB + lane_off;
#pragma unroll
            for (int k = 0; k < NH; k++) w[k] = ld_relaxed_sys(e + (long long)k * LPC * 4);
        } else {
#pragma unroll
            for (int k = 0; k < NH; k++) w[k] = src_tag;
        }
    };
    fetch(0, pf0);
    fetch(1, pf1);
    block_sync(n_sync);
    const unsigned x_last = x_s + (unsigned)last_e * EX;
    unsigned long long t_wait = 0;
#pragma unroll 1
    for (int s = 0; s < T; s++) {
        int spins = 0;
        // ---- publish the state the last chain left after step s - 1
        if (exp && dst && s > 0) {
            xch_read<NR, LPC>(x_last + ((s & 1) ? x_stride : 0), li.sl, a);
            uint32_t w[NH];
            state_to_words<NR>(a, dst_tag, w);
            uint8_t *e = dst + (long long)(s - 1) * EB + lane_off;
#pragma unroll
            for (int k = 0; k < NH; k++) {
                st_relaxed_sys(e + (long long)k * LPC * 4, w[k]);
            }
        }
        // ---- deliver the predecessor's state after step s for step s + 1 (read two entries ahead)
        uint32_t w[NH];
#pragma unroll
        for (int k = 0; k < NH; k++) { w[k] = pf0[k]; pf0[k] = pf1[k]; }
        fetch(s + 2, pf1);
        if (src && s < T - 1) { // warp-uniform
            for (;;) {
                uint32_t bad = 0u;
#pragma unroll
                for (int k = 0; k < NH; k++) bad |= (w[k] ^ src_tag) & src_mask;
                if (!__any_sync(kFull, imp && bad != 0u)) break;
                // bounded by time, not polls (a predecessor may be queued behind other kernels, or a tool may slow everything
                // down): a broken pipeline must never hang the device -- report it and carry on with what is there
                if (spins++ == 0) t_wait = global_ns();
                else if ((spins & 255) == 0 && global_ns() - t_wait > kWaitNs) {
                    if (lane == 0) atomicOr(status, kStatusSpinTimeout);
                    src_mask = 0u;
                    break;
                }
                __nanosleep(32);
                if (imp) {
                    const uint8_t *e = src + (long long)s * EB + lane_off;
#pragma unroll
                    for (int k = 0; k < NH; k++) w[k] = ld_relaxed_sys(e + (long long)k * LPC * 4);
                }
            }
            if (imp) {
                words_to_state<NR, LPC, FULL>(w, li, a);
                xch_write<NR, LPC>(x_s + ((s & 1) ? 0 : x_stride), li.sl, a);
            }
        }
        block_sync(n_sync);
    }
    if (exp && dst) { // the state after the last step
        xch_read<NR, LPC>(x_last + ((T & 1) ? x_stride : 0), li.sl, a);
        uint32_t w[NH];
        state_to_words<NR>(a, dst_tag, w);
        uint8_t *e = dst + (long long)(T - 1) * EB + lane_off;
#pragma unroll
        for (int k = 0; k < NH; k++) {
            st_relaxed_sys(e + (long long)k * LPC * 4, w[k]);
        }
    }
}

template <int NR, int LPC, bool FULL, bool IL>
__global__ void __launch_bounds__(NR <= 6 ? 1024 : NR <= 8 ? 896 : (sweep_warps_max(NR) + 2) * 32) // (<= 64 registers: a block of another rig's match or fuse kernel fits beside a sweep block)
    k_sgm_sweeps(const uint8_t *__restrict__ fused, Dims d, SweepPlan pl, uint8_t *__restrict__ vols, uint8_t *__restrict__ mailbox, int *__restrict__ status, unsigned one)
{
    constexpr int CPW = 32 / LPC;
    constexpr int EX = NR * LPC * 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // ---- which sweep, which block of it: blocks are dealt round-robin over the sweeps that still have blocks at that
    // position, so that every sweep's blocks are dispatched in pipeline order
    int lvl = 0;
#pragma unroll
    for (int k = 1; k < 4; k++) lvl += (int)blockIdx.x >= pl.lvl_vid[k];
    const int rel = (int)blockIdx.x - pl.lvl_vid[lvl];
    const int s_id = pl.lvl_sweep[lvl][rel % pl.lvl_n[lvl]];
    const int b = pl.lvl_pos[lvl] + rel / pl.lvl_n[lvl];
    const SweepGeo &g = pl.g[s_id];
    const int nw = g.nw, CH = nw * CPW;
    const int slots_left = g.n1 - g.n0 + g.lead - b * CH; // slots from this block's first to the sweep's last chain
    const int last_e = slots_left < CH ? slots_left : CH;
    const int n_live = (last_e + CPW - 1) / CPW;
    const int n_sync = (n_live + 2) * 32;
    LaneInfo<NR, LPC, FULL> li;
    li.init(lane, d.D);
    opaque(li.up_mask); opaque(li.dn_mask);
    li.one = one; // a kernel argument: the only 1 neither nvvm nor ptxas can fold (add_fma)
    const unsigned x_s = (unsigned)__cvta_generic_to_shared(smem_raw);
    const int x_stride = (CH + 1) * EX;
    if (warp >= pl.nw_max) {
        mailbox_warp<NR, LPC, FULL>(li, g, b, CH, last_e, mailbox, pl.tagword, x_s, x_stride, lane, n_sync, status, warp - pl.nw_max);
        return;
    }
    if (warp >= n_live) return; // (also the warps between this sweep's nw and nw_max)
    const int sub = lane / LPC;
    const int cl = warp * CPW + sub;        // slot within the block
    const int u = b * CH + cl;              // slot within the sweep
    int n = g.n0 + u - g.lead;
    const bool alive = u >= g.lead && n < g.n1;
    n = n < g.n0 ? g.n0 : n >= g.n1 ? g.n1 - 1 : n; // dead slots shadow a real chain (their stores are off)
    const bool car = alive && n >= g.car0 && n < g.car1;
    const unsigned ring_s = x_s + 2u * (unsigned)x_stride + (unsigned)warp * (kRing * CPW * 2 * NR * LPC);
    uint8_t *vol = vols + (size_t)g.vol * (size_t)pl.vol_stride;
    const bool first_line_warp = g.row && g.n0 == 0 && b == 0 && warp == 0; // holds chain 0 in its last sub-chain, the others are dead
    const bool any_car = __any_sync(kFull, car);
#ifdef SISTER_DEBUG_HOOKS
    if (threadIdx.x == 0 && blockIdx.x < 2048) { g_sweep_times[blockIdx.x][0] = global_ns(); g_sweep_times[blockIdx.x][2] = sm_id(); g_sweep_times[blockIdx.x][3] = (unsigned long long)s_id * 65536ull + (unsigned)b; }
#endif
    if (first_line_warp) sweep_warp<NR, LPC, FULL, IL, 2>(fused, vol, li, g, d.D, n, alive, car, cl, ring_s, x_s, x_stride, lane, n_sync);
    else if (!any_car) sweep_warp<NR, LPC, FULL, IL, 1>(fused, vol, li, g, d.D, n, alive, car, cl, ring_s, x_s, x_stride, lane, n_sync);
    else sweep_warp<NR, LPC, FULL, IL, 0>(fused, vol, li, g, d.D, n, alive, car, cl, ring_s, x_s, x_stride, lane, n_sync);
#ifdef SISTER_DEBUG_HOOKS
    if (threadIdx.x == 0 && blockIdx.x < 2048) g_sweep_times[blockIdx.x][1] = global_ns();
#endif
}

// ---------------------------------------------------------------------------------------------- final sum + WTA + encode

__device__ __forceinline__ uint2 ldg8(const uint8_t *p) { return __ldg(reinterpret_cast<const uint2 *>(p)); }

// S = nC * C + the four pair bytes; WTALeft_SSE with uniqueness 1 (hpp:283): first-index argmin over d <= min(j, D-1); then
// convertTo(CV_16UC1), crop Rect(D, D, W, H) and * 255 with saturation (hpp:111-118).
// Eight lanes per pixel, each lane owns 8-byte chunks sub, sub + 8, ... of the pixel's D bytes in all five volumes.
// grid-stride over groups of 4 pixels per warp.
template <bool IL>
__global__ void __launch_bounds__(256, 6) k_sgm_final(const uint8_t *__restrict__ fused, const uint8_t *__restrict__ vols, size_t vol_stride, Dims d, Roi roi,
                                                   uint16_t *__restrict__ sum, int16_t *__restrict__ raw_disp, uint16_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31, sub = lane & 7, grp = lane >> 3;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const int D = d.D, nchunk = D >> 3;
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
                uint2 qq[4];
#pragma unroll
                for (int v = 0; v < 4; v++) qq[v] = ldg8(vols + (size_t)v * vol_stride + off);
                uint32_t S[4]; // 8 cells as packed u16: S[k] = bytes 2k, 2k + 1 of the chunk
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t cw = h ? cc.y : cc.x;
                    uint32_t lo = __byte_perm(cw, 0u, 0x4140) << sh, hi = __byte_perm(cw, 0u, 0x4342) << sh;
#pragma unroll
                    for (int v = 0; v < 4; v++) {
                        const uint32_t w = h ? qq[v].y : qq[v].x; // a pair byte may reach 255 on a first line: widen, then add
                        lo += __byte_perm(w, 0u, 0x4140);
                        hi += __byte_perm(w, 0u, 0x4342);
                    }
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
// even, chosen so that D fits in lpc lanes.
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

static Roi make_roi(const Dims &d, bool full_frame)
{
    Roi roi;
    if (full_frame) { roi.r0 = 0; roi.r1 = d.Hp; roi.c0 = 0; roi.c1 = d.Wp; }
    else { roi.r0 = d.D; roi.r1 = d.D + d.H; roi.c0 = d.D; roi.c1 = d.D + d.W; } // Rect(D, D, W, H), hpp:116-118
    return roi;
}

// bytes of a mailbox / band-state entry: one chain state, a byte per disparity slot of the chain's lanes
static inline long long entry_bytes(const Dims &d) { return 2LL * d.nr * d.lpc; }

// The four sweeps of a frame restricted to the region `r` and the row band [b0, b1); sweeps not in `mask` get no blocks.
static void plan_sweeps(const Dims &d, const Roi &r, int b0, int b1, unsigned mask, const int (&nw)[4], SweepPlan &pl)
{
    const int D8 = d.D >> 3, Wp = d.Wp, Hp = d.Hp, cpw = 32 / d.lpc;
    long long entries = 0;
    for (int s = 0; s < 4; s++) {
        SweepGeo &g = pl.g[s];
        g = SweepGeo();
        g.vol = s;
        g.row = (s == 0 || s == 2);
        if (s == 0) {        // rows 0 .. r1-1 (the band's), columns 0 .. c1-1
            const int lo = b0 > 0 ? b0 : 0, hi = b1 < r.r1 ? b1 : r.r1;
            g.n0 = lo; g.n1 = hi; g.t0 = 0; g.t1 = r.c1;
            g.car0 = r.r0; g.car1 = r.r1; g.ts0 = r.c0; g.ts1 = r.c1;
            g.base8 = 0; g.sn8 = Wp * D8; g.st8 = D8;
        } else if (s == 2) { // rows Hp-1 .. r0 (the band's), columns Wp-1 .. c0
            const int lo = b0 > r.r0 ? b0 : r.r0, hi = b1 < Hp ? b1 : Hp;
            g.n0 = Hp - hi; g.n1 = Hp - lo; g.t0 = 0; g.t1 = Wp - r.c0;
            g.car0 = Hp - r.r1; g.car1 = Hp - r.r0; g.ts0 = Wp - r.c1; g.ts1 = Wp - r.c0;
            g.base8 = ((Hp - 1) * Wp + Wp - 1) * D8; g.sn8 = -Wp * D8; g.st8 = -D8;
        } else if (s == 1) { // columns Wp-1 .. c0, rows 0 .. r1-1 (the band's)
            const int lo = b0 > 0 ? b0 : 0, hi = b1 < r.r1 ? b1 : r.r1;
            g.n0 = 0; g.n1 = Wp - r.c0; g.t0 = lo; g.t1 = hi;
            g.car0 = Wp - r.c1; g.car1 = Wp - r.c0; g.ts0 = r.r0; g.ts1 = r.r1;
            g.base8 = (Wp - 1) * D8; g.sn8 = -D8; g.st8 = Wp * D8;
        } else {             // columns 0 .. c1-1, rows Hp-1 .. r0 (the band's)
            const int lo = b0 > r.r0 ? b0 : r.r0, hi = b1 < Hp ? b1 : Hp;
            g.n0 = 0; g.n1 = r.c1; g.t0 = Hp - hi; g.t1 = Hp - lo;
            g.car0 = r.c0; g.car1 = r.c1; g.ts0 = Hp - r.r1; g.ts1 = Hp - r.r0;
            g.base8 = (Hp - 1) * Wp * D8; g.sn8 = D8; g.st8 = -Wp * D8;
        }
        const bool empty = !((mask >> s) & 1u) || g.n1 <= g.n0 || g.t1 <= g.t0;
        g.lead = (g.row && g.n0 == 0) ? cpw - 1 : 0;
        g.nw = nw[s];
        g.nblk = empty ? 0 : (g.n1 - g.n0 + g.lead + nw[s] * cpw - 1) / (nw[s] * cpw);
        g.mb_off = entries;
        entries += (long long)g.nblk * (g.t1 - g.t0);
    }
    pl.entries = entries;
    pl.nw_max = std::max(std::max(nw[0], nw[1]), std::max(nw[2], nw[3]));
    // deal the blocks round-robin: level L covers the positions at which the same set of sweeps still has blocks
    int order[4] = {0, 1, 2, 3};
    std::sort(order, order + 4, [&](int a, int b) { return pl.g[a].nblk < pl.g[b].nblk; });
    int vid = 0, pos = 0;
    for (int L = 0; L < 4; L++) {
        pl.lvl_vid[L] = vid;
        pl.lvl_pos[L] = pos;
        int na = 0;
        for (int k = L; k < 4; k++) pl.lvl_sweep[L][na++] = order[k];
        std::sort(pl.lvl_sweep[L], pl.lvl_sweep[L] + na);
        for (int k = na; k < 4; k++) pl.lvl_sweep[L][k] = pl.lvl_sweep[L][0];
        pl.lvl_n[L] = na;
        const int upto = pl.g[order[L]].nblk;
        vid += (upto - pos) * na;
        pos = upto;
    }
    pl.lvl_vid[4] = vid;
    pl.total_blocks = vid;
}

size_t sgm_mailbox_bytes(int max_w, int max_h, int max_d, int band_rows)
{
    const long long Wp = max_w + 2LL * max_d, Hp = band_rows > 0 ? band_rows : max_h + 2LL * max_d; // a band runs its rows only
    long long eb = 0;
    for (int D = 8; D <= max_d; D += 8) {
        Dims d;
        d.D = D;
        set_cell_order(d);
        eb = std::max(eb, entry_bytes(d));
    }
    const long long ch = kSweepWarpsMin * 2; // at least two chains per warp
    const long long entries = 2 * (((Hp + 1 + ch - 1) / ch) * Wp + ((Wp + ch - 1) / ch) * Hp);
    return (size_t)(entries * eb);
}

size_t sgm_band_state_bytes(const Dims &d) { return (size_t)(3LL * d.Wp * entry_bytes(d)); }

template <int NR, int LPC, bool FULL, bool IL>
static void launch_sweeps_t(const uint8_t *fused, const Dims &d, const Roi &roi, int b0, int b1, unsigned mask, const uint8_t *band_in, uint8_t *band_out, const BandStream *bs,
                            SgmScratch &sc, int *status, cudaStream_t st, LaunchCounter &lc)
{
    constexpr int CPW = 32 / LPC;
    const void *kernel = (const void *)k_sgm_sweeps<NR, LPC, FULL, IL>;
    auto smem_for = [&](int nw) { return (size_t)2 * (nw * CPW + 1) * NR * LPC * 4 + (size_t)nw * kRing * CPW * 2 * NR * LPC; };
    constexpr int kWarpsMax = sweep_warps_max(NR);
    if (smem_for(kWarpsMax) > 48 * 1024) lc.fail(optin_dynamic_smem(kernel, smem_for(kWarpsMax)));
    // Compute warps per block, per sweep. The blocks of a sweep are a pipeline that runs at the pace of its slowest block, and
    // a block's pace is its number of warps (they share one SM's issue slots); a sweep of T steps with c warps per block
    // therefore takes about T * c. The choice: the smallest time budget B such that c_s = B / T_s warps per block (within
    // the limits) need no more blocks than are resident at once -- long sweeps get small blocks, short sweeps large ones,
    // and all of them finish together. (A block that has to wait for a slot starts its sweep late, which costs up to a whole
    // sweep of time, not just its own share.)
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    SweepPlan pl;
    int nw[4] = {kWarpsMax, kWarpsMax, kWarpsMax, kWarpsMax};
    plan_sweeps(d, roi, b0, b1, mask, nw, pl); // chains and steps of every sweep (they do not depend on the block size)
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, (kWarpsMax + 2) * 32, smem_for(kWarpsMax)) != cudaSuccess || per_sm < 1) per_sm = 1;
    const long long capacity = (long long)per_sm * n_sm;
    long long best_cost = -1;
    for (int s0 = 0; s0 < 4; s0++) {
        if (pl.g[s0].nblk == 0) continue;
        for (int c0 = kSweepWarpsMin; c0 <= kWarpsMax; c0++) {
            const long long budget = (long long)(pl.g[s0].t1 - pl.g[s0].t0) * c0;
            int cand[4];
            long long blocks = 0, cost = 0;
            for (int k = 0; k < 4; k++) {
                const SweepGeo &g = pl.g[k];
                cand[k] = kWarpsMax;
                if (g.nblk == 0) continue;
                const long long T = g.t1 - g.t0;
                cand[k] = (int)std::min<long long>(kWarpsMax, std::max<long long>(kSweepWarpsMin, budget / T));
                blocks += (g.n1 - g.n0 + g.lead + cand[k] * CPW - 1) / (cand[k] * CPW);
                cost = std::max(cost, T * cand[k]);
            }
            if (blocks > capacity) continue;
            cost *= (blocks + n_sm - 1) / n_sm; // blocks that share an SM share its issue slots
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; for (int k = 0; k < 4; k++) nw[k] = cand[k]; }
        }
    }
#ifdef SISTER_DEBUG_HOOKS
    if (const char *e = getenv("SISTER_DEBUG_SWEEP_NW")) { // measurement aid: "row,col" compute warps per block
        int a = 0, b = 0;
        if (sscanf(e, "%d,%d", &a, &b) == 2) { nw[0] = nw[2] = std::min(std::max(a, kSweepWarpsMin), kWarpsMax); nw[1] = nw[3] = std::min(std::max(b, kSweepWarpsMin), kWarpsMax); }
    }
#endif
    plan_sweeps(d, roi, b0, b1, mask, nw, pl);
#ifdef SISTER_DEBUG_HOOKS
    if (getenv("SISTER_DEBUG_SWEEP_STRIDE0")) for (int k = 0; k < 4; k++) pl.g[k].st8 = 0; // measurement aid: every step on the chain's first cell (no DRAM)
#endif
    if (pl.total_blocks == 0) return;
    const long long eb = entry_bytes(d);
    if ((size_t)(pl.entries * eb) > sc.mailbox_bytes) { lc.fail(cudaErrorMemoryAllocation); return; }
    // band states (sgm_band_state_bytes): Wp entries of the row sweep's riders, then 2 Wp entries of the column sweep
    for (int s = 0; s < 4; s++) {
        SweepGeo &g = pl.g[s];
        const long long off = g.row ? 0 : (long long)d.Wp * eb;
        const bool first = g.row ? g.n0 == 0 : g.t0 == 0; // starts at the pass's first line: nothing to continue
        g.band_in = (band_in && !first) ? band_in + off : nullptr;
        g.band_out = band_out ? band_out + off : nullptr;
        g.band_tag = 0u;
        if (bs && g.row) { // row sweeps streamed between the GPUs of two neighbouring bands: one mailbox per sweep and direction
            g.band_in = first ? nullptr : bs->in[s];
            g.band_out = bs->out[s];
            const unsigned e = bs->tag & 15u;
            g.band_tag = ((e & 1u) << 7) | (((e >> 1) & 1u) << 15) | (((e >> 2) & 1u) << 23) | (((e >> 3) & 1u) << 31);
        }
    }
    // epoch tags: every entry a launch reads is written by that launch, so the tag only has to differ from the previous
    // launch of the same geometry; anything else clears the mailbox first
    const unsigned long long key = ((unsigned long long)d.W << 48) ^ ((unsigned long long)d.H << 32) ^ ((unsigned long long)d.D << 20) ^
                                   ((unsigned long long)roi.r0 << 10) ^ ((unsigned long long)(unsigned)b0 * 0x9E3779B97F4A7C15ull) ^
                                   ((unsigned long long)(unsigned)b1 * 0xC2B2AE3D27D4EB4Full) ^ ((unsigned long long)mask << 4) ^ (unsigned long long)(nw[0] * 32 + nw[1]);
    if (key != sc.geo_key || sc.epoch == 0) {
        lc.fail(cudaMemsetAsync(sc.mailbox, 0, (size_t)(pl.entries * eb), st));
        sc.geo_key = key;
        sc.epoch = 0;
    }
    sc.epoch = sc.epoch % 15 + 1;
    const unsigned e = sc.epoch;
    pl.tagword = ((e & 1u) << 7) | (((e >> 1) & 1u) << 15) | (((e >> 2) & 1u) << 23) | (((e >> 3) & 1u) << 31);
    pl.vol_stride = sc.vol_stride ? (long long)sc.vol_stride : d.cells;
    k_sgm_sweeps<NR, LPC, FULL, IL><<<(unsigned)pl.total_blocks, (pl.nw_max + 2) * 32, smem_for(pl.nw_max), st>>>(fused, d, pl, sc.vols - sc.row_shift, sc.mailbox, status, 1u);
    lc.add();
}

template <int LPC, int NRMAX>
static void launch_sweeps_lpc(const uint8_t *fused, const Dims &d, const Roi &roi, int b0, int b1, unsigned mask, const uint8_t *band_in, uint8_t *band_out, const BandStream *bs,
                              SgmScratch &sc, int *status, cudaStream_t st, LaunchCounter &lc)
{
    const int nr = d.nr;
    const bool full = d.D == 2 * LPC * nr;
#define SISTER_SWEEPS_CASE(N)                                                                                                  \
    case N:                                                                                                                    \
        if constexpr (N <= NRMAX) {                                                                                            \
            if constexpr (N % 4 == 2) {                                                                                        \
                if (d.interleaved) { launch_sweeps_t<N, LPC, true, true>(fused, d, roi, b0, b1, mask, band_in, band_out, bs, sc, status, st, lc); break; } \
            }                                                                                                                  \
            if (full) launch_sweeps_t<N, LPC, true, false>(fused, d, roi, b0, b1, mask, band_in, band_out, bs, sc, status, st, lc);  \
            else launch_sweeps_t<N, LPC, false, false>(fused, d, roi, b0, b1, mask, band_in, band_out, bs, sc, status, st, lc);      \
        }                                                                                                                      \
        break;
    switch (nr) {
        SISTER_SWEEPS_CASE(2) SISTER_SWEEPS_CASE(4) SISTER_SWEEPS_CASE(6) SISTER_SWEEPS_CASE(8)
        SISTER_SWEEPS_CASE(10) SISTER_SWEEPS_CASE(12) SISTER_SWEEPS_CASE(14) SISTER_SWEEPS_CASE(16)
    }
#undef SISTER_SWEEPS_CASE
}

static void launch_sweeps(const uint8_t *fused, const Dims &d, const Roi &roi, int b0, int b1, unsigned mask, const uint8_t *band_in, uint8_t *band_out, const BandStream *bs,
                          SgmScratch &sc, int *status, cudaStream_t st, LaunchCounter &lc)
{
#ifdef SISTER_DEBUG_HOOKS
    if (const char *e = getenv("SISTER_DEBUG_SWEEP_MASK")) mask &= (unsigned)atoi(e); // measurement aid: run some of the sweeps only
#endif
    if (d.lpc == 8) launch_sweeps_lpc<8, 12>(fused, d, roi, b0, b1, mask, band_in, band_out, bs, sc, status, st, lc); // four chains per warp
    else launch_sweeps_lpc<16, 16>(fused, d, roi, b0, b1, mask, band_in, band_out, bs, sc, status, st, lc);           // two (D <= 512, check_shape)
}

static void launch_final(const uint8_t *fused, const uint8_t *vols, size_t vol_stride, const Dims &d, const Roi &roi, uint16_t *sum, int16_t *raw_disp,
                         uint16_t *out, cudaStream_t st)
{
    const long long groups = ((long long)(roi.r1 - roi.r0) * (roi.c1 - roi.c0) + 3) / 4;
    if (groups <= 0) return;
    long long blocks = (groups + 7) / 8;
    if (blocks > 148LL * 64) blocks = 148LL * 64;
    if (d.interleaved) k_sgm_final<true><<<(unsigned)blocks, 256, 0, st>>>(fused, vols, vol_stride, d, roi, sum, raw_disp, out);
    else k_sgm_final<false><<<(unsigned)blocks, 256, 0, st>>>(fused, vols, vol_stride, d, roi, sum, raw_disp, out);
}

void launch_sgm(const uint8_t *fused, const Dims &d, bool full_frame, SgmScratch &sc, uint16_t *sum, int16_t *raw_disp, uint16_t *out,
                int *status, cudaStream_t st, LaunchCounter &lc)
{
    const Roi roi = make_roi(d, full_frame);
    launch_sweeps(fused, d, roi, 0, d.Hp, 0xFu, nullptr, nullptr, nullptr, sc, status, st, lc);
    launch_final(fused, sc.vols - sc.row_shift, sc.vol_stride ? sc.vol_stride : (size_t)d.cells, d, roi, sum, raw_disp, out, st);
    lc.add();
}

// ---- row bands (one band per GPU; crop-only aggregation). what: 1 = the two sweeps of pass 0 inside the band (state_in from
// the band above, state_out for the band below), 2 = those of pass 1 (state_in from the band below, state_out for the band
// above), 3 = final sum / WTA / encode of the band's rows of the crop; 5 / 6 = the column sweep of pass 0 / 1 alone (same
// state layout; the row sweep's part of it is not touched).
void launch_sgm_band(int what, const uint8_t *fused, const Dims &d, int band_r0, int band_r1, const uint8_t *state_in, uint8_t *state_out,
                     SgmScratch &sc, int16_t *raw_disp, uint16_t *out, int *status, cudaStream_t st, LaunchCounter &lc)
{
    Roi roi = make_roi(d, false);
    if (what == 3) {
        roi.r0 = roi.r0 > band_r0 ? roi.r0 : band_r0;
        roi.r1 = roi.r1 < band_r1 ? roi.r1 : band_r1;
        launch_final(fused, sc.vols - sc.row_shift, sc.vol_stride ? sc.vol_stride : (size_t)d.cells, d, roi, nullptr, raw_disp, out, st);
        lc.add();
    } else {
        const unsigned mask = what == 1 ? 0x3u : what == 2 ? 0xCu : what == 5 ? 0x2u : 0x8u;
        launch_sweeps(fused, d, roi, band_r0, band_r1, mask, state_in, state_out, nullptr, sc, status, st, lc);
    }
}

// The row sweeps of the band (mask: 1 = pass 0, 4 = pass 1, 5 = both in one launch) with their rider states STREAMED between
// the bands: bs.in[s] is this band's mailbox for sweep s (one entry per step, written by the neighbouring band's GPU while
// this kernel runs, NULL on the first band of that direction), bs.out[s] the neighbour's mailbox this band writes into
// (NULL on the last band). All bands run at the same time, one step behind each other -- the hand-over between two blocks
// of a sweep, over NVLink.
void launch_sgm_band_rows(const uint8_t *fused, const Dims &d, int band_r0, int band_r1, unsigned mask, const BandStream &bs, SgmScratch &sc,
                          int *status, cudaStream_t st, LaunchCounter &lc)
{
    const Roi roi = make_roi(d, false);
    launch_sweeps(fused, d, roi, band_r0, band_r1, mask & 0x5u, nullptr, nullptr, &bs, sc, status, st, lc);
}

} // namespace sister
