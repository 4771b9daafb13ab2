// Host-side launch wrappers of the sm_100a kernels (definitions in stage.cu / match.cu / sgm.cu).
#pragma once
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace sister {

// Opt a kernel in to `bytes` of dynamic shared memory (above the 48 KB default). cudaFuncSetAttribute applies to the
// CURRENT device only, so the opt-in is remembered per (kernel, device): a second context on another GPU of the same
// process gets its own (one context per GPU in one process is a supported set-up).
inline cudaError_t optin_dynamic_smem(const void *kernel, size_t bytes)
{
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, size_t> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    const auto key = std::make_pair(kernel, dev);
    const auto it = done.find(key);
    if (it != done.end() && it->second >= bytes) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) done[key] = bytes;
    return e;
}

struct LaunchCounter {
    unsigned long long total = 0;
    int stage[16] = {0};
    int cur_stage = 0;
    cudaError_t err = cudaSuccess; // first failure of a launcher's own set-up (allocation, attribute); checked by the caller
    inline void add(int n = 1) { total += n; stage[cur_stage] += n; }
    inline void fail(cudaError_t e) { if (e != cudaSuccess && err == cudaSuccess) err = e; }
};

// ---- stage.cu ----
// grey + replicate pad + re-orientation: 5 input views -> 8 oriented view-frame images (hpp:29-70)
void launch_prep(const uint8_t *in, size_t view_stride, int row_stride, int channels, const Dims &d,
                 uint8_t *oriented, cudaStream_t st, LaunchCounter &lc);
// 9x7 centre-symmetric census with the bit-63 carry, on the 8 oriented images (census.cpp:30-51)
void launch_census(const uint8_t *oriented, const Dims &d, unsigned long long *census, cudaStream_t st, LaunchCounter &lc);

// ---- match.cu ----
// raw Hamming cost + WTA-left + WTA-right for the views in view_mask (census.cpp:54-146, postprocess.cpp:74-315)
// share / n_shares: the rows [hv * share / n, hv * (share + 1) / n) of every view only (row bands; 0 / 1 = all rows)
void launch_match_wta(const unsigned long long *census, const Dims &d, unsigned view_mask, int16_t *wtaL, int16_t *wtaR,
                      cudaStream_t st, LaunchCounter &lc, int share = 0, int n_shares = 1);
// in-place-semantics 3x3 median (postprocess.cpp:15-71 with src == dst), LRC (postprocess.cpp:318-341), masks (hpp:201-251)
void launch_median_lrc_mask(const int16_t *wtaL, const int16_t *wtaR, const Dims &d, unsigned view_mask, int16_t *medL,
                            int16_t *medR, int16_t *lr_final, uint8_t *masks, int *status, cudaStream_t st, LaunchCounter &lc);
// fused volume C = sum_v mask_v * cost_v as uint8 (hpp:255-277); view_mask selects the mode's views
// row_lo / row_hi: image rows to produce (whole tiles; the full frame is 0, Hp)
// fused_h / fused_v (both or neither, with view_mask 0xF): the same pass also writes the fused volumes of the horizontal
// pair (mode 1) and of the vertical pair (mode 2); `fused` then holds their sum, the multiview volume of mode 0
void launch_fuse(const unsigned long long *census, const uint8_t *masks, const Dims &d, unsigned view_mask, uint8_t *fused,
                 int *status, cudaStream_t st, LaunchCounter &lc, int row_lo = 0, int row_hi = -1, uint8_t *fused_h = nullptr,
                 uint8_t *fused_v = nullptr);

// ---- sgm.cu ----
// Aggregation scratch of one slot: the four pair volumes (4 * cells bytes) and the mailbox through which the blocks of a
// sweep hand the rider states on (sgm_mailbox_bytes), with the epoch of its tags.
struct SgmScratch {
    uint8_t *vols = nullptr;
    size_t vol_stride = 0;  // bytes between two pair volumes (0: d.cells, the whole frame)
    size_t row_shift = 0;   // a band context holds the rows from band_row0 on only: band_row0 * Wp * D, subtracted from the bases
    uint8_t *mailbox = nullptr;
    size_t mailbox_bytes = 0;
    unsigned epoch = 0;
    unsigned long long geo_key = 0;
};
// mailbox size that serves every rig a context of this capacity accepts
size_t sgm_mailbox_bytes(int max_w, int max_h, int max_d, int band_rows = 0);
// 8-path SGM (sgm.cpp:26-455) on the uint8 fused volume: four two-path sweeps write four one-byte pair volumes, then
// S = nC * C + sum of the pair bytes, the final WTA-left (hpp:283) and convertTo/crop/*255 (hpp:111-118) in one sweep.
// sum (uint16 [Hp][Wp][D]) is written only when non-null (test tap); raw_disp / out may be null.
// full_frame = false aggregates for the crop Rect(D, D, W, H) only (all the caller ever sees, hpp:116-118): rows and
// columns behind it are not run, those in front of it run only the diagonal path that will enter it, bytes exist only
// inside it. raw_disp and sum are then written inside the crop only.
void launch_sgm(const uint8_t *fused, const Dims &d, bool full_frame, SgmScratch &sc, uint16_t *sum, int16_t *raw_disp, uint16_t *out,
                int *status, cudaStream_t st, LaunchCounter &lc);

// WTARight_SSE on the aggregated volume (hpp:138): int16 map Hp x Wp
void launch_wta_right_sum(const uint16_t *sum, const Dims &d, int16_t *outR, cudaStream_t st, LaunchCounter &lc);

// One row band [band_r0, band_r1) of the padded frame (a large frame split over several GPUs),
// crop-only aggregation. what: 1 / 2 = the two sweeps of pass 0 / pass 1 inside the band, continued from state_in (the
// neighbouring band's state_out; null on the first band of the pass) and leaving their state in state_out (null on the
// last), 3 = final sum / WTA / encode of the band's rows. State: sgm_band_state_bytes.
size_t sgm_band_state_bytes(const Dims &d);
// row sweeps streamed between the bands: per sweep (0: pass 0, 2: pass 1) this band's mailbox and the neighbour's, and the
// tag (1..15) every word of this frame carries; see launch_sgm_band_rows in sgm.cu
struct BandStream {
    const uint8_t *in[4] = {nullptr, nullptr, nullptr, nullptr};
    uint8_t *out[4] = {nullptr, nullptr, nullptr, nullptr};
    unsigned tag = 0;
};
void launch_sgm_band_rows(const uint8_t *fused, const Dims &d, int band_r0, int band_r1, unsigned mask, const BandStream &bs, SgmScratch &sc,
                          int *status, cudaStream_t st, LaunchCounter &lc);
void launch_sgm_band(int what, const uint8_t *fused, const Dims &d, int band_r0, int band_r1, const uint8_t *state_in, uint8_t *state_out,
                     SgmScratch &sc, int16_t *raw_disp, uint16_t *out, int *status, cudaStream_t st, LaunchCounter &lc);

} // namespace sister
