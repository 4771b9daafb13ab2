/*
 * sister_b200 -- C ABI of the B200-native (sm_100a) 5-view disparity path.
 *
 * This is the drop-in boundary for the hot path of CVLAB-Unibo/sister:
 *   SisterMultiviewDisparities::compute_disparities   cpp/include/sister/SisterMultiviewDisparities.hpp:26-119
 *     -> doMultiStereo                                hpp:152-295
 *        -> ad_census / WTA / median / LRC / sgm      cpp/src/sister/{census,postprocess,sgm}.cpp
 * Plain pointers and sizes only; no C++/torch types. Every entry point returns 0 or a negative
 * SISTER_E_* code, never throws, never prints (the reference prints three timing lines per call,
 * hpp:79,84,89, and reports errors by aborting).
 *
 * Terminology follows the reference: a "rig" is one set of five views (center, right, top, left,
 * bottom -- the constructor order of hpp:22); disp_count is the reference's dispCount (number of
 * disparity hypotheses, max disparity = disp_count - 1, hpp:155); the three outputs are
 * disp_multiview / disp_horizontal / disp_vertical of hpp:26, CV_16UC1 H x W holding
 * saturate_u16(disparity * 255) (hpp:111-118).
 *
 * There is no CPU fallback: every function that computes needs a CUDA device of compute
 * capability 10.x and fails with SISTER_E_CUDA / SISTER_E_DEVICE otherwise.
 */
#ifndef SISTER_B200_H
#define SISTER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SISTER_B200_VERSION 100 /* 0.1.0 */

enum {
    SISTER_OK = 0,
    SISTER_E_ARG = -1,      /* null pointer, bad channel count, bad slot ...                          */
    SISTER_E_SHAPE = -2,    /* violates the reference's implicit preconditions (see below)             */
    SISTER_E_CAPACITY = -3, /* larger than the max_w / max_h / max_disp given to sister_create         */
    SISTER_E_CUDA = -4,     /* a CUDA runtime call failed; sister_last_error() has the text            */
    SISTER_E_DEVICE = -5,   /* no usable sm_100 device                                                 */
    SISTER_E_NOMEM = -6,    /* device or pinned-host allocation failed                                 */
    SISTER_E_INTERNAL = -7, /* an invariant the kernels rely on was violated (reported, never silent)  */
    SISTER_E_BUSY = -8      /* slot already has work in flight                                         */
};

/* mode_mask bits: which of the three doMultiStereo runs to perform (hpp:77,82,87). */
#define SISTER_MODE_MULTIVIEW 1u  /* mode 0: right + left + top + bottom */
#define SISTER_MODE_HORIZONTAL 2u /* mode 1: right + left                 */
#define SISTER_MODE_VERTICAL 4u   /* mode 2: top + bottom                 */
#define SISTER_MODE_ALL 7u

typedef struct sister_ctx sister_ctx;

/*
 * Shape preconditions, identical to the reference's implicit ones but reported
 * instead of crashing:   disp_count % 8 == 0   (sgm.cpp:268, postprocess.cpp:97)
 *                        (w + 2*disp_count) % 4 == 0 and (h + 2*disp_count) % 4 == 0   (postprocess.cpp:18)
 * Lifted limits: disp_count up to 512 (reference: < 272, postprocess.cpp:193) and 64-bit volume
 * indexing (reference: int32, types.h:31-34).
 */

/* Create a context on CUDA device `device` able to process rigs up to max_w x max_h with up to
 * max_disp disparities, with n_slots rigs in flight (each slot owns a stream, pinned staging and all
 * scratch volumes: about (5.3 * cells + 140 * pixels) bytes, cells = (w+2D)(h+2D)D: the uint8 fused volume,
 * four uint8 pair volumes of the aggregation and its block-to-block mailbox; 2.6 GB at 1280 x 960 x 192; a call that
 * asks for several modes adds 2 * cells on first use). */
int sister_create(sister_ctx **ctx, int device, int max_w, int max_h, int max_disp, int n_slots);
int sister_destroy(sister_ctx *ctx);

/*
 * Synchronous drop-in for compute_disparities (hpp:26) on host buffers.
 *   views       5 pointers: center, right, top, left, bottom (hpp:22 order)
 *   channels    3 = packed BGR as produced by cv::imread (compute_disp.cpp:19-23), converted with
 *               OpenCV-4's fixed-point BGR2GRAY (hpp:29-33); 1 = already grey
 *   row_stride  bytes between rows of each view (cv::Mat::step)
 *   out         3 pointers (multiview, horizontal, vertical), each H x W uint16, dense; entries whose
 *               mode bit is clear may be NULL and are not touched
 *   raw_disp    optional: 3 x (h+2D) x (w+2D) int16, the un-encoded integer disparity of the whole
 *               padded frame per mode (what WTALeft_SSE wrote at hpp:283). Asking for it makes this call
 *               aggregate the whole padded frame (see sister_set_full_frame).
 */
int sister_compute(sister_ctx *ctx, const uint8_t *const views[5], int w, int h, int channels,
                   size_t row_stride, int disp_count, unsigned mode_mask, uint16_t *const out[3],
                   int16_t *raw_disp);

/* Same for n_rigs rigs (the batched / service shape of ros/README.md:31-68): views holds 5*n_rigs
 * pointers, out 3*n_rigs. Rigs are pipelined over the context's slots: H2D copy, kernels and D2H copy
 * of different rigs overlap. All rigs share one shape. */
int sister_compute_batch(sister_ctx *ctx, int n_rigs, const uint8_t *const *views, int w, int h,
                         int channels, size_t row_stride, int disp_count, unsigned mode_mask,
                         uint16_t *const *out);

/* Asynchronous halves of sister_compute for callers that pipeline themselves. */
int sister_submit(sister_ctx *ctx, int slot, const uint8_t *const views[5], int w, int h, int channels,
                  size_t row_stride, int disp_count, unsigned mode_mask);
int sister_wait(sister_ctx *ctx, int slot, uint16_t *const out[3], int16_t *raw_disp);

/* Device-resident variant: views_dev are device pointers (same layout as the host views, dense rows
 * of w*channels bytes), out_dev device pointers to H x W uint16 (may be NULL per cleared mode bit).
 * Enqueues kernels only on the slot's stream; no copies, no synchronisation. */
int sister_submit_device(sister_ctx *ctx, int slot, const uint8_t *const views_dev[5], int w, int h,
                         int channels, int disp_count, unsigned mode_mask, uint16_t *const out_dev[3]);
int sister_sync(sister_ctx *ctx, int slot); /* slot < 0: all slots */

/* compute_disparities returns the crop Rect(D, D, W, H) of each map (hpp:116-118), and nothing else of the
 * padded frame ever leaves the reference. By default the aggregation therefore runs for that crop only: SGM
 * paths still start at the borders of the padded frame (their state flows into the crop exactly as in
 * sgm.cpp:26-455), but paths that never reach the crop are skipped, the others stop once they have left it,
 * and the aggregated cost is formed inside it only. The three output maps are bit-identical either way.
 * enabled = 1 aggregates the whole padded frame on later submits, which the raw_disp argument of
 * sister_wait and SISTER_TAP_RAW_DISP / SISTER_TAP_SUM need (sister_compute with raw_disp and
 * sister_set_test_taps switch it on by themselves). */
int sister_set_full_frame(sister_ctx *ctx, int enabled);

/*
 * The two-view path of the reference (doStereo, hpp:122-150): AD-census cost of
 * (center, side) (census.cpp:149-158), SGM on that raw volume (hpp:135), WTA left and right on the aggregated volume
 * (hpp:137-138), in-place 3x3 median on both maps (hpp:139-140), LRC with threshold 5 (hpp:143).
 *   center, side   grey uint8 images, h rows of w pixels, row_stride bytes apart; `side` is the view whose content
 *                  appears shifted towards smaller columns (the "right" view of the rig, hpp:181)
 *   out_left       h x w float: the disparity of `center`, -10 where the left-right check rejects it (doLRCheck)
 *   out_right      h x w float or NULL: the median-filtered disparity of `side`
 * No padding and no crop: the frame is the image, w % 4 == 0, h % 4 == 0, disp_count % 8 == 0. Runs on slot 0,
 * synchronous. The frame must fit the context: w * h <= (max_w + 2 max_disp)(max_h + 2 max_disp).
 */
int sister_stereo(sister_ctx *ctx, const uint8_t *center, const uint8_t *side, int w, int h, size_t row_stride,
                  int disp_count, float *out_left, float *out_right);

/*
 * Row bands: ONE large frame split over several GPUs (BASELINE.json configs[3]).
 * Each GPU (one context per GPU, one process per GPU) owns the rows [band_row0, band_row1) of the PADDED frame
 * (0 .. h + 2 * disp_count). What is split is everything that is a volume: the fused cost (hpp:255-277), the
 * four SGM pair volumes and the final sum / WTA (sgm.cpp:26-455, hpp:283) -- most of the time and, in a context made
 * with sister_create_band, most of the memory (a context made with sister_create keeps whole-frame volumes and merely
 * fills the band's rows). Staging, census, raw-cost WTA and the masks are computed for the whole frame by every band
 * (their neighbourhoods reach D rows across a band border for the vertical views, and the recursive median
 * (postprocess.cpp:15-71 in place) is a whole-map recurrence).
 * Every SGM path but the horizontal one crosses the bands, and the horizontal path is paired with a diagonal one
 * (sister_b200/csrc/sgm.cu): pass 0 runs top to bottom, pass 1 bottom to top, and a band continues each path from the state
 * the neighbouring band left -- the exact recurrence, no approximate overlap. That state is sister_band_state_bytes() of
 * device memory per pass; moving it between the GPUs (NCCL send / recv, cudaMemcpyPeer) is the caller's job, see
 * sister_b200/bands.py for the schedule (pass 0 flows down the ranks while pass 1 flows up).
 *
 *   sister_band_submit    staging .. masks for the whole frame, fused cost for the band; mode: 0 multiview,
 *                         1 horizontal, 2 vertical (one map per call)
 *   sister_band_vertical  the four paths of one pass inside the band. state_in_dev: what the band above (pass 0) /
 *                         below (pass 1) wrote, NULL on the first band of the pass; state_out_dev: receives the state
 *                         for the next band, NULL on the last band (SISTER_E_ARG when a state that is needed is missing)
 *   sister_band_finish    final sum + WTA + encode of the band's rows of the crop into out_dev, a full H x W
 *                         uint16 map of which only the band's rows are written
 * All three only enqueue on the slot's stream; sister_sync(slot) completes them. Outputs are bit-identical to
 * sister_compute on one GPU (tests/test_bands_gpu.py).
 *
 * The raw-cost WTA (census.cpp:63-88 + postprocess.cpp:103-141 for both maps of every view) is the largest of the stages a
 * band would otherwise repeat for the whole frame, and its rows are independent, so n bands can share it:
 *   sister_band_submit_share  like sister_band_submit up to the WTA, but only for the rows [hv * share / n, hv * (share + 1) / n)
 *                             of every view (hv: the view frame's rows); the results are packed into share_out_dev,
 *                             sister_band_share_bytes() of device memory (the same size for every share)
 *   (the caller all-gathers the n shares, in share order, into one buffer of n * sister_band_share_bytes())
 *   sister_band_submit_rest   unpacks the gathered shares into the slot's maps, then the masks and the band's fused cost:
 *                             the slot is where sister_band_submit would have left it
 */
size_t sister_band_state_bytes(int w, int h, int disp_count);
/* A context for row bands only: like sister_create, but the volumes of a slot -- the fused cost and the four SGM pair
 * volumes, 5 bytes per cell -- hold max_band_rows rows of the padded frame instead of all h + 2 * disp_count, so that a frame
 * whose volumes do not fit one GPU can be split over several (or run band after band on one). The per-pixel buffers (images,
 * census codes, per-view maps: about 140 bytes per padded pixel) still cover the whole frame. Such a context accepts the
 * sister_band_* calls only (anything else: SISTER_E_CAPACITY). */
int sister_create_band(sister_ctx **ctx, int device, int max_w, int max_h, int max_disp, int n_slots, int max_band_rows);
int sister_band_submit(sister_ctx *ctx, int slot, const uint8_t *const views_dev[5], int w, int h, int channels,
                       int disp_count, int mode, int band_row0, int band_row1);
size_t sister_band_share_bytes(int w, int h, int disp_count, int n_shares);
int sister_band_submit_share(sister_ctx *ctx, int slot, const uint8_t *const views_dev[5], int w, int h, int channels,
                             int disp_count, int mode, int band_row0, int band_row1, int share, int n_shares,
                             uint8_t *share_out_dev);
int sister_band_submit_rest(sister_ctx *ctx, int slot, const uint8_t *shares_dev, int n_shares);
int sister_band_vertical(sister_ctx *ctx, int slot, int pass, const uint8_t *state_in_dev, uint8_t *state_out_dev);
/* The same aggregation with the ROW sweeps of all bands running at the same time. A row sweep takes as many steps as the
 * frame is wide however few rows the band has, so handing its state over only when a band has finished (sister_band_vertical)
 * serialises G sweeps of full length. Instead the band below reads the rider states from a mailbox in its own memory WHILE the
 * band above writes them there, one step behind -- exactly the hand-over between two thread blocks of a sweep, across NVLink:
 *   sister_band_rows     passes: 1 = pass 0, 2 = pass 1, 3 = both in one launch. in_pass0: this band's mailbox for pass 0
 *                        (written by the band above), out_pass0: the mailbox of the band below (peer memory, sister_ipc_open);
 *                        in_pass1 / out_pass1 likewise with below / above. A mailbox is sister_band_state_bytes() of zero-
 *                        initialised device memory; NULL on the side where the band has no neighbour. tag: 1 .. 15, the same
 *                        on all bands of a frame and different from the previous frame's (every word carries it, so a
 *                        reader never takes a word of the frame before for one of this frame).
 *   sister_band_columns  the column sweep of one pass alone, states as in sister_band_vertical (still handed over when the
 *                        band has finished: a column sweep has only as many steps as the band has rows). It runs on a second
 *                        stream of the slot, beside the row sweeps; sister_band_columns_wait() blocks until the column sweeps
 *                        enqueued so far are done (state_out_dev complete) without waiting for the row sweeps, and
 *                        sister_band_finish orders the final sweep behind both. (The second stream has a hand-over mailbox of
 *                        its own, allocated on the slot's first sister_band_columns call: as large as the slot's first one.)
 * The bands' sister_band_rows must all be enqueued before any of them is waited for; a band whose neighbour never starts
 * reports SISTER_E_INTERNAL (status bit "spin timeout") after a few seconds instead of hanging. */
int sister_band_rows(sister_ctx *ctx, int slot, int passes, const uint8_t *in_pass0, uint8_t *out_pass0, const uint8_t *in_pass1,
                     uint8_t *out_pass1, unsigned tag);
int sister_band_columns(sister_ctx *ctx, int slot, int pass, const uint8_t *state_in_dev, uint8_t *state_out_dev);
int sister_band_columns_wait(sister_ctx *ctx, int slot);
int sister_band_finish(sister_ctx *ctx, int slot, uint16_t *out_dev);

/* Plain device memory helpers so that a host language needs no CUDA binding of its own. */
int sister_dev_alloc(sister_ctx *ctx, size_t bytes, void **dev_ptr);
int sister_dev_free(sister_ctx *ctx, void *dev_ptr);
int sister_dev_memset(sister_ctx *ctx, void *dev_ptr, int value, size_t bytes);
/* Memory of sister_dev_alloc made visible to the other processes of the box (one process per GPU, row bands): the owner
 * exports a 64-byte handle (cudaIpcMemHandle_t), a peer opens it into its own address space (peer access over NVLink) and
 * closes it before the owner frees the memory. */
int sister_ipc_export(sister_ctx *ctx, void *dev_ptr, unsigned char handle_out[64]);
int sister_ipc_open(sister_ctx *ctx, const unsigned char handle[64], void **dev_ptr);
int sister_ipc_close(sister_ctx *ctx, void *dev_ptr);
/* Page-locked host memory. sister_submit / sister_compute(_batch) copy a dense view that lives in page-locked memory
 * (from here, cudaHostAlloc or cudaHostRegister) to the device directly; other buffers are staged through the slot's
 * own pinned area with one extra memcpy. */
int sister_host_alloc(sister_ctx *ctx, size_t bytes, void **host_ptr);
int sister_host_free(sister_ctx *ctx, void *host_ptr);
int sister_dev_upload(sister_ctx *ctx, void *dev_dst, const void *host_src, size_t bytes);
int sister_dev_download(sister_ctx *ctx, void *host_dst, const void *dev_src, size_t bytes);

/* ---- measurement hooks (bench.py) ---- */
/* Per-stage device time of the LAST submit on `slot`, measured with CUDA events on the slot's stream
 * (enable with sister_set_profiling before submitting; adds 2 event records per stage). */
enum {
    SISTER_STAGE_PREP = 0,  /* grey + replicate pad + re-orientation       (hpp:29-70)        */
    SISTER_STAGE_CENSUS,    /* 8 census maps                               (census.cpp:38-51) */
    SISTER_STAGE_MATCH,     /* raw Hamming cost + WTA left/right, 4 views  (census.cpp:54-146, postprocess.cpp:74-315) */
    SISTER_STAGE_MASK,      /* recursive median, LRC, confidence masks     (hpp:193-252)      */
    SISTER_STAGE_FUSE,      /* confidence-weighted fused volume            (hpp:255-277)      */
    SISTER_STAGE_AGGREGATE, /* SGM, both passes, + final WTA + encode/crop (sgm.cpp:26-455, hpp:283,111-118): the path
                               kernel and the sum/WTA kernel                                   */
    SISTER_STAGE_SELECT,    /* (folded into AGGREGATE: the final WTA reads the path volumes directly; always 0) */
    SISTER_STAGE_COUNT
};
int sister_set_profiling(sister_ctx *ctx, int enabled);
/* Device-time bracket over ALL slots' streams (CUDA events): begin waits for idle streams and fences them behind a
 * start event; end joins every slot stream into an end event, synchronises and returns the elapsed milliseconds. */
int sister_region_begin(sister_ctx *ctx);
int sister_region_end(sister_ctx *ctx, float *elapsed_ms);
int sister_get_stage_ms(sister_ctx *ctx, int slot, float *ms, int n);      /* summed over the modes run */
int sister_get_stage_launches(sister_ctx *ctx, int slot, int *count, int n); /* kernel launches per stage, last submit */
uint64_t sister_get_launch_count(sister_ctx *ctx);                          /* kernels launched since create */

/* ---- test taps (tests/ only): copy an intermediate product of the last submit on `slot` to the host ---- */
enum {
    SISTER_TAP_ORIENTED = 0, /* 8 x px uint8: view-frame images, order C0,R,C180,L,C90,T,C270,B          */
    SISTER_TAP_CENSUS,       /* 8 x px uint64: census codes of the same 8 images                          */
    SISTER_TAP_WTA_L,        /* 4 x px int16: raw-volume WTA-left per view (right,left,top,bottom), view frame */
    SISTER_TAP_WTA_R,        /* 4 x px int16: raw-volume WTA-right                                        */
    SISTER_TAP_LR_FINAL,     /* 4 x px int16: left map after median + LRC (-10 = rejected), view frame    */
    SISTER_TAP_MASKS,        /* 4 x px uint8: confidence masks in the image frame                         */
    SISTER_TAP_FUSED,        /* cells uint8: fused volume of the LAST mode run, [row][col][d]             */
    SISTER_TAP_SUM,          /* cells uint16: aggregated volume of the LAST mode run (needs sister_set_test_taps) */
    SISTER_TAP_RAW_DISP      /* 3 x px int16                                                              */
};
int sister_debug_fetch(sister_ctx *ctx, int slot, int what, void *host_dst, size_t bytes);
/* The aggregated volume is not materialised by the product path (the final WTA consumes the path volumes directly);
 * enabling the taps allocates it (2 * cells bytes per slot) and makes later submits write it for SISTER_TAP_SUM. */
int sister_set_test_taps(sister_ctx *ctx, int enabled);

/* Stage-level entry points for known-answer tests: run ONE stage on caller data (host pointers). */
/* sgm(): fused volume uint8 [h][w][D] (values <= 252) -> aggregated uint16 [h][w][D] (sgm.cpp:457). */
int sister_test_sgm(sister_ctx *ctx, const uint8_t *fused, int w, int h, int disp_count, uint16_t *sum,
                    int16_t *disp /* optional WTA-left of the sum, h*w */);

const char *sister_strerror(int code);
const char *sister_last_error(sister_ctx *ctx); /* detail text of the last failure on this context */
int sister_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SISTER_B200_H */
