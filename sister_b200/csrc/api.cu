// sister_b200 / api.cu -- the C ABI (include/sister_b200.h): context, slots, pipeline orchestration.
//
// Replaces, for the caller, SisterMultiviewDisparities::compute_disparities (hpp:26-119): staging (hpp:29-70), the
// three doMultiStereo runs (hpp:77-89) and the output encoding (hpp:111-118). Census, per-view WTA and the
// confidence masks are identical across the three modes (the reference recomputes them, hpp:181-252); here they are
// computed once per rig and only fuse -> SGM -> WTA run per mode.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/sister_b200.h"
#include "kernels.cuh"

using namespace sister;

namespace {

struct Slot {
    cudaStream_t st = nullptr;
    uint8_t *d_in = nullptr, *h_in = nullptr;
    uint8_t *d_oriented = nullptr;
    unsigned long long *d_census = nullptr;
    int16_t *d_wtaL = nullptr, *d_wtaR = nullptr, *d_medL = nullptr, *d_medR = nullptr, *d_lr = nullptr;
    uint8_t *d_masks = nullptr, *d_fused = nullptr;
    const uint8_t *last_fused = nullptr; // fused volume of the last mode run (SISTER_TAP_FUSED)
    uint8_t *d_fused_hv = nullptr; // the horizontal-pair and vertical-pair volumes of a call that asks for several modes (allocated on first use)
    SgmScratch sgm;             // four one-byte pair volumes + the sweeps' mailbox (sgm.cu)
    uint16_t *d_sum = nullptr;  // aggregated volume, allocated only while the test taps are enabled
    int16_t *d_raw = nullptr;
    uint16_t *d_out = nullptr, *h_out = nullptr;
    int *d_status = nullptr, *h_status = nullptr;
    Dims dims{};
    unsigned mode_mask = 0;
    bool busy = false, host_io = false;
    bool full_frame = false; // the last run aggregated the whole padded frame (raw_disp is complete)
    int band_r0 = 0, band_r1 = 0; // row band in progress (sister_band_*)
    // column sweeps of a band beside its streamed row sweeps (sister_band_columns): their own stream and mailbox, made on first use
    cudaStream_t st_cols = nullptr;
    SgmScratch sgm_cols;
    cudaEvent_t ev_fused = nullptr, ev_cols = nullptr; // the band's fused cost is ready (on st) / the column sweeps so far are done (on st_cols)
    bool cols_pending = false;
    // profiling
    std::vector<cudaEvent_t> ev_b, ev_e;
    std::vector<int> ev_stage;
    int n_ev = 0;
    int stage_launches[SISTER_STAGE_COUNT] = {0};
};

} // namespace

struct sister_ctx {
    int device = 0;
    int max_w = 0, max_h = 0, max_d = 0;
    long long px_max = 0, cells_max = 0;
    int band_rows = 0;       // > 0: a band context (sister_create_band): the volumes hold this many rows of the padded frame
    long long vol_cells = 0; // cells a slot's fused volume (and each pair volume) holds
    size_t in_bytes_max = 0;
    std::vector<Slot> slots;
    bool profiling = false;
    bool taps = false; // keep the aggregated volume for SISTER_TAP_SUM (tests only)
    bool full_frame = false; // aggregate the whole padded frame (raw_disp, taps) instead of the crop the caller sees
    cudaEvent_t region_b = nullptr, region_e = nullptr;
    std::vector<cudaEvent_t> region_join;
    LaunchCounter lc;
    std::string err;
};

namespace {

int fail_cuda(sister_ctx *ctx, cudaError_t e, const char *what)
{
    if (ctx) ctx->err = std::string(what) + ": " + cudaGetErrorString(e);
    return (e == cudaErrorMemoryAllocation) ? SISTER_E_NOMEM : SISTER_E_CUDA;
}

#define SCK(call)                                                        \
    do {                                                                 \
        cudaError_t e__ = (call);                                        \
        if (e__ != cudaSuccess) return fail_cuda(ctx, e__, #call);       \
    } while (0)

// a launcher's own set-up failure (shared-memory opt-in, table allocation), reported once
cudaError_t take_launch_error(sister_ctx *ctx)
{
    const cudaError_t e = ctx->lc.err;
    ctx->lc.err = cudaSuccess;
    return e;
}

int check_shape(sister_ctx *ctx, int w, int h, int D, Dims &d, bool band_call = false)
{
    if (w <= 0 || h <= 0 || D <= 0) { ctx->err = "non-positive size"; return SISTER_E_ARG; }
    if (D % 8 != 0) { ctx->err = "disp_count must be a multiple of 8 (sgm.cpp:268)"; return SISTER_E_SHAPE; }
    if (D > 512) { ctx->err = "disp_count above 512 is not supported"; return SISTER_E_SHAPE; }
    if ((w + 2 * D) % 4 != 0 || (h + 2 * D) % 4 != 0) { ctx->err = "(w + 2*disp_count) and (h + 2*disp_count) must be multiples of 4 (postprocess.cpp:18)"; return SISTER_E_SHAPE; }
    if (w + 2 * D < 16 || h + 2 * D < 16) { ctx->err = "padded frame too small for the 9x7 census"; return SISTER_E_SHAPE; }
    d.W = w; d.H = h; d.D = D; d.Wp = w + 2 * D; d.Hp = h + 2 * D;
    d.px = (long long)d.Wp * d.Hp;
    d.cells = d.px * D;
    set_cell_order(d);
    if (d.cells / 8 > 0x7F000000LL) { ctx->err = "cost volume above 1.7e10 cells (32-bit cursor offsets in sgm.cu)"; return SISTER_E_SHAPE; }
    if (w > ctx->max_w || h > ctx->max_h || D > ctx->max_d || d.px > ctx->px_max || d.cells > ctx->cells_max) {
        ctx->err = "rig larger than the capacity given to sister_create";
        return SISTER_E_CAPACITY;
    }
    if (ctx->band_rows > 0 && !band_call) {
        ctx->err = "a band context (sister_create_band) holds a band of the volumes only: use sister_band_*";
        return SISTER_E_CAPACITY;
    }
    return SISTER_OK;
}

void begin_stage(sister_ctx *ctx, Slot &s, int stage)
{
    ctx->lc.cur_stage = stage;
    if (!ctx->profiling) return;
    if (s.n_ev == (int)s.ev_b.size()) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        s.ev_b.push_back(a); s.ev_e.push_back(b); s.ev_stage.push_back(stage);
    }
    s.ev_stage[s.n_ev] = stage;
    cudaEventRecord(s.ev_b[s.n_ev], s.st);
}
void end_stage(sister_ctx *ctx, Slot &s)
{
    if (!ctx->profiling) return;
    cudaEventRecord(s.ev_e[s.n_ev], s.st);
    s.n_ev++;
}

// Enqueue the whole path for one rig on the slot's stream. `in` is a device pointer to 5 views.
int run_pipeline(sister_ctx *ctx, Slot &s, const uint8_t *const in_views[5], int row_stride, int channels, const Dims &d,
                 unsigned mode_mask, uint16_t *const out_dev[3])
{
    s.n_ev = 0;
    s.sgm.row_shift = 0;
    int before[16];
    memcpy(before, ctx->lc.stage, sizeof(before));
    // the status word ACCUMULATES (kernels atomicOr into it) over every submit queued on the slot until finish_slot has
    // read and cleared it: a bit raised by an earlier rig is not overwritten by a later one
    const unsigned view_mask = (mode_mask & SISTER_MODE_MULTIVIEW) ? 0xFu
                               : (((mode_mask & SISTER_MODE_HORIZONTAL) ? 0x3u : 0u) | ((mode_mask & SISTER_MODE_VERTICAL) ? 0xCu : 0u));
    // the 5 views may live anywhere on the device: express them relative to view 0 when they are equally spaced,
    // otherwise run the prep per view group (kept simple: require equal spacing, which both callers provide)
    const ptrdiff_t stride = in_views[1] - in_views[0];
    for (int k = 2; k < 5; k++)
        if (in_views[k] - in_views[k - 1] != stride) { ctx->err = "device views must be equally spaced"; return SISTER_E_ARG; }
    if (stride <= 0) { ctx->err = "device views must be in ascending address order"; return SISTER_E_ARG; }

    begin_stage(ctx, s, SISTER_STAGE_PREP);
    launch_prep(in_views[0], (size_t)stride, row_stride, channels, d, s.d_oriented, s.st, ctx->lc);
    end_stage(ctx, s);
    begin_stage(ctx, s, SISTER_STAGE_CENSUS);
    launch_census(s.d_oriented, d, s.d_census, s.st, ctx->lc);
    end_stage(ctx, s);
    begin_stage(ctx, s, SISTER_STAGE_MATCH);
    launch_match_wta(s.d_census, d, view_mask, s.d_wtaL, s.d_wtaR, s.st, ctx->lc);
    end_stage(ctx, s);
    begin_stage(ctx, s, SISTER_STAGE_MASK);
    launch_median_lrc_mask(s.d_wtaL, s.d_wtaR, d, view_mask, s.d_medL, s.d_medR, s.d_lr, s.d_masks, s.d_status, s.st, ctx->lc);
    end_stage(ctx, s);
    // Several modes in one call (the reference class always runs all three, hpp:77-89): one fuse pass evaluates every Hamming
    // distance once and writes the horizontal, the vertical and the multiview (their sum, hpp:262-276) volume together.
    const bool several = (mode_mask & (mode_mask - 1)) != 0;
    if (several && !s.d_fused_hv) {
        const cudaError_t e = cudaMalloc((void **)&s.d_fused_hv, (size_t)2 * ctx->cells_max);
        if (e != cudaSuccess) { cudaGetLastError(); s.d_fused_hv = nullptr; } // no room: fall back to one pass per mode
    }
    const bool triple = several && s.d_fused_hv;
    if (triple) {
        begin_stage(ctx, s, SISTER_STAGE_FUSE);
        launch_fuse(s.d_census, s.d_masks, d, 0xFu, s.d_fused, s.d_status, s.st, ctx->lc, 0, -1, s.d_fused_hv, s.d_fused_hv + (size_t)ctx->cells_max);
        end_stage(ctx, s);
    }
    for (int mode = 0; mode < 3; mode++) {
        if (!(mode_mask & (1u << mode))) continue;
        const unsigned vm = mode == 0 ? 0xFu : mode == 1 ? 0x3u : 0xCu; // hpp:262-276
        const uint8_t *fused_of_mode = !triple || mode == 0 ? s.d_fused : s.d_fused_hv + (size_t)(mode - 1) * (size_t)ctx->cells_max;
        if (!triple) {
            begin_stage(ctx, s, SISTER_STAGE_FUSE);
            launch_fuse(s.d_census, s.d_masks, d, vm, s.d_fused, s.d_status, s.st, ctx->lc);
            end_stage(ctx, s);
        }
        s.last_fused = fused_of_mode;
        begin_stage(ctx, s, SISTER_STAGE_AGGREGATE);
        launch_sgm(fused_of_mode, d, ctx->full_frame || ctx->taps, s.sgm, ctx->taps ? s.d_sum : nullptr, s.d_raw + (size_t)mode * d.px,
                   out_dev ? out_dev[mode] : nullptr, s.d_status, s.st, ctx->lc);
        end_stage(ctx, s);
    }
    SCK(cudaMemcpyAsync(s.h_status, s.d_status, sizeof(int), cudaMemcpyDeviceToHost, s.st));
    SCK(cudaGetLastError());
    SCK(take_launch_error(ctx));
    for (int k = 0; k < SISTER_STAGE_COUNT; k++) s.stage_launches[k] = ctx->lc.stage[k] - before[k];
    s.dims = d;
    s.mode_mask = mode_mask;
    s.full_frame = ctx->full_frame || ctx->taps;
    return SISTER_OK;
}

void free_slot(Slot &s)
{
    if (s.st) cudaStreamSynchronize(s.st);
    cudaFree(s.d_in); cudaFreeHost(s.h_in); cudaFree(s.d_oriented); cudaFree(s.d_census);
    cudaFree(s.d_wtaL); cudaFree(s.d_wtaR); cudaFree(s.d_medL); cudaFree(s.d_medR); cudaFree(s.d_lr);
    cudaFree(s.d_masks); cudaFree(s.d_fused); cudaFree(s.d_fused_hv); cudaFree(s.sgm.vols); cudaFree(s.sgm.mailbox); cudaFree(s.d_sum); cudaFree(s.d_raw); cudaFree(s.d_out);
    cudaFreeHost(s.h_out); cudaFree(s.d_status); cudaFreeHost(s.h_status);
    for (auto e : s.ev_b) cudaEventDestroy(e);
    for (auto e : s.ev_e) cudaEventDestroy(e);
    if (s.st) cudaStreamDestroy(s.st);
    if (s.st_cols) cudaStreamDestroy(s.st_cols);
    if (s.ev_fused) cudaEventDestroy(s.ev_fused);
    if (s.ev_cols) cudaEventDestroy(s.ev_cols);
    cudaFree(s.sgm_cols.mailbox);
    s = Slot();
}

int slot_ok(sister_ctx *ctx, int slot)
{
    if (!ctx) return SISTER_E_ARG;
    if (slot < 0 || slot >= (int)ctx->slots.size()) { ctx->err = "bad slot index"; return SISTER_E_ARG; }
    return SISTER_OK;
}

int finish_slot(sister_ctx *ctx, Slot &s)
{
    SCK(cudaStreamSynchronize(s.st));
    if (*s.h_status != 0) {
        char buf[96];
        snprintf(buf, sizeof buf, "kernel invariant violated, status bits 0x%x", *s.h_status);
        ctx->err = buf;
        *s.h_status = 0;
        cudaMemsetAsync(s.d_status, 0, sizeof(int), s.st); // reported: start the next rig from a clean word
        return SISTER_E_INTERNAL;
    }
    return SISTER_OK;
}

} // namespace

extern "C" {

int sister_version(void) { return SISTER_B200_VERSION; }

const char *sister_strerror(int code)
{
    switch (code) {
    case SISTER_OK: return "ok";
    case SISTER_E_ARG: return "invalid argument";
    case SISTER_E_SHAPE: return "shape violates the path's preconditions";
    case SISTER_E_CAPACITY: return "rig exceeds the context capacity";
    case SISTER_E_CUDA: return "CUDA error";
    case SISTER_E_DEVICE: return "no usable sm_100 device";
    case SISTER_E_NOMEM: return "out of memory";
    case SISTER_E_INTERNAL: return "kernel invariant violated";
    case SISTER_E_BUSY: return "slot busy";
    }
    return "unknown error";
}

const char *sister_last_error(sister_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

static int create_impl(sister_ctx **out, int device, int max_w, int max_h, int max_disp, int n_slots, int band_rows)
{
    if (!out || max_w <= 0 || max_h <= 0 || max_disp <= 0 || n_slots <= 0 || n_slots > 64 || band_rows < 0) return SISTER_E_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return SISTER_E_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SISTER_E_DEVICE;
    if (prop.major != 10) return SISTER_E_DEVICE; // sm_100a code only; no fallback
    sister_ctx *ctx = new (std::nothrow) sister_ctx();
    if (!ctx) return SISTER_E_NOMEM;
    ctx->device = device;
    ctx->max_w = max_w; ctx->max_h = max_h; ctx->max_d = max_disp;
    ctx->px_max = (long long)(max_w + 2 * max_disp) * (max_h + 2 * max_disp);
    ctx->cells_max = ctx->px_max * max_disp;
    ctx->band_rows = band_rows;
    ctx->vol_cells = band_rows > 0 ? (long long)(max_w + 2 * max_disp) * band_rows * max_disp : ctx->cells_max;
    ctx->in_bytes_max = (size_t)5 * max_w * max_h * 3;
    int rc = SISTER_OK;
    auto bail = [&](int code) { for (auto &s : ctx->slots) free_slot(s); delete ctx; return code; };
    if (cudaSetDevice(device) != cudaSuccess) return bail(SISTER_E_DEVICE);
    ctx->slots.resize(n_slots);
    const size_t px = (size_t)ctx->px_max, wh = (size_t)max_w * max_h;
    for (auto &s : ctx->slots) {
        cudaError_t e = cudaSuccess;
        auto A = [&](void **p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
        auto H = [&](void **p, size_t bytes) { if (e == cudaSuccess) e = cudaMallocHost(p, bytes); };
        e = cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking);
        A((void **)&s.d_in, ctx->in_bytes_max); H((void **)&s.h_in, ctx->in_bytes_max);
        A((void **)&s.d_oriented, 8 * px); A((void **)&s.d_census, 8 * px * 8);
        A((void **)&s.d_wtaL, 4 * px * 2); A((void **)&s.d_wtaR, 4 * px * 2);
        A((void **)&s.d_medL, 4 * px * 2); A((void **)&s.d_medR, 4 * px * 2); A((void **)&s.d_lr, 4 * px * 2);
        const size_t vol = (size_t)ctx->vol_cells;
        A((void **)&s.d_masks, 4 * px); A((void **)&s.d_fused, vol); A((void **)&s.sgm.vols, 4 * vol);
        s.sgm.vol_stride = band_rows > 0 ? vol : 0;
        s.sgm.mailbox_bytes = sgm_mailbox_bytes(max_w, max_h, max_disp, band_rows);
        A((void **)&s.sgm.mailbox, s.sgm.mailbox_bytes);
        A((void **)&s.d_raw, 3 * px * 2); A((void **)&s.d_out, 3 * wh * 2); H((void **)&s.h_out, 3 * wh * 2);
        A((void **)&s.d_status, sizeof(int)); H((void **)&s.h_status, sizeof(int));
        if (e != cudaSuccess) { rc = fail_cuda(nullptr, e, "alloc"); return bail(rc); }
        *s.h_status = 0;
        cudaMemsetAsync(s.d_status, 0, sizeof(int), s.st);
        cudaMemsetAsync(s.d_masks, 0, 4 * px, s.st);
        cudaMemsetAsync(s.d_raw, 0, 3 * px * 2, s.st);
    }
    cudaDeviceSynchronize();
    *out = ctx;
    return SISTER_OK;
}

int sister_create(sister_ctx **out, int device, int max_w, int max_h, int max_disp, int n_slots)
{
    return create_impl(out, device, max_w, max_h, max_disp, n_slots, 0);
}

int sister_create_band(sister_ctx **out, int device, int max_w, int max_h, int max_disp, int n_slots, int max_band_rows)
{
    if (max_band_rows <= 0) return SISTER_E_ARG;
    return create_impl(out, device, max_w, max_h, max_disp, n_slots, max_band_rows);
}

int sister_destroy(sister_ctx *ctx)
{
    if (!ctx) return SISTER_E_ARG;
    cudaSetDevice(ctx->device);
    for (auto &s : ctx->slots) free_slot(s);
    if (ctx->region_b) { cudaEventDestroy(ctx->region_b); cudaEventDestroy(ctx->region_e); }
    for (auto e : ctx->region_join) cudaEventDestroy(e);
    delete ctx;
    return SISTER_OK;
}

int sister_submit(sister_ctx *ctx, int slot, const uint8_t *const views[5], int w, int h, int channels, size_t row_stride,
                  int disp_count, unsigned mode_mask)
{
    int rc = slot_ok(ctx, slot);
    if (rc) return rc;
    if (!views || (channels != 1 && channels != 3) || !(mode_mask & SISTER_MODE_ALL)) { ctx->err = "bad views/channels/mode_mask"; return SISTER_E_ARG; }
    for (int k = 0; k < 5; k++) if (!views[k]) { ctx->err = "null view"; return SISTER_E_ARG; }
    if (row_stride < (size_t)w * channels) { ctx->err = "row_stride smaller than a row"; return SISTER_E_ARG; }
    Dims d;
    rc = check_shape(ctx, w, h, disp_count, d);
    if (rc) return rc;
    Slot &s = ctx->slots[slot];
    if (s.busy) { ctx->err = "slot busy"; return SISTER_E_BUSY; }
    SCK(cudaSetDevice(ctx->device));
    const size_t row = (size_t)w * channels, view_bytes = row * h;
    for (int k = 0; k < 5; k++) {
        // a dense view in page-locked memory (sister_host_alloc, cudaHostAlloc, cudaHostRegister) is copied straight from
        // the caller's buffer; anything else goes through the slot's pinned staging area first
        const uint8_t *src = views[k];
        cudaPointerAttributes attr;
        const bool pinned = row_stride == row && cudaPointerGetAttributes(&attr, views[k]) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        if (!pinned) {
            cudaGetLastError(); // a failed query must not poison the stream of error checks
            uint8_t *dst = s.h_in + k * view_bytes;
            if (row_stride == row) memcpy(dst, views[k], view_bytes);
            else for (int i = 0; i < h; i++) memcpy(dst + i * row, views[k] + i * row_stride, row);
            src = dst;
        }
        SCK(cudaMemcpyAsync(s.d_in + k * view_bytes, src, view_bytes, cudaMemcpyHostToDevice, s.st));
    }
    const uint8_t *dv[5];
    for (int k = 0; k < 5; k++) dv[k] = s.d_in + k * view_bytes;
    const size_t wh = (size_t)w * h;
    uint16_t *od[3] = {s.d_out, s.d_out + wh, s.d_out + 2 * wh};
    rc = run_pipeline(ctx, s, dv, (int)row, channels, d, mode_mask, od);
    if (rc) return rc;
    for (int m = 0; m < 3; m++)
        if (mode_mask & (1u << m)) SCK(cudaMemcpyAsync(s.h_out + m * wh, s.d_out + m * wh, wh * 2, cudaMemcpyDeviceToHost, s.st));
    s.busy = true;
    s.host_io = true;
    return SISTER_OK;
}

int sister_wait(sister_ctx *ctx, int slot, uint16_t *const out[3], int16_t *raw_disp)
{
    int rc = slot_ok(ctx, slot);
    if (rc) return rc;
    Slot &s = ctx->slots[slot];
    if (!s.busy || !s.host_io) { ctx->err = "nothing submitted on this slot"; return SISTER_E_ARG; }
    SCK(cudaSetDevice(ctx->device));
    rc = finish_slot(ctx, s);
    s.busy = false;
    if (rc) return rc;
    const size_t wh = (size_t)s.dims.W * s.dims.H;
    for (int m = 0; m < 3; m++)
        if ((s.mode_mask & (1u << m)) && out && out[m]) memcpy(out[m], s.h_out + m * wh, wh * 2);
    if (raw_disp) {
        if (!s.full_frame) { ctx->err = "raw_disp needs sister_set_full_frame(ctx, 1) before the submit"; return SISTER_E_ARG; }
        SCK(cudaMemcpy(raw_disp, s.d_raw, (size_t)3 * s.dims.px * 2, cudaMemcpyDeviceToHost));
    }
    return SISTER_OK;
}

int sister_compute(sister_ctx *ctx, const uint8_t *const views[5], int w, int h, int channels, size_t row_stride,
                   int disp_count, unsigned mode_mask, uint16_t *const out[3], int16_t *raw_disp)
{
    if (!ctx) return SISTER_E_ARG;
    const bool keep = ctx->full_frame;
    if (raw_disp) ctx->full_frame = true; // the whole padded map is asked for: aggregate the whole frame
    int rc = sister_submit(ctx, 0, views, w, h, channels, row_stride, disp_count, mode_mask);
    ctx->full_frame = keep;
    if (rc) return rc;
    return sister_wait(ctx, 0, out, raw_disp);
}

int sister_compute_batch(sister_ctx *ctx, int n_rigs, const uint8_t *const *views, int w, int h, int channels,
                         size_t row_stride, int disp_count, unsigned mode_mask, uint16_t *const *out)
{
    if (!ctx || n_rigs < 0 || !views || !out) return SISTER_E_ARG;
    const int ns = (int)ctx->slots.size();
    int err = SISTER_OK;
    std::string msg;
    for (int k = 0; k < n_rigs && !err; k++) {
        const int slot = k % ns;
        if (k >= ns) err = sister_wait(ctx, slot, out + 3 * (size_t)(k - ns), nullptr); // retire the rig that held this slot
        if (!err) err = sister_submit(ctx, slot, views + 5 * (size_t)k, w, h, channels, row_stride, disp_count, mode_mask);
    }
    if (err) msg = ctx->err;
    for (int k = (n_rigs > ns ? n_rigs - ns : 0); k < n_rigs; k++) {
        Slot &s = ctx->slots[k % ns];
        if (!s.busy) continue;
        if (!err) err = sister_wait(ctx, k % ns, out + 3 * (size_t)k, nullptr);
        else { cudaStreamSynchronize(s.st); s.busy = false; } // drain after a failure
        if (err && msg.empty()) msg = ctx->err;
    }
    if (err) ctx->err = msg;
    return err;
}

int sister_submit_device(sister_ctx *ctx, int slot, const uint8_t *const views_dev[5], int w, int h, int channels,
                         int disp_count, unsigned mode_mask, uint16_t *const out_dev[3])
{
    int rc = slot_ok(ctx, slot);
    if (rc) return rc;
    if (!views_dev || (channels != 1 && channels != 3) || !(mode_mask & SISTER_MODE_ALL)) { ctx->err = "bad views/channels/mode_mask"; return SISTER_E_ARG; }
    Dims d;
    rc = check_shape(ctx, w, h, disp_count, d);
    if (rc) return rc;
    Slot &s = ctx->slots[slot];
    SCK(cudaSetDevice(ctx->device));
    if (s.busy && s.host_io) { ctx->err = "slot has an un-waited host submit"; return SISTER_E_BUSY; }
    rc = run_pipeline(ctx, s, views_dev, w * channels, channels, d, mode_mask, out_dev);
    if (rc) return rc;
    s.host_io = false; // device submits are ordered by the slot's stream; nothing to hand back on the host
    return SISTER_OK;
}

// ---- row bands: one frame split over several GPUs (include/sister_b200.h, "Row bands") ----

size_t sister_band_state_bytes(int w, int h, int disp_count)
{
    if (w <= 0 || h <= 0 || disp_count <= 0 || disp_count % 8 != 0 || disp_count > 512) return 0;
    Dims d;
    d.W = w; d.H = h; d.D = disp_count; d.Wp = w + 2 * disp_count; d.Hp = h + 2 * disp_count;
    d.px = (long long)d.Wp * d.Hp;
    d.cells = d.px * disp_count;
    set_cell_order(d);
    return sgm_band_state_bytes(d);
}

// the packed WTA maps of one share: map i = 2 * view + (0 left, 1 right), each with room for ceil(hv / n) rows
static size_t band_share_layout(const Dims &d, int n, size_t off[9])
{
    size_t o = 0;
    for (int i = 0; i < 8; i++) {
        const int v = i >> 1, hv = view_rows(d, v), wv = view_cols(d, v);
        off[i] = o;
        o += (size_t)((hv + n - 1) / n) * wv * sizeof(int16_t);
    }
    off[8] = o;
    return o;
}

size_t sister_band_share_bytes(int w, int h, int disp_count, int n_shares)
{
    if (w <= 0 || h <= 0 || disp_count <= 0 || n_shares <= 0) return 0;
    Dims d;
    d.W = w; d.H = h; d.D = disp_count; d.Wp = w + 2 * disp_count; d.Hp = h + 2 * disp_count;
    d.px = (long long)d.Wp * d.Hp;
    size_t off[9];
    return band_share_layout(d, n_shares, off);
}

// validation + staging + census; the slot remembers the band
static int band_begin(sister_ctx *ctx, int slot, const uint8_t *const views_dev[5], int w, int h, int channels, int disp_count, int mode,
                      int band_row0, int band_row1)
{
    int rc = slot_ok(ctx, slot);
    if (rc) return rc;
    if (!views_dev || (channels != 1 && channels != 3) || mode < 0 || mode > 2) { ctx->err = "bad views/channels/mode"; return SISTER_E_ARG; }
    Dims d;
    rc = check_shape(ctx, w, h, disp_count, d, true);
    if (rc) return rc;
    if (band_row0 < 0 || band_row1 > d.Hp || band_row0 >= band_row1) { ctx->err = "band rows must satisfy 0 <= row0 < row1 <= h + 2 * disp_count"; return SISTER_E_ARG; }
    if (ctx->band_rows > 0 && band_row1 - band_row0 > ctx->band_rows) { ctx->err = "band taller than the max_band_rows given to sister_create_band"; return SISTER_E_CAPACITY; }
    Slot &s = ctx->slots[slot];
    SCK(cudaSetDevice(ctx->device));
    if (s.busy && s.host_io) { ctx->err = "slot has an un-waited host submit"; return SISTER_E_BUSY; }
    const ptrdiff_t stride = views_dev[1] - views_dev[0];
    for (int k = 2; k < 5; k++)
        if (views_dev[k] - views_dev[k - 1] != stride) { ctx->err = "device views must be equally spaced"; return SISTER_E_ARG; }
    if (stride <= 0) { ctx->err = "device views must be in ascending address order"; return SISTER_E_ARG; }
    launch_prep(views_dev[0], (size_t)stride, w * channels, channels, d, s.d_oriented, s.st, ctx->lc);
    launch_census(s.d_oriented, d, s.d_census, s.st, ctx->lc);
    s.dims = d;
    s.mode_mask = 1u << mode;
    s.full_frame = false;
    s.band_r0 = band_row0; s.band_r1 = band_row1;
    s.host_io = false;
    return SISTER_OK;
}

// masks (whole frame) + the band's fused cost, from the WTA maps in the slot
static int band_rest(sister_ctx *ctx, Slot &s)
{
    const Dims &d = s.dims;
    const unsigned vm = s.mode_mask == 1u ? 0xFu : s.mode_mask == 2u ? 0x3u : 0xCu; // hpp:262-276
    launch_median_lrc_mask(s.d_wtaL, s.d_wtaR, d, vm, s.d_medL, s.d_medR, s.d_lr, s.d_masks, s.d_status, s.st, ctx->lc);
    // a band context's volumes start with the band's first row: every kernel addresses rows of the padded frame, so the
    // bases are moved back by the rows that are not there (only the band's rows are ever touched)
    s.sgm.row_shift = ctx->band_rows > 0 ? (size_t)s.band_r0 * (size_t)d.Wp * (size_t)d.D : 0;
    launch_fuse(s.d_census, s.d_masks, d, vm, s.d_fused - s.sgm.row_shift, s.d_status, s.st, ctx->lc, s.band_r0, s.band_r1);
    s.last_fused = s.d_fused;
    if (s.ev_fused) SCK(cudaEventRecord(s.ev_fused, s.st));
    s.cols_pending = false;
    SCK(cudaGetLastError());
    SCK(take_launch_error(ctx));
    return SISTER_OK;
}

int sister_band_submit(sister_ctx *ctx, int slot, const uint8_t *const views_dev[5], int w, int h, int channels, int disp_count, int mode,
                       int band_row0, int band_row1)
{
    // staging, census, raw-cost WTA and the masks need neighbours far outside the band (D rows for the vertical views,
    // the whole map for the recursive median): every band computes them for the whole frame; the volumes -- fused cost,
    // path bytes, i.e. all the memory and most of the time -- exist for the band's rows only
    int rc = band_begin(ctx, slot, views_dev, w, h, channels, disp_count, mode, band_row0, band_row1);
    if (rc) return rc;
    Slot &s = ctx->slots[slot];
    const unsigned vm = mode == 0 ? 0xFu : mode == 1 ? 0x3u : 0xCu;
    launch_match_wta(s.d_census, s.dims, vm, s.d_wtaL, s.d_wtaR, s.st, ctx->lc);
    return band_rest(ctx, s);
}

int sister_band_submit_share(sister_ctx *ctx, int slot, const uint8_t *const views_dev[5], int w, int h, int channels, int disp_count,
                             int mode, int band_row0, int band_row1, int share, int n_shares, uint8_t *share_out_dev)
{
    if (n_shares <= 0 || share < 0 || share >= n_shares || !share_out_dev) { if (ctx) ctx->err = "0 <= share < n_shares and a share buffer are required"; return SISTER_E_ARG; }
    int rc = band_begin(ctx, slot, views_dev, w, h, channels, disp_count, mode, band_row0, band_row1);
    if (rc) return rc;
    Slot &s = ctx->slots[slot];
    const Dims &d = s.dims;
    const unsigned vm = mode == 0 ? 0xFu : mode == 1 ? 0x3u : 0xCu;
    launch_match_wta(s.d_census, d, vm, s.d_wtaL, s.d_wtaR, s.st, ctx->lc, share, n_shares);
    size_t off[9];
    band_share_layout(d, n_shares, off);
    for (int i = 0; i < 8; i++) {
        const int v = i >> 1, hv = view_rows(d, v), wv = view_cols(d, v);
        if (!((vm >> v) & 1u)) continue;
        const int r0 = share_row0(hv, share, n_shares), r1 = share_row0(hv, share + 1, n_shares);
        const int16_t *src = ((i & 1) ? s.d_wtaR : s.d_wtaL) + (size_t)v * d.px + (size_t)r0 * wv;
        if (r1 > r0) SCK(cudaMemcpyAsync(share_out_dev + off[i], src, (size_t)(r1 - r0) * wv * sizeof(int16_t), cudaMemcpyDeviceToDevice, s.st));
    }
    SCK(cudaGetLastError());
    SCK(take_launch_error(ctx));
    return SISTER_OK;
}

int sister_band_submit_rest(sister_ctx *ctx, int slot, const uint8_t *shares_dev, int n_shares)
{
    int rc = slot_ok(ctx, slot);
    if (rc) return rc;
    Slot &s = ctx->slots[slot];
    if (!shares_dev || n_shares <= 0 || s.band_r1 <= s.band_r0) { ctx->err = "sister_band_submit_share first; the gathered shares must not be null"; return SISTER_E_ARG; }
    SCK(cudaSetDevice(ctx->device));
    const Dims &d = s.dims;
    const unsigned vm = s.mode_mask == 1u ? 0xFu : s.mode_mask == 2u ? 0x3u : 0xCu;
    size_t off[9];
    const size_t share_bytes = band_share_layout(d, n_shares, off);
    for (int k = 0; k < n_shares; k++)
        for (int i = 0; i < 8; i++) {
            const int v = i >> 1, hv = view_rows(d, v), wv = view_cols(d, v);
            if (!((vm >> v) & 1u)) continue;
            const int r0 = share_row0(hv, k, n_shares), r1 = share_row0(hv, k + 1, n_shares);
            int16_t *dst = ((i & 1) ? s.d_wtaR : s.d_wtaL) + (size_t)v * d.px + (size_t)r0 * wv;
            if (r1 > r0) SCK(cudaMemcpyAsync(dst, shares_dev + (size_t)k * share_bytes + off[i], (size_t)(r1 - r0) * wv * sizeof(int16_t), cudaMemcpyDeviceToDevice, s.st));
        }
    return band_rest(ctx, s);
}

int sister_band_vertical(sister_ctx *ctx, int slot, int pass, const uint8_t *state_in_dev, uint8_t *state_out_dev)
{
    int rc = slot_ok(ctx, slot);
    if (rc) return rc;
    Slot &s = ctx->slots[slot];
    if (pass < 0 || pass > 1 || s.band_r1 <= s.band_r0) { ctx->err = "sister_band_submit first; pass is 0 or 1"; return SISTER_E_ARG; }
    SCK(cudaSetDevice(ctx->device));
    // a band that is not the first of its pass continues chains: without the neighbour's state they would silently restart
    const bool first_of_pass = pass == 0 ? s.band_r0 == 0 : s.band_r1 == s.dims.Hp;
    const bool last_of_pass = pass == 0 ? s.band_r1 == s.dims.Hp : s.band_r0 == 0;
    if (!first_of_pass && !state_in_dev) { ctx->err = "state_in_dev is NULL but the band is not the first of this pass"; return SISTER_E_ARG; }
    if (!last_of_pass && !state_out_dev) { ctx->err = "state_out_dev is NULL but the band is not the last of this pass"; return SISTER_E_ARG; }
    launch_sgm_band(1 + pass, s.d_fused - s.sgm.row_shift, s.dims, s.band_r0, s.band_r1, state_in_dev, state_out_dev, s.sgm, nullptr, nullptr, s.d_status, s.st, ctx->lc);
    SCK(cudaGetLastError());
    SCK(take_launch_error(ctx));
    return SISTER_OK;
}

int sister_band_rows(sister_ctx *ctx, int slot, int passes, const uint8_t *in_pass0, uint8_t *out_pass0, const uint8_t *in_pass1, uint8_t *out_pass1,
                     unsigned tag)
{
    int rc = slot_ok(ctx, slot);
    if (rc) return rc;
    Slot &s = ctx->slots[slot];
    if (passes < 1 || passes > 3 || s.band_r1 <= s.band_r0) { ctx->err = "sister_band_submit first; passes is 1 (pass 0), 2 (pass 1) or 3 (both)"; return SISTER_E_ARG; }
    if (tag < 1 || tag > 15) { ctx->err = "the frame tag must be 1 .. 15 and differ from the previous frame's"; return SISTER_E_ARG; }
    const bool top = s.band_r0 == 0, bottom = s.band_r1 == s.dims.Hp;
    // pass 0 flows down (in from the band above, out to the band below), pass 1 up; a missing mailbox would silently restart
    // or drop the diagonal paths at the band border
    if (((passes & 1) && ((!top && !in_pass0) || (!bottom && !out_pass0))) || ((passes & 2) && ((!bottom && !in_pass1) || (!top && !out_pass1)))) {
        ctx->err = "a mailbox is NULL although the band has a neighbour on that side";
        return SISTER_E_ARG;
    }
    SCK(cudaSetDevice(ctx->device));
    BandStream bs;
    bs.in[0] = top ? nullptr : in_pass0;     bs.out[0] = bottom ? nullptr : out_pass0;
    bs.in[2] = bottom ? nullptr : in_pass1;  bs.out[2] = top ? nullptr : out_pass1;
    bs.tag = tag;
    launch_sgm_band_rows(s.d_fused - s.sgm.row_shift, s.dims, s.band_r0, s.band_r1, ((passes & 1) ? 1u : 0u) | ((passes & 2) ? 4u : 0u), bs, s.sgm,
                         s.d_status, s.st, ctx->lc);
    SCK(cudaGetLastError());
    SCK(take_launch_error(ctx));
    return SISTER_OK;
}

int sister_band_columns(sister_ctx *ctx, int slot, int pass, const uint8_t *state_in_dev, uint8_t *state_out_dev)
{
    int rc = slot_ok(ctx, slot);
    if (rc) return rc;
    Slot &s = ctx->slots[slot];
    if (pass < 0 || pass > 1 || s.band_r1 <= s.band_r0) { ctx->err = "sister_band_submit first; pass is 0 or 1"; return SISTER_E_ARG; }
    SCK(cudaSetDevice(ctx->device));
    const bool first_of_pass = pass == 0 ? s.band_r0 == 0 : s.band_r1 == s.dims.Hp;
    const bool last_of_pass = pass == 0 ? s.band_r1 == s.dims.Hp : s.band_r0 == 0;
    if (!first_of_pass && !state_in_dev) { ctx->err = "state_in_dev is NULL but the band is not the first of this pass"; return SISTER_E_ARG; }
    if (!last_of_pass && !state_out_dev) { ctx->err = "state_out_dev is NULL but the band is not the last of this pass"; return SISTER_E_ARG; }
    // beside the row sweeps, not behind them: a stream and a mailbox of their own (the pair volumes are the slot's; the two
    // launches write different ones)
    if (!s.st_cols) {
        SCK(cudaStreamCreateWithFlags(&s.st_cols, cudaStreamNonBlocking));
        SCK(cudaEventCreateWithFlags(&s.ev_fused, cudaEventDisableTiming));
        SCK(cudaEventCreateWithFlags(&s.ev_cols, cudaEventDisableTiming));
        s.sgm_cols.mailbox_bytes = s.sgm.mailbox_bytes;
        if (cudaMalloc((void **)&s.sgm_cols.mailbox, s.sgm_cols.mailbox_bytes) != cudaSuccess) { cudaGetLastError(); ctx->err = "no memory for the column sweeps' mailbox"; return SISTER_E_NOMEM; }
        SCK(cudaEventRecord(s.ev_fused, s.st)); // (the first frame: everything enqueued on the slot's stream so far)
    }
    s.sgm_cols.vols = s.sgm.vols;
    s.sgm_cols.vol_stride = s.sgm.vol_stride;
    s.sgm_cols.row_shift = s.sgm.row_shift;
    if (!s.cols_pending) SCK(cudaStreamWaitEvent(s.st_cols, s.ev_fused, 0));
    launch_sgm_band(5 + pass, s.d_fused - s.sgm.row_shift, s.dims, s.band_r0, s.band_r1, state_in_dev, state_out_dev, s.sgm_cols, nullptr, nullptr, s.d_status,
                    s.st_cols, ctx->lc);
    SCK(cudaEventRecord(s.ev_cols, s.st_cols));
    s.cols_pending = true;
    SCK(cudaGetLastError());
    SCK(take_launch_error(ctx));
    return SISTER_OK;
}

int sister_band_columns_wait(sister_ctx *ctx, int slot)
{
    int rc = slot_ok(ctx, slot);
    if (rc) return rc;
    Slot &s = ctx->slots[slot];
    SCK(cudaSetDevice(ctx->device));
    if (s.st_cols) SCK(cudaStreamSynchronize(s.st_cols));
    return SISTER_OK;
}

int sister_band_finish(sister_ctx *ctx, int slot, uint16_t *out_dev)
{
    int rc = slot_ok(ctx, slot);
    if (rc) return rc;
    Slot &s = ctx->slots[slot];
    if (!out_dev || s.band_r1 <= s.band_r0) { ctx->err = "sister_band_submit first; out_dev must not be null"; return SISTER_E_ARG; }
    SCK(cudaSetDevice(ctx->device));
    if (s.cols_pending) { SCK(cudaStreamWaitEvent(s.st, s.ev_cols, 0)); s.cols_pending = false; }
    launch_sgm_band(3, s.d_fused - s.sgm.row_shift, s.dims, s.band_r0, s.band_r1, nullptr, nullptr, s.sgm, nullptr, out_dev, s.d_status, s.st, ctx->lc);
    SCK(cudaMemcpyAsync(s.h_status, s.d_status, sizeof(int), cudaMemcpyDeviceToHost, s.st));
    SCK(cudaGetLastError());
    s.band_r0 = s.band_r1 = 0; // sister_sync(slot) completes the band and checks the status word
    return SISTER_OK;
}

int sister_sync(sister_ctx *ctx, int slot)
{
    if (!ctx) return SISTER_E_ARG;
    SCK(cudaSetDevice(ctx->device));
    int rc = SISTER_OK;
    for (int k = 0; k < (int)ctx->slots.size(); k++) {
        if (slot >= 0 && k != slot) continue;
        Slot &s = ctx->slots[k];
        int r = finish_slot(ctx, s);
        if (r && !rc) rc = r;
    }
    return rc;
}

int sister_dev_alloc(sister_ctx *ctx, size_t bytes, void **dev_ptr)
{
    if (!ctx || !dev_ptr) return SISTER_E_ARG;
    SCK(cudaSetDevice(ctx->device));
    SCK(cudaMalloc(dev_ptr, bytes));
    return SISTER_OK;
}
int sister_dev_free(sister_ctx *ctx, void *dev_ptr)
{
    if (!ctx) return SISTER_E_ARG;
    SCK(cudaSetDevice(ctx->device));
    SCK(cudaFree(dev_ptr));
    return SISTER_OK;
}
// ---- device memory shared between the processes of one box (one process per GPU): the band mailboxes of the row sweeps
int sister_ipc_export(sister_ctx *ctx, void *dev_ptr, unsigned char handle_out[64])
{
    if (!ctx || !dev_ptr || !handle_out) return SISTER_E_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the C ABI passes the handle as 64 bytes");
    SCK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    SCK(cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle_out, &h, 64);
    return SISTER_OK;
}
int sister_ipc_open(sister_ctx *ctx, const unsigned char handle[64], void **dev_ptr)
{
    if (!ctx || !handle || !dev_ptr) return SISTER_E_ARG;
    SCK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    SCK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SISTER_OK;
}
int sister_ipc_close(sister_ctx *ctx, void *dev_ptr)
{
    if (!ctx || !dev_ptr) return SISTER_E_ARG;
    SCK(cudaSetDevice(ctx->device));
    SCK(cudaIpcCloseMemHandle(dev_ptr));
    return SISTER_OK;
}
int sister_dev_memset(sister_ctx *ctx, void *dev_ptr, int value, size_t bytes)
{
    if (!ctx || !dev_ptr) return SISTER_E_ARG;
    SCK(cudaSetDevice(ctx->device));
    SCK(cudaMemset(dev_ptr, value, bytes));
    return SISTER_OK;
}
int sister_host_alloc(sister_ctx *ctx, size_t bytes, void **host_ptr)
{
    if (!ctx || !host_ptr) return SISTER_E_ARG;
    SCK(cudaSetDevice(ctx->device));
    SCK(cudaHostAlloc(host_ptr, bytes, cudaHostAllocDefault));
    return SISTER_OK;
}
int sister_host_free(sister_ctx *ctx, void *host_ptr)
{
    if (!ctx) return SISTER_E_ARG;
    SCK(cudaSetDevice(ctx->device));
    SCK(cudaFreeHost(host_ptr));
    return SISTER_OK;
}
int sister_dev_upload(sister_ctx *ctx, void *dev_dst, const void *host_src, size_t bytes)
{
    if (!ctx) return SISTER_E_ARG;
    SCK(cudaSetDevice(ctx->device));
    SCK(cudaMemcpy(dev_dst, host_src, bytes, cudaMemcpyHostToDevice));
    return SISTER_OK;
}
int sister_dev_download(sister_ctx *ctx, void *host_dst, const void *dev_src, size_t bytes)
{
    if (!ctx) return SISTER_E_ARG;
    SCK(cudaSetDevice(ctx->device));
    SCK(cudaMemcpy(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost));
    return SISTER_OK;
}

int sister_set_full_frame(sister_ctx *ctx, int enabled)
{
    if (!ctx) return SISTER_E_ARG;
    ctx->full_frame = enabled != 0;
    return SISTER_OK;
}

int sister_set_test_taps(sister_ctx *ctx, int enabled)
{
    if (!ctx) return SISTER_E_ARG;
    SCK(cudaSetDevice(ctx->device));
    for (auto &s : ctx->slots) {
        SCK(cudaStreamSynchronize(s.st));
        if (enabled && !s.d_sum) SCK(cudaMalloc((void **)&s.d_sum, (size_t)ctx->cells_max * 2));
        if (!enabled && s.d_sum) { SCK(cudaFree(s.d_sum)); s.d_sum = nullptr; }
    }
    ctx->taps = enabled != 0;
    return SISTER_OK;
}

int sister_set_profiling(sister_ctx *ctx, int enabled)
{
    if (!ctx) return SISTER_E_ARG;
    ctx->profiling = enabled != 0;
    return SISTER_OK;
}

int sister_region_begin(sister_ctx *ctx)
{
    if (!ctx) return SISTER_E_ARG;
    SCK(cudaSetDevice(ctx->device));
    if (!ctx->region_b) {
        SCK(cudaEventCreate(&ctx->region_b));
        SCK(cudaEventCreate(&ctx->region_e));
        ctx->region_join.resize(ctx->slots.size());
        for (auto &e : ctx->region_join) SCK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    for (auto &s : ctx->slots) SCK(cudaStreamSynchronize(s.st));
    SCK(cudaEventRecord(ctx->region_b, ctx->slots[0].st));
    for (size_t k = 1; k < ctx->slots.size(); k++) SCK(cudaStreamWaitEvent(ctx->slots[k].st, ctx->region_b, 0));
    return SISTER_OK;
}

int sister_region_end(sister_ctx *ctx, float *elapsed_ms)
{
    if (!ctx || !elapsed_ms || !ctx->region_b) return SISTER_E_ARG;
    SCK(cudaSetDevice(ctx->device));
    for (size_t k = 1; k < ctx->slots.size(); k++) {
        SCK(cudaEventRecord(ctx->region_join[k], ctx->slots[k].st));
        SCK(cudaStreamWaitEvent(ctx->slots[0].st, ctx->region_join[k], 0));
    }
    SCK(cudaEventRecord(ctx->region_e, ctx->slots[0].st));
    SCK(cudaEventSynchronize(ctx->region_e));
    SCK(cudaEventElapsedTime(elapsed_ms, ctx->region_b, ctx->region_e));
    return SISTER_OK;
}

int sister_get_stage_ms(sister_ctx *ctx, int slot, float *ms, int n)
{
    int rc = slot_ok(ctx, slot);
    if (rc) return rc;
    if (!ms || n < SISTER_STAGE_COUNT) return SISTER_E_ARG;
    Slot &s = ctx->slots[slot];
    SCK(cudaSetDevice(ctx->device));
    SCK(cudaStreamSynchronize(s.st));
    for (int k = 0; k < n; k++) ms[k] = 0.f;
    for (int k = 0; k < s.n_ev; k++) {
        float t = 0.f;
        SCK(cudaEventElapsedTime(&t, s.ev_b[k], s.ev_e[k]));
        ms[s.ev_stage[k]] += t;
    }
    return SISTER_OK;
}

int sister_get_stage_launches(sister_ctx *ctx, int slot, int *count, int n)
{
    int rc = slot_ok(ctx, slot);
    if (rc) return rc;
    if (!count || n < SISTER_STAGE_COUNT) return SISTER_E_ARG;
    for (int k = 0; k < SISTER_STAGE_COUNT; k++) count[k] = ctx->slots[slot].stage_launches[k];
    return SISTER_OK;
}

uint64_t sister_get_launch_count(sister_ctx *ctx) { return ctx ? ctx->lc.total : 0; }

// the fused volume is stored in the path kernel's cell order (Dims, common.cuh); taps and test volumes are in natural order
static void reorder_cells(const Dims &d, uint8_t *buf, size_t ncells, bool to_natural)
{
    if (!d.interleaved) return;
    std::vector<int> disp_of(d.D);
    for (int p = 0; p < d.D; p++) disp_of[p] = cell_disp(d, p);
    std::vector<uint8_t> tmp(d.D);
    for (size_t c = 0; c < ncells; c++) {
        uint8_t *cell = buf + c * d.D;
        if (to_natural) { for (int p = 0; p < d.D; p++) tmp[disp_of[p]] = cell[p]; }
        else { for (int p = 0; p < d.D; p++) tmp[p] = cell[disp_of[p]]; }
        memcpy(cell, tmp.data(), d.D);
    }
}

int sister_debug_fetch(sister_ctx *ctx, int slot, int what, void *host_dst, size_t bytes)
{
    int rc = slot_ok(ctx, slot);
    if (rc) return rc;
    if (!host_dst) return SISTER_E_ARG;
    Slot &s = ctx->slots[slot];
    SCK(cudaSetDevice(ctx->device));
    SCK(cudaStreamSynchronize(s.st));
    const size_t px = (size_t)s.dims.px, cells = (size_t)s.dims.cells;
    const void *src = nullptr;
    size_t have = 0;
    switch (what) {
    case SISTER_TAP_ORIENTED: src = s.d_oriented; have = 8 * px; break;
    case SISTER_TAP_CENSUS: src = s.d_census; have = 8 * px * 8; break;
    case SISTER_TAP_WTA_L: src = s.d_wtaL; have = 4 * px * 2; break;
    case SISTER_TAP_WTA_R: src = s.d_wtaR; have = 4 * px * 2; break;
    case SISTER_TAP_LR_FINAL: src = s.d_lr; have = 4 * px * 2; break;
    case SISTER_TAP_MASKS: src = s.d_masks; have = 4 * px; break;
    case SISTER_TAP_FUSED: src = s.last_fused ? s.last_fused : s.d_fused; have = cells; break;
    case SISTER_TAP_SUM:
        if (!ctx->taps || !s.d_sum) { ctx->err = "SISTER_TAP_SUM needs sister_set_test_taps(ctx, 1) before the submit"; return SISTER_E_ARG; }
        src = s.d_sum; have = cells * 2; break;
    case SISTER_TAP_RAW_DISP:
        if (!s.full_frame) { ctx->err = "SISTER_TAP_RAW_DISP needs sister_set_full_frame(ctx, 1) or the test taps before the submit"; return SISTER_E_ARG; }
        src = s.d_raw; have = 3 * px * 2; break;
    default: ctx->err = "unknown tap"; return SISTER_E_ARG;
    }
    if (bytes > have) { ctx->err = "tap smaller than requested"; return SISTER_E_ARG; }
    SCK(cudaMemcpy(host_dst, src, bytes, cudaMemcpyDeviceToHost));
    if (what == SISTER_TAP_FUSED) reorder_cells(s.dims, static_cast<uint8_t *>(host_dst), bytes / (size_t)s.dims.D, true);
    return SISTER_OK;
}

// The two-view path of the reference (doStereo, hpp:122-150): AD-census cost of (center, side), SGM on that raw volume,
// WTA left and right on the aggregated volume, in-place median on both maps, LRC. No padding, no crop: the frame is the
// image. Every kernel is one of the 5-view path's, run for view 0 ("right", rotation 0) only.
int sister_stereo(sister_ctx *ctx, const uint8_t *center, const uint8_t *side, int w, int h, size_t row_stride, int disp_count,
                  float *out_left, float *out_right)
{
    if (!ctx || !center || !side || !out_left) return SISTER_E_ARG;
    if (w <= 0 || h <= 0 || disp_count <= 0 || row_stride < (size_t)w) { ctx->err = "bad stereo arguments"; return SISTER_E_ARG; }
    if (disp_count % 8 != 0) { ctx->err = "disp_count must be a multiple of 8 (sgm.cpp:268)"; return SISTER_E_SHAPE; }
    if (disp_count > 512) { ctx->err = "disp_count above 512 is not supported"; return SISTER_E_SHAPE; }
    if (w % 4 != 0 || h % 4 != 0) { ctx->err = "w and h must be multiples of 4 (postprocess.cpp:18)"; return SISTER_E_SHAPE; }
    if (w < 16 || h < 16) { ctx->err = "frame too small for the 9x7 census"; return SISTER_E_SHAPE; }
    Dims d;
    d.W = 0; d.H = 0; d.D = disp_count; d.Wp = w; d.Hp = h;
    d.px = (long long)w * h;
    d.cells = d.px * disp_count;
    set_cell_order(d);
    if (d.cells / 8 > 0x7F000000LL) { ctx->err = "cost volume above 1.7e10 cells (32-bit cursor offsets in sgm.cu)"; return SISTER_E_SHAPE; }
    if (d.px > ctx->px_max || d.cells > ctx->cells_max || (size_t)2 * d.px > ctx->in_bytes_max) {
        ctx->err = "stereo pair larger than the capacity given to sister_create";
        return SISTER_E_CAPACITY;
    }
    Slot &s = ctx->slots[0];
    if (s.busy) { ctx->err = "slot busy"; return SISTER_E_BUSY; }
    SCK(cudaSetDevice(ctx->device));
    const size_t px = (size_t)d.px;
    for (int k = 0; k < 2; k++) {
        const uint8_t *src = k ? side : center;
        uint8_t *dst = s.h_in + k * px;
        if (row_stride == (size_t)w) memcpy(dst, src, px);
        else for (int i = 0; i < h; i++) memcpy(dst + (size_t)i * w, src + (size_t)i * row_stride, (size_t)w);
    }
    // oriented images 0 / 1 = the pair as it is (view 0 is the identity orientation, hpp:56-58)
    SCK(cudaMemcpyAsync(s.d_oriented, s.h_in, 2 * px, cudaMemcpyHostToDevice, s.st));
    SCK(cudaMemsetAsync(s.d_status, 0, sizeof(int), s.st));
    if (!s.d_sum) SCK(cudaMalloc((void **)&s.d_sum, (size_t)ctx->cells_max * 2));
    ctx->lc.cur_stage = SISTER_STAGE_CENSUS;
    launch_census(s.d_oriented, d, s.d_census, s.st, ctx->lc);
    // the raw volume of ad_census (census.cpp:54-146, 255 markers and all) is the "fused" volume of view 0 under an
    // all-ones mask: k_fuse evaluates the literal per-cell formula wherever the fast path does not apply
    ctx->lc.cur_stage = SISTER_STAGE_FUSE;
    SCK(cudaMemsetAsync(s.d_masks, 1, px, s.st));
    launch_fuse(s.d_census, s.d_masks, d, 0x1u, s.d_fused, s.d_status, s.st, ctx->lc);
    s.last_fused = s.d_fused;
    // SGM (hpp:135) + WTA-left (hpp:137) in the final sweep; the aggregated volume is kept for WTA-right (hpp:138)
    ctx->lc.cur_stage = SISTER_STAGE_AGGREGATE;
    launch_sgm(s.d_fused, d, true, s.sgm, s.d_sum, s.d_wtaL, nullptr, s.d_status, s.st, ctx->lc);
    launch_wta_right_sum(s.d_sum, d, s.d_wtaR, s.st, ctx->lc);
    // median on both maps (hpp:139-140), LRC (hpp:143)
    ctx->lc.cur_stage = SISTER_STAGE_MASK;
    launch_median_lrc_mask(s.d_wtaL, s.d_wtaR, d, 0x1u, s.d_medL, s.d_medR, s.d_lr, s.d_masks, s.d_status, s.st, ctx->lc);
    SCK(cudaMemcpyAsync(s.h_status, s.d_status, sizeof(int), cudaMemcpyDeviceToHost, s.st));
    SCK(cudaGetLastError());
    SCK(cudaStreamSynchronize(s.st));
    SCK(cudaMemsetAsync(s.d_masks, 0, px, s.st)); // the 5-view path expects untouched masks to be 0
    if (!ctx->taps) { SCK(cudaFree(s.d_sum)); s.d_sum = nullptr; }
    s.dims = d;
    s.full_frame = true;
    const int bits = *s.h_status; // the word accumulates (run_pipeline): leave it clean for the next submit on this slot
    *s.h_status = 0;
    SCK(cudaMemsetAsync(s.d_status, 0, sizeof(int), s.st));
    if (bits & ~kStatusFusedOverflow) { ctx->err = "kernel invariant violated"; return SISTER_E_INTERNAL; }
    std::vector<int16_t> tmp(px);
    SCK(cudaMemcpy(tmp.data(), s.d_lr, px * 2, cudaMemcpyDeviceToHost));
    for (size_t k = 0; k < px; k++) out_left[k] = (float)tmp[k];
    if (out_right) {
        SCK(cudaMemcpy(tmp.data(), s.d_medR, px * 2, cudaMemcpyDeviceToHost));
        for (size_t k = 0; k < px; k++) out_right[k] = (float)tmp[k];
    }
    return SISTER_OK;
}

int sister_test_sgm(sister_ctx *ctx, const uint8_t *fused, int w, int h, int disp_count, uint16_t *sum, int16_t *disp)
{
    if (!ctx || !fused || !sum) return SISTER_E_ARG;
    if (disp_count <= 0 || disp_count % 8 != 0 || disp_count > 512 || w < 12 || h < 12) { ctx->err = "bad sgm test shape"; return SISTER_E_SHAPE; }
    Dims d;
    d.W = 0; d.H = 0; d.D = disp_count; d.Wp = w; d.Hp = h;
    d.px = (long long)w * h;
    d.cells = d.px * disp_count;
    set_cell_order(d);
    if (d.px > ctx->px_max || d.cells > ctx->cells_max) { ctx->err = "sgm test volume exceeds capacity"; return SISTER_E_CAPACITY; }
    Slot &s = ctx->slots[0];
    if (s.busy) return SISTER_E_BUSY;
    SCK(cudaSetDevice(ctx->device));
    if (d.interleaved) {
        std::vector<uint8_t> ordered(fused, fused + (size_t)d.cells);
        reorder_cells(d, ordered.data(), (size_t)d.px, false);
        SCK(cudaMemcpy(s.d_fused, ordered.data(), (size_t)d.cells, cudaMemcpyHostToDevice));
    } else {
        SCK(cudaMemcpyAsync(s.d_fused, fused, (size_t)d.cells, cudaMemcpyHostToDevice, s.st));
    }
    SCK(cudaMemsetAsync(s.d_status, 0, sizeof(int), s.st));
    if (!s.d_sum) SCK(cudaMalloc((void **)&s.d_sum, (size_t)ctx->cells_max * 2));
    ctx->lc.cur_stage = SISTER_STAGE_AGGREGATE;
    launch_sgm(s.d_fused, d, true, s.sgm, s.d_sum, s.d_raw, nullptr, s.d_status, s.st, ctx->lc);
    SCK(cudaMemcpyAsync(s.h_status, s.d_status, sizeof(int), cudaMemcpyDeviceToHost, s.st));
    SCK(cudaGetLastError());
    SCK(cudaStreamSynchronize(s.st));
    SCK(cudaMemcpy(sum, s.d_sum, (size_t)d.cells * 2, cudaMemcpyDeviceToHost));
    if (disp) SCK(cudaMemcpy(disp, s.d_raw, (size_t)d.px * 2, cudaMemcpyDeviceToHost));
    if (!ctx->taps) { SCK(cudaFree(s.d_sum)); s.d_sum = nullptr; }
    s.dims = d;
    const int bits = *s.h_status;
    *s.h_status = 0;
    SCK(cudaMemsetAsync(s.d_status, 0, sizeof(int), s.st));
    return bits ? SISTER_E_INTERNAL : SISTER_OK;
}

} // extern "C"
