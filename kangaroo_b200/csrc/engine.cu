// Fused per-frame engine: the whole path of applications/stereo2/main.cpp:375-454
//   Census x2 -> [right-reference WTA on raw costs] -> Hamming cost -> SGM sweeps -> WTA/parabola
//   (epilogue of the last sweep) -> LeftRightCheck x2
// on engine-owned scratch in the internal disparity-innermost layout, for a batch of stereo pairs
// per launch.  Host side is plain C++ behind the C ABI of include/roo_b200.h.
#include <new>
#include <thread>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"
#include "sgm_step.cuh"

namespace roo_b200 {
std::atomic<unsigned long long> g_launches{0};
std::atomic<int> g_ieee_div{0};
std::atomic<int> g_insweep_cost{1};
}  // namespace roo_b200

using namespace roo_b200;

struct roo_engine {
    roo_pipeline_params_t p;
    int device = 0;
    int DP = 0, words = 0;
    int ieee = 0;          // fp mode of THIS engine (roo_pipeline_params_t.fp_mode; resolved at creation)
    size_t npx = 0;
    // scratch (device)
    unsigned long long* cen[2] = {nullptr, nullptr};  // [batch][h][w][words]
    unsigned long long* cen_base[2] = {nullptr, nullptr};   // the allocations: cen[] sits CEN_PAD elements inside (in-sweep
                                                            // cost reads strips that stick out of a row by < 2 * 256 + 32 pixels)
    unsigned char* c8 = nullptr;                      // [batch][h][w][DP]
    float* H = nullptr;                               // [batch][h][w][DP]
    float* dispR = nullptr;                           // [batch][h][w]
    float* med = nullptr;                             // [batch][h][w] output of the median stage / FilterDispGrad snapshot
    float* imgf = nullptr;                            // [batch][h][w] adaptive-P2 intensity (u8 * img_scale)
    float* edge = nullptr;                            // fused vertical groups: band-to-band state rows
    int* flags = nullptr;                             //                        and their progress flags
    // optional front end (roo_engine_set_front_end): raw frames of (w << level) x (h << level) -> [Warp] -> BoxHalf x level
    int fe_level = 0;
    bool fe_rectify = false;
    roo_image_t fe_lut[2] = {};                       // float2 lookup tables (caller-owned device memory)
    unsigned char* fe_rect[2] = {nullptr, nullptr};   // [batch][raw h][raw w] rectified frames
    unsigned char* fe_pyr[2] = {nullptr, nullptr};    // pyramid levels 1..level, one after the other, each [batch][h_l][w_l]
    size_t in_npx = 0;                                // pixels of one INPUT frame (= npx without a front end)
    SgmPlan plan{};                                   // fused vertical groups where possible
    SgmPlan plan_sep{};                               // one pass per path
    int n_bands = 0;
    // staging for run_host (device) and its streams
    unsigned char* in_dev[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [buffer][side]
    float* out_dev[2] = {nullptr, nullptr};
    cudaStream_t s_compute = nullptr, s_in = nullptr, s_out = nullptr;
    bool host_ready = false;   // streams, events and staging buffers below all exist
    long long ticket = 0;   // groups submitted through the host-buffer path so far
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    size_t scratch_bytes = 0;
    int last_batch = 0;
    // optional per-kernel timing (roo_engine_set_profiling): one event after every launch
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;
    std::vector<int> prof_kinds;   // kind of the launch that ENDS at event i (-1 = group start marker)
    std::vector<int> prof_pass;    // aggregation pass index of that launch (-1: not an aggregation pass)
    size_t prof_used = 0;
    double prof_ms[ROO_PROF_KINDS] = {0};
    long long prof_n[ROO_PROF_KINDS] = {0};
};

// fuse_vertical == 0 (auto): minimum number of (band, pair) CTAs per launch for the fused passes to be chosen
constexpr long long FUSE_MIN_CTAS = 100;
// ... unless the group is so small that single-path sweeps are not bandwidth-bound either (pixel*disparity units)
constexpr long long FUSE_MIN_UNITS = 100000000;
constexpr size_t CEN_PAD = 1024;   // elements of padding before and after the census arrays

static void prof_mark(roo_engine* e, int kind, cudaStream_t st, int pass = -1) {
    if (!e->profiling) return;
    if (e->prof_used == e->prof_events.size()) {
        cudaEvent_t ev;
        if (cudaEventCreate(&ev) != cudaSuccess) return;
        e->prof_events.push_back(ev);
        e->prof_kinds.push_back(kind);
        e->prof_pass.push_back(pass);
    }
    e->prof_kinds[e->prof_used] = kind;
    e->prof_pass[e->prof_used] = pass;
    cudaEventRecord(e->prof_events[e->prof_used++], st);
}

static int engine_group(roo_engine* e, const unsigned char* left, const unsigned char* right, float* disp, int batch,
                        cudaStream_t st) {
    const roo_pipeline_params_t& p = e->p;
    const int w = p.w, h = p.h;
    const size_t npx = e->npx;
    int rc;
    prof_mark(e, -1, st);
    // front end (stereo2/main.cpp:360-375): Warp(rectify) -> BoxReduce down to the working level
    if (e->fe_rectify || e->fe_level > 0) {
        const unsigned char* src[2] = {left, right};
        for (int side = 0; side < 2; ++side) {
            int cw = w << e->fe_level, ch = h << e->fe_level;
            if (e->fe_rectify) {
                rc = launch_warp_u8(e->fe_rect[side], src[side], cw, ch, batch, e->fe_lut[side], st);
                if (rc) return rc;
                src[side] = e->fe_rect[side];
            }
            unsigned char* dst = e->fe_pyr[side];
            for (int l = 1; l <= e->fe_level; ++l) {
                rc = launch_box_half_u8(dst, src[side], cw >> 1, ch >> 1, cw, ch, batch, st);
                if (rc) return rc;
                cw >>= 1; ch >>= 1;
                src[side] = dst;
                dst += (size_t)e->p.max_batch * cw * ch;
            }
        }
        left = src[0]; right = src[1];
        prof_mark(e, ROO_PROF_CENSUS, st);
    }
    for (int side = 0; side < 2; ++side) {
        rc = launch_census((char*)e->cen[side], (size_t)w * e->words * 8, npx * e->words * 8,
                           (const char*)(side == 0 ? left : right), (size_t)w, npx, w, h, batch, p.window, ROO_IMG_U8, st);
        if (rc) return rc;
        prof_mark(e, ROO_PROF_CENSUS, st);
    }
    if (p.lrcheck) {
        rc = launch_census_wta(e->dispR, e->cen[1], e->cen[0], w, h, batch, p.max_disp, e->words, p.popc_mode, p.subpix, +1, e->ieee, st);
        if (rc) return rc;
        prof_mark(e, ROO_PROF_WTA, st);
    }
    // Plan per group.  A fused pass walks the rows of a band serially (~1.1 ms per pass at 720 rows whatever the
    // batch) and pays off once there are enough (band, pair) CTAs to fill the GPU; a single pair is faster with one
    // pass per path (1280x720x128, 8 paths: 1 pair 1.8 ms vs 3.0 ms).
    // Measured on B200, 8 paths, fused / separate: 640x480x64 x1 1.8 / 1.9 ms (sweeps of so few columns are
    // latency-bound too), 1280x720x128 x2 3.3 / 3.2, x3 3.7 / 4.6, 1920x1080x256 x1 6.4 / 7.4, 3840x2160x256 x1 17.9 / 27.1.
    const long long units = (long long)w * h * e->DP * batch;
    const bool use_fused = e->DP <= ROO_MAX_DISP_FUSED && (p.fuse_vertical > 0 ||
                           (p.fuse_vertical == 0 && ((long long)batch * e->n_bands >= FUSE_MIN_CTAS || units < FUSE_MIN_UNITS)));
    const SgmPlan& plan = use_fused ? e->plan : e->plan_sep;
    const int ndir = plan.n;
    if (ndir == 0) {
        rc = launch_census_wta(disp, e->cen[0], e->cen[1], w, h, batch, p.max_disp, e->words, p.popc_mode, p.subpix, -1, e->ieee, st);
        if (rc) return rc;
        prof_mark(e, ROO_PROF_WTA, st);
    } else {
        // In-sweep matching cost (COST_CEN32): a pass that can recompute popc(L ^ R) from the census words does not
        // read the u8 cost volume; when every pass of the plan can, the volume is never built (north_star (3)).
        // Eligible: one-word descriptors under the reference's 32-bit popcount (9x7 window, hamming_distance.h:40-44).
        const bool cen_ok = e->words == 1 && p.popc_mode == ROO_POPC32_COMPAT && g_insweep_cost.load(std::memory_order_relaxed);
        auto pass_in_sweep = [&](const SgmPass& ps) {
            // The fused vertical groups and the bulk-copy horizontal kernel always can.  The generic single-path sweep stages
            // the 32*DPL census words of every step with 4-byte cp.async: measured faster than the u8 volume only up to 64
            // disparities (640x480x64 4-path: 5967 -> 6973 pairs/s; 1242x375x128: 2989 -> 2645, so it keeps the volume there).
            if (!cen_ok) return false;
            if (ps.fused || (ps.dy == 0 && g_use_hsweep.load(std::memory_order_relaxed))) return true;
            return e->DP <= 64;
        };
        bool need_c8 = false;
        for (int i = 0; i < ndir; ++i) need_c8 |= !pass_in_sweep(plan.pass[i]);
        if (need_c8) {
            rc = launch_cost_u8(e->c8, e->cen[0], e->cen[1], w, h, batch, e->DP, p.max_disp, e->words, p.popc_mode, st);
            if (rc) return rc;
        }
        rc = launch_image_to_f32(e->imgf, left, (size_t)w, npx, ROO_IMG_U8, w, h, batch, p.img_scale, st);
        if (rc) return rc;
        prof_mark(e, ROO_PROF_COST, st);
        SweepArgs a{};
        a.H = e->H; a.h_pair = npx * e->DP; a.C = e->c8; a.c_pair = npx * e->DP;
        a.img = e->imgf; a.img_pair = npx; a.cost_scale = 1.0f / (float)(e->words * 64);
        a.w = w; a.h = h; a.DP = e->DP; a.maxDisp = p.max_disp; a.batch = batch;
        a.P1 = p.P1; a.P2 = p.P2; a.ieee = e->ieee; a.subpix = p.subpix;
        a.cenL = e->cen[0]; a.cenR = e->cen[1]; a.cen_pair = npx; a.disp = disp; a.disp_pair = npx;
        for (int i = 0; i < ndir; ++i) {
            a.first = i == 0;
            a.cost_kind = pass_in_sweep(plan.pass[i]) ? COST_CEN32 : COST_U8;
            a.epi = i + 1 < ndir ? EPI_NONE : (p.keep_volume ? EPI_WTA_WRITE : EPI_WTA_ONLY);
            rc = launch_pass(a, plan.pass[i], e->edge, e->flags, st);
            if (rc) return rc;
            prof_mark(e, plan.pass[i].fused ? ROO_PROF_VGROUP : ROO_PROF_SWEEP, st, i);
        }
    }
    // MedianFilterRejectNegativeNxN(disp[di], disp[di], maxbad) x iters on every disparity image (main.cpp:438-444),
    // out of place into scratch and copied back
    if (p.median_size > 0) {
        float* imgs[2] = {disp, p.lrcheck ? e->dispR : nullptr};
        for (int di = 0; di < 2 && imgs[di]; ++di)
            for (int it = 0; it < p.median_iters; ++it) {
                rc = launch_median(e->med, (size_t)w * 4, npx * 4, imgs[di], (size_t)w * 4, npx * 4, w, h, batch, p.median_size,
                                   p.median_maxbad, st);
                if (rc) return rc;
                ROO_CUDA_TRY(cudaMemcpyAsync(imgs[di], e->med, (size_t)batch * npx * 4, cudaMemcpyDeviceToDevice, st));
                prof_mark(e, ROO_PROF_WTA, st);
            }
    }
    if (p.lrcheck) {
        // LeftRightCheck(disp[1], disp[0], +1, maxdiff); LeftRightCheck(disp[0], disp[1], -1, maxdiff) (main.cpp:451-454)
        rc = launch_lr_check_f32(e->dispR, (size_t)w * 4, disp, (size_t)w * 4, w, h, batch, npx * 4, npx * 4, +1.0f, p.lr_maxdiff, st);
        if (rc) return rc;
        prof_mark(e, ROO_PROF_LRCHECK, st);
        rc = launch_lr_check_f32(disp, (size_t)w * 4, e->dispR, (size_t)w * 4, w, h, batch, npx * 4, npx * 4, -1.0f, p.lr_maxdiff, st);
        if (rc) return rc;
        prof_mark(e, ROO_PROF_LRCHECK, st);
    }
    // FilterDispGrad(disp[0], disp[0], filtgradthresh) (main.cpp:456-458): gradient of a snapshot of the disparities
    if (p.filtgrad_threshold > 0.0f) {
        ROO_CUDA_TRY(cudaMemcpyAsync(e->med, disp, (size_t)batch * npx * 4, cudaMemcpyDeviceToDevice, st));
        const roo_image_t o{(size_t)w * 4, disp, (size_t)w, (size_t)h}, g{(size_t)w * 4, e->med, (size_t)w, (size_t)h};
        rc = launch_filter_disp_grad(o, g, g, p.filtgrad_threshold, st, batch, npx * 4, npx * 4, npx * 4);
        if (rc) return rc;
        prof_mark(e, ROO_PROF_LRCHECK, st);
    }
    e->last_batch = batch;
    return ROO_OK;
}

static void host_streams_free(roo_engine* e);

static void engine_free(roo_engine* e) {
    cudaFree(e->cen_base[0]); cudaFree(e->cen_base[1]); cudaFree(e->c8); cudaFree(e->H); cudaFree(e->dispR); cudaFree(e->med); cudaFree(e->imgf);
    for (int sd = 0; sd < 2; ++sd) { cudaFree(e->fe_rect[sd]); cudaFree(e->fe_pyr[sd]); } cudaFree(e->edge); cudaFree(e->flags);
    for (cudaEvent_t ev : e->prof_events) cudaEventDestroy(ev);
    host_streams_free(e);
}

extern "C" int roo_engine_create(roo_engine_t** out, const roo_pipeline_params_t* params) {
    if (!out || !params) return ROO_ERR_INVALID_ARGUMENT;
    const roo_pipeline_params_t& p = *params;
    if (p.w <= 0 || p.h <= 0 || p.max_disp <= 0 || p.max_batch <= 0 || p.window < 0 || p.window > 2)
        return ROO_ERR_INVALID_ARGUMENT;
    if (p.max_disp > ROO_MAX_DISP) return ROO_ERR_UNSUPPORTED;
    if (p.fp_mode < ROO_FP_DEFAULT || p.fp_mode > ROO_FP_IEEE) return ROO_ERR_INVALID_ARGUMENT;
    if (p.median_size != 0 && p.median_size != 5 && p.median_size != 7 && p.median_size != 9) return ROO_ERR_UNSUPPORTED;
    if (p.median_size != 0 && p.median_iters < 0) return ROO_ERR_INVALID_ARGUMENT;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return ROO_ERR_NO_DEVICE;
    roo_engine* e = new (std::nothrow) roo_engine();
    if (!e) return ROO_ERR_OUT_OF_MEMORY;
    e->p = p;
    cudaGetDevice(&e->device);
    e->ieee = p.fp_mode == ROO_FP_DEFAULT ? (g_ieee_div.load() != 0) : (p.fp_mode == ROO_FP_IEEE);
    e->DP = disp_padded(p.max_disp);
    e->words = p.window == ROO_WIN_9x7 ? 1 : (p.window == ROO_WIN_11x11 ? 2 : 4);
    e->npx = (size_t)p.w * p.h;
    e->in_npx = e->npx;
    const size_t B = (size_t)p.max_batch, npx = e->npx;
    auto alloc = [&](void** ptr, size_t bytes) -> bool {
        if (cudaMalloc(ptr, bytes) != cudaSuccess) return false;
        e->scratch_bytes += bytes;
        return true;
    };
    e->plan = sgm_plan(p.dohoriz, p.dovert, p.doreverse, p.dodiag, (p.fuse_vertical >= 0 && e->DP <= ROO_MAX_DISP_FUSED) ? 1 : 0);
    e->plan_sep = sgm_plan(p.dohoriz, p.dovert, p.doreverse, p.dodiag, 0);
    e->n_bands = vgroup_bands(p.w, p.h, e->DP);
    const int ndir = e->plan.n;
    bool fused = false;
    for (int i = 0; i < ndir; ++i) fused |= e->plan.pass[i].fused != 0;
    bool ok = alloc((void**)&e->cen_base[0], (B * npx * e->words + 2 * CEN_PAD) * 8) &&
              alloc((void**)&e->cen_base[1], (B * npx * e->words + 2 * CEN_PAD) * 8);
    if (ok) {
        for (int sd = 0; sd < 2; ++sd) {
            e->cen[sd] = e->cen_base[sd] + CEN_PAD;
            cudaMemset(e->cen_base[sd], 0, CEN_PAD * 8);
            cudaMemset(e->cen[sd] + B * npx * e->words, 0, CEN_PAD * 8);
        }
    }
    if (ok && ndir > 0)
        ok = alloc((void**)&e->c8, B * npx * e->DP) && alloc((void**)&e->H, B * npx * e->DP * 4) &&
             alloc((void**)&e->imgf, B * npx * 4);
    if (ok && p.lrcheck) ok = alloc((void**)&e->dispR, B * npx * 4);
    if (ok && (p.median_size > 0 || p.filtgrad_threshold > 0.0f)) ok = alloc((void**)&e->med, B * npx * 4);
    if (ok && fused)
        ok = alloc((void**)&e->edge, B * vgroup_edge_floats(p.w, p.h, e->DP) * 4) &&
             alloc((void**)&e->flags, B * (size_t)vgroup_bands(p.w, p.h, e->DP) * 4 + 256);   // + debug counters (VG_TIMING builds)
    if (!ok) {
        cudaGetLastError();
        engine_free(e);
        delete e;
        return ROO_ERR_OUT_OF_MEMORY;
    }
    *out = e;
    return ROO_OK;
}

extern "C" int roo_engine_destroy(roo_engine_t* e) {
    if (!e) return ROO_ERR_INVALID_ARGUMENT;
    cudaDeviceSynchronize();
    engine_free(e);
    delete e;
    return ROO_OK;
}

// the engine's scratch lives on the device that was current at creation: calls from another device are refused
static bool on_engine_device(const roo_engine* e) {
    int dev = -1;
    return cudaGetDevice(&dev) == cudaSuccess && dev == e->device;
}

extern "C" int roo_engine_set_front_end(roo_engine_t* e, int level, const roo_image_t* lookup_left,
                                        const roo_image_t* lookup_right) {
    if (!e || level < 0 || level > 8 || (lookup_left == nullptr) != (lookup_right == nullptr) || !on_engine_device(e))
        return ROO_ERR_INVALID_ARGUMENT;
    if (e->host_ready || e->fe_rectify || e->fe_level > 0) return ROO_ERR_INVALID_ARGUMENT;   // once, before the first host run
    const size_t rw = (size_t)e->p.w << level, rh = (size_t)e->p.h << level, B = (size_t)e->p.max_batch;
    if (lookup_left) {
        const roo_image_t* lut[2] = {lookup_left, lookup_right};
        for (int sd = 0; sd < 2; ++sd) {
            if (!valid_image(lut[sd], 8) || lut[sd]->w < rw || lut[sd]->h < rh || (((uintptr_t)lut[sd]->ptr | lut[sd]->pitch) & 7))
                return ROO_ERR_INVALID_ARGUMENT;
            e->fe_lut[sd] = *lut[sd];
        }
    }
    size_t pyr = 0;
    for (int l = 1; l <= level; ++l) pyr += B * (rw >> l) * (rh >> l);
    for (int sd = 0; sd < 2; ++sd) {
        if (lookup_left && cudaMalloc((void**)&e->fe_rect[sd], B * rw * rh) != cudaSuccess) return ROO_ERR_OUT_OF_MEMORY;
        if (pyr && cudaMalloc((void**)&e->fe_pyr[sd], pyr) != cudaSuccess) return ROO_ERR_OUT_OF_MEMORY;
        e->scratch_bytes += (lookup_left ? B * rw * rh : 0) + pyr;
    }
    e->fe_level = level;
    e->fe_rectify = lookup_left != nullptr;
    e->in_npx = rw * rh;
    return ROO_OK;
}

extern "C" size_t roo_engine_scratch_bytes(const roo_engine_t* e) { return e ? e->scratch_bytes : 0; }

extern "C" int roo_engine_run_device(roo_engine_t* e, const uint8_t* left, const uint8_t* right, float* disp, int n_pairs,
                                     void* stream) {
    if (!e || !left || !right || !disp || n_pairs < 0 || !on_engine_device(e)) return ROO_ERR_INVALID_ARGUMENT;
    cudaStream_t st = as_stream(stream);
    for (int g = 0; g < n_pairs; g += e->p.max_batch) {
        const int batch = n_pairs - g < e->p.max_batch ? n_pairs - g : e->p.max_batch;
        const int rc = engine_group(e, left + (size_t)g * e->in_npx, right + (size_t)g * e->in_npx, disp + (size_t)g * e->npx, batch, st);
        if (rc) return rc;
    }
    return ROO_OK;
}

// Host buffers: upload group g+1 and download group g-1 while group g computes (three streams, two
// staging buffers).  Pinned host memory makes the copies truly asynchronous.
static void host_streams_free(roo_engine_t* e) {
    for (int b = 0; b < 2; ++b) {
        cudaFree(e->in_dev[b][0]); cudaFree(e->in_dev[b][1]); cudaFree(e->out_dev[b]);
        e->in_dev[b][0] = e->in_dev[b][1] = nullptr; e->out_dev[b] = nullptr;
        if (e->ev_in[b]) cudaEventDestroy(e->ev_in[b]);
        if (e->ev_done[b]) cudaEventDestroy(e->ev_done[b]);
        if (e->ev_out[b]) cudaEventDestroy(e->ev_out[b]);
        e->ev_in[b] = e->ev_done[b] = e->ev_out[b] = nullptr;
    }
    if (e->s_compute) cudaStreamDestroy(e->s_compute);
    if (e->s_in) cudaStreamDestroy(e->s_in);
    if (e->s_out) cudaStreamDestroy(e->s_out);
    e->s_compute = e->s_in = e->s_out = nullptr;
}

// Streams, events and staging buffers of the host-buffer path, created on first use.  `host_ready` is set only after
// EVERY allocation succeeded; a failure (e.g. out of memory for the 2 x max_batch staging frames) rolls everything back,
// so a later call starts from scratch instead of running on half-initialised state.
static int host_streams_init(roo_engine_t* e) {
    if (e->host_ready) return ROO_OK;
    const size_t npx = e->npx, B = (size_t)e->p.max_batch;
    auto attempt = [&]() -> int {
        ROO_CUDA_TRY(cudaStreamCreateWithFlags(&e->s_compute, cudaStreamNonBlocking));
        ROO_CUDA_TRY(cudaStreamCreateWithFlags(&e->s_in, cudaStreamNonBlocking));
        ROO_CUDA_TRY(cudaStreamCreateWithFlags(&e->s_out, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            ROO_CUDA_TRY(cudaMalloc((void**)&e->in_dev[b][0], B * e->in_npx));
            ROO_CUDA_TRY(cudaMalloc((void**)&e->in_dev[b][1], B * e->in_npx));
            ROO_CUDA_TRY(cudaMalloc((void**)&e->out_dev[b], B * npx * 4));
            ROO_CUDA_TRY(cudaEventCreateWithFlags(&e->ev_in[b], cudaEventDisableTiming));
            ROO_CUDA_TRY(cudaEventCreateWithFlags(&e->ev_done[b], cudaEventDisableTiming));
            ROO_CUDA_TRY(cudaEventCreateWithFlags(&e->ev_out[b], cudaEventDisableTiming));
        }
        return ROO_OK;
    };
    const int rc = attempt();
    if (rc != ROO_OK) {
        cudaGetLastError();
        host_streams_free(e);
        return rc;
    }
    e->scratch_bytes += 2 * (2 * B * e->in_npx + B * npx * 4);
    e->host_ready = true;
    return ROO_OK;
}

// One group (<= max_batch pairs) through staging slot ticket & 1: H2D on s_in, the path on s_compute, D2H on s_out.
// The slot's previous user (ticket - 2) is waited for on the HOST first, so its events are never re-recorded
// while somebody may still wait on them and its staging buffers are free.
static int submit_group(roo_engine_t* e, const uint8_t* left_host, const uint8_t* right_host, float* disp_host, int batch) {
    const size_t npx = e->npx;
    const int b = (int)(e->ticket & 1);
    if (e->ticket >= 2) ROO_CUDA_TRY(cudaEventSynchronize(e->ev_out[b]));
    ROO_CUDA_TRY(cudaMemcpyAsync(e->in_dev[b][0], left_host, (size_t)batch * e->in_npx, cudaMemcpyHostToDevice, e->s_in));
    ROO_CUDA_TRY(cudaMemcpyAsync(e->in_dev[b][1], right_host, (size_t)batch * e->in_npx, cudaMemcpyHostToDevice, e->s_in));
    ROO_CUDA_TRY(cudaEventRecord(e->ev_in[b], e->s_in));
    ROO_CUDA_TRY(cudaStreamWaitEvent(e->s_compute, e->ev_in[b], 0));
    const int rc = engine_group(e, e->in_dev[b][0], e->in_dev[b][1], e->out_dev[b], batch, e->s_compute);
    if (rc) return rc;
    ROO_CUDA_TRY(cudaEventRecord(e->ev_done[b], e->s_compute));
    ROO_CUDA_TRY(cudaStreamWaitEvent(e->s_out, e->ev_done[b], 0));
    ROO_CUDA_TRY(cudaMemcpyAsync(disp_host, e->out_dev[b], (size_t)batch * npx * 4, cudaMemcpyDeviceToHost, e->s_out));
    ROO_CUDA_TRY(cudaEventRecord(e->ev_out[b], e->s_out));
    ++e->ticket;
    return ROO_OK;
}

extern "C" int roo_engine_submit_host(roo_engine_t* e, const uint8_t* left_host, const uint8_t* right_host, float* disp_host,
                                      int n_pairs, long long* ticket) {
    if (!e || !left_host || !right_host || !disp_host || !ticket || n_pairs <= 0 || n_pairs > e->p.max_batch ||
        !on_engine_device(e))
        return ROO_ERR_INVALID_ARGUMENT;
    const int rc0 = host_streams_init(e);
    if (rc0) return rc0;
    const int rc = submit_group(e, left_host, right_host, disp_host, n_pairs);
    if (rc) return rc;
    *ticket = e->ticket - 1;
    return ROO_OK;
}

extern "C" int roo_engine_wait(roo_engine_t* e, long long ticket) {
    if (!e || ticket < 0 || ticket >= e->ticket) return ROO_ERR_INVALID_ARGUMENT;
    if (ticket + 2 < e->ticket) return ROO_OK;   // its slot was reused: submit_group already waited for it
    ROO_CUDA_TRY(cudaEventSynchronize(e->ev_out[ticket & 1]));
    return ROO_OK;
}

extern "C" int roo_engine_run_host(roo_engine_t* e, const uint8_t* left_host, const uint8_t* right_host, float* disp_host,
                                   int n_pairs) {
    if (!e || !left_host || !right_host || !disp_host || n_pairs < 0 || !on_engine_device(e)) return ROO_ERR_INVALID_ARGUMENT;
    const int rc0 = host_streams_init(e);
    if (rc0) return rc0;
    const size_t npx = e->npx;
    const int B = e->p.max_batch;
    for (int g = 0; g < n_pairs; g += B) {
        const int batch = n_pairs - g < B ? n_pairs - g : B;
        const int rc = submit_group(e, left_host + (size_t)g * e->in_npx, right_host + (size_t)g * e->in_npx, disp_host + (size_t)g * npx, batch);
        if (rc) return rc;
    }
    ROO_CUDA_TRY(cudaStreamSynchronize(e->s_out));
    ROO_CUDA_TRY(cudaStreamSynchronize(e->s_compute));
    return ROO_OK;
}

extern "C" int roo_engine_export_volume(roo_engine_t* e, int slot, const roo_volume_t* volH, void* stream) {
    if (!e || !e->H || !e->p.keep_volume || slot < 0 || slot >= e->last_batch) return ROO_ERR_INVALID_ARGUMENT;
    if (!valid_volume(volH, 4) || (int)volH->w != e->p.w || (int)volH->h != e->p.h || (int)volH->d < e->p.max_disp)
        return ROO_ERR_INVALID_ARGUMENT;
    return launch_internal_to_vol(volH, e->H + (size_t)slot * e->npx * e->DP, e->DP, e->p.max_disp, as_stream(stream));
}

extern "C" int roo_engine_export_census(roo_engine_t* e, int slot, int side, const roo_image_t* census, void* stream) {
    if (!e || slot < 0 || slot >= e->last_batch || side < 0 || side > 1) return ROO_ERR_INVALID_ARGUMENT;
    const size_t esz = (size_t)e->words * 8;
    if (!valid_image(census, esz) || (int)census->w != e->p.w || (int)census->h != e->p.h) return ROO_ERR_INVALID_ARGUMENT;
    ROO_CUDA_TRY(cudaMemcpy2DAsync(census->ptr, census->pitch, e->cen[side] + (size_t)slot * e->npx * e->words,
                                   (size_t)e->p.w * esz, (size_t)e->p.w * esz, (size_t)e->p.h, cudaMemcpyDeviceToDevice,
                                   as_stream(stream)));
    return ROO_OK;
}

extern "C" int roo_engine_set_profiling(roo_engine_t* e, int on) {
    if (!e) return ROO_ERR_INVALID_ARGUMENT;
    e->profiling = on != 0;
    e->prof_used = 0;
    for (int k = 0; k < ROO_PROF_KINDS; ++k) { e->prof_ms[k] = 0; e->prof_n[k] = 0; }
    return ROO_OK;
}

extern "C" int roo_engine_get_profile(roo_engine_t* e, double* ms_by_kind, long long* launches_by_kind) {
    if (!e || !ms_by_kind || !launches_by_kind) return ROO_ERR_INVALID_ARGUMENT;
    // fold the events recorded since the last call (the caller has synchronised the stream)
    for (size_t i = 1; i < e->prof_used; ++i) {
        const int kind = e->prof_kinds[i];
        if (kind < 0) continue;
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, e->prof_events[i - 1], e->prof_events[i]) == cudaSuccess) {
            e->prof_ms[kind] += ms;
            e->prof_n[kind] += 1;
            const int pass = e->prof_pass[i];
            if (pass >= 0 && pass < 8) { e->prof_ms[ROO_PROF_PASS0 + pass] += ms; e->prof_n[ROO_PROF_PASS0 + pass] += 1; }
        }
    }
    e->prof_used = 0;
    for (int k = 0; k < ROO_PROF_KINDS; ++k) { ms_by_kind[k] = e->prof_ms[k]; launches_by_kind[k] = e->prof_n[k]; }
    return ROO_OK;
}

// ------------------------------------------------------------------------------------------------
// Multi-GPU: stereo pairs are independent, so a batch is sharded across the GPUs of the box by pair index --
// one engine, one host thread and one set of streams per device, no collective, nothing exchanged
// (SURVEY.md 8e).  Host buffers in, host buffers out.
// ------------------------------------------------------------------------------------------------
struct roo_multi_engine {
    std::vector<int> devices;
    std::vector<roo_engine*> engines;
    size_t npx = 0;
};

extern "C" int roo_multi_engine_create(roo_multi_engine_t** out, const roo_pipeline_params_t* params, const int* devices,
                                       int n_devices) {
    if (!out || !params) return ROO_ERR_INVALID_ARGUMENT;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return ROO_ERR_NO_DEVICE;
    if (n_devices <= 0) n_devices = ndev;          // all visible devices
    if (n_devices > ndev && !devices) return ROO_ERR_INVALID_ARGUMENT;
    roo_multi_engine* m = new (std::nothrow) roo_multi_engine();
    if (!m) return ROO_ERR_OUT_OF_MEMORY;
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = ROO_OK;
    for (int i = 0; i < n_devices && rc == ROO_OK; ++i) {
        const int dev = devices ? devices[i] : i;
        if (dev < 0 || dev >= ndev) { rc = ROO_ERR_INVALID_ARGUMENT; break; }
        if (cudaSetDevice(dev) != cudaSuccess) { rc = ROO_ERR_NO_DEVICE; break; }
        roo_engine* e = nullptr;
        rc = roo_engine_create(&e, params);
        if (rc == ROO_OK) { m->devices.push_back(dev); m->engines.push_back(e); }
    }
    cudaSetDevice(prev);
    if (rc != ROO_OK) {
        for (size_t i = 0; i < m->engines.size(); ++i) { cudaSetDevice(m->devices[i]); roo_engine_destroy(m->engines[i]); }
        cudaSetDevice(prev);
        delete m;
        return rc;
    }
    m->npx = (size_t)params->w * params->h;
    *out = m;
    return ROO_OK;
}

extern "C" int roo_multi_engine_destroy(roo_multi_engine_t* m) {
    if (!m) return ROO_ERR_INVALID_ARGUMENT;
    int prev = 0;
    cudaGetDevice(&prev);
    for (size_t i = 0; i < m->engines.size(); ++i) { cudaSetDevice(m->devices[i]); roo_engine_destroy(m->engines[i]); }
    cudaSetDevice(prev);
    delete m;
    return ROO_OK;
}

extern "C" int roo_multi_engine_device_count(const roo_multi_engine_t* m) { return m ? (int)m->engines.size() : 0; }

// contiguous block of pairs for device i of n (blocks differ by at most one pair)
static void shard_of(int n_pairs, int n, int i, int* begin, int* count) {
    const int base = n_pairs / n, rem = n_pairs % n;
    *begin = i * base + (i < rem ? i : rem);
    *count = base + (i < rem ? 1 : 0);
}

extern "C" int roo_multi_engine_run_host(roo_multi_engine_t* m, const uint8_t* left_host, const uint8_t* right_host,
                                         float* disp_host, int n_pairs) {
    if (!m || !left_host || !right_host || !disp_host || n_pairs < 0) return ROO_ERR_INVALID_ARGUMENT;
    const int n = (int)m->engines.size();
    std::vector<int> status(n, ROO_OK);
    std::vector<std::thread> workers;
    bool spawn_failed = false;
    for (int i = 0; i < n && !spawn_failed; ++i) {
        try {   // std::thread may throw (resource exhaustion): nothing may propagate through the extern "C" boundary
        workers.emplace_back([&, i]() {
            int begin = 0, count = 0;
            shard_of(n_pairs, n, i, &begin, &count);
            if (count == 0) return;
            if (cudaSetDevice(m->devices[i]) != cudaSuccess) { status[i] = ROO_ERR_NO_DEVICE; return; }
            const size_t off = (size_t)begin * m->npx;
            status[i] = roo_engine_run_host(m->engines[i], left_host + off, right_host + off, disp_host + off, count);
        });
        } catch (...) { spawn_failed = true; }
    }
    for (auto& t : workers) t.join();
    if (spawn_failed) return ROO_ERR_OUT_OF_MEMORY;
    for (int i = 0; i < n; ++i)
        if (status[i] != ROO_OK) return status[i];
    return ROO_OK;
}

// Development aid: cycle counters written by a -DVG_TIMING build of sgm_fused.cu (zeros otherwise).
extern "C" int roo_engine_debug_counters(roo_engine_t* e, unsigned long long* out, int n, int reset) {
    if (!e || !e->flags || !out || n <= 0 || n > 32) return ROO_ERR_INVALID_ARGUMENT;
    char* base = reinterpret_cast<char*>(e->flags) + ((size_t)e->p.max_batch * vgroup_bands(e->p.w, e->p.h, e->DP) + 2) * 4;
    ROO_CUDA_TRY(cudaMemcpy(out, base, (size_t)n * 8, cudaMemcpyDeviceToHost));
    if (reset) ROO_CUDA_TRY(cudaMemset(base, 0, 256));
    return ROO_OK;
}

extern "C" const char* roo_b200_version(void) { return "kangaroo_b200 0.1 (sm_100a)"; }

extern "C" unsigned long long roo_launch_count(void) { return g_launches.load(); }

extern "C" void roo_set_ieee_division(int on) { g_ieee_div.store(on ? 1 : 0); }

extern "C" int roo_set_tuning(int knob, int value) {
    switch (knob) {
        case ROO_TUNE_HSWEEP: g_use_hsweep.store(value ? 1 : 0); return ROO_OK;
        case ROO_TUNE_INSWEEP_COST: g_insweep_cost.store(value ? 1 : 0); return ROO_OK;
        case ROO_TUNE_STRIP_CTAS_PER_SM: g_strip_ctas_per_sm.store(value < 0 ? 0 : value); return ROO_OK;
        case ROO_TUNE_SOLO_GEOMETRY: g_solo_geometry.store(value ? 1 : 0); return ROO_OK;
        case ROO_TUNE_GUIDED_SCRATCH_MIB: g_guided_scratch_mib.store(value < 1 ? 1 : value); return ROO_OK;
        default: return ROO_ERR_INVALID_ARGUMENT;
    }
}

extern "C" const char* roo_status_string(int status) {
    switch (status) {
        case ROO_OK: return "ok";
        case ROO_ERR_INVALID_ARGUMENT: return "invalid argument";
        case ROO_ERR_UNSUPPORTED: return "unsupported configuration";
        case ROO_ERR_OUT_OF_MEMORY: return "out of memory";
        case ROO_ERR_NO_DEVICE: return "no CUDA device";
        default: return status > 0 ? cudaGetErrorString((cudaError_t)status) : "unknown";
    }
}
