// Row-strip split of ONE stereo pair across the GPUs of a box (BASELINE config 5: a single 3840x2160 pair,
// 256 disparities).  The reference is single-device by construction
// (src/cu_semi_global_matching.cu:69-84 launches on the current device), so this is new design.
//
//   * GPU k owns the image rows [y0_k, y1_k): its strip of the aggregate H (fp32, disparity innermost), of the u8
//     matching cost and of the disparity images never leaves it.  Each GPU receives its rows of the input frames plus
//     eight halo rows above and below (the reach of the largest census window) straight from the host and computes the
//     census descriptors of those rows itself, so no census halo is exchanged between GPUs.
//   * Horizontal paths (+1,0), (-1,0), the winner-takes-all epilogue, the right-reference disparity and the left-right
//     check are local to a row: no exchange at all.
//   * The six paths that travel in y (down, down-right, down-left, then up, up-left, up-right -- one sweep each, every
//     scanline a warp) cross the strips.  A path leaving strip k through its last row hands its state -- the aggregate
//     row it just wrote (DP floats), lastBestCr and the pixel intensity -- to the path entering strip k+1: the exporting
//     warp writes the record straight into the NEXT GPU's memory over NVLink (peer stores) and publishes it with
//     st.release.sys; the importing warp polls the record's sequence word with ld.acquire.sys (sgm.cu).  Scanlines are
//     independent, so the strips of one sweep run as a pipeline offset by a few scanline groups, not one after the other:
//     no collective, no host round trip, 1 MB..4 MB per sweep and boundary.
//   * Two strips on the SAME device (devices = {0, 0}: how the single-GPU test box exercises this code) cannot poll each
//     other from inside a kernel -- the consumer could occupy every SM before the producer is resident -- so there the
//     hand-off is ordered by a CUDA event between the two streams instead.
// Results are bit-identical to the single-GPU engine: same kernels, same order of operations per path.
#include <new>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"
#include "sgm_step.cuh"

using namespace roo_b200;

namespace {
// restores the caller's current device on every exit path
struct DeviceGuard {
    int prev = 0;
    DeviceGuard() { cudaGetDevice(&prev); }
    ~DeviceGuard() { cudaSetDevice(prev); }
};
constexpr size_t CEN_PAD = 1024;   // elements of padding around the census arrays (in-sweep cost strips stick out of a row)

struct Strip {
    int device = 0;
    int y0 = 0, hl = 0;                       // first row and number of rows
    cudaStream_t st = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    std::vector<cudaEvent_t> ev_sweep;        // completion of crossing sweep i on this strip (same-device ordering)
    unsigned char* frame[2] = {nullptr, nullptr};     // whole frames [h][w]
    unsigned long long* cen_base[2] = {nullptr, nullptr};
    unsigned long long* cen[2] = {nullptr, nullptr};  // whole-frame census descriptors
    float* imgf = nullptr;                    // [hl][w] adaptive-P2 intensities of the strip
    unsigned char* c8 = nullptr;              // [hl][w][DP]
    float* H = nullptr;                       // [hl][w][DP]
    float* disp = nullptr;                    // [hl][w]
    float* dispR = nullptr;                   // [hl][w]
    std::vector<float*> import;               // per crossing sweep: [w][DP+4] records written by the upstream strip
};
}  // namespace

struct roo_split_engine {
    roo_pipeline_params_t p;
    int DP = 0, words = 0, ieee = 0;
    SgmPlan plan{};                           // one pass per path (the fused vertical groups do not split by rows)
    std::vector<Strip> strips;
    long long frame = 0;
    float last_ms = 0.0f;
    size_t exchanged_bytes = 0;               // bytes written to peer memory per frame
};

static void split_free(roo_split_engine* e) {
    int prev = 0;
    cudaGetDevice(&prev);
    for (Strip& s : e->strips) {
        cudaSetDevice(s.device);
        if (s.st) cudaStreamSynchronize(s.st);
        cudaFree(s.frame[0]); cudaFree(s.frame[1]); cudaFree(s.cen_base[0]); cudaFree(s.cen_base[1]);
        cudaFree(s.imgf); cudaFree(s.c8); cudaFree(s.H); cudaFree(s.disp); cudaFree(s.dispR);
        for (float* b : s.import) cudaFree(b);
        for (cudaEvent_t ev : s.ev_sweep)
            if (ev) cudaEventDestroy(ev);
        if (s.ev_begin) cudaEventDestroy(s.ev_begin);
        if (s.ev_end) cudaEventDestroy(s.ev_end);
        if (s.st) cudaStreamDestroy(s.st);
    }
    cudaSetDevice(prev);
    cudaGetLastError();
}

extern "C" int roo_split_engine_create(roo_split_engine_t** out, const roo_pipeline_params_t* params, const int* devices,
                                       int n_devices) {
    if (!out || !params) return ROO_ERR_INVALID_ARGUMENT;
    const roo_pipeline_params_t& p = *params;
    if (p.w <= 0 || p.h <= 0 || p.max_disp <= 0 || p.window < 0 || p.window > 2) return ROO_ERR_INVALID_ARGUMENT;
    if (p.max_disp > ROO_MAX_DISP) return ROO_ERR_UNSUPPORTED;
    if (p.median_size != 0 || p.filtgrad_threshold > 0.0f) return ROO_ERR_UNSUPPORTED;   // these stages need halo rows: single-GPU engine only
    if (p.fp_mode < ROO_FP_DEFAULT || p.fp_mode > ROO_FP_IEEE) return ROO_ERR_INVALID_ARGUMENT;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return ROO_ERR_NO_DEVICE;
    if (n_devices <= 0) { n_devices = ndev; devices = nullptr; }
    if (n_devices > p.h) return ROO_ERR_INVALID_ARGUMENT;
    roo_split_engine* e = new (std::nothrow) roo_split_engine();
    if (!e) return ROO_ERR_OUT_OF_MEMORY;
    e->p = p;
    e->DP = disp_padded(p.max_disp);
    e->words = p.window == ROO_WIN_9x7 ? 1 : (p.window == ROO_WIN_11x11 ? 2 : 4);
    e->ieee = p.fp_mode == ROO_FP_DEFAULT ? (g_ieee_div.load() != 0) : (p.fp_mode == ROO_FP_IEEE);
    e->plan = sgm_plan(p.dohoriz, p.dovert, p.doreverse, p.dodiag, 0);
    const int G = n_devices, w = p.w, h = p.h;
    const size_t npx = (size_t)w * h;
    int prev = 0;
    cudaGetDevice(&prev);
    e->strips.resize(G);
    int rc = ROO_OK;
    for (int k = 0; k < G && rc == ROO_OK; ++k) {
        Strip& s = e->strips[k];
        s.device = devices ? devices[k] : k;
        if (s.device < 0 || s.device >= ndev) { rc = ROO_ERR_INVALID_ARGUMENT; break; }
        s.y0 = (int)((long long)h * k / G);
        s.hl = (int)((long long)h * (k + 1) / G) - s.y0;
        if (cudaSetDevice(s.device) != cudaSuccess) { rc = ROO_ERR_NO_DEVICE; break; }
        const size_t spx = (size_t)w * s.hl;
        bool ok = cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaEventCreate(&s.ev_begin) == cudaSuccess && cudaEventCreate(&s.ev_end) == cudaSuccess;
        for (int sd = 0; sd < 2 && ok; ++sd) {
            ok = cudaMalloc((void**)&s.frame[sd], npx) == cudaSuccess &&
                 cudaMalloc((void**)&s.cen_base[sd], (npx * e->words + 2 * CEN_PAD) * 8) == cudaSuccess;
            if (ok) {
                s.cen[sd] = s.cen_base[sd] + CEN_PAD;
                cudaMemset(s.cen_base[sd], 0, CEN_PAD * 8);
                cudaMemset(s.cen[sd] + npx * e->words, 0, CEN_PAD * 8);
            }
        }
        ok = ok && cudaMalloc((void**)&s.imgf, spx * 4) == cudaSuccess && cudaMalloc((void**)&s.disp, spx * 4) == cudaSuccess;
        if (ok && e->plan.n > 0)
            ok = cudaMalloc((void**)&s.c8, spx * e->DP) == cudaSuccess && cudaMalloc((void**)&s.H, spx * e->DP * 4) == cudaSuccess;
        if (ok && p.lrcheck) ok = cudaMalloc((void**)&s.dispR, spx * 4) == cudaSuccess;
        s.import.assign(e->plan.n, nullptr);
        s.ev_sweep.assign(e->plan.n, nullptr);
        for (int i = 0; i < e->plan.n && ok; ++i) {
            if (e->plan.pass[i].dy == 0) continue;
            const size_t bytes = (size_t)w * strip_rec_floats(e->DP) * 4;
            ok = cudaMalloc((void**)&s.import[i], bytes) == cudaSuccess && cudaMemset(s.import[i], 0xff, bytes) == cudaSuccess &&
                 cudaEventCreateWithFlags(&s.ev_sweep[i], cudaEventDisableTiming) == cudaSuccess;
        }
        if (!ok) rc = ROO_ERR_OUT_OF_MEMORY;
    }
    // neighbours write into each other's import buffers
    for (int k = 0; k + 1 < G && rc == ROO_OK; ++k) {
        const int a = e->strips[k].device, b = e->strips[k + 1].device;
        if (a == b) continue;
        int can_ab = 0, can_ba = 0;
        cudaDeviceCanAccessPeer(&can_ab, a, b);
        cudaDeviceCanAccessPeer(&can_ba, b, a);
        if (!can_ab || !can_ba) { rc = ROO_ERR_UNSUPPORTED; break; }
        cudaSetDevice(a);
        cudaError_t ea = cudaDeviceEnablePeerAccess(b, 0);
        cudaSetDevice(b);
        cudaError_t eb = cudaDeviceEnablePeerAccess(a, 0);
        if ((ea != cudaSuccess && ea != cudaErrorPeerAccessAlreadyEnabled) || (eb != cudaSuccess && eb != cudaErrorPeerAccessAlreadyEnabled))
            rc = ROO_ERR_UNSUPPORTED;
        cudaGetLastError();
    }
    for (Strip& s : e->strips) { cudaSetDevice(s.device); cudaDeviceSynchronize(); }
    cudaSetDevice(prev);
    if (rc != ROO_OK) {
        cudaGetLastError();
        split_free(e);
        delete e;
        return rc;
    }
    *out = e;
    return ROO_OK;
}

extern "C" int roo_split_engine_destroy(roo_split_engine_t* e) {
    if (!e) return ROO_ERR_INVALID_ARGUMENT;
    split_free(e);
    delete e;
    return ROO_OK;
}

extern "C" int roo_split_engine_strip_count(const roo_split_engine_t* e) { return e ? (int)e->strips.size() : 0; }

// One pair: whole frames in host memory (pinned recommended), the disparity image back in host memory.
extern "C" int roo_split_engine_run_host(roo_split_engine_t* e, const uint8_t* left_host, const uint8_t* right_host,
                                         float* disp_host) {
    if (!e || !left_host || !right_host || !disp_host) return ROO_ERR_INVALID_ARGUMENT;
    const roo_pipeline_params_t& p = e->p;
    const int G = (int)e->strips.size(), w = p.w, h = p.h, DP = e->DP, ndir = e->plan.n;
    DeviceGuard guard;
    ++e->frame;
    e->exchanged_bytes = 0;
    // in-sweep matching cost (no u8 cost volume at all) with the one-word descriptor under the reference's popcount
    const bool cen_ok = e->words == 1 && p.popc_mode == ROO_POPC32_COMPAT && g_insweep_cost.load(std::memory_order_relaxed);
    // (the generic single-path sweep recomputes the cost in the sweep only up to 64 disparities, see engine.cu)
    const bool hs = g_use_hsweep.load(std::memory_order_relaxed) != 0;
    auto in_sweep = [&](const SgmPass& ps) { return cen_ok && ((ps.dy == 0 && hs) || DP <= 64); };
    bool need_c8 = false;
    for (int i = 0; i < ndir; ++i) need_c8 |= !in_sweep(e->plan.pass[i]);
    int rc = ROO_OK;
    // Stage 1 on every device: upload, census, cost, intensities, right-reference disparity.
    // Stage 2: the sweeps in plan order; the strips are visited in the travel direction of each crossing sweep so that a
    // producer's launch (and its completion event, for strips that share a device) is always issued before its consumer's.
    for (int k = 0; k < G && rc == ROO_OK; ++k) {
        Strip& s = e->strips[k];
        cudaSetDevice(s.device);
        const size_t off = (size_t)s.y0 * w;
        // Only the strip's rows plus the census window's reach above and below (8 rows covers all three windows) are
        // uploaded and transformed: the clamp-to-edge of the census then acts at the true image border or inside the
        // halo, never on a row this strip keeps.  Buffers stay whole-frame sized so that every row sits at its own offset.
        const int ya = s.y0 - 8 > 0 ? s.y0 - 8 : 0, yb = s.y0 + s.hl + 8 < h ? s.y0 + s.hl + 8 : h;
        const size_t hoff = (size_t)ya * w, hbytes = (size_t)(yb - ya) * w;
        ROO_CUDA_TRY(cudaMemcpyAsync(s.frame[0] + hoff, left_host + hoff, hbytes, cudaMemcpyHostToDevice, s.st));
        ROO_CUDA_TRY(cudaMemcpyAsync(s.frame[1] + hoff, right_host + hoff, hbytes, cudaMemcpyHostToDevice, s.st));
        ROO_CUDA_TRY(cudaEventRecord(s.ev_begin, s.st));
        for (int sd = 0; sd < 2 && rc == ROO_OK; ++sd)
            rc = launch_census((char*)(s.cen[sd] + hoff * e->words), (size_t)w * e->words * 8, 0, (const char*)(s.frame[sd] + hoff),
                               (size_t)w, 0, w, yb - ya, 1, p.window, ROO_IMG_U8, s.st);
        if (rc == ROO_OK && p.lrcheck)
            rc = launch_census_wta(s.dispR, s.cen[1] + off * e->words, s.cen[0] + off * e->words, w, s.hl, 1, p.max_disp, e->words,
                                   p.popc_mode, p.subpix, +1, e->ieee, s.st);
        if (rc == ROO_OK && ndir == 0)
            rc = launch_census_wta(s.disp, s.cen[0] + off * e->words, s.cen[1] + off * e->words, w, s.hl, 1, p.max_disp, e->words,
                                   p.popc_mode, p.subpix, -1, e->ieee, s.st);
        if (rc == ROO_OK && ndir > 0 && need_c8)
            rc = launch_cost_u8(s.c8, s.cen[0] + off * e->words, s.cen[1] + off * e->words, w, s.hl, 1, DP, p.max_disp, e->words,
                                p.popc_mode, s.st);
        if (rc == ROO_OK && ndir > 0)
            rc = launch_image_to_f32(s.imgf, s.frame[0] + off, (size_t)w, 0, ROO_IMG_U8, w, s.hl, 1, p.img_scale, s.st);
    }
    for (int i = 0; i < ndir && rc == ROO_OK; ++i) {
        const SgmPass& ps = e->plan.pass[i];
        const bool crossing = ps.dy != 0;
        for (int t = 0; t < G && rc == ROO_OK; ++t) {
            const int k = (crossing && ps.dy < 0) ? G - 1 - t : t;            // travel order
            const int up = ps.dy > 0 ? k - 1 : k + 1, down = ps.dy > 0 ? k + 1 : k - 1;
            Strip& s = e->strips[k];
            cudaSetDevice(s.device);
            const size_t off = (size_t)s.y0 * w;
            SweepArgs a{};
            a.H = s.H; a.h_pair = 0; a.C = s.c8; a.c_pair = 0; a.img = s.imgf; a.img_pair = 0;
            a.cost_scale = 1.0f / (float)(e->words * 64);
            a.w = w; a.h = s.hl; a.DP = DP; a.maxDisp = p.max_disp; a.batch = 1; a.P1 = p.P1; a.P2 = p.P2;
            a.dx = ps.dx; a.dy = ps.dy; a.first = i == 0; a.ieee = e->ieee; a.subpix = p.subpix;
            a.cost_kind = in_sweep(ps) ? COST_CEN32 : COST_U8;
            a.cenL = s.cen[0] + off; a.cenR = s.cen[1] + off; a.cen_pair = 0;
            a.epi = i + 1 < ndir ? EPI_NONE : (p.keep_volume ? EPI_WTA_WRITE : EPI_WTA_ONLY);
            a.disp = s.disp; a.disp_pair = 0;
            a.strip_seq = (int)((e->frame * 8 + i) & 0x3fffffff);
            if (crossing) {
                if (up >= 0 && up < G) {
                    a.strip_import = s.import[i];
                    if (e->strips[up].device == s.device)   // same device: order the two streams, kernels cannot poll each other safely
                        ROO_CUDA_TRY(cudaStreamWaitEvent(s.st, e->strips[up].ev_sweep[i], 0));
                }
                if (down >= 0 && down < G) {
                    a.strip_export = e->strips[down].import[i];
                    e->exchanged_bytes += (size_t)w * strip_rec_floats(DP) * 4;
                }
            }
            rc = launch_sweep(a, s.st);
            if (rc == ROO_OK && crossing) ROO_CUDA_TRY(cudaEventRecord(s.ev_sweep[i], s.st));
        }
    }
    for (int k = 0; k < G && rc == ROO_OK; ++k) {
        Strip& s = e->strips[k];
        cudaSetDevice(s.device);
        if (p.lrcheck) {
            // LeftRightCheck(disp[1], disp[0], +1, maxdiff); LeftRightCheck(disp[0], disp[1], -1, maxdiff) (main.cpp:451-454)
            rc = launch_lr_check_f32(s.dispR, (size_t)w * 4, s.disp, (size_t)w * 4, w, s.hl, 1, 0, 0, +1.0f, p.lr_maxdiff, s.st);
            if (rc == ROO_OK)
                rc = launch_lr_check_f32(s.disp, (size_t)w * 4, s.dispR, (size_t)w * 4, w, s.hl, 1, 0, 0, -1.0f, p.lr_maxdiff, s.st);
        }
        ROO_CUDA_TRY(cudaEventRecord(s.ev_end, s.st));
        ROO_CUDA_TRY(cudaMemcpyAsync(disp_host + (size_t)s.y0 * w, s.disp, (size_t)w * s.hl * 4, cudaMemcpyDeviceToHost, s.st));
    }
    float ms_max = 0.0f;
    for (int k = 0; k < G; ++k) {
        Strip& s = e->strips[k];
        cudaSetDevice(s.device);
        const cudaError_t se = cudaStreamSynchronize(s.st);
        if (se != cudaSuccess && rc == ROO_OK) rc = (int)se;
        float ms = 0.0f;
        if (se == cudaSuccess && cudaEventElapsedTime(&ms, s.ev_begin, s.ev_end) == cudaSuccess && ms > ms_max) ms_max = ms;
    }
    e->last_ms = ms_max;
    return rc;
}

// Device time of the last frame (census .. left-right check; max over the strips, CUDA events on each strip's stream)
// and the bytes handed from strip to strip through peer memory.
extern "C" int roo_split_engine_last_stats(const roo_split_engine_t* e, float* device_ms, unsigned long long* exchanged_bytes) {
    if (!e) return ROO_ERR_INVALID_ARGUMENT;
    if (device_ms) *device_ms = e->last_ms;
    if (exchanged_bytes) *exchanged_bytes = (unsigned long long)e->exchanged_bytes;
    return ROO_OK;
}
