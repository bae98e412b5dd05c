// C ABI of libllicti_b200.so: context, geometry, weight packing, workspace and the host-side
// sequencing of the kernels for the full compress / decompress path.  See include/llicti.h.
#include <cstdlib>
#include <stdarg.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace llicti {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int device_sm_count(llicti_ctx *ctx, int *out) {
    if (!ctx->sm_count) {
        LLICTI_CUDA(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, ctx->cfg.device));
        LLICTI_CUDA(cudaDeviceGetAttribute(&ctx->regs_per_sm, cudaDevAttrMaxRegistersPerMultiprocessor, ctx->cfg.device));
    }
    *out = ctx->sm_count;
    return LLICTI_OK;
}

int num_substreams(int64_t n, int sub_len) {
    if (sub_len <= 0) return 1;
    int64_t s = std::max<int64_t>(1, (n + sub_len - 1) / sub_len);
    if (s > 32) s = (s + 31) / 32 * 32;
    return (int)std::min<int64_t>(s, 65535);
}

int make_plan(const llicti_config &cfg, int H, int W, Plan *out) {
    LLICTI_REQUIRE(cfg.num_scales >= 1 && cfg.num_scales <= LLICTI_MAX_SCALES, "num_scales %d out of range", cfg.num_scales);
    LLICTI_REQUIRE(cfg.num_mixtures == kM, "num_mixtures must be %d", kM);
    LLICTI_REQUIRE(cfg.sub_len >= 0 && cfg.sub_len <= 16384, "sub_len %d out of range [0, 16384]", cfg.sub_len);
    Plan &p = *out;
    memset(&p, 0, sizeof(p));
    llicti_geom &g = p.g;
    g.H = H; g.W = W; g.num_scales = cfg.num_scales;
    int hin = H, win = W;
    for (int s = 0; s < cfg.num_scales; ++s) {
        LLICTI_REQUIRE(hin >= 2 && win >= 2, "image %dx%d too small for %d scales", H, W, cfg.num_scales);
        g.Hs[s] = (hin + 1) / 2; g.Ws[s] = (win + 1) / 2;
        g.padH[s] = g.Hs[s] > hin / 2; g.padW[s] = g.Ws[s] > win / 2;
        g.pad_int = 4 * g.pad_int + 2 * g.padH[s] + g.padW[s];
        for (int b = 0; b < 3; ++b) {
            g.crop_h[s][b] = (b == 0 || b == 2) ? g.Hs[s] - g.padH[s] : g.Hs[s];
            g.crop_w[s][b] = (b == 0 || b == 1) ? g.Ws[s] - g.padW[s] : g.Ws[s];
            // container version 2: substreams (= serial decoder chains) of sub_len symbols at the finest scale, half of that
            // at the coarser ones, which have a quarter and less of the symbols and would otherwise have too few chains
            g.num_sub[s][b] = num_substreams((int64_t)g.crop_h[s][b] * g.crop_w[s][b],
                                             cfg.sub_len > 0 && s > 0 ? std::max(cfg.sub_len / 2, 1) : cfg.sub_len);
        }
        g.positions += (int64_t)g.Hs[s] * g.Ws[s];
        p.plane_elems[s] = 12ll * g.Hs[s] * g.Ws[s];
        hin = g.Hs[s]; win = g.Ws[s];
    }
    // the reference stores the coarsest size in two uint8 (LLICTI_nets.py:347)
    LLICTI_REQUIRE(g.Hs[cfg.num_scales - 1] <= 255 && g.Ws[cfg.num_scales - 1] <= 255,
                   "coarsest band %dx%d does not fit the header's uint8 fields", g.Hs[cfg.num_scales - 1],
                   g.Ws[cfg.num_scales - 1]);
    p.n_streams = 9 * cfg.num_scales;
    int64_t sym = 0, slot = 0;
    int sub = 0;
    int k = 0;
    for (int s = cfg.num_scales - 1; s >= 0; --s)
        for (int b = 0; b < 3; ++b)
            for (int c = 0; c < 3; ++c, ++k) {
                StreamDesc &d = p.sd[k];
                d.scale = s; d.band = b; d.clr = c;
                d.Hs = g.Hs[s]; d.Ws = g.Ws[s];
                d.crop_h = g.crop_h[s][b]; d.crop_w = g.crop_w[s][b];
                d.n_sym = d.crop_h * d.crop_w;
                d.S = g.num_sub[s][b];
                d.sub_first = sub;
                const int per = (d.n_sym + d.S - 1) / d.S;
                d.slot_bytes = (2 * per + per / 64 + 16 + 15) / 16 * 16;
                d.sym_off = sym;
                d.slot_off = slot;
                sym += d.n_sym;
                sub += d.S;
                slot += (int64_t)d.S * d.slot_bytes;
                g.max_stream_bytes += (cfg.sub_len > 0 ? 2 + 2 * d.S : 0) + (int64_t)d.S * d.slot_bytes;
            }
    g.symbols = sym;
    g.substreams = sub;
    p.scratch_bytes = slot;
    return LLICTI_OK;
}

// Layer-0 branches in the order of llicti_weights.l0_*: source phase, kernel size, left/top pad.
struct BranchDef { int band, phase, kh, kw, padl, padt; };
static const BranchDef kBranches[6] = {
    {0, 0, 4, 4, 1, 1},   // layer0_00_11  LLICTI_nets.py:651-652
    {1, 0, 3, 4, 1, 1},   // layer0_00_01  :659-660
    {1, 1, 4, 3, 1, 2},   // layer0_11_01  :661-662
    {2, 0, 4, 3, 1, 1},   // layer0_00_10  :670-671
    {2, 1, 3, 4, 2, 1},   // layer0_11_10  :672-673
    {2, 2, 4, 4, 2, 1},   // layer0_01_10  :674-675
};

static int upload(const std::vector<float> &h, float **d) {
    LLICTI_CUDA(cudaMalloc((void **)d, h.size() * sizeof(float)));
    LLICTI_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    return LLICTI_OK;
}

static int pack_weights(llicti_ctx *ctx, const llicti_weights &w) {
    const int G = ctx->cfg.chs, Ch = 4 * G, GP = G + 8;
    for (int band = 0; band < 3; ++band) {
        TapTable &t = ctx->taps[band];
        memset(&t, 0, sizeof(t));
        int K0 = 0;
        for (int br = 0; br < 6; ++br)
            if (kBranches[br].band == band) K0 += 3 * kBranches[br].kh * kBranches[br].kw;
        LLICTI_REQUIRE(K0 <= 128, "layer-0 depth %d too large", K0);
        t.K0 = K0;
        std::vector<float> w0((size_t)4 * K0 * GP, 0.f), b0(Ch, 0.f);
        int k = 0;
        for (int br = 0; br < 6; ++br) {
            const BranchDef &bd = kBranches[br];
            if (bd.band != band) continue;
            LLICTI_REQUIRE(w.l0_w[br] && w.l0_b[br], "missing layer-0 weights of branch %d", br);
            for (int c = 0; c < 3; ++c)
                for (int dy = 0; dy < bd.kh; ++dy)
                    for (int dx = 0; dx < bd.kw; ++dx, ++k) {
                        t.phase[k] = (int8_t)bd.phase; t.chan[k] = (int8_t)c;
                        t.dy[k] = (int8_t)(dy - bd.padt); t.dx[k] = (int8_t)(dx - bd.padl);
                        for (int ch = 0; ch < Ch; ++ch)
                            w0[((size_t)(ch / G) * K0 + k) * GP + ch % G] =
                                w.l0_w[br][(((size_t)ch * 3 + c) * bd.kh + dy) * bd.kw + dx];
                    }
            for (int ch = 0; ch < Ch; ++ch) b0[ch] += w.l0_b[br][ch];
        }
        LLICTI_REQUIRE(w.l1_w[band] && w.l1_b[band] && w.l2_w[band] && w.l2_b[band], "missing 1x1 weights of band %d", band);
        std::vector<float> w1((size_t)4 * G * GP, 0.f), b1(w.l1_b[band], w.l1_b[band] + Ch);
        for (int g = 0; g < 4; ++g)
            for (int o = 0; o < G; ++o)
                for (int i = 0; i < G; ++i) w1[((size_t)g * G + i) * GP + o] = w.l1_w[band][(size_t)(g * G + o) * G + i];
        std::vector<float> w2((size_t)4 * G * 16, 0.f), b2(w.l2_b[band], w.l2_b[band] + kParamCh);
        for (int g = 0; g < 4; ++g)
            for (int o = 0; o < 15; ++o)
                for (int i = 0; i < G; ++i) w2[((size_t)g * G + i) * 16 + o] = w.l2_w[band][(size_t)(g * 15 + o) * G + i];
        BandWeightsF32 &bw = ctx->wf32[band];
        bw.K0 = K0;
        int rc;
        if ((rc = upload(w0, &bw.w0)) || (rc = upload(b0, &bw.b0)) || (rc = upload(w1, &bw.w1)) ||
            (rc = upload(b1, &bw.b1)) || (rc = upload(w2, &bw.w2)) || (rc = upload(b2, &bw.b2)))
            return rc;
    }
    return LLICTI_OK;
}

static void drop_decode_graph(llicti_ctx *ctx) {
    if (ctx->dec_graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)ctx->dec_graph_exec);
    ctx->dec_graph_exec = nullptr;
    ctx->dec_key_seen = false;
}

static void free_workspace(llicti_ctx *ctx) {
    drop_decode_graph(ctx);          // it holds the workspace pointers
    cudaFree(ctx->d_sd); ctx->d_sd = nullptr;
    cudaFree(ctx->d_rgb); ctx->d_rgb = nullptr;
    for (auto &p : ctx->d_planes) { cudaFree(p); p = nullptr; }
    cudaFree(ctx->d_minmax); ctx->d_minmax = nullptr;
    cudaFree(ctx->d_minmax16); ctx->d_minmax16 = nullptr;
    cudaFree(ctx->d_params); ctx->d_params = nullptr;
    cudaFree(ctx->d_bounds); ctx->d_bounds = nullptr;
    cudaFree(ctx->d_scratch); ctx->d_scratch = nullptr;
    cudaFree(ctx->d_sublen); ctx->d_sublen = nullptr;
    cudaFree(ctx->d_suboff); ctx->d_suboff = nullptr;
    cudaFree(ctx->d_stream_bytes); ctx->d_stream_bytes = nullptr;
    cudaFree(ctx->d_stream_off); ctx->d_stream_off = nullptr;
    cudaFree(ctx->d_blob); ctx->d_blob = nullptr;
    cudaFree(ctx->d_x00); ctx->d_x00 = nullptr;
    cudaFree(ctx->d_items); ctx->d_items = nullptr; ctx->items_cap = 0;
    cudaFree(ctx->d_item_flags); ctx->d_item_flags = nullptr;
    cudaFree(ctx->d_chain_state_raw); ctx->d_chain_state_raw = nullptr; ctx->wave_ws = false;
    cudaFree(ctx->d_syms); ctx->d_syms = nullptr; ctx->sym_cap = 0;
    ctx->ws_images = 0;
}

static int check_batch(llicti_ctx *ctx, int n, int H, int W) {
    LLICTI_REQUIRE(ctx, "null context");
    LLICTI_REQUIRE(ctx->ws_images > 0, "llicti_reserve has not been called");
    LLICTI_REQUIRE(n >= 1 && n <= ctx->ws_images && H == ctx->ws_H && W == ctx->ws_W,
                   "batch %d x %dx%d does not fit the reserved workspace %d x %dx%d", n, H, W, ctx->ws_images,
                   ctx->ws_H, ctx->ws_W);
    return LLICTI_OK;
}

static int read_status(llicti_ctx *ctx, cudaStream_t st) {
    int32_t s = 0;
    LLICTI_CUDA(cudaMemcpyAsync(&s, ctx->d_status, sizeof(s), cudaMemcpyDeviceToHost, st));
    LLICTI_CUDA(cudaStreamSynchronize(st));
    if (s != 0) {
        int32_t zero = 0;
        cudaMemcpyAsync(ctx->d_status, &zero, sizeof(zero), cudaMemcpyHostToDevice, st);
        cudaStreamSynchronize(st);
        set_error(s == LLICTI_E_NOMEM ? "device-side capacity overflow while coding"
                  : s == LLICTI_E_TIMEOUT ? "decode kernels that hand work to each other were not resident together and gave up"
                                          : "malformed bitstream container");
    }
    return s;
}

static int cnn(llicti_ctx *ctx, int band, const int16_t *planes, int n, int Hs, int Ws, float *params, cudaStream_t st) {
    if (ctx->cfg.cnn_impl == LLICTI_CNN_FP32) return launch_cnn_fp32(ctx, band, planes, n, Hs, Ws, params, st);
    if (ctx->cfg.cnn_impl == LLICTI_CNN_TCGEN05) return launch_cnn_tc(ctx, band, planes, n, Hs, Ws, params, st);
    set_error("unknown cnn_impl %d", ctx->cfg.cnn_impl);
    return LLICTI_E_ARG;
}

}  // namespace llicti

// ---- batches of mixed sizes ------------------------------------------------------------------
// Indices of `n` descriptors grouped by (H, W), first-seen order of the sizes, original order inside a group.
template <class Item>
static std::vector<std::vector<int>> group_by_size(const Item *items, int n) {
    std::vector<std::vector<int>> groups;
    for (int i = 0; i < n; ++i) {
        size_t g = 0;
        for (; g < groups.size(); ++g)
            if (items[groups[g][0]].H == items[i].H && items[groups[g][0]].W == items[i].W) break;
        if (g == groups.size()) groups.emplace_back();
        groups[g].push_back(i);
    }
    return groups;
}

static int reserve_for(llicti_ctx *ctx, int n, int H, int W) {
    if (ctx->ws_images >= n && ctx->ws_H == H && ctx->ws_W == W) return LLICTI_OK;
    return llicti_reserve(ctx, std::max(n, ctx->ws_H == H && ctx->ws_W == W ? ctx->ws_images : 0), H, W);
}

using namespace llicti;

extern "C" {

int llicti_abi_version(void) { return LLICTI_ABI_VERSION; }
const char *llicti_last_error(void) { return g_err; }

int llicti_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int llicti_geometry(const llicti_config *cfg, int H, int W, llicti_geom *out) {
    LLICTI_REQUIRE(cfg && out, "null argument");
    Plan p;
    const int rc = make_plan(*cfg, H, W, &p);
    if (rc) return rc;
    *out = p.g;
    return LLICTI_OK;
}

int llicti_create(const llicti_config *cfg, const llicti_weights *w, llicti_ctx **out) {
    LLICTI_REQUIRE(cfg && w && out, "null argument");
    LLICTI_REQUIRE(cfg->chs == 88 || cfg->chs == 60, "chs must be 88 or 60 (got %d)", cfg->chs);
    LLICTI_REQUIRE(cfg->num_mixtures == kM, "num_mixtures must be %d", kM);
    LLICTI_REQUIRE(cfg->num_scales >= 1 && cfg->num_scales <= LLICTI_MAX_SCALES, "num_scales out of range");
    LLICTI_REQUIRE(cfg->numerics == LLICTI_NUM_TORCH_CUDA || cfg->numerics == LLICTI_NUM_TORCH_CPU, "bad numerics profile");
    if (llicti_device_count() <= 0) {
        set_error("no CUDA device: libllicti_b200 has no CPU fallback");
        return LLICTI_E_NODEVICE;
    }
    LLICTI_CUDA(cudaSetDevice(cfg->device));
    llicti_ctx *ctx = new llicti_ctx();
    ctx->cfg = *cfg;
    ctx->num.div255_recip = cfg->numerics == LLICTI_NUM_TORCH_CUDA;
    ctx->num.sum_ilp4 = cfg->numerics == LLICTI_NUM_TORCH_CUDA;
    memset(ctx->wf32, 0, sizeof(ctx->wf32));
    int rc = pack_weights(ctx, *w);
    if (rc == LLICTI_OK) rc = tc_pack_weights(ctx, *w);
    if (rc == LLICTI_OK) {
        cudaStream_t s2 = nullptr;
        cudaEvent_t e1 = nullptr, e2 = nullptr;
        if (cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&e1, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&e2, cudaEventDisableTiming) != cudaSuccess) {
            set_error("cannot create the side stream of the wavefront decode");
            rc = LLICTI_E_CUDA;
        }
        ctx->side_stream = s2; ctx->ev_fork = e1; ctx->ev_join = e2;
        if (rc == LLICTI_OK) rc = probe_concurrent_kernels(ctx, &ctx->concurrent_kernels);
        if (rc == LLICTI_OK) rc = apply_decode_test_knobs();
    }
    if (rc == LLICTI_OK) {
        cudaError_t e = cudaMalloc((void **)&ctx->d_status, sizeof(int32_t));
        if (e == cudaSuccess) e = cudaMemset(ctx->d_status, 0, sizeof(int32_t));
        if (e != cudaSuccess) { set_error("cudaMalloc(status): %s", cudaGetErrorString(e)); rc = LLICTI_E_CUDA; }
    }
    if (rc != LLICTI_OK) { llicti_destroy(ctx); return rc; }
    *out = ctx;
    return LLICTI_OK;
}

void llicti_destroy(llicti_ctx *ctx) {
    if (!ctx) return;
    free_workspace(ctx);
    for (auto &b : ctx->wf32) { cudaFree(b.w0); cudaFree(b.b0); cudaFree(b.w1); cudaFree(b.b1); cudaFree(b.w2); cudaFree(b.b2); }
    tc_free_weights(ctx);
    train_free(ctx);
    if (ctx->ev_fork) cudaEventDestroy((cudaEvent_t)ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy((cudaEvent_t)ctx->ev_join);
    if (ctx->side_stream) cudaStreamDestroy((cudaStream_t)ctx->side_stream);
    for (void *&e : ctx->ev_pipe) { if (e) cudaEventDestroy((cudaEvent_t)e); e = nullptr; }
    if (ctx->copy_in) cudaStreamDestroy((cudaStream_t)ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy((cudaStream_t)ctx->copy_out);
    drop_decode_graph(ctx);
    if (ctx->dec_capture_stream) cudaStreamDestroy((cudaStream_t)ctx->dec_capture_stream);
    cudaFree(ctx->d_status);
    delete ctx;
}

int llicti_reserve(llicti_ctx *ctx, int max_images, int H, int W) {
    LLICTI_REQUIRE(ctx && max_images >= 1, "bad argument");
    Plan p;
    int rc = make_plan(ctx->cfg, H, W, &p);
    if (rc) return rc;
    free_workspace(ctx);
    ctx->plan = p;
    const llicti_geom &g = p.g;
    const size_t n = (size_t)max_images;
    const int S = g.num_scales;
    LLICTI_CUDA(cudaMalloc((void **)&ctx->d_sd, sizeof(StreamDesc) * kMaxStreams));
    LLICTI_CUDA(cudaMemcpy(ctx->d_sd, p.sd, sizeof(StreamDesc) * kMaxStreams, cudaMemcpyHostToDevice));
    LLICTI_CUDA(cudaMalloc((void **)&ctx->d_rgb, n * 3 * H * W));
    for (int s = 0; s < S; ++s) LLICTI_CUDA(cudaMalloc((void **)&ctx->d_planes[s], n * p.plane_elems[s] * sizeof(int16_t)));
    LLICTI_CUDA(cudaMalloc((void **)&ctx->d_minmax, n * 4 * sizeof(int32_t)));
    LLICTI_CUDA(cudaMalloc((void **)&ctx->d_minmax16, n * 6 * sizeof(int16_t)));
    // the wavefront decode keeps the network outputs, windows and symbols of all three bands of a scale in flight
    const size_t wb = (size_t)wave_bands_in_workspace(ctx->cfg, max_images);
    LLICTI_CUDA(cudaMalloc((void **)&ctx->d_params, wb * n * kParamCh * (size_t)g.Hs[0] * g.Ws[0] * sizeof(float)));
    LLICTI_CUDA(cudaMalloc((void **)&ctx->d_bounds, n * (size_t)g.symbols * sizeof(uint32_t)));
    LLICTI_CUDA(cudaMalloc((void **)&ctx->d_scratch, n * (size_t)p.scratch_bytes));
    LLICTI_CUDA(cudaMalloc((void **)&ctx->d_sublen, n * (size_t)g.substreams * sizeof(uint32_t)));
    LLICTI_CUDA(cudaMalloc((void **)&ctx->d_suboff, n * (size_t)g.substreams * sizeof(uint64_t)));
    LLICTI_CUDA(cudaMalloc((void **)&ctx->d_stream_bytes, (n * p.n_streams + 1) * sizeof(uint64_t)));
    LLICTI_CUDA(cudaMalloc((void **)&ctx->d_stream_off, (n * p.n_streams + 4) * sizeof(uint64_t)));   // + 3: every part of the pipelined host path ends with a total of its own
    ctx->blob_cap = n * (size_t)g.max_stream_bytes;
    LLICTI_CUDA(cudaMalloc((void **)&ctx->d_blob, ctx->blob_cap));
    LLICTI_CUDA(cudaMalloc((void **)&ctx->d_x00, n * 3 * (size_t)g.Hs[S - 1] * g.Ws[S - 1]));
    // window items, their flags and the compact symbol arrays serve the windowed decode schedules: every band of
    // torchac-compatible streams, and the bands of a substream container with too few chains for the group schedule
    const bool windows = ctx->cfg.decode_impl != 1;
    ctx->wave_ws = false;
    if (windows) {
        ctx->items_cap = std::max<int64_t>(decode_items_capacity(ctx->cfg, p, max_images), 1);
        LLICTI_CUDA(cudaMalloc(&ctx->d_items, wb * (size_t)ctx->items_cap * decode_item_bytes()));
        LLICTI_CUDA(cudaMalloc((void **)&ctx->d_item_flags, (size_t)decode_flag_words(wb * ctx->items_cap) * sizeof(uint32_t)));
        LLICTI_CUDA(cudaMalloc(&ctx->d_chain_state_raw, n * 9 * 64));
        ctx->wave_ws = wb == 3;
        ctx->sym_cap = ((int64_t)g.Hs[0] * g.Ws[0] + 63) / 64 * 64;
        LLICTI_CUDA(cudaMalloc((void **)&ctx->d_syms, wb * n * 3 * (size_t)ctx->sym_cap * sizeof(int16_t)));
    }
    ctx->ws_images = max_images; ctx->ws_H = H; ctx->ws_W = W;
    return LLICTI_OK;
}

int64_t llicti_launch_count(const llicti_ctx *ctx) { return ctx ? ctx->launches : 0; }

int llicti_cnn_operands(const llicti_ctx *ctx) {
    if (!ctx) return -1;
    return ctx->cfg.cnn_impl == LLICTI_CNN_TCGEN05 ? tc_operand_type(ctx) : 0;
}

int llicti_status(llicti_ctx *ctx, void *stream) {
    LLICTI_REQUIRE(ctx, "null context");
    const int rc = read_status(ctx, (cudaStream_t)stream);
    if (rc == LLICTI_E_TIMEOUT && !ctx->no_coresidency) {      // the next decode takes the schedule without hand-overs
        ctx->no_coresidency = true;
        drop_decode_graph(ctx);
    }
    return rc;
}

int llicti_profile(llicti_ctx *ctx, int enable) {
    LLICTI_REQUIRE(ctx, "null context");
    for (cudaEvent_t e : ctx->prof_ev) cudaEventDestroy(e);
    ctx->prof_ev.clear();
    ctx->prof_cls.clear();
    ctx->prof_on = enable != 0;
    return LLICTI_OK;
}

int llicti_profile_read(llicti_ctx *ctx, double *ms, int64_t *count) {
    LLICTI_REQUIRE(ctx && ms && count, "null argument");
    for (int i = 0; i < LLICTI_KERNEL_CLASSES; ++i) { ms[i] = 0; count[i] = 0; }
    for (size_t i = 0; i < ctx->prof_cls.size(); ++i) {
        LLICTI_CUDA(cudaEventSynchronize(ctx->prof_ev[2 * i + 1]));
        float t = 0.f;
        LLICTI_CUDA(cudaEventElapsedTime(&t, ctx->prof_ev[2 * i], ctx->prof_ev[2 * i + 1]));
        ms[ctx->prof_cls[i]] += t;
        count[ctx->prof_cls[i]] += 1;
        if (getenv("LLICTI_PROF_DUMP")) fprintf(stderr, "[llicti profile] scope %zu class %d %.4f ms\n", i, ctx->prof_cls[i], t);
    }
    for (cudaEvent_t e : ctx->prof_ev) cudaEventDestroy(e);
    ctx->prof_ev.clear();
    ctx->prof_cls.clear();
    return LLICTI_OK;
}

int llicti_decode_stats(llicti_ctx *ctx, uint64_t *out8, int reset) {
    LLICTI_REQUIRE(ctx && out8, "null argument");
    LLICTI_CUDA(cudaDeviceSynchronize());
    return read_decode_stats(out8, reset);
}

int llicti_selftest_fdiv(llicti_ctx *ctx, int64_t n_pairs, uint64_t seed, uint64_t *mismatches) {
    LLICTI_REQUIRE(ctx && mismatches && n_pairs > 0, "bad argument");
    unsigned long long *d = nullptr, h = 0;
    LLICTI_CUDA(cudaMalloc((void **)&d, sizeof(h)));
    LLICTI_CUDA(cudaMemset(d, 0, sizeof(h)));
    int rc = launch_selftest_fdiv(ctx, n_pairs, seed, d, nullptr);
    if (rc == LLICTI_OK) {
        cudaError_t e = cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { set_error("selftest: %s", cudaGetErrorString(e)); rc = LLICTI_E_CUDA; }
    }
    cudaFree(d);
    *mismatches = h;
    return rc;
}

// ---- stage-level ------------------------------------------------------------------------
int llicti_color_split(llicti_ctx *ctx, const uint8_t *rgb_dev, int n, int H, int W, int16_t *const *planes_dev,
                       int32_t *minmax_dev, void *stream) {
    LLICTI_REQUIRE(ctx && rgb_dev && planes_dev && minmax_dev && n >= 1, "bad argument");
    Plan p;
    int rc = make_plan(ctx->cfg, H, W, &p);
    if (rc) return rc;
    return launch_color_split(ctx, p, rgb_dev, n, planes_dev, minmax_dev, (cudaStream_t)stream);
}

int llicti_merge_color(llicti_ctx *ctx, const int16_t *planes0_dev, int n, int H, int W, uint8_t *rgb_dev, void *stream) {
    LLICTI_REQUIRE(ctx && planes0_dev && rgb_dev && n >= 1, "bad argument");
    Plan p;
    int rc = make_plan(ctx->cfg, H, W, &p);
    if (rc) return rc;
    return launch_merge_color(ctx, p, planes0_dev, n, rgb_dev, (cudaStream_t)stream);
}

int llicti_cnn_params(llicti_ctx *ctx, int band, const int16_t *planes_dev, int n, int Hs, int Ws, float *params_dev,
                      void *stream) {
    LLICTI_REQUIRE(ctx && planes_dev && params_dev && band >= 0 && band < 3 && n >= 1 && Hs >= 1 && Ws >= 1, "bad argument");
    return cnn(ctx, band, planes_dev, n, Hs, Ws, params_dev, (cudaStream_t)stream);
}

int llicti_cdf_table(llicti_ctx *ctx, const float *params_dev, const int16_t *yband_dev, int clr, int min_val, int max_val,
                     int P, int16_t *table_dev, void *stream) {
    LLICTI_REQUIRE(ctx && params_dev && yband_dev && table_dev && clr >= 0 && clr < 3 && max_val >= min_val && P >= 1, "bad argument");
    return launch_cdf_table(ctx, params_dev, yband_dev, clr, min_val, max_val, P, table_dev, (cudaStream_t)stream);
}

int llicti_cdf_bounds(llicti_ctx *ctx, const float *params_dev, const int16_t *yband_dev, int clr, int min_val, int max_val,
                      int P, uint32_t *bounds_dev, void *stream) {
    LLICTI_REQUIRE(ctx && params_dev && yband_dev && bounds_dev && clr >= 0 && clr < 3 && max_val >= min_val && P >= 1, "bad argument");
    return launch_cdf_bounds_flat(ctx, params_dev, yband_dev, clr, min_val, max_val, P, bounds_dev, (cudaStream_t)stream);
}

int llicti_ac_encode_bounds(llicti_ctx *ctx, const uint32_t *bounds_dev, int n_sym, int S, uint8_t *out_dev, int slot_bytes,
                            uint32_t *lens_dev, void *stream) {
    LLICTI_REQUIRE(ctx && bounds_dev && out_dev && lens_dev && n_sym >= 1 && S >= 1 && slot_bytes >= 16 && slot_bytes % 4 == 0,
                   "bad argument (slot_bytes must be a multiple of 4, >= 16)");
    return launch_encode_flat(ctx, bounds_dev, n_sym, S, out_dev, slot_bytes, lens_dev, (cudaStream_t)stream);
}

int llicti_ac_decode_table(llicti_ctx *ctx, const int16_t *table_dev, int n_sym, int Lp, int S, const uint8_t *in_dev,
                           const uint32_t *offs_dev, int16_t *sym_dev, void *stream) {
    LLICTI_REQUIRE(ctx && table_dev && in_dev && offs_dev && sym_dev && n_sym >= 1 && Lp >= 2 && S >= 1, "bad argument");
    return launch_decode_table(ctx, table_dev, n_sym, Lp, S, in_dev, offs_dev, sym_dev, (cudaStream_t)stream);
}

// ---- full path ----------------------------------------------------------------------------
int llicti_encode_dev(llicti_ctx *ctx, const uint8_t *rgb_dev, int n, int H, int W, uint8_t *out_dev, size_t out_cap,
                      uint64_t *stream_off_dev, int16_t *minmax_dev, void *stream) {
    int rc = check_batch(ctx, n, H, W);
    if (rc) return rc;
    LLICTI_REQUIRE(rgb_dev && out_dev && stream_off_dev && minmax_dev, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const Plan &p = ctx->plan;
    const llicti_geom &g = p.g;
    if ((rc = launch_color_split(ctx, p, rgb_dev, n, ctx->d_planes, ctx->d_minmax, st))) return rc;
    for (int s = g.num_scales - 1; s >= 0; --s)
        for (int b = 0; b < 3; ++b) {
            if ((rc = cnn(ctx, b, ctx->d_planes[s], n, g.Hs[s], g.Ws[s], ctx->d_params, st))) return rc;
            if ((rc = launch_band_bounds(ctx, p, s, b, ctx->d_params, ctx->d_planes[s], ctx->d_minmax, n, ctx->d_bounds,
                                         g.symbols, st)))
                return rc;
        }
    if ((rc = launch_encode_all(ctx, p, ctx->d_bounds, g.symbols, n, ctx->d_scratch, p.scratch_bytes, ctx->d_sublen, st))) return rc;
    if ((rc = launch_compact(ctx, p, n, ctx->d_scratch, p.scratch_bytes, ctx->d_sublen, ctx->d_stream_bytes, stream_off_dev,
                             out_dev, out_cap, st)))
        return rc;
    return launch_minmax16(ctx, ctx->d_minmax, minmax_dev, n, st);
}

// LLICTI.forward for a batch: float lifting + un-padded pyramid, then per scale and band the CNN and the point
// likelihood of every sample.  Uses the workspace's int16 planes and parameter buffer.
int llicti_forward_dev(llicti_ctx *ctx, const uint8_t *rgb_dev, int n, int H, int W, float *const *fplanes_dev,
                       float *const *sinfo_dev, void *stream) {
    int rc = check_batch(ctx, n, H, W);
    if (rc) return rc;
    LLICTI_REQUIRE(rgb_dev && fplanes_dev && sinfo_dev, "null argument");
    const Plan &p = ctx->plan;
    const llicti_geom &g = p.g;
    LLICTI_REQUIRE(H % (1 << g.num_scales) == 0 && W % (1 << g.num_scales) == 0,
                   "forward() needs H and W to be multiples of 2^num_scales = %d (the reference's un-padded lazyDWT, "
                   "LLICTI_nets.py:218-241, concatenates phases of equal size; validate() pads its images, "
                   "agents/llicti_agent.py:105-116)", 1 << g.num_scales);
    cudaStream_t st = (cudaStream_t)stream;
    for (int s = 0; s < g.num_scales; ++s) LLICTI_REQUIRE(fplanes_dev[s] && sinfo_dev[s], "null plane pointer");
    if ((rc = launch_color_split_float(ctx, p, rgb_dev, n, fplanes_dev, ctx->d_planes, st))) return rc;
    for (int s = 0; s < g.num_scales; ++s)
        for (int b = 0; b < 3; ++b) {
            // fp32 contexts feed the CNN the float lifting's own values (what the reference's convolutions see); the tcgen05
            // CNN reads the int16 planes (value * 255, the same numbers up to an ulp of the quotient)
            if (ctx->cfg.cnn_impl == LLICTI_CNN_FP32) rc = launch_cnn_forward_train(ctx, b, fplanes_dev[s], n, g.Hs[s], g.Ws[s], ctx->d_params, st);
            else rc = cnn(ctx, b, ctx->d_planes[s], n, g.Hs[s], g.Ws[s], ctx->d_params, st);
            if (rc) return rc;
            if ((rc = launch_self_info(ctx, ctx->d_params, fplanes_dev[s], b, n, g.Hs[s] * g.Ws[s], sinfo_dev[s], st))) return rc;
        }
    return LLICTI_OK;
}

// ---- training step ------------------------------------------------------------------
// New fp32 weights from device memory (the optimizer's parameters): repacked in place by three small kernels per band.
// Only for contexts of the fp32 CNN: the tcgen05 operand pack, the fp16-range proof and the stream fingerprint belong to
// the weights given at creation.
int llicti_set_weights_dev(llicti_ctx *ctx, const llicti_weights *w_dev, void *stream) {
    LLICTI_REQUIRE(ctx && w_dev, "null argument");
    LLICTI_REQUIRE(ctx->cfg.cnn_impl == LLICTI_CNN_FP32,
                   "llicti_set_weights_dev needs a context created with cnn_impl = LLICTI_CNN_FP32 (the training context); "
                   "create a new context to code with the trained weights");
    ctx->weights_from_device = true;
    return launch_train_layouts(ctx, *w_dev, true, (cudaStream_t)stream);
}

// Offset (in floats) of band b of scale s in the kept-parameter buffer of a training step: blocks [n][60][P_s] in
// (scale, band) order.
static size_t kept_offset(const llicti_geom &g, int n, int s, int b) {
    size_t pos = 0;
    for (int t = 0; t < s; ++t) pos += (size_t)3 * g.Hs[t] * g.Ws[t];
    return (size_t)kParamCh * n * (pos + (size_t)b * g.Hs[s] * g.Ws[s]);
}

// forward() of a training step: as llicti_forward_dev on an fp32 context, but every band's 60 network outputs stay in
// params_keep_dev (180 * n * sum_s Hs*Ws floats) for the backward pass, which then neither repeats the colour split nor the CNN.
int llicti_train_forward_dev(llicti_ctx *ctx, const uint8_t *rgb_dev, int n, int H, int W, float *const *fplanes_dev,
                             float *const *sinfo_dev, float *params_keep_dev, void *stream) {
    int rc = check_batch(ctx, n, H, W);
    if (rc) return rc;
    LLICTI_REQUIRE(rgb_dev && fplanes_dev && sinfo_dev && params_keep_dev, "null argument");
    LLICTI_REQUIRE(ctx->cfg.cnn_impl == LLICTI_CNN_FP32, "llicti_train_forward_dev needs a context created with cnn_impl = LLICTI_CNN_FP32");
    const Plan &p = ctx->plan;
    const llicti_geom &g = p.g;
    LLICTI_REQUIRE(H % (1 << g.num_scales) == 0 && W % (1 << g.num_scales) == 0,
                   "forward() needs H and W to be multiples of 2^num_scales = %d", 1 << g.num_scales);
    cudaStream_t st = (cudaStream_t)stream;
    for (int s = 0; s < g.num_scales; ++s) LLICTI_REQUIRE(fplanes_dev[s] && sinfo_dev[s], "null plane pointer");
    if ((rc = launch_color_split_float(ctx, p, rgb_dev, n, fplanes_dev, ctx->d_planes, st))) return rc;
    for (int s = 0; s < g.num_scales; ++s)
        for (int b = 0; b < 3; ++b) {
            float *params = params_keep_dev + kept_offset(g, n, s, b);
            if ((rc = launch_cnn_forward_train(ctx, b, fplanes_dev[s], n, g.Hs[s], g.Ws[s], params, st))) return rc;
            if ((rc = launch_self_info(ctx, params, fplanes_dev[s], b, n, g.Hs[s] * g.Ws[s], sinfo_dev[s], st))) return rc;
        }
    return LLICTI_OK;
}

// d loss / d weights for a loss whose gradient with respect to forward()'s self-informations is gsinfo_dev.  Per band:
// the network outputs (kept by llicti_train_forward_dev, or -- params_kept_dev == NULL -- recomputed: colour split and fp32
// CNN again), the likelihood differentiated in place over them, the CNN backward accumulating packed weight gradients,
// which are written to grads_dev in PyTorch's layouts at the end.
int llicti_backward_dev(llicti_ctx *ctx, const uint8_t *rgb_dev, int n, int H, int W, float *const *fplanes_dev,
                        const float *const *gsinfo_dev, float *params_kept_dev, const llicti_weights *grads_dev, void *stream) {
    int rc = check_batch(ctx, n, H, W);
    if (rc) return rc;
    LLICTI_REQUIRE(rgb_dev && fplanes_dev && gsinfo_dev && grads_dev, "null argument");
    LLICTI_REQUIRE(ctx->cfg.cnn_impl == LLICTI_CNN_FP32, "llicti_backward_dev needs a context created with cnn_impl = LLICTI_CNN_FP32");
    const Plan &p = ctx->plan;
    const llicti_geom &g = p.g;
    LLICTI_REQUIRE(H % (1 << g.num_scales) == 0 && W % (1 << g.num_scales) == 0,
                   "backward() needs H and W to be multiples of 2^num_scales = %d, as forward()", 1 << g.num_scales);
    cudaStream_t st = (cudaStream_t)stream;
    for (int s = 0; s < g.num_scales; ++s) LLICTI_REQUIRE(fplanes_dev[s] && gsinfo_dev[s], "null plane pointer");
    if ((rc = launch_train_zero_grads(ctx, st))) return rc;
    if (!params_kept_dev && (rc = launch_color_split_float(ctx, p, rgb_dev, n, fplanes_dev, ctx->d_planes, st))) return rc;
    for (int s = 0; s < g.num_scales; ++s)
        for (int b = 0; b < 3; ++b) {
            float *params = params_kept_dev ? params_kept_dev + kept_offset(g, n, s, b) : ctx->d_params;
            if (!params_kept_dev && (rc = launch_cnn_forward_train(ctx, b, fplanes_dev[s], n, g.Hs[s], g.Ws[s], params, st))) return rc;
            if ((rc = launch_self_info_grad(ctx, params, fplanes_dev[s], gsinfo_dev[s], b, n, g.Hs[s] * g.Ws[s], st))) return rc;
            if ((rc = launch_cnn_backward(ctx, b, fplanes_dev[s], n, g.Hs[s], g.Ws[s], params, st))) return rc;
        }
    return launch_train_layouts(ctx, *grads_dev, false, st);
}

// ---- host-buffer entry points, pipelined ----------------------------------------------------------
// A large batch of the substream container is coded as two half batches: the host->device copy of the second half runs
// while the first half is coded, the device->host copy of the first half while the second is coded (two copy streams next
// to the caller's).  The workspace is shared (the halves' kernels are serialised on the caller's stream); only the input,
// output and offset buffers are split.  Everything the caller sees (contiguous blob, global offsets) is as from one batch.
static bool host_pipelined(const llicti_ctx *ctx, int n, int H, int W) {
    const char *force = getenv("LLICTI_HOST_PIPELINE");              // "1": whenever possible (tests), "0": never
    if (force && *force) return atoi(force) != 0 && ctx->cfg.sub_len > 0 && n >= 2;
    return ctx->cfg.sub_len > 0 && n >= 16 && (size_t)n * 3 * H * W >= ((size_t)64 << 20);
}
static int ensure_copy_streams(llicti_ctx *ctx) {
    if (ctx->copy_in) return LLICTI_OK;
    cudaStream_t a = nullptr, b = nullptr;
    LLICTI_CUDA(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
    LLICTI_CUDA(cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking));
    ctx->copy_in = a; ctx->copy_out = b;
    for (void *&e : ctx->ev_pipe) {
        cudaEvent_t ev = nullptr;
        LLICTI_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        e = ev;
    }
    return LLICTI_OK;
}

static int encode_host_pipelined(llicti_ctx *ctx, const uint8_t *rgb, int n, int H, int W, uint8_t *out, size_t out_cap,
                                 uint64_t *stream_off, int16_t *minmax, cudaStream_t st) {
    int rc = ensure_copy_streams(ctx);
    if (rc) return rc;
    cudaStream_t cin = (cudaStream_t)ctx->copy_in, cout = (cudaStream_t)ctx->copy_out;
    constexpr int kMaxParts = 4;
    cudaEvent_t e_start = (cudaEvent_t)ctx->ev_pipe[0];
    cudaEvent_t eh[kMaxParts], ec[kMaxParts];
    for (int k = 0; k < kMaxParts; ++k) { eh[k] = (cudaEvent_t)ctx->ev_pipe[1 + k]; ec[k] = (cudaEvent_t)ctx->ev_pipe[5 + k]; }
    // parts of at least 16 images (the kernels of a part should still fill the machine), at most four
    const int parts = n >= 64 ? 4 : n >= 48 ? 3 : 2;
    const int ns = ctx->plan.n_streams;
    const size_t img = (size_t)3 * H * W, cap_img = (size_t)ctx->plan.g.max_stream_bytes;
    int first[kMaxParts + 1];
    for (int k = 0; k <= parts; ++k) first[k] = (int)((long long)n * k / parts);
    // part k's offsets live at d_stream_off + first[k] * ns + k (each part ends with its own total: one extra entry per part)
    LLICTI_CUDA(cudaMemsetAsync(ctx->d_status, 0, sizeof(int32_t), st));
    LLICTI_CUDA(cudaEventRecord(e_start, st));                       // the copy streams start after whatever precedes on st
    LLICTI_CUDA(cudaStreamWaitEvent(cin, e_start, 0));
    LLICTI_CUDA(cudaStreamWaitEvent(cout, e_start, 0));
    for (int k = 0; k < parts; ++k) {
        LLICTI_CUDA(cudaMemcpyAsync(ctx->d_rgb + first[k] * img, rgb + first[k] * img, (size_t)(first[k + 1] - first[k]) * img,
                                    cudaMemcpyHostToDevice, cin));
        LLICTI_CUDA(cudaEventRecord(eh[k], cin));
    }
    uint64_t total = 0;                                              // bytes of the parts whose offsets have arrived
    bool fits = true;
    for (int k = 0; k <= parts; ++k) {
        if (k < parts) {                                             // code part k (its input has arrived) ...
            const int n_k = first[k + 1] - first[k];
            LLICTI_CUDA(cudaStreamWaitEvent(st, eh[k], 0));
            if ((rc = llicti_encode_dev(ctx, ctx->d_rgb + first[k] * img, n_k, H, W, ctx->d_blob + first[k] * cap_img, n_k * cap_img,
                                        ctx->d_stream_off + ((size_t)first[k] * ns + k), ctx->d_minmax16 + (size_t)first[k] * 6, st)))
                return rc;
            LLICTI_CUDA(cudaEventRecord(ec[k], st));
        }
        if (k > 0) {                                                 // ... while part k - 1 leaves: offsets first (the host needs its total)
            const int j = k - 1, n_j = first[j + 1] - first[j];
            uint64_t *dst = stream_off + (size_t)first[j] * ns;      // global offsets: part j's entry 0 coincides with the total so far
            LLICTI_CUDA(cudaStreamWaitEvent(cout, ec[j], 0));
            LLICTI_CUDA(cudaMemcpyAsync(dst + 1, ctx->d_stream_off + ((size_t)first[j] * ns + j) + 1, (size_t)n_j * ns * sizeof(uint64_t),
                                        cudaMemcpyDeviceToHost, cout));
            LLICTI_CUDA(cudaMemcpyAsync(minmax + (size_t)first[j] * 6, ctx->d_minmax16 + (size_t)first[j] * 6, (size_t)n_j * 6 * sizeof(int16_t),
                                        cudaMemcpyDeviceToHost, cout));
            LLICTI_CUDA(cudaStreamSynchronize(cout));
            dst[0] = total;
            const uint64_t total_j = dst[(size_t)n_j * ns];
            for (size_t i = 1; i <= (size_t)n_j * ns; ++i) dst[i] += total;
            fits = fits && total + total_j <= out_cap;
            if (fits) LLICTI_CUDA(cudaMemcpyAsync(out + total, ctx->d_blob + first[j] * cap_img, total_j, cudaMemcpyDeviceToHost, cout));
            total += total_j;
        }
    }
    rc = read_status(ctx, st);
    LLICTI_CUDA(cudaStreamSynchronize(cout));
    if (rc) return rc;
    if (!fits) {
        set_error("output needs %llu bytes, capacity is %llu", (unsigned long long)total, (unsigned long long)out_cap);
        return LLICTI_E_NOMEM;
    }
    return LLICTI_OK;
}

static int decode_host_pipelined(llicti_ctx *ctx, const uint8_t *blob, const uint64_t *stream_off, const int16_t *minmax,
                                 const uint8_t *x00_rgb, int n, int H, int W, uint8_t *rgb_out, cudaStream_t st, bool *done) {
    *done = false;
    const Plan &p = ctx->plan;
    const int ns = p.n_streams, S = p.g.num_scales, n0 = (n + 1) / 2, n1 = n - n0;
    const size_t img = (size_t)3 * H * W, cap0 = (size_t)n0 * (size_t)p.g.max_stream_bytes;
    const size_t x00b = (size_t)3 * p.g.Hs[S - 1] * p.g.Ws[S - 1];
    const uint64_t total = stream_off[(size_t)n * ns], total0 = stream_off[(size_t)n0 * ns], total1 = total - total0;
    if (total0 > cap0 || total1 > ctx->blob_cap - cap0) return LLICTI_OK;          // (a half larger than any valid encoding: let the plain path report it)
    int rc = ensure_copy_streams(ctx);
    if (rc) return rc;
    cudaStream_t cin = (cudaStream_t)ctx->copy_in, cout = (cudaStream_t)ctx->copy_out;
    cudaEvent_t e_start = (cudaEvent_t)ctx->ev_pipe[0], eh0 = (cudaEvent_t)ctx->ev_pipe[1], eh1 = (cudaEvent_t)ctx->ev_pipe[2],
                ec0 = (cudaEvent_t)ctx->ev_pipe[5], ec1 = (cudaEvent_t)ctx->ev_pipe[6];
    std::vector<uint64_t> off1((size_t)n1 * ns + 1);                 // the second half's offsets, relative to its own bytes
    for (size_t i = 0; i < off1.size(); ++i) off1[i] = stream_off[(size_t)n0 * ns + i] - total0;
    uint64_t *off1_dev = ctx->d_stream_off + ((size_t)n0 * ns + 1);
    LLICTI_CUDA(cudaMemsetAsync(ctx->d_status, 0, sizeof(int32_t), st));
    LLICTI_CUDA(cudaEventRecord(e_start, st));
    LLICTI_CUDA(cudaStreamWaitEvent(cin, e_start, 0));
    LLICTI_CUDA(cudaStreamWaitEvent(cout, e_start, 0));
    LLICTI_CUDA(cudaMemcpyAsync(ctx->d_blob, blob, total0, cudaMemcpyHostToDevice, cin));
    LLICTI_CUDA(cudaMemcpyAsync(ctx->d_stream_off, stream_off, ((size_t)n0 * ns + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, cin));
    LLICTI_CUDA(cudaMemcpyAsync(ctx->d_minmax16, minmax, (size_t)n * 6 * sizeof(int16_t), cudaMemcpyHostToDevice, cin));
    LLICTI_CUDA(cudaMemcpyAsync(ctx->d_x00, x00_rgb, (size_t)n * x00b, cudaMemcpyHostToDevice, cin));
    LLICTI_CUDA(cudaEventRecord(eh0, cin));
    LLICTI_CUDA(cudaMemcpyAsync(ctx->d_blob + cap0, blob + total0, total1, cudaMemcpyHostToDevice, cin));
    LLICTI_CUDA(cudaMemcpyAsync(off1_dev, off1.data(), off1.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, cin));
    LLICTI_CUDA(cudaEventRecord(eh1, cin));
    LLICTI_CUDA(cudaStreamWaitEvent(st, eh0, 0));
    if ((rc = llicti_decode_dev(ctx, ctx->d_blob, ctx->d_stream_off, ctx->d_minmax16, ctx->d_x00, n0, H, W, ctx->d_rgb, st))) return rc;
    LLICTI_CUDA(cudaEventRecord(ec0, st));
    LLICTI_CUDA(cudaStreamWaitEvent(cout, ec0, 0));
    LLICTI_CUDA(cudaMemcpyAsync(rgb_out, ctx->d_rgb, n0 * img, cudaMemcpyDeviceToHost, cout));
    LLICTI_CUDA(cudaStreamWaitEvent(st, eh1, 0));
    if ((rc = llicti_decode_dev(ctx, ctx->d_blob + cap0, off1_dev, ctx->d_minmax16 + (size_t)n0 * 6, ctx->d_x00 + n0 * x00b, n1, H, W,
                                ctx->d_rgb + n0 * img, st)))
        return rc;
    LLICTI_CUDA(cudaEventRecord(ec1, st));
    LLICTI_CUDA(cudaStreamWaitEvent(cout, ec1, 0));
    LLICTI_CUDA(cudaMemcpyAsync(rgb_out + n0 * img, ctx->d_rgb + n0 * img, n1 * img, cudaMemcpyDeviceToHost, cout));
    rc = read_status(ctx, st);
    LLICTI_CUDA(cudaStreamSynchronize(cin));
    LLICTI_CUDA(cudaStreamSynchronize(cout));
    *done = true;
    return rc;
}

int llicti_encode_host(llicti_ctx *ctx, const uint8_t *rgb, int n, int H, int W, uint8_t *out, size_t out_cap,
                       uint64_t *stream_off, int16_t *minmax, void *stream) {
    int rc = check_batch(ctx, n, H, W);
    if (rc) return rc;
    LLICTI_REQUIRE(rgb && out && stream_off && minmax, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int ns = ctx->plan.n_streams;
    if (host_pipelined(ctx, n, H, W)) return encode_host_pipelined(ctx, rgb, n, H, W, out, out_cap, stream_off, minmax, st);
    LLICTI_CUDA(cudaMemsetAsync(ctx->d_status, 0, sizeof(int32_t), st));     // a flag left by an earlier *_dev call is not this call's
    LLICTI_CUDA(cudaMemcpyAsync(ctx->d_rgb, rgb, (size_t)n * 3 * H * W, cudaMemcpyHostToDevice, st));
    if ((rc = llicti_encode_dev(ctx, ctx->d_rgb, n, H, W, ctx->d_blob, ctx->blob_cap, ctx->d_stream_off, ctx->d_minmax16, st)))
        return rc;
    LLICTI_CUDA(cudaMemcpyAsync(stream_off, ctx->d_stream_off, ((size_t)n * ns + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    LLICTI_CUDA(cudaMemcpyAsync(minmax, ctx->d_minmax16, (size_t)n * 6 * sizeof(int16_t), cudaMemcpyDeviceToHost, st));
    if ((rc = read_status(ctx, st))) return rc;
    const uint64_t total = stream_off[(size_t)n * ns];
    if (total > out_cap) {
        set_error("output needs %llu bytes, capacity is %llu", (unsigned long long)total, (unsigned long long)out_cap);
        return LLICTI_E_NOMEM;
    }
    LLICTI_CUDA(cudaMemcpyAsync(out, ctx->d_blob, total, cudaMemcpyDeviceToHost, st));
    LLICTI_CUDA(cudaStreamSynchronize(st));
    return LLICTI_OK;
}

static int decode_dev_launches(llicti_ctx *ctx, const uint8_t *blob_dev, const uint64_t *stream_off_dev, const int16_t *minmax_dev,
                               const uint8_t *x00_rgb_dev, int n, uint8_t *rgb_out_dev, cudaStream_t st);

int llicti_decode_dev(llicti_ctx *ctx, const uint8_t *blob_dev, const uint64_t *stream_off_dev, const int16_t *minmax_dev,
                      const uint8_t *x00_rgb_dev, int n, int H, int W, uint8_t *rgb_out_dev, void *stream) {
    int rc = check_batch(ctx, n, H, W);
    if (rc) return rc;
    LLICTI_REQUIRE(blob_dev && stream_off_dev && minmax_dev && x00_rgb_dev && rgb_out_dev, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    // A decode is some hundred small launches on two streams; replayed as one graph launch it does not depend on how
    // fast the host thread issues them (eight ranks on one box: 38 -> 32 ms per batch on the slowest rank).  Not while
    // the per-class profile is on (its events are created per call).
    const bool use_graph = !ctx->prof_on && !getenv("LLICTI_NO_GRAPH");
    if (!use_graph) return decode_dev_launches(ctx, blob_dev, stream_off_dev, minmax_dev, x00_rgb_dev, n, rgb_out_dev, st);
    const void *kp[5] = {blob_dev, stream_off_dev, minmax_dev, x00_rgb_dev, rgb_out_dev};
    // the schedule knobs of kernels_decode.cu (debugging aids read from the environment) are part of the key
    static const char *const knobs[] = {"LLICTI_WAVE_STRIP_ROWS", "LLICTI_WAVE_MAX_STRIPS", "LLICTI_NO_WAVE", "LLICTI_NO_PIPE",
                                        "LLICTI_WAVE_CHAINS_PER_CTA", "LLICTI_WAVE_SHARE_SMS", "LLICTI_WAVE_PATTERN",
                                        "LLICTI_WAVE_PRODUCER_CTAS_PER_SM", "LLICTI_PIPE_CONS_PER_SM", "LLICTI_PIPE_CTAS_PER_SM",
                                        "LLICTI_WAVE_DEBUG", "LLICTI_NO_HALVES"};
    unsigned long long h = 1469598103934665603ull;
    for (const char *k : knobs) {
        const char *v = getenv(k);
        for (const char *c = v ? v : "-"; *c; ++c) h = (h ^ (unsigned char)*c) * 1099511628211ull;
        h = (h ^ 0xFFu) * 1099511628211ull;
    }
    h = (h ^ (ctx->no_coresidency ? 0x55u : 0xAAu)) * 1099511628211ull;
    const long long kd[4] = {n, H, W, (long long)h};
    bool same = ctx->dec_key_seen;
    for (int i = 0; i < 5; ++i) same = same && kp[i] == ctx->dec_key_ptr[i];
    for (int i = 0; i < 4; ++i) same = same && kd[i] == ctx->dec_key_dim[i];
    if (!same) {                                  // first call with these arguments: run it, remember them
        drop_decode_graph(ctx);
        for (int i = 0; i < 5; ++i) ctx->dec_key_ptr[i] = kp[i];
        for (int i = 0; i < 4; ++i) ctx->dec_key_dim[i] = kd[i];
        ctx->dec_key_seen = true;
        return decode_dev_launches(ctx, blob_dev, stream_off_dev, minmax_dev, x00_rgb_dev, n, rgb_out_dev, st);
    }
    if (!ctx->dec_graph_exec) {                   // second call: capture (on a stream of our own: `st` may be the legacy stream)
        if (!ctx->dec_capture_stream) {
            cudaStream_t cs = nullptr;
            LLICTI_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
            ctx->dec_capture_stream = cs;
        }
        cudaStream_t cs = (cudaStream_t)ctx->dec_capture_stream;
        const int64_t l0 = ctx->launches;
        LLICTI_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed));
        rc = decode_dev_launches(ctx, blob_dev, stream_off_dev, minmax_dev, x00_rgb_dev, n, rgb_out_dev, cs);
        cudaGraph_t graph = nullptr;
        const cudaError_t e = cudaStreamEndCapture(cs, &graph);
        ctx->dec_graph_launches = ctx->launches - l0;
        ctx->launches = l0;
        if (rc || e != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            if (rc) return rc;
            return decode_dev_launches(ctx, blob_dev, stream_off_dev, minmax_dev, x00_rgb_dev, n, rgb_out_dev, st);   // not capturable here
        }
        cudaGraphExec_t exec = nullptr;
        const cudaError_t e2 = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e2 != cudaSuccess) {
            cudaGetLastError();
            return decode_dev_launches(ctx, blob_dev, stream_off_dev, minmax_dev, x00_rgb_dev, n, rgb_out_dev, st);
        }
        ctx->dec_graph_exec = exec;
    }
    LLICTI_CUDA(cudaGraphLaunch((cudaGraphExec_t)ctx->dec_graph_exec, st));
    ctx->launches += ctx->dec_graph_launches;
    return LLICTI_OK;
}

static int decode_dev_launches(llicti_ctx *ctx, const uint8_t *blob_dev, const uint64_t *stream_off_dev, const int16_t *minmax_dev,
                               const uint8_t *x00_rgb_dev, int n, uint8_t *rgb_out_dev, cudaStream_t st) {
    int rc;
    const Plan &p = ctx->plan;
    const llicti_geom &g = p.g;
    const int S = g.num_scales;
    if ((rc = launch_minmax32(ctx, minmax_dev, ctx->d_minmax, n, st))) return rc;
    // no valid encoding of n images exceeds n * max_stream_bytes (the encoder's own capacity): offsets beyond it are malformed
    if ((rc = launch_index_streams(ctx, p, n, blob_dev, (uint64_t)n * (uint64_t)g.max_stream_bytes, stream_off_dev, ctx->d_suboff,
                                   ctx->d_sublen, st)))
        return rc;
    if ((rc = launch_x00_from_header(ctx, p, x00_rgb_dev, n, ctx->d_planes[S - 1], st))) return rc;
    for (int s = S - 1; s >= 0; --s) {
        if (wave_eligible(ctx, p, s, n)) {       // the three bands concurrently, two strips of rows apart
            if ((rc = launch_decode_scale_wave(ctx, p, s, ctx->d_planes[s], ctx->d_minmax, n, blob_dev, ctx->d_suboff,
                                               ctx->d_sublen, st)))
                return rc;
        } else {
            for (int b = 0; b < 3; ++b) {
                if ((rc = cnn(ctx, b, ctx->d_planes[s], n, g.Hs[s], g.Ws[s], ctx->d_params, st))) return rc;
                if ((rc = launch_decode_band(ctx, p, s, b, ctx->d_params, ctx->d_planes[s], ctx->d_minmax, n, blob_dev,
                                             ctx->d_suboff, ctx->d_sublen, st)))
                    return rc;
            }
        }
        if (s > 0 && (rc = launch_interleave(ctx, p, s, ctx->d_planes[s], ctx->d_planes[s - 1], n, st))) return rc;
    }
    if ((rc = launch_merge_color(ctx, p, ctx->d_planes[0], n, rgb_out_dev, st))) return rc;
    // schedules whose kernels wait on one another end by reporting an abandoned wait (LLICTI_E_TIMEOUT)
    if (ctx->cfg.sub_len == 0 && ctx->cfg.decode_impl != 1 && !ctx->no_coresidency) return launch_abort_check(ctx, st);
    return LLICTI_OK;
}

int llicti_decode_host(llicti_ctx *ctx, const uint8_t *blob, const uint64_t *stream_off, const int16_t *minmax,
                       const uint8_t *x00_rgb, int n, int H, int W, uint8_t *rgb_out, void *stream) {
    int rc = check_batch(ctx, n, H, W);
    if (rc) return rc;
    LLICTI_REQUIRE(blob && stream_off && minmax && x00_rgb && rgb_out, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const Plan &p = ctx->plan;
    const int ns = p.n_streams, S = p.g.num_scales;
    const uint64_t total = stream_off[(size_t)n * ns];
    LLICTI_REQUIRE(total <= ctx->blob_cap, "bitstream of %llu bytes exceeds the reserved %llu", (unsigned long long)total,
                   (unsigned long long)ctx->blob_cap);
    if (stream_off[0] != 0) { set_error("malformed stream offsets: the first offset is not 0"); return LLICTI_E_STREAM; }
    for (size_t i = 0; i < (size_t)n * ns; ++i)
        if (stream_off[i] > stream_off[i + 1]) {
            set_error("malformed stream offsets: offset %zu decreases", i + 1);
            return LLICTI_E_STREAM;
        }
    // (Decoding in halves pays only when forced: the decoder's chain-parallel kernels need the whole batch's chains to fill
    // the machine -- c2: 1813 -> 1317 MP/s end to end in halves -- and the decoded pixels exist only at the very end.)
    if (host_pipelined(ctx, n, H, W) && getenv("LLICTI_HOST_PIPELINE")) {
        bool done = false;
        rc = decode_host_pipelined(ctx, blob, stream_off, minmax, x00_rgb, n, H, W, rgb_out, st, &done);
        if (rc || done) return rc;
    }
    LLICTI_CUDA(cudaMemsetAsync(ctx->d_status, 0, sizeof(int32_t), st));
    LLICTI_CUDA(cudaMemcpyAsync(ctx->d_blob, blob, total, cudaMemcpyHostToDevice, st));
    LLICTI_CUDA(cudaMemcpyAsync(ctx->d_stream_off, stream_off, ((size_t)n * ns + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    LLICTI_CUDA(cudaMemcpyAsync(ctx->d_minmax16, minmax, (size_t)n * 6 * sizeof(int16_t), cudaMemcpyHostToDevice, st));
    LLICTI_CUDA(cudaMemcpyAsync(ctx->d_x00, x00_rgb, (size_t)n * 3 * p.g.Hs[S - 1] * p.g.Ws[S - 1], cudaMemcpyHostToDevice, st));
    if ((rc = llicti_decode_dev(ctx, ctx->d_blob, ctx->d_stream_off, ctx->d_minmax16, ctx->d_x00, n, H, W, ctx->d_rgb, st)))
        return rc;
    rc = read_status(ctx, st);
    if (rc == LLICTI_E_TIMEOUT && !ctx->no_coresidency) {
        // the producer / consumer kernels were not resident together: decode again with the schedule whose kernels
        // never wait on each other, and stay with it
        ctx->no_coresidency = true;
        drop_decode_graph(ctx);
        if ((rc = llicti_decode_dev(ctx, ctx->d_blob, ctx->d_stream_off, ctx->d_minmax16, ctx->d_x00, n, H, W, ctx->d_rgb, st)))
            return rc;
        rc = read_status(ctx, st);
    }
    if (rc) return rc;
    LLICTI_CUDA(cudaMemcpyAsync(rgb_out, ctx->d_rgb, (size_t)n * 3 * H * W, cudaMemcpyDeviceToHost, st));
    LLICTI_CUDA(cudaStreamSynchronize(st));
    return LLICTI_OK;
}

int llicti_encode_batch_host(llicti_ctx *ctx, const llicti_encode_item *items, int n, void *stream) {
    LLICTI_REQUIRE(ctx && items && n >= 1, "bad argument");
    for (int i = 0; i < n; ++i)
        LLICTI_REQUIRE(items[i].rgb && items[i].out && items[i].stream_off && items[i].minmax, "null pointer in item %d", i);
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<uint64_t> off;
    for (const std::vector<int> &grp : group_by_size(items, n)) {
        const int m = (int)grp.size(), H = items[grp[0]].H, W = items[grp[0]].W;
        int rc = reserve_for(ctx, m, H, W);
        if (rc) return rc;
        const int ns = ctx->plan.n_streams;
        const size_t img_bytes = (size_t)3 * H * W;
        LLICTI_CUDA(cudaMemsetAsync(ctx->d_status, 0, sizeof(int32_t), st));
        for (int k = 0; k < m; ++k)
            LLICTI_CUDA(cudaMemcpyAsync(ctx->d_rgb + (size_t)k * img_bytes, items[grp[k]].rgb, img_bytes, cudaMemcpyHostToDevice, st));
        if ((rc = llicti_encode_dev(ctx, ctx->d_rgb, m, H, W, ctx->d_blob, ctx->blob_cap, ctx->d_stream_off, ctx->d_minmax16, st))) return rc;
        off.resize((size_t)m * ns + 1);
        LLICTI_CUDA(cudaMemcpyAsync(off.data(), ctx->d_stream_off, off.size() * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        if ((rc = read_status(ctx, st))) return rc;
        for (int k = 0; k < m; ++k) {
            const llicti_encode_item &it = items[grp[k]];
            const uint64_t b0 = off[(size_t)k * ns], b1 = off[(size_t)(k + 1) * ns];
            if (b1 - b0 > it.out_cap) {
                set_error("item %d needs %llu bytes, capacity is %llu", grp[k], (unsigned long long)(b1 - b0), (unsigned long long)it.out_cap);
                return LLICTI_E_NOMEM;
            }
            for (int j = 0; j <= ns; ++j) it.stream_off[j] = off[(size_t)k * ns + j] - b0;
            LLICTI_CUDA(cudaMemcpyAsync(it.out, ctx->d_blob + b0, b1 - b0, cudaMemcpyDeviceToHost, st));
            LLICTI_CUDA(cudaMemcpyAsync(it.minmax, ctx->d_minmax16 + (size_t)k * 6, 6 * sizeof(int16_t), cudaMemcpyDeviceToHost, st));
        }
        LLICTI_CUDA(cudaStreamSynchronize(st));
    }
    return LLICTI_OK;
}

int llicti_decode_batch_host(llicti_ctx *ctx, const llicti_decode_item *items, int n, void *stream) {
    LLICTI_REQUIRE(ctx && items && n >= 1, "bad argument");
    for (int i = 0; i < n; ++i)
        LLICTI_REQUIRE(items[i].blob && items[i].stream_off && items[i].minmax && items[i].x00_rgb && items[i].rgb_out,
                       "null pointer in item %d", i);
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<uint64_t> off;
    std::vector<uint8_t> blob, x00;
    std::vector<int16_t> mm;
    std::vector<uint8_t> rgb;
    for (const std::vector<int> &grp : group_by_size(items, n)) {
        const int m = (int)grp.size(), H = items[grp[0]].H, W = items[grp[0]].W;
        int rc = reserve_for(ctx, m, H, W);
        if (rc) return rc;
        const Plan &p = ctx->plan;
        const int ns = p.n_streams, S = p.g.num_scales;
        const size_t x00_bytes = (size_t)3 * p.g.Hs[S - 1] * p.g.Ws[S - 1], img_bytes = (size_t)3 * H * W;
        // the uniform entry point takes one blob with running offsets: lay the group's streams end to end
        off.assign((size_t)m * ns + 1, 0);
        uint64_t pos = 0;
        for (int k = 0; k < m; ++k) {
            const llicti_decode_item &it = items[grp[k]];
            for (int j = 0; j < ns; ++j) {
                if (it.stream_off[j] > it.stream_off[j + 1]) { set_error("item %d: stream offsets decrease", grp[k]); return LLICTI_E_STREAM; }
                off[(size_t)k * ns + j] = pos + (it.stream_off[j] - it.stream_off[0]);
            }
            pos += it.stream_off[ns] - it.stream_off[0];
        }
        off[(size_t)m * ns] = pos;
        blob.resize(std::max<uint64_t>(pos, 1));
        x00.resize((size_t)m * x00_bytes);
        mm.resize((size_t)m * 6);
        for (int k = 0; k < m; ++k) {
            const llicti_decode_item &it = items[grp[k]];
            memcpy(blob.data() + off[(size_t)k * ns], it.blob + it.stream_off[0], it.stream_off[ns] - it.stream_off[0]);
            memcpy(x00.data() + (size_t)k * x00_bytes, it.x00_rgb, x00_bytes);
            memcpy(mm.data() + (size_t)k * 6, it.minmax, 6 * sizeof(int16_t));
        }
        rgb.resize((size_t)m * img_bytes);
        if ((rc = llicti_decode_host(ctx, blob.data(), off.data(), mm.data(), x00.data(), m, H, W, rgb.data(), st))) return rc;
        for (int k = 0; k < m; ++k) memcpy(items[grp[k]].rgb_out, rgb.data() + (size_t)k * img_bytes, img_bytes);
    }
    return LLICTI_OK;
}

}  // extern "C"
