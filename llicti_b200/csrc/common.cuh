// Shared declarations of libllicti_b200.so (sm_100a).  Internal; the public surface is
// include/llicti.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/llicti.h"

namespace llicti {

constexpr int kM = 5;            // mixtures per colour channel (config.num_mixtures)
constexpr int kParamCh = 12 * kM;  // 60 output channels of the interpolator CNN
constexpr int kMaxStreams = 9 * LLICTI_MAX_SCALES;

void set_error(const char *fmt, ...);
int device_sm_count(llicti_ctx *ctx, int *out);

#define LLICTI_CUDA(call)                                                                    \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            ::llicti::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,               \
                                cudaGetErrorString(e__));                                    \
            return LLICTI_E_CUDA;                                                            \
        }                                                                                    \
    } while (0)

#define LLICTI_REQUIRE(cond, ...)                \
    do {                                         \
        if (!(cond)) {                           \
            ::llicti::set_error(__VA_ARGS__);    \
            return LLICTI_E_ARG;                 \
        }                                        \
    } while (0)

// One coded stream = (scale, band, colour channel) of one image.  Same for every image of a
// batch (uniform H x W), so it lives in a small device table indexed by stream order k.
struct StreamDesc {
    int32_t scale, band, clr;
    int32_t Hs, Ws;            // plane size of the scale
    int32_t crop_h, crop_w;    // coded region
    int32_t n_sym;             // crop_h * crop_w
    int32_t S;                 // interleaved substreams
    int32_t sub_first;         // index of substream 0 among the image's substreams
    int32_t slot_bytes;        // encoder scratch bytes per substream
    int32_t pad_;
    int64_t sym_off;           // first symbol among the image's symbols
    int64_t slot_off;          // first scratch byte among the image's scratch
};

struct Plan {
    llicti_geom g;
    int n_streams;                 // 9 * num_scales
    StreamDesc sd[kMaxStreams];
    int64_t plane_elems[LLICTI_MAX_SCALES];  // 12 * Hs * Ws
    int64_t scratch_bytes;         // encoder scratch per image
    int hdr_per_stream(int k) const { return 0; }
};

int make_plan(const llicti_config &cfg, int H, int W, Plan *out);
int num_substreams(int64_t n, int sub_len);

// Packed weights of one band on the device (fp32 path).
struct BandWeightsF32 {
    float *w0;   // [K0][Ch]   K0 = 3 * taps, k = ((branch, c, dy, dx))
    float *b0;   // [Ch]       sum of the branch biases
    float *w1;   // [4][g][g]  [group][in][out]
    float *b1;   // [Ch]
    float *w2;   // [4][g][15] [group][in][out]
    float *b2;   // [60]
    int K0;
};

// Tap table of layer 0: source phase, colour channel and offset of every k.
struct TapTable {
    int K0;
    int8_t phase[128], chan[128], dy[128], dx[128];
};

struct NumericsProfile {
    int div255_recip;   // 1: x/255 evaluated as x * (1/255f)  (ATen CUDA); 0: IEEE division
    int sum_ilp4;       // 1: (((t0+t4)+t1)+t2)+t3 (ATen 4-accumulator); 0: left to right
};

}  // namespace llicti

struct llicti_ctx {
    llicti_config cfg;
    llicti::NumericsProfile num;
    llicti::BandWeightsF32 wf32[3];
    llicti::TapTable taps[3];
    void *tc_weights = nullptr;       // tcgen05 path (packed bf16 operands), see cnn_tc.cu
    int64_t launches = 0;

    // optional per-kernel-class timing with CUDA events on the launching stream (llicti_profile)
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;   // pairs (begin, end)
    std::vector<int> prof_cls;

    // workspace (llicti_reserve)
    int ws_images = 0, ws_H = 0, ws_W = 0;
    llicti::Plan plan;
    llicti::StreamDesc *d_sd = nullptr;
    uint8_t *d_rgb = nullptr;
    int16_t *d_planes[LLICTI_MAX_SCALES] = {};
    int32_t *d_minmax = nullptr;       // [n][4]
    int16_t *d_minmax16 = nullptr;     // [n][6] header words
    float *d_params = nullptr;         // [n][60][P0]
    uint32_t *d_bounds = nullptr;      // [n][symbols]
    uint8_t *d_scratch = nullptr;      // [n][scratch_bytes]
    uint32_t *d_sublen = nullptr;      // [n][substreams]
    uint64_t *d_suboff = nullptr;      // [n][substreams] (decode)
    uint64_t *d_stream_bytes = nullptr;  // [n*streams + 1]
    uint64_t *d_stream_off = nullptr;    // [n*streams + 1]
    uint8_t *d_blob = nullptr;         // compacted output / decode input
    size_t blob_cap = 0;
    uint8_t *d_x00 = nullptr;          // [n][3][h_last][w_last]
    void *d_items = nullptr;           // decode windows: one item slot per 32 steps of one chain (kernels_decode.cu)
    int64_t items_cap = 0;             // in items
    int16_t *d_syms = nullptr;         // [n][3][sym_cap] compact decoded symbols of the band in flight
    int64_t sym_cap = 0;
    void *d_chain_state_raw = nullptr;
    void *side_stream = nullptr, *ev_fork = nullptr, *ev_join = nullptr;   // second stream of the wavefront decode (consumer kernel)
    void *copy_in = nullptr, *copy_out = nullptr, *ev_pipe[9] = {};        // host-buffer entry points: copies of one part of a batch overlap the other parts' kernels (api.cu)
    bool no_coresidency = false;       // a decode gave up waiting (LLICTI_E_TIMEOUT): only schedules without kernel-to-kernel hand-overs from now on
    bool concurrent_kernels = false;   // two kernels on two streams really overlap (false under kernel-serialising profilers)
    bool wave_ws = false;              // workspace holds three bands' worth of decode buffers (wavefront schedule)
    uint32_t *d_item_flags = nullptr;  // [items_cap] readiness flags of the piped decode schedule

    // llicti_decode_dev as a CUDA graph: the launch sequence of a decode is fixed by its arguments, so the second
    // call with the same arguments captures it and later calls replay it with one launch (api.cu)
    void *dec_capture_stream = nullptr, *dec_graph_exec = nullptr;
    const void *dec_key_ptr[5] = {};
    long long dec_key_dim[4] = {};
    bool dec_key_seen = false;
    int64_t dec_graph_launches = 0;
    int32_t *d_status = nullptr;       // device-side error flag

    // properties of the context's device and of the kernels on it (filled on first use; per context, not per
    // process: two contexts of one process may sit on two devices)
    int sm_count = 0, regs_per_sm = 0;
    int pipe_resident = 0;             // one-warp CTAs of decode_band_pipe_kernel per SM
    int prod_resident = 0, cons_regs = 0, prod_regs = 0;   // wavefront kernels
    size_t tc_attr_smem = 0;           // dynamic shared memory opted in for the tcgen05 CNN kernels
    void *train_state = nullptr;       // packed gradient accumulators of the training step (kernels_train.cu), on first use
    bool weights_from_device = false;  // llicti_set_weights_dev replaced the fp32 weights given at creation
};

namespace llicti {

enum KernelClass { KC_SPLIT = 0, KC_CNN = 1, KC_BOUNDS = 2, KC_ENCODE = 3, KC_COMPACT = 4, KC_INDEX = 5,
                   KC_DECODE = 6, KC_MERGE = 7, KC_WINDOW = 8, KC_COUNT = 9 };

// Brackets the kernels launched in its lifetime with two events when profiling is on.
struct ProfScope {
    llicti_ctx *ctx;
    cudaStream_t st;
    cudaEvent_t b = nullptr, e = nullptr;
    ProfScope(llicti_ctx *c, int cls, cudaStream_t s) : ctx(c), st(s) {
        if (!ctx->prof_on) return;
        cudaEventCreate(&b);
        cudaEventCreate(&e);
        cudaEventRecord(b, st);
        ctx->prof_cls.push_back(cls);
    }
    ~ProfScope() {
        if (!b) return;
        cudaEventRecord(e, st);
        ctx->prof_ev.push_back(b);
        ctx->prof_ev.push_back(e);
    }
};

// kernels_color.cu
int launch_color_split(llicti_ctx *ctx, const Plan &p, const uint8_t *rgb, int n, int16_t *const *planes,
                       int32_t *minmax, cudaStream_t st);
int launch_color_split_float(llicti_ctx *ctx, const Plan &p, const uint8_t *rgb, int n, float *const *fplanes,
                             int16_t *const *planes, cudaStream_t st);
int launch_self_info(llicti_ctx *ctx, const float *params, const float *fplanes, int band, int n, int P, float *sinfo,
                     cudaStream_t st);
int launch_merge_color(llicti_ctx *ctx, const Plan &p, const int16_t *planes0, int n, uint8_t *rgb,
                       cudaStream_t st);
int launch_x00_from_header(llicti_ctx *ctx, const Plan &p, const uint8_t *x00_rgb, int n, int16_t *planes_last,
                           cudaStream_t st);
int launch_interleave(llicti_ctx *ctx, const Plan &p, int scale_from, const int16_t *planes_from,
                      int16_t *planes_to, int n, cudaStream_t st);
int launch_minmax16(llicti_ctx *ctx, const int32_t *minmax, int16_t *minmax16, int n, cudaStream_t st);
int launch_minmax32(llicti_ctx *ctx, const int16_t *minmax16, int32_t *minmax, int n, cudaStream_t st);

// kernels_cnn_fp32.cu
int launch_cnn_fp32(llicti_ctx *ctx, int band, const int16_t *planes, int n, int Hs, int Ws, float *params,
                    cudaStream_t st);

// kernels_train.cu
int launch_self_info_grad(llicti_ctx *ctx, float *params, const float *fplanes, const float *gsinfo, int band, int n, int P,
                          cudaStream_t st);
int launch_cnn_backward(llicti_ctx *ctx, int band, const float *fplanes, int n, int Hs, int Ws, const float *dparams, cudaStream_t st);
int launch_cnn_forward_train(llicti_ctx *ctx, int band, const float *fplanes, int n, int Hs, int Ws, float *params, cudaStream_t st);
int launch_train_zero_grads(llicti_ctx *ctx, cudaStream_t st);
int launch_train_layouts(llicti_ctx *ctx, const llicti_weights &tw, bool to_packed, cudaStream_t st);
void train_free(llicti_ctx *ctx);

// cnn_tc.cu
int tc_pack_weights(llicti_ctx *ctx, const llicti_weights &w);
void tc_free_weights(llicti_ctx *ctx);
int tc_operand_type(const llicti_ctx *ctx);
int launch_cnn_tc(llicti_ctx *ctx, int band, const int16_t *planes, int n, int Hs, int Ws, float *params,
                  cudaStream_t st, int row0 = 0, int nrows = -1);

// kernels_coder.cu
int launch_cdf_table(llicti_ctx *ctx, const float *params, const int16_t *yband, int clr, int min_val,
                     int max_val, int P, int16_t *table, cudaStream_t st);
int launch_cdf_bounds_flat(llicti_ctx *ctx, const float *params, const int16_t *yband, int clr, int min_val,
                           int max_val, int P, uint32_t *bounds, cudaStream_t st);
int launch_band_bounds(llicti_ctx *ctx, const Plan &p, int scale, int band, const float *params,
                       const int16_t *planes, const int32_t *minmax, int n, uint32_t *bounds, int64_t sym_stride,
                       cudaStream_t st);
int launch_encode_all(llicti_ctx *ctx, const Plan &p, const uint32_t *bounds, int64_t sym_stride, int n,
                      uint8_t *scratch, int64_t scratch_stride, uint32_t *sublen, cudaStream_t st);
int launch_encode_flat(llicti_ctx *ctx, const uint32_t *bounds, int n_sym, int S, uint8_t *out, int slot_bytes,
                       uint32_t *lens, cudaStream_t st);
int launch_compact(llicti_ctx *ctx, const Plan &p, int n, const uint8_t *scratch, int64_t scratch_stride,
                   const uint32_t *sublen, uint64_t *stream_bytes, uint64_t *stream_off, uint8_t *out,
                   size_t out_cap, cudaStream_t st);
int launch_index_streams(llicti_ctx *ctx, const Plan &p, int n, const uint8_t *blob, uint64_t blob_bytes,
                         const uint64_t *stream_off, uint64_t *suboff, uint32_t *sublen, cudaStream_t st);
int launch_decode_band(llicti_ctx *ctx, const Plan &p, int scale, int band, const float *params, int16_t *planes,
                       const int32_t *minmax, int n, const uint8_t *blob, const uint64_t *suboff,
                       const uint32_t *sublen, cudaStream_t st);
int64_t decode_items_per_image(const Plan &p);
int64_t decode_items_capacity(const llicti_config &cfg, const Plan &p, int max_images);
size_t decode_item_bytes();
bool wave_eligible(const llicti_ctx *ctx, const llicti::Plan &p, int scale, int n);
int wave_bands_in_workspace(const llicti_config &cfg, int max_images);
int probe_concurrent_kernels(llicti_ctx *ctx, bool *ok);
int launch_decode_scale_wave(llicti_ctx *ctx, const llicti::Plan &p, int scale, int16_t *planes, const int32_t *minmax, int n,
                             const uint8_t *blob, const uint64_t *suboff, const uint32_t *sublen, cudaStream_t st);
int64_t decode_flag_words(int64_t items_cap);
int launch_abort_check(llicti_ctx *ctx, cudaStream_t st);
int apply_decode_test_knobs();
int read_decode_stats(uint64_t *out, int reset);
int launch_selftest_fdiv(llicti_ctx *ctx, long long n_pairs, uint64_t seed, unsigned long long *mismatches_dev, cudaStream_t st);
int launch_decode_table(llicti_ctx *ctx, const int16_t *table, int n_sym, int Lp, int S, const uint8_t *in,
                        const uint32_t *offs, int16_t *sym, cudaStream_t st);

}  // namespace llicti
