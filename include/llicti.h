/*
 * include/llicti.h -- C ABI of libllicti_b200.so (sm_100a).
 *
 * The reference (kamisli-icpl/LLICTI) is pure Python/PyTorch and has no FFI seam of
 * its own; the seam fixed here is the one its hot path would bind through ctypes:
 * every entry point replaces a stretch of graphs/models/LLICTI_nets.py or
 * graphs/layers/entropy_layer_nets.py (file:line given per function, paths relative
 * to the reference root).  INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *   - plain C types only; no torch / C++ types cross the boundary;
 *   - every function returns 0 on success or a negative LLICTI_E_* code; the text
 *     of the last error on the calling thread is llicti_last_error();
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - "dev" pointers are device pointers owned by the caller; stage-level calls
 *     neither allocate nor synchronise.  Full-path calls use the context's
 *     workspace (llicti_reserve) and synchronise only in their *_host variants;
 *   - there is NO CPU fallback: without a CUDA device every compute call fails.
 *
 * Layouts
 *   rgb      uint8  [n][3][H][W]                       planar (what ToTensor()*255 holds)
 *   planes_s int16  [n][12][Hs][Ws]  per scale s       phase order x00,x11,x01,x10; 3 colour
 *                                                      channels each: Y-127, Co, Cg
 *   params   float  [n][12*M][Hs*Ws]                   channel order of the reference's conv
 *                                                      output (sigma | mu | w | a b d)
 *   bounds   uint32 per coded symbol                   c_low | (c_high-1) << 16
 *   tables   int16  [P][Lp]                            the reference's integer CDF rows
 */
#ifndef LLICTI_H
#define LLICTI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define LLICTI_API __attribute__((visibility("default")))
#else
#define LLICTI_API
#endif

#define LLICTI_ABI_VERSION 1
#define LLICTI_MAX_SCALES 8

enum {
    LLICTI_OK = 0,
    LLICTI_E_ARG = -1,      /* bad argument / unsupported configuration   */
    LLICTI_E_CUDA = -2,     /* a CUDA runtime call failed                 */
    LLICTI_E_NOMEM = -3,    /* workspace / output capacity too small      */
    LLICTI_E_STREAM = -4,   /* malformed bitstream                        */
    LLICTI_E_NODEVICE = -5, /* no usable CUDA device (no CPU fallback)    */
    LLICTI_E_TIMEOUT = -6   /* kernels that hand work to each other inside one decode (torchac-compatible streams) were not
                               resident together and gave up; the context has switched to the schedule without such
                               hand-overs: llicti_decode_host retries by itself, llicti_decode_dev must be called again */
};

/* Floating-point conventions of the GMM-CDF stage (see DESIGN.md "numerics profile"). */
enum {
    LLICTI_NUM_TORCH_CUDA = 0, /* x/255 as x*(1/255), reductions in ATen's 4-accumulator order */
    LLICTI_NUM_TORCH_CPU = 1   /* x/255 as IEEE division, reductions left to right              */
};

enum { LLICTI_CNN_FP32 = 0, LLICTI_CNN_TCGEN05 = 1 };

typedef struct llicti_ctx llicti_ctx;

/* The subset of configs/llicti_*.json the eval_model path reads
 * (LLICTI_nets.py:19-26, 260-278, 590-603). */
typedef struct {
    int32_t num_scales;   /* len(dwtlevels); levels must be 0..num_scales-1        */
    int32_t chs;          /* config.chs[0]: hidden width of each of the 4 sub-nets  */
    int32_t num_mixtures; /* 5                                                     */
    int32_t sub_len;      /* 0: one torchac-compatible stream per (scale,band,channel);
                             >0: interleaved substreams of about sub_len symbols    */
    int32_t numerics;     /* LLICTI_NUM_*                                          */
    int32_t cnn_impl;     /* LLICTI_CNN_*                                          */
    int32_t device;       /* CUDA device ordinal                                   */
    int32_t decode_impl;  /* 0 (default): substream container -> one lane group per chain evaluates the few table entries
                             it needs; torchac-compatible streams -> CDF windows + serial coder chains;
                             1: legacy, one warp per chain; 2: CDF windows + chains for every container (A/B) */
} llicti_config;

/* fp32 host pointers in PyTorch's own layouts (state_dict keys in SURVEY.md 8b).
 * Branch order of l0_*: 00_11 | 00_01, 11_01 | 00_10, 11_10, 01_10. */
typedef struct {
    const float *l0_w[6]; /* (4*chs, 3, kh, kw)                      */
    const float *l0_b[6]; /* (4*chs)                                 */
    const float *l1_w[3]; /* per band (4*chs, chs)   groups=4 1x1    */
    const float *l1_b[3]; /* (4*chs)                                 */
    const float *l2_w[3]; /* per band (12*M, chs)    groups=4 1x1    */
    const float *l2_b[3]; /* (12*M)                                  */
} llicti_weights;

/* Geometry of one image size under a configuration (pure host arithmetic;
 * restates lazyDWT's shape logic LLICTI_nets.py:218-241 and the crop rules :396-397). */
typedef struct {
    int32_t H, W, num_scales;
    int32_t Hs[LLICTI_MAX_SCALES], Ws[LLICTI_MAX_SCALES];     /* x00 size per scale          */
    int32_t padH[LLICTI_MAX_SCALES], padW[LLICTI_MAX_SCALES]; /* replicate-pad flags         */
    int32_t pad_int;                                          /* header word (:230)          */
    int32_t crop_h[LLICTI_MAX_SCALES][3], crop_w[LLICTI_MAX_SCALES][3]; /* coded region per band */
    int32_t num_sub[LLICTI_MAX_SCALES][3];                    /* substreams per stream       */
    int64_t positions;                                        /* sum Hs*Ws                   */
    int64_t symbols;                                          /* coded symbols per image     */
    int64_t substreams;                                       /* substreams per image (all 9*S streams) */
    int64_t max_stream_bytes;                                 /* worst-case bytes of all streams of one image */
} llicti_geom;

LLICTI_API int llicti_abi_version(void);
LLICTI_API const char *llicti_last_error(void);

/* Number of CUDA devices visible to the library (0 if none; never fails). */
LLICTI_API int llicti_device_count(void);

LLICTI_API int llicti_geometry(const llicti_config *cfg, int H, int W, llicti_geom *out);

/* Replaces LLICTI(config).to(device) + load_state_dict (LLICTI_nets.py:93-99, agents/base.py:51-81):
 * packs the weights into the kernels' layouts on the device. */
LLICTI_API int llicti_create(const llicti_config *cfg, const llicti_weights *w, llicti_ctx **out);
LLICTI_API void llicti_destroy(llicti_ctx *ctx);

/* Size the context's device workspace for batches of up to max_images images of H x W. */
LLICTI_API int llicti_reserve(llicti_ctx *ctx, int max_images, int H, int W);

/* ---- stage-level entry points (device pointers, asynchronous) ------------------------ */

/* get_YCoCg_R_from_RGB__intOps + min/max + Y-127 + lazyDWT(pad=True)
 * (LLICTI_nets.py:62-74, 137-139, 143, 218-241).  planes_dev[s] -> int16 [n][12][Hs][Ws];
 * minmax_dev int32 [n][4] = minCo, minCg, maxCo, maxCg. */
LLICTI_API int llicti_color_split(llicti_ctx *ctx, const uint8_t *rgb_dev, int n, int H, int W,
                       int16_t *const *planes_dev, int32_t *minmax_dev, void *stream);

/* Inverse lazy DWT of the finest scale + Y+127 + get_RGB_from_YCoCg_R__intOps
 * (LLICTI_nets.py:501-509, 174-175, 77-88): planes0_dev int16 [n][12][Hs0][Ws0] -> rgb. */
LLICTI_API int llicti_merge_color(llicti_ctx *ctx, const int16_t *planes0_dev, int n, int H, int W,
                       uint8_t *rgb_dev, void *stream);

/* LLICTIEntropyModel4.get_params (LLICTI_nets.py:822-825, 721-753, 695-712) for band 0..2:
 * planes_dev int16 [n][12][Hs][Ws] (phases 0..band read) -> params_dev float [n][12*M][Hs*Ws]. */
LLICTI_API int llicti_cnn_params(llicti_ctx *ctx, int band, const int16_t *planes_dev, int n, int Hs, int Ws,
                      float *params_dev, void *stream);

/* Mean coupling + get_cdfs(int_cdf=True) (LLICTI_nets.py:389-392, 938-952, 955-983;
 * entropy_layer_nets.py:185-204) as a dense table, for parity checks only:
 * params_dev float [12*M][P], yband_dev int16 [3][P] (centred values of the band being
 * coded) -> table_dev int16 [P][Lp], Lp = max_val - min_val + 2. */
LLICTI_API int llicti_cdf_table(llicti_ctx *ctx, const float *params_dev, const int16_t *yband_dev, int clr,
                     int min_val, int max_val, int P, int16_t *table_dev, void *stream);

/* Same stage, table-free: only the two entries the coder needs per symbol.
 * sym = yband + shift (LLICTI_nets.py:544-557); bounds_dev[i] = c_low | (c_high-1)<<16. */
LLICTI_API int llicti_cdf_bounds(llicti_ctx *ctx, const float *params_dev, const int16_t *yband_dev, int clr,
                      int min_val, int max_val, int P, uint32_t *bounds_dev, void *stream);

/* torchac.encode_int16_normalized_cdf (call site LLICTI_nets.py:406-407) over S interleaved
 * substreams (S = 1: the torchac bitstream).  Substream j codes symbols j, j+S, ...  into
 * out_dev + j*slot_bytes; lens_dev[j] receives its byte count. */
LLICTI_API int llicti_ac_encode_bounds(llicti_ctx *ctx, const uint32_t *bounds_dev, int n_sym, int S,
                            uint8_t *out_dev, int slot_bytes, uint32_t *lens_dev, void *stream);

/* torchac.decode_int16_normalized_cdf (call site LLICTI_nets.py:492-493) from a dense table:
 * table_dev int16 [n_sym][Lp]; substream j is in_dev[offs_dev[j] .. offs_dev[j+1]). */
LLICTI_API int llicti_ac_decode_table(llicti_ctx *ctx, const int16_t *table_dev, int n_sym, int Lp, int S,
                           const uint8_t *in_dev, const uint32_t *offs_dev, int16_t *sym_dev,
                           void *stream);

/* ---- full path ---------------------------------------------------------------------- */

/* LLICTI.compress for a batch (LLICTI_nets.py:125-159, 344-413).
 *   rgb          uint8 [n][3][H][W], host (pinned or pageable) or device memory
 *   out          receives the 9*num_scales stream blobs of every image, image-major, in
 *                bytestream_list order (scale S-1..0, index 3*band+clr)
 *   stream_off   uint64 [n*9*S + 1] byte offsets into out (last = total)
 *   minmax       int16 [n][6] header words (0,minCo,minCg,255,maxCo,maxCg)   (:139, :348)
 * The header's remaining pieces (dims, pad word, raw x00) are functions of the input and
 * the geometry and are assembled by the host wrapper.  *_host copies in and out on
 * `stream` and returns after the stream is idle; *_dev takes device pointers and is
 * asynchronous (stream_off / minmax then are device pointers too).
 * llicti_encode_host codes a large batch of the substream container (>= 16 images, >= 64 MB of pixels) in two to four
 * parts on two copy streams of its own, so that one part's host<->device copies run while another part is coded; the
 * caller sees what one batch gives (contiguous bytes, global offsets).  LLICTI_HOST_PIPELINE=0 turns that off. */
LLICTI_API int llicti_encode_host(llicti_ctx *ctx, const uint8_t *rgb, int n, int H, int W, uint8_t *out,
                       size_t out_cap, uint64_t *stream_off, int16_t *minmax, void *stream);
LLICTI_API int llicti_encode_dev(llicti_ctx *ctx, const uint8_t *rgb_dev, int n, int H, int W, uint8_t *out_dev,
                      size_t out_cap, uint64_t *stream_off_dev, int16_t *minmax_dev, void *stream);

/* The same two calls for a batch of images of DIFFERENT sizes, described per image (host pointers).  Images are coded
 * independently (the reference's eval loop walks them with batch size 1, agents/llicti_agent.py:129-149), so the
 * library groups the descriptors by size, reserves the workspace per group and runs each group through the uniform
 * path above; results land in the caller's per-image buffers.  Streams of an image do not depend on its neighbours. */
typedef struct {
    const uint8_t *rgb;      /* in : uint8 [3][H][W] planar                                              */
    int32_t H, W;
    uint8_t *out;            /* out: the image's 9*num_scales streams back to back                      */
    size_t out_cap;          /*      capacity of out (llicti_geom.max_stream_bytes is always enough)     */
    uint64_t *stream_off;    /* out: uint64 [9*num_scales + 1] byte offsets into out (first 0, last = total) */
    int16_t *minmax;         /* out: int16 [6] header words                                             */
} llicti_encode_item;
typedef struct {
    const uint8_t *blob;          /* in : the image's streams back to back                               */
    const uint64_t *stream_off;   /* in : uint64 [9*num_scales + 1] offsets into blob                    */
    const int16_t *minmax;        /* in : int16 [6]                                                      */
    const uint8_t *x00_rgb;       /* in : uint8 [3][h_last][w_last] raw coarsest band                    */
    int32_t H, W;
    uint8_t *rgb_out;             /* out: uint8 [3][H][W]                                                */
} llicti_decode_item;
LLICTI_API int llicti_encode_batch_host(llicti_ctx *ctx, const llicti_encode_item *items, int n, void *stream);
LLICTI_API int llicti_decode_batch_host(llicti_ctx *ctx, const llicti_decode_item *items, int n, void *stream);

/* LLICTI.forward for a batch (LLICTI_nets.py:101-123, 318-342, 802-811, 827-935; entropy_layer_nets.py:160-183,
 * 121-139): the rate-estimation path of validate() / training -- no coding, -log2 of the probability mass of every
 * sample.  Float lifting of :40-49 (NOT the integer transform of compress), un-padded lazyDWT (H and W must be
 * multiples of 2^num_scales), the same CNN kernels, then the point likelihood.
 *   fplanes_dev[s]  float [n][12][Hs][Ws]  scratch / by-product: the fp32 phase planes of scale s
 *   sinfo_dev[s]    float [n][9][Hs][Ws]   self-information in bits, index 3*band + clr (what forward returns) */
LLICTI_API int llicti_forward_dev(llicti_ctx *ctx, const uint8_t *rgb_dev, int n, int H, int W,
                       float *const *fplanes_dev, float *const *sinfo_dev, void *stream);

/* The device half of the reference's training step (agents/llicti_agent.py:48-83: `self.model(x)`, TrainRLossList,
 * `.backward()`).  The optimizer, gradient clipping and the checkpoint stay with the caller, as in the
 * reference's agent.  Both calls need a context created with cnn_impl = LLICTI_CNN_FP32 (the reference trains in fp32).
 *
 * llicti_set_weights_dev: replace the context's weights by the ones at these DEVICE pointers (PyTorch layouts, the
 *   shapes of llicti_weights) -- the parameters after an optimizer step.
 * llicti_train_forward_dev: llicti_forward_dev that keeps every band's 60 network outputs for the backward pass.
 *   params_keep_dev float [180 * n * sum_s Hs*Ws]  (blocks [n][60][Hs*Ws] in (scale, band) order; llicti_geom.positions = sum_s Hs*Ws)
 * llicti_backward_dev: gradients of a loss L over forward()'s outputs with respect to every weight.
 *   fplanes_dev[s]  float [n][12][Hs][Ws]  the planes llicti_train_forward_dev wrote (scratch when params_kept_dev is NULL)
 *   gsinfo_dev[s]   float [n][9][Hs][Ws]   dL / d self-information (what autograd hands the backward of forward())
 *   params_kept_dev the buffer llicti_train_forward_dev filled for this batch (consumed: overwritten with the gradients
 *                   of the network outputs), or NULL: colour split and CNN are run again
 *   grads_dev       DEVICE pointers, layouts and order of llicti_weights; overwritten with dL / d weight (each of a
 *                   band's branch biases receives the gradient of their sum)
 * compressai's LowerBound gradient rule (pass where x >= bound or where the step would raise x) applies to the spreads,
 * the mixture weights and the likelihood (entropy_layer_nets.py:135, 176, 181). */
LLICTI_API int llicti_set_weights_dev(llicti_ctx *ctx, const llicti_weights *w_dev, void *stream);
LLICTI_API int llicti_train_forward_dev(llicti_ctx *ctx, const uint8_t *rgb_dev, int n, int H, int W, float *const *fplanes_dev,
                             float *const *sinfo_dev, float *params_keep_dev, void *stream);
LLICTI_API int llicti_backward_dev(llicti_ctx *ctx, const uint8_t *rgb_dev, int n, int H, int W, float *const *fplanes_dev,
                        const float *const *gsinfo_dev, float *params_kept_dev, const llicti_weights *grads_dev, void *stream);

/* LLICTI.decompres for a batch (LLICTI_nets.py:161-179, 415-509).
 *   blob / stream_off  as produced by encode
 *   minmax             int16 [n][6]
 *   x00_rgb            uint8 [n][3][h_last][w_last] raw coarsest band (:350, :429)
 *   rgb_out            uint8 [n][3][H][W]
 * The launch sequence of a decode depends only on its arguments (pointers, n, H, W), not on the stream contents: from
 * the third consecutive call with the same arguments on it is replayed as one CUDA graph (LLICTI_NO_GRAPH=1 in the
 * environment keeps every call eager).  Nothing is retained beyond the pointers passed to the call being made. */
LLICTI_API int llicti_decode_host(llicti_ctx *ctx, const uint8_t *blob, const uint64_t *stream_off,
                       const int16_t *minmax, const uint8_t *x00_rgb, int n, int H, int W,
                       uint8_t *rgb_out, void *stream);
LLICTI_API int llicti_decode_dev(llicti_ctx *ctx, const uint8_t *blob_dev, const uint64_t *stream_off_dev,
                      const int16_t *minmax_dev, const uint8_t *x00_rgb_dev, int n, int H, int W,
                      uint8_t *rgb_out_dev, void *stream);

/* Arithmetic type of the CNN's operands in this context: 0 fp32 (LLICTI_CNN_FP32), 1 bf16, 2 fp16 (LLICTI_CNN_TCGEN05;
 * fp16 when the weights prove that no hidden activation can leave fp16's range, bf16 otherwise; accumulation is fp32
 * either way).  Part of a stream's fingerprint: encoder and decoder must agree on it. */
LLICTI_API int llicti_cnn_operands(const llicti_ctx *ctx);

/* Device-side error flag of the asynchronous *_dev entry points (an encoder that ran out of output
 * capacity: LLICTI_E_NOMEM; a malformed container or stream offsets: LLICTI_E_STREAM).  Waits for
 * `stream`, returns the flag (0 = none) and clears it.  The *_host entry points do this themselves. */
LLICTI_API int llicti_status(llicti_ctx *ctx, void *stream);

/* Kernel launches issued by this context since creation (bench.py's gpu_launches). */
LLICTI_API int64_t llicti_launch_count(const llicti_ctx *ctx);

/* Per-kernel-class device time, measured with CUDA events recorded on the launching stream
 * around every launch while profiling is enabled.  Classes: 0 colour+pyramid split, 1 CNN,
 * 2 CDF bounds, 3 range encode, 4 container compaction, 5 container indexing, 6 range decode
 * (serial coder chains), 7 inverse pyramid / colour merge, 8 decode-side CDF windows.  llicti_profile_read waits for the
 * recorded events, returns summed milliseconds and launch-group counts, and clears them. */
#define LLICTI_KERNEL_CLASSES 9
LLICTI_API int llicti_profile(llicti_ctx *ctx, int enable);
LLICTI_API int llicti_profile_read(llicti_ctx *ctx, double *ms, int64_t *count);

/* Self-test of the CDF stage's hoisted IEEE division (csrc/gmm.cuh: fdiv_hoisted) against div.rn.f32 on
 * n_pairs random operand pairs from the stage's domain; *mismatches must come back 0. */
LLICTI_API int llicti_selftest_fdiv(llicti_ctx *ctx, int64_t n_pairs, uint64_t seed, uint64_t *mismatches);

/* Decode-side counters since the last reset (synchronises the device): out8[0] symbols that
 * fell outside their pre-computed CDF window and took the full analytic search, out8[1] polls
 * of consumer warps waiting for windows (piped schedule), out8[3] 8-symbol chunks whose
 * branch-free decode was redone with the careful path.  Diagnostics for bench.py; not on the data path. */
LLICTI_API int llicti_decode_stats(llicti_ctx *ctx, uint64_t *out8, int reset);

#ifdef __cplusplus
}
#endif
#endif /* LLICTI_H */
