/*
 * depthmatch.h -- C ABI of libdepthmatch.so, the B200 (sm_100a) dense-matching library.
 *
 * This is the drop-in boundary for the matching hot path of
 * MichaelMathieu/depth-estimation.  Every entry point replaces one native
 * interface the reference's Lua code binds today; the file:line of that
 * interface is cited next to each declaration.  LuaJIT `ffi.cdef` can take the
 * declarations below verbatim (INTEGRATION.md shows the binding), Python uses
 * ctypes, C/C++ hosts include this header.
 *
 * Conventions
 *  - All functions return DM_OK (0) or a negative dm_status; dm_last_error()
 *    returns a thread-local message.  Nothing throws or longjmps.
 *  - Data pointers may be HOST or DEVICE pointers (detected with
 *    cudaPointerGetAttributes).  Host buffers are staged through buffers owned by
 *    the context and the call is complete at return.  If every pointer is a
 *    device pointer the work is enqueued on the context's stream and the call
 *    returns immediately; use dm_synchronize().
 *  - Tensors are fp32, "Long" tensors are int64_t, row-major, innermost stride 1.
 *  - Indices are 1-based exactly where the reference's are (Torch convention).
 *  - There is NO CPU fallback: without a CUDA device dm_create fails.
 */
#ifndef DEPTHMATCH_H
#define DEPTHMATCH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library itself is built with -fvisibility=hidden */
#endif

#define DM_VERSION 100

typedef struct dm_ctx dm_ctx;

typedef enum dm_status {
  DM_OK = 0,
  DM_ERR_INVALID = -1,     /* bad argument (the reference would luaL_error / THError) */
  DM_ERR_CUDA = -2,        /* CUDA runtime/driver failure, message in dm_last_error() */
  DM_ERR_UNSUPPORTED = -3, /* valid request this build has no kernel for */
  DM_ERR_NOMEM = -4
} dm_status;

/* ---- context ------------------------------------------------------------ */
int dm_version(void);
const char *dm_last_error(void);
int dm_create(int device, dm_ctx **ctx);
int dm_destroy(dm_ctx *ctx);
int dm_synchronize(dm_ctx *ctx);
/* run on the caller's stream (cudaStream_t as void*; NULL = CUDA's legacy default stream),
 * or go back to the context's own non-blocking stream */
int dm_set_stream(dm_ctx *ctx, void *cuda_stream);
int dm_reset_stream(dm_ctx *ctx);
void *dm_get_stream(dm_ctx *ctx);
/* pinned host memory for callers that want full-speed staging (cudaHostAlloc) */
int dm_host_alloc(void **ptr, size_t bytes);
int dm_host_free(void *ptr);
/* number of CUDA kernels this context has launched so far (bench.py: gpu_launches) */
int64_t dm_launch_count(dm_ctx *ctx);
/* CUDA-event timing of the dominant kernel of a call (the matching sweep): switch it on,
 * make calls, read the duration of the most recent sweep kernel in milliseconds */
int dm_set_profiling(dm_ctx *ctx, int on);
int dm_last_kernel_ms(dm_ctx *ctx, float *ms);
/* tuning / diagnostic switches (none is needed for normal use).  The environment variables
 * DM_SSD_FORM, DM_NO_SMALL_TILES, DM_NO_PIPELINE, DM_PIPE_CHUNK, DM_VOLUME_DEBUG, DM_DEBUG_TODO,
 * DM_CONV_TILE, DM_SWEEP, DM_CONV, DM_VOLUME_KERNEL are read once, in dm_create; this changes them on a live context.  Names:
 * "ssd_form" = "auto" | "diff" | "dot"; "no_small_tiles", "no_pipeline", "debug_todo" = "0" | "1";
 * "pipe_chunk", "volume_debug", "sweep" = integer; "conv_tile" = "<candidate>,<CTAs per SM>";
 * "sweep" = 2 | 3: the two-rows-per-warp dot sweep (match_sweep2.cuh); "conv" = 2: the feature
 * extractor's layers on the tensor cores (tcgen05, filter_tc.cu).  Both variants meet the parity
 * bars and are measured slower than the defaults on B200, hence opt-in.  "volume_kernel" = 0: volume
 * mode picks the strip kernel (whole pixel streams written by bulk copies, match_volume_px.cuh) where it
 * fits and pays, 1: always the tiled kernel with sector stores, 2: the strip kernel wherever it fits. */
int dm_set_option(dm_ctx *ctx, const char *name, const char *value);
/* the "near-tie pixels are logged" part of the parity contract: how many pixels of the most
 * recent dm_match_extract on this context were handed to the entry-by-entry rescore (window entries
 * the fast form could not order, zero-flow near-ties) and how many went through the exact
 * thresholded-extraction pass.  Synchronises the context's stream.  -1 = not applicable. */
int dm_last_counts(dm_ctx *ctx, int64_t *rescored, int64_t *exact_pass);

/* ---- inputs ------------------------------------------------------------- */
/* Two stacks of feature maps.  in1 is the (already window-cropped) frame-1 map
 * the reference's prepareInput returns (opticalflow_model.lua:131-151): a strided
 * view, not a copy.  in1[n][k][y][x] = in1 + n*in1_stride_n + k*in1_stride_c +
 * y*in1_stride_y + x; likewise in2.  A stride of 0 means "contiguous default". */
typedef struct dm_pair {
  const float *in1;
  const float *in2;
  int32_t n_pairs, channels;
  int32_t h1, w1; /* in1 / output size */
  int32_t h2, w2; /* in2 size; h2 >= h1+maxh-1, w2 >= w1+maxw-1 */
  int64_t in1_stride_n, in1_stride_c, in1_stride_y;
  int64_t in2_stride_n, in2_stride_c, in2_stride_y;
} dm_pair;

/* ---- K1: cost volume ---------------------------------------------------- */
typedef enum dm_volume_mode {
  DM_VOLUME_SSD = 0,        /* nn.SpatialMatching output */
  DM_VOLUME_NEG_SOFTMAX = 1, /* ... -> nn.Minus -> nn.SoftMax over the window */
  DM_VOLUME_EXACT = 0x100    /* OR-ed in: unfused mul+add, bit-exact SSD (see DM_FLAG_EXACT_SSD) */
} dm_volume_mode;

/* Replaces input[1].nn.SpatialMatching_updateOutput(self, in1, in2) (out-of-tree
 * nnx; constructed at opticalflow_model.lua:93, opticalflow_model_multiscale.lua:216,
 * groundtruth_opticalflow.lua:73) and, with DM_VOLUME_NEG_SOFTMAX, the Minus +
 * SoftMax stages of getModel (opticalflow_model.lua:94-109).
 * out: [n_pairs][h1][w1][maxh][maxw].  maxw == 1 is nn.SpatialRadialMatching(maxh)
 * (radial/radial_opticalflow_network.lua:32-34). */
int dm_match_volume(dm_ctx *ctx, const dm_pair *in, int maxh, int maxw, int mode, float *out);

/* ---- K1+K2 fused: match + extract, volume never written ------------------ */
#define DM_FLAG_TIE_MIDDLE 1u  /* zero-flow tie rule, opticalflow_model.lua:157-159 */
#define DM_FLAG_EXACT_SSD 2u   /* unfused mul+add (bit-exact with the CPU path) instead of FMA */
#define DM_FLAG_DIFF_SSD 8u    /* always form the SSD as sum (a-b)^2.  Default: |a|^2+|b|^2-2a.b (half the
                                  FP32 work, absolute error a few ulp of |a|^2+|b|^2) whenever the
                                  largest norms keep that error under the parity bar, decided on
                                  the device per call; min_ssd is always re-scored as a difference */
#define DM_FLAG_ASYNC 4u       /* host-buffer call: return once the copies and kernels are queued;
                                  inputs and outputs belong to the library until dm_synchronize(ctx).
                                  Page-locked buffers (dm_host_alloc) make the copies overlap. */

typedef struct dm_extract_out {
  /* every pointer may be NULL = not wanted.  Shapes are [n_pairs][h1][w1] unless noted. */
  int64_t *index;      /* getOutputConfidences(threshold=nil) index, 1-based (:153-161) */
  float *min_ssd;      /* SSD of the winner */
  float *pmax;         /* softmax probability of the winner */
  float *flow_full;    /* processOutput().full: [n][2][h_img][w_img], y-flow then x-flow (:201-252) */
  int64_t *index_thr;  /* extractOutput(prob, prob_threshold) ret; 0 where untouched (:162-168) */
  float *score_thr;    /* extractOutput scores; 0 where untouched */
  float *soft_yx;      /* [n][2][h1][w1]: OutputExtractor y, x (1-based, sub-pixel) (OutputExtractor.lua:21-35) */
  int64_t *n_untouched; /* [n_pairs] pixels extractOutput would have left untouched */
  float *conf_marginal; /* getOutputConfidences2's confidence (opticalflow_model.lua:186-195): 1 where
                           some row marginal sum_dx p(dy, dx) exceeds prob_threshold, else 0.
                           Needs soft_yx (the 'mean' extraction it belongs to) */
} dm_extract_out;

/* Replaces prepareInput + model:forward + processOutput
 * (depth_estimation_api.lua:164-168, test_opticalflow.lua:347-355). */
int dm_match_extract(dm_ctx *ctx, const dm_pair *in, int maxh, int maxw, unsigned flags,
                     double prob_threshold, int h_img, int w_img, const dm_extract_out *out);

/* Radial variant: getTesterNetwork's matcher + argmin
 * (radial/radial_opticalflow_network.lua:56-74, radial/test_radial_opticalflow.lua:204-207).
 * flow[n][h1][w] = argmin_d - 1 as float (the reference copies idx into a FloatTensor). */
int dm_radial_match_extract(dm_ctx *ctx, const dm_pair *in, int h_win, float *flow, float *min_ssd);

/* ---- K2 stand-alone pieces (module-level drop-ins) ---------------------- */
/* nn.Minus + nn.SoftMax over the last dim (opticalflow_model.lua:94-109); in-place allowed */
int dm_neg_softmax(dm_ctx *ctx, const float *vol, int64_t rows, int k, float *out);
/* getOutputConfidences un-thresholded (opticalflow_model.lua:153-161); middle<=0: no tie rule;
 * take_min != 0: argmin (radial/radial_opticalflow_groundtruth.lua:88-95) */
int dm_argmax_tie(dm_ctx *ctx, const float *vol, int64_t rows, int k, int middle, int take_min,
                  int64_t *index, float *value);
/* extractoutput.extractOutput(input, scores, threshold, ret) (extract_output.cpp:63-155).
 * Untouched pixels keep the caller's ret/scores like the reference; n_untouched (host
 * pointer, may be NULL) receives their count. */
int dm_extract_output(dm_ctx *ctx, const float *input, int h, int w, int n, double threshold,
                      int64_t *ret, float *scores, int64_t *n_untouched);
/* the same extractOutput applied to the RAW SSD volume of a frame pair, volume never materialised: what the
 * ground-truth generators compute with `extractOutput(output, scores, 0.21, ret)` on the matcher's output
 * (radial/radial_opticalflow_groundtruth.lua:105, version2/groundtruth.lua:103: M = 4, the first four SSDs
 * above the threshold in scan order).  ret / scores: [n_pairs][h1][w1]; a pixel with no SSD above the
 * threshold gets ret = 0, score = 0 and is counted in n_untouched ([n_pairs], may be NULL). */
int dm_match_extract_raw_ssd(dm_ctx *ctx, const dm_pair *in, int maxh, int maxw, double threshold,
                             int64_t *ret, float *scores, int64_t *n_untouched);
/* extractoutput.extractOutputMarginalized (extract_output.cpp:157-255); retgd is zeroed */
int dm_extract_output_marginalized(dm_ctx *ctx, const float *input, int h, int w, int n,
                                   double threshold, double threshold_acc, int64_t *ret,
                                   int64_t *retgd);
/* nn.OutputExtractor:updateOutput (OutputExtractor.lua:21-35): y, x [rows] */
int dm_soft_mean(dm_ctx *ctx, const float *prob, int64_t rows, int maxh, int maxw, float *ymean,
                 float *xmean);
/* input:reshape(h,w,maxh,maxw):sum(4) (opticalflow_model.lua:191): [rows][maxh] */
int dm_marginal_x(dm_ctx *ctx, const float *prob, int64_t rows, int maxh, int maxw, float *pm);
/* x2yx + centre offset + canvas (opticalflow_model.lua:16-25,208-212,227-250) */
int dm_flow_canvas(dm_ctx *ctx, const int64_t *index, int h1, int w1, int maxh, int maxw,
                   int h_img, int w_img, float *full);

/* ---- K3: multiscale ----------------------------------------------------- */
/* Replaces inline.load("x2yxMulti2.c")(x, maxh, maxw, ratios, retx, rety)
 * (opticalflow_model_multiscale.lua:72-81, x2yxMulti2.c:1-95).  bug_compat=0 follows
 * the Lua spec x2yxMultiNumber (:83-132); bug_compat=1 reproduces the C file's
 * divergences (entries it falls through on are left untouched). */
int dm_x2yx_multi(dm_ctx *ctx, const int64_t *x, int h, int w, int maxh, int maxw,
                  const int *ratios, int nratios, int bug_compat, int64_t *rety, int64_t *retx);
/* nn.CascadingAddTable:updateOutput (CascadingAddTable.lua:108-135), forward only.
 * in/out: [nratios][rows][kh][kw]. */
int dm_cascade_add(dm_ctx *ctx, const float *in, int64_t rows, int kh, int kw, const int *ratios,
                   int nratios, float *out);
/* getModelMultiscale on prefiltered per-scale maps + processOutput
 * (opticalflow_model_multiscale.lua:175-333, opticalflow_model.lua:204-207):
 * per scale s: in1[s] is [C][H/r][W/r], in2[s] is [C][H/r+maxh-1][W/r+maxw-1]
 * (contiguous).  Outputs per full-resolution pixel: ring index (1-based, length-L
 * vector argmax with the middle-index tie rule) and decoded flow (Lua spec). */
int dm_multiscale_extract(dm_ctx *ctx, const float *const *in1, const float *const *in2,
                          int channels, int h, int w, int maxh, int maxw, const int *ratios,
                          int nratios, int64_t *index, int64_t *flow_y, int64_t *flow_x);
/* nn.SpatialDownSampling(r, r) (opticalflow_model_multiscale.lua:145) */
int dm_downsample_avg(dm_ctx *ctx, const float *in, int c, int h, int w, int r, float *out);

/* ---- K4: radial / polar remap ------------------------------------------- */
/* cartesian2polar(img, getC2PMask(...)) (radial/cartesian2polar.lua:4-49,91-93;
 * radial/test_radial_opticalflow.lua:184-194): analytic coordinates, no LUT in HBM.
 * dst: [c][hdst][wdst+lpad+rpad]. */
int dm_polar_remap(dm_ctx *ctx, const float *src, int c, int hsrc, int wsrc, double xcenter,
                   double ycenter, double rmax, double alpha, int lpad, int rpad, float *dst,
                   int hdst, int wdst);
/* cartesian2polar(polar, getP2CMask(...)) (radial/cartesian2polar.lua:51-89;
 * radial/test_radial_opticalflow.lua:217-218): polar -> cartesian. dst: [c][hdst][wdst]. */
int dm_polar_unmap(dm_ctx *ctx, const float *src, int c, int hsrc, int wsrc, double xcenter,
                   double ycenter, double rmax, double alpha, float *dst, int hdst, int wdst);
/* image.warp(src, field, 'bilinear', false) with an explicit LUT (field: [2][hd][wd], y then x) */
int dm_warp_bilinear(dm_ctx *ctx, const float *src, int c, int hs, int ws, const float *field,
                     int hd, int wd, float *dst);
/* the two LUT builders themselves, for callers that keep masks around */
int dm_c2p_mask(dm_ctx *ctx, int wdst, int hdst, double xcenter, double ycenter, int lpad,
                int rpad, double rmax, double alpha, float *mask);
int dm_p2c_mask(dm_ctx *ctx, int wsrc, int hsrc, int wdst, int hdst, double xcenter,
                double ycenter, double rmax, double alpha, float *mask);
/* flow2depth inline C (radial/radial_opticalflow_display.lua:6-58), before the /infty */
int dm_flow2depth(dm_ctx *ctx, const float *flow, int h, int w, float xcenter, float ycenter,
                  float infty, float *depth, float *confs);

/* ---- next rows: the steps right after the matching path ----------- */
/* postProcessImage(input, mask, winsize, method) (opticalflow_model.lua:323-472): input/output
 * [2][h][w] (y-flow, x-flow); method_max = 0: masked k x k median ('med'), 1: mode of the rounded
 * flow ('max', 16 x 16 histogram semantics).  winsize <= 5 (the reference's scratch size). */
int dm_post_process_image(dm_ctx *ctx, const float *input, const float *mask, int h, int w,
                          int winsize, int method_max, float *output);
/* enlargeMask(mask, ix, iy) (depth_estimation_api.lua:76-132), in place */
int dm_enlarge_mask(dm_ctx *ctx, float *mask, int h, int w, int ix, int iy);
/* radial(geometry, flow, mh, mw) (test_opticalflow.lua:143-193): flow [2][h][w] -> depth, conf */
int dm_radial_depth(dm_ctx *ctx, const float *flow, int h, int w, float mh, float mw, float infty,
                    float *ret, float *conf);
/* ARdroneAPI::computeDepthMapFromFlow (ardrone/ardrone_api.cpp:99-140): 6 x 6 masked mode filter
 * of the rounded x-flow, then depth = m * |x - w/2| / |flow| */
int dm_depth_from_xflow(dm_ctx *ctx, const float *xflow, const float *mask, int h, int w, float m,
                        float *depth, float *conf);

/* sfm2.removeEgoMotion(im, K, R) (out-of-tree sfm2; depth_estimation_api.lua:147,
 * radial/test_radial_opticalflow.lua:192): homography gather.  dst[k][y][x] = bilinear sample of
 * src[k] at hmat * (x, y, 1) (hmat: 9 doubles row-major, HOST pointer, e.g. K * R * K^-1);
 * outside the source frame dst = 0 and mask (optional, [hd][wd]) = 0, else mask = 1. */
int dm_warp_homography(dm_ctx *ctx, const float *src, int c, int hs, int ws, const double *hmat,
                       int hd, int wd, float *dst, float *mask);

/* ---- next row 3: the feature extractor in front of the path ---------------------- */
/* One layer of getFilter (opticalflow_model.lua:45-79; radial/radial_opticalflow_network.lua:6-31).
 * n_conn == 0: nn.SpatialConvolution(n_in, n_out, kw, kh), weight [n_out][n_in][kh][kw].
 * n_conn  > 0: nn.SpatialConvolutionMap(conn, kw, kh), conn = n_conn rows of 1-based (from, to)
 *              as nn.tables.random builds them, weight [n_conn][kh][kw].
 * bias [n_out]; tanh_after = an nn.Tanh follows.  weight, bias and conn are HOST pointers, read
 * at dm_filter_create only. */
typedef struct dm_conv_layer {
  int32_t n_in, n_out, kh, kw;
  int32_t n_conn;
  int32_t tanh_after;
  const int32_t *conn;
  const float *weight;
  const float *bias;
} dm_conv_layer;
typedef struct dm_filter dm_filter;
/* getFilter(geometry): the weights are packed and kept on the context's device */
int dm_filter_create(dm_ctx *ctx, const dm_conv_layer *layers, int n_layers, dm_filter **out);
int dm_filter_destroy(dm_filter *filter);
int dm_filter_output_size(const dm_filter *filter, int h, int w, int pad_l, int pad_r, int pad_t,
                          int pad_b, int *channels, int *hout, int *wout);
/* filter:forward on n_img images [n_img][n_in][h][w] (both frames of a pair share the weights,
 * opticalflow_model.lua:85-90), zero-padded first like nn.SpatialZeroPadding(l, r, t, b)
 * (opticalflow_model_multiscale.lua:136-146) -> [n_img][n_out][hout][wout] */
int dm_filter_forward(dm_ctx *ctx, const dm_filter *filter, const float *in, int n_img, int h, int w,
                      int pad_l, int pad_r, int pad_t, int pad_b, float *out);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif
