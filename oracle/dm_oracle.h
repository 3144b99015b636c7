/*
 * dm_oracle.h -- CPU oracle for the dense-matching hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / CPU baseline.
 *
 * Every function is a plain-C restatement of one piece of the reference
 * (MichaelMathieu/depth-estimation, Torch7/Lua) and cites the file:line it
 * follows.  Arithmetic that lives in un-vendored, un-pinned third-party code
 * (Torch7 nn / nnx / image, 2012) is restated from the reference's call sites
 * and tests; for those functions PARITY IS UNPINNED (no golden vector exists in
 * the reference) and the header of each function says so.  Functions whose
 * source is in the reference tree (extract_output.cpp, x2yxMulti2.c) are pinned
 * against that source compiled as-is into oracle/_ref/ (see oracle/Makefile).
 *
 * Conventions: tensors are contiguous row-major fp32 unless noted; "Long"
 * tensors are int64_t; indices returned to the caller are 1-based exactly where
 * the reference's are.  Compile with -ffp-contract=off: the reference was built
 * for pre-FMA x86, every multiply and add rounds separately.
 */
#ifndef DM_ORACLE_H
#define DM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- matching.c ------------------------------------------------------- */

/* nn.SpatialMatching(maxh,maxw,false):updateOutput({in1,in2})
 * (out-of-tree nnx; call sites opticalflow_model.lua:93, semantics pinned
 * behaviourally by tests/test_multiscale.lua:149-166, tests/test_patches.lua:46-60).
 * out[y][x][dy][dx] = sum_k (in1[k,y,x] - in2[k,y+dy,x+dx])^2, k ascending, fp32
 * accumulator, OpenMP over y.  Requires H2 >= H1+maxh-1, W2 >= W1+maxw-1.
 * PARITY UNPINNED (no numeric fixture in the reference). */
void orc_spatial_matching(const float *in1, const float *in2, int C, int H1, int W1,
                          int H2, int W2, int maxh, int maxw, float *out, int nthreads);

/* nn.SpatialRadialMatching(hWin) (author's nnx fork, out-of-tree; call site
 * radial/radial_opticalflow_network.lua:32-34): out[y][x][d] =
 * sum_k (in1[k,y,x]-in2[k,y+d,x])^2.  PARITY UNPINNED. */
void orc_radial_matching(const float *in1, const float *in2, int C, int H1, int W,
                         int H2, int hWin, float *out, int nthreads);

/* nn.Minus + nn.SoftMax over the last dimension (opticalflow_model.lua:94-109).
 * p = exp(max - (-v))... i.e. softmax(-v): subtract the row max, exp, sum in
 * double, scale by 1/sum (Torch7 nn/generic/SoftMax.c of 2012, recollection).
 * exp_mode 0 = exact expf (canonical); 1 = THExpMinusApprox
 * (degree-4 polynomial to the 8th power, 0 beyond 13: what 2012 Torch7 shipped;
 * recollection, kept to bound the effect).  PARITY UNPINNED. */
void orc_neg_softmax(const float *vol, int64_t rows, int K, int exp_mode, float *out,
                     int nthreads);

/* getOutputConfidences, un-thresholded branch (opticalflow_model.lua:153-161):
 * TH max (strict >, first occurrence) + zero-flow tie rule.  middle is 1-based;
 * idx out is 1-based. */
void orc_argmax_tie(const float *prob, int64_t rows, int K, int middle, int64_t *idx,
                    float *maxval);

/* Same on raw distances with min (radial_opticalflow_groundtruth.lua:88-95,
 * radial/test_radial_opticalflow.lua:205-207): first-occurrence argmin,
 * optional tie rule (middle <= 0 disables).  idx 1-based. */
void orc_argmin_tie(const float *vol, int64_t rows, int K, int middle, int64_t *idx,
                    float *minval);

/* Gap between the two largest entries of each row, relative to the largest
 * (north_star near-tie rule: indices may differ where this is < 1e-5). */
void orc_top2_relgap(const float *prob, int64_t rows, int K, float *relgap);

/* nn.OutputExtractor / getOutputConfidences2 (OutputExtractor.lua:21-35,
 * opticalflow_model.lua:171-185): x = sum_k p_k*col_k, y = sum_k p_k*row_k
 * (1-based), product in fp32, TH sum in double. */
void orc_soft_mean(const float *prob, int64_t rows, int maxh, int maxw, float *ymean,
                   float *xmean);

/* marginal over x (opticalflow_model.lua:191): pm[r][i] = sum_j p[r][i*maxw+j]. */
void orc_marginal_x(const float *prob, int64_t rows, int maxh, int maxw, float *pm);

/* x2yx + centre offset + canvas embedding (opticalflow_model.lua:16-25,208-212,
 * 227-250).  idx is 1-based h1*w1; full is 2*hImg*wImg zero-filled then pasted at
 * (floor((hImg-h1)/2), floor((wImg-w1)/2)); full[0]=y-flow, full[1]=x-flow. */
void orc_flow_canvas(const int64_t *idx, int h1, int w1, int maxh, int maxw, int hImg,
                     int wImg, float *full);

/* ---- extract.c -------------------------------------------------------- */

/* extractoutput.extractOutput (extract_output.cpp:63-155 ==
 * version2/extract_output.cpp:63-155).  Pixels with nothing above threshold are
 * left untouched, exactly like the reference.  Returns the number of pixels
 * written.  PINNED against oracle/_ref (the reference source compiled as-is). */
int64_t orc_extract_output(const float *input, int h, int w, int n, double threshold,
                           int64_t *ret, float *scores);

/* extractoutput.extractOutputMarginalized (version2/extract_output.cpp:157-255).
 * retgd is zeroed first (:167), ret is not.  PINNED against oracle/_ref. */
int64_t orc_extract_output_marginalized(const float *input, int h, int w, int n,
                                        double threshold, double threshold_acc,
                                        int64_t *ret, int64_t *retgd);

/* ---- multiscale.c ----------------------------------------------------- */

/* yx2xMulti (opticalflow_model_multiscale.lua:10-52): (dy,dx) in full-res pixels
 * -> 1-based ring index; returns 0 where the Lua asserts. */
int64_t orc_yx2x_multi(int maxh, int maxw, const int *ratios, int nratios, double y,
                       double x);

/* x2yxMultiNumber (opticalflow_model_multiscale.lua:83-132), the scalar Lua spec.
 * Returns 0 on success, -1 where the Lua asserts. */
int orc_x2yx_multi_number(int maxh, int maxw, const int *ratios, int nratios, int64_t x,
                          int64_t *outy, int64_t *outx);

/* x2yxMulti2.c:1-95 semantics INCLUDING its divergences from the Lua spec
 * (ratios shifted by one, integer ceil, lengths without *d, strict <); entries
 * the C falls through on are left untouched.  PINNED against oracle/_ref. */
void orc_x2yx_multi2_bugcompat(const int64_t *xim, int h, int w, int maxh, int maxw,
                               const int *ratios, int nratios, int64_t *retx,
                               int64_t *rety);

/* Vector of L entries length for a geometry (multiscale.lua:293-333). */
int orc_multiscale_length(int maxh, int maxw, const int *ratios, int nratios);

/* CascadingAddTable:updateOutput (CascadingAddTable.lua:108-135), forward only,
 * normalisers disabled as in the reference (:29,:46,:61).  in/out: nratios
 * tensors, each rows*Kh*Kw, concatenated scale-major.  SpatialReSamplingEx
 * 'average' when upsampling by an integer factor = nearest replication (pinned
 * by tests/test_multiscale.lua:180-187). */
void orc_cascade_add(const float *in, int64_t rows, int Kh, int Kw, const int *ratios,
                     int nratios, float *out);

/* Ring extraction + join (multiscale.lua:293-333): rows x L vector. */
void orc_ring_join(const float *casc, int64_t rows, int maxh, int maxw, const int *ratios,
                   int nratios, float *outvec);

/* nn.SpatialDownSampling(r,r): r x r average (multiscale.lua:145), and nearest
 * upsampling of a per-pixel K-vector map (pyramid's SpatialUpSampling). */
void orc_downsample_avg(const float *in, int C, int H, int W, int r, float *out);
void orc_upsample_nearest_rows(const float *in, int h, int w, int K, int r, float *out);

/* ---- radial.c --------------------------------------------------------- */

/* getC2PMask (radial/cartesian2polar.lua:4-49): mask is 2 x hdst x (wdst+lpad+rpad),
 * mask[0]=y, mask[1]=x, columns circularly padded. */
void orc_c2p_mask(int wdst, int hdst, double xcenter, double ycenter, int lpad, int rpad,
                  double rmax, double alpha, float *mask);

/* getP2CMask (radial/cartesian2polar.lua:51-89): 2 x hdst x wdst. */
void orc_p2c_mask(int wsrc, int hsrc, int wdst, int hdst, double xcenter, double ycenter,
                  double rmax, double alpha, float *mask);

/* getRMax (radial/radial_opticalflow_polar.lua:4-10). */
double orc_get_rmax(int h, int w, double ex, double ey);

/* image.warp(src, field, 'bilinear', false) (out-of-tree Torch7 image package,
 * 2012: coordinates clamped to the border, 4-neighbour weights, MIN on the +1
 * neighbours; recollection).  field[0]=y, field[1]=x absolute coordinates.
 * PARITY UNPINNED. */
void orc_warp_bilinear(const float *src, int C, int hs, int ws, const float *field, int hd,
                       int wd, float *dst);

/* sfm2.removeEgoMotion as a homography gather (out-of-tree sfm2; depth_estimation_api.lua:147):
 * dst(x,y) = bilinear src(hmat*(x,y,1)); outside: 0, mask 0.  PARITY UNPINNED. */
void orc_warp_homography(const float *src, int C, int hs, int ws, const double *hmat, int hd, int wd,
                         float *dst, float *mask);

/* flow2depth inline C (radial/radial_opticalflow_display.lua:6-58). */
void orc_flow2depth(const float *flow, int h, int w, float xcenter, float ycenter,
                    float infty, float *depth, float *confs);

/* ---- postprocess.c ("next" rows: the steps right after the matching path) ---- */

/* postProcessImage(input, mask, winsize, method) (opticalflow_model.lua:323-472): input/output
 * [2][h][w] (y-flow, x-flow), masked k x k median ('med') or mode ('max') filter.
 * PINNED against the reference's inline C (oracle/_ref). */
void orc_post_process_image(const float *input, const float *mask, int h, int w, int k, int method_max,
                            float *output);
void orc_pp_median(const float *flow, const float *mask, int h, int w, int k, float *ret);
void orc_pp_mode(const float *flow, const float *mask, int h, int w, int k, float *ret);
/* enlargeMask (depth_estimation_api.lua:76-132), in place.  PINNED against oracle/_ref. */
void orc_enlarge_mask(float *mask, int h, int w, int ix, int iy);
/* radial() (test_opticalflow.lua:143-193).  PINNED against oracle/_ref. */
void orc_radial_depth(const float *flow, int h, int w, float mh, float mw, float infty, float *ret,
                      float *conf);
/* ARdroneAPI::computeDepthMapFromFlow (ardrone/ardrone_api.cpp:99-140).  PARITY UNPINNED. */
void orc_depth_from_xflow(const float *xflow, const float *mask, int h, int w, float m, float *depth,
                          float *conf);

/* ---- filter.c ("next" row 3: the feature extractor in front of the path) ---- */

/* One layer of getFilter (opticalflow_model.lua:45-79): nn.SpatialConvolution (conn == NULL,
 * weight [n_out][n_in][kh][kw]) or nn.SpatialConvolutionMap (conn = n_conn 1-based (from,to)
 * rows as nn.tables.random makes them, weight [n_conn][kh][kw]), valid cross-correlation on the
 * zero-padded input (nn.SpatialZeroPadding, multiscale.lua:146), optional nn.Tanh.
 * out: n_out x (h+pad_t+pad_b-kh+1) x (w+pad_l+pad_r-kw+1).  PARITY UNPINNED (Torch7 nn). */
void orc_conv_layer(const float *in, int n_in, int h, int w, const float *weight, const float *bias,
                    int n_out, int kh, int kw, const int32_t *conn, int n_conn, int pad_l, int pad_r,
                    int pad_t, int pad_b, int tanh_after, float *out, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
