-- depthmatch_ffi.lua -- LuaJIT FFI binding of libdepthmatch.so (include/depthmatch.h).
--
-- NOT EXECUTED in this repository's CI: the build image has no Lua/LuaJIT/Torch7.  The
-- declarations below are the header's, verbatim; the tested host layer with the same names
-- is the Python package ../depthmatch (tests/ call the same C ABI through ctypes).
-- Requires Torch7 on LuaJIT (torch.data(tensor) returns an FFI pointer).
local ffi = require 'ffi'

ffi.cdef[[
typedef struct dm_ctx dm_ctx;
typedef struct dm_pair {
  const float *in1; const float *in2;
  int32_t n_pairs, channels, h1, w1, h2, w2;
  int64_t in1_stride_n, in1_stride_c, in1_stride_y;
  int64_t in2_stride_n, in2_stride_c, in2_stride_y;
} dm_pair;
typedef struct dm_extract_out {
  int64_t *index; float *min_ssd; float *pmax; float *flow_full;
  int64_t *index_thr; float *score_thr; float *soft_yx; int64_t *n_untouched;
  float *conf_marginal;
} dm_extract_out;
int dm_version(void);
const char *dm_last_error(void);
int dm_create(int device, dm_ctx **ctx);
int dm_destroy(dm_ctx *ctx);
int dm_synchronize(dm_ctx *ctx);
int dm_match_volume(dm_ctx *ctx, const dm_pair *in, int maxh, int maxw, int mode, float *out);
int dm_match_extract(dm_ctx *ctx, const dm_pair *in, int maxh, int maxw, unsigned flags,
                     double prob_threshold, int h_img, int w_img, const dm_extract_out *out);
int dm_radial_match_extract(dm_ctx *ctx, const dm_pair *in, int h_win, float *flow, float *min_ssd);
int dm_neg_softmax(dm_ctx *ctx, const float *vol, int64_t rows, int k, float *out);
int dm_argmax_tie(dm_ctx *ctx, const float *vol, int64_t rows, int k, int middle, int take_min,
                  int64_t *index, float *value);
int dm_extract_output(dm_ctx *ctx, const float *input, int h, int w, int n, double threshold,
                      int64_t *ret, float *scores, int64_t *n_untouched);
int dm_match_extract_raw_ssd(dm_ctx *ctx, const dm_pair *in, int maxh, int maxw, double threshold,
                             int64_t *ret, float *scores, int64_t *n_untouched);
int dm_extract_output_marginalized(dm_ctx *ctx, const float *input, int h, int w, int n,
                                   double threshold, double threshold_acc, int64_t *ret,
                                   int64_t *retgd);
int dm_soft_mean(dm_ctx *ctx, const float *prob, int64_t rows, int maxh, int maxw, float *ymean,
                 float *xmean);
int dm_x2yx_multi(dm_ctx *ctx, const int64_t *x, int h, int w, int maxh, int maxw,
                  const int *ratios, int nratios, int bug_compat, int64_t *rety, int64_t *retx);
int dm_cascade_add(dm_ctx *ctx, const float *in, int64_t rows, int kh, int kw, const int *ratios,
                   int nratios, float *out);
int dm_multiscale_extract(dm_ctx *ctx, const float *const *in1, const float *const *in2,
                          int channels, int h, int w, int maxh, int maxw, const int *ratios,
                          int nratios, int64_t *index, int64_t *flow_y, int64_t *flow_x);
int dm_polar_remap(dm_ctx *ctx, const float *src, int c, int hsrc, int wsrc, double xcenter,
                   double ycenter, double rmax, double alpha, int lpad, int rpad, float *dst,
                   int hdst, int wdst);
int dm_polar_unmap(dm_ctx *ctx, const float *src, int c, int hsrc, int wsrc, double xcenter,
                   double ycenter, double rmax, double alpha, float *dst, int hdst, int wdst);
int dm_warp_bilinear(dm_ctx *ctx, const float *src, int c, int hs, int ws, const float *field,
                     int hd, int wd, float *dst);
int dm_flow2depth(dm_ctx *ctx, const float *flow, int h, int w, float xcenter, float ycenter,
                  float infty, float *depth, float *confs);
int dm_warp_homography(dm_ctx *ctx, const float *src, int c, int hs, int ws, const double *hmat,
                       int hd, int wd, float *dst, float *mask);
int dm_post_process_image(dm_ctx *ctx, const float *input, const float *mask, int h, int w,
                          int winsize, int method_max, float *output);
int dm_enlarge_mask(dm_ctx *ctx, float *mask, int h, int w, int ix, int iy);
int dm_radial_depth(dm_ctx *ctx, const float *flow, int h, int w, float mh, float mw, float infty,
                    float *ret, float *conf);
int dm_depth_from_xflow(dm_ctx *ctx, const float *xflow, const float *mask, int h, int w, float m,
                        float *depth, float *conf);
typedef struct dm_conv_layer {
  int32_t n_in, n_out, kh, kw;
  int32_t n_conn;
  int32_t tanh_after;
  const int32_t *conn;
  const float *weight;
  const float *bias;
} dm_conv_layer;
typedef struct dm_filter dm_filter;
int dm_filter_create(dm_ctx *ctx, const dm_conv_layer *layers, int n_layers, dm_filter **out);
int dm_filter_destroy(dm_filter *filter);
int dm_filter_output_size(const dm_filter *filter, int h, int w, int pad_l, int pad_r, int pad_t,
                          int pad_b, int *channels, int *hout, int *wout);
int dm_filter_forward(dm_ctx *ctx, const dm_filter *filter, const float *in, int n_img, int h, int w,
                      int pad_l, int pad_r, int pad_t, int pad_b, float *out);
]]

local M = {}
M.C = ffi.load(os.getenv('DEPTHMATCH_SO') or 'depthmatch')
M.DM_VOLUME_SSD, M.DM_VOLUME_NEG_SOFTMAX = 0, 1
M.DM_FLAG_TIE_MIDDLE, M.DM_FLAG_EXACT_SSD, M.DM_FLAG_ASYNC, M.DM_FLAG_DIFF_SSD = 1, 2, 4, 8

function M.check(status)
   if status ~= 0 then
      error('libdepthmatch: ' .. ffi.string(M.C.dm_last_error()), 2)
   end
end

local ctxp = ffi.new('dm_ctx*[1]')
M.check(M.C.dm_create(tonumber(os.getenv('DEPTHMATCH_DEVICE') or 0), ctxp))
M.ctx = ffi.gc(ctxp[0], M.C.dm_destroy)

-- {in1, in2}: C x H x W FloatTensors, innermost stride 1 (narrow()ed views are fine)
function M.pair(in1, in2)
   assert(in1:stride(3) == 1 and in2:stride(3) == 1, 'innermost stride must be 1')
   local p = ffi.new('dm_pair')
   p.in1, p.in2 = torch.data(in1), torch.data(in2)
   p.n_pairs, p.channels = 1, in1:size(1)
   p.h1, p.w1, p.h2, p.w2 = in1:size(2), in1:size(3), in2:size(2), in2:size(3)
   p.in1_stride_n, p.in1_stride_c, p.in1_stride_y = 0, in1:stride(1), in1:stride(2)
   p.in2_stride_n, p.in2_stride_c, p.in2_stride_y = 0, in2:stride(1), in2:stride(2)
   return p
end

return M
