/*
 * ref_inline_tu.c -- wraps the reference's `inline.load` C bodies (extracted verbatim by
 * oracle/extract_inline.py into oracle/_ref/*.inc) the way Torch's `inline` package does:
 * each body becomes `int f(lua_State *L) { <body> return 0; }`.  TEST INFRASTRUCTURE ONLY.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "luaT.h"
#include "TH/TH.h"

#include "pp_preamble.inc"

static int body_pp_fmax(lua_State *L) {
#include "pp_fmax.inc"
  return 0;
}
static int body_pp_fmed(lua_State *L) {
#include "pp_fmed.inc"
  return 0;
}
static int body_radial_depth(lua_State *L) {
#include "radial_depth.inc"
#undef min
  return 0;
}
static int body_enlarge_mask(lua_State *L) {
#include "enlarge_mask.inc"
  return 0;
}
static int body_c2p_mask(lua_State *L) {
#include "c2p_mask.inc"
  return 0;
}
static int body_p2c_mask(lua_State *L) {
#include "p2c_mask.inc"
  return 0;
}
static int body_flow2depth(lua_State *L) {
#include "flow2depth.inc"
#undef square
  return 0;
}

static THFloatTensor *f2(float *d, long a, long b) {
  const long s[2] = {a, b};
  return shim_float_view(d, 2, s);
}
static THFloatTensor *f3(float *d, long a, long b, long c) {
  const long s[3] = {a, b, c};
  return shim_float_view(d, 3, s);
}

/* postProcessImage's two kernels: flow [2][h][w], mask [h][w], ret [2][h][w] (pre-zeroed by caller) */
int ref_pp_filter(int method_max, const float *flow, const float *mask, int k, long h, long w, float *ret) {
  THFloatTensor *tf = f3((float *)flow, 2, h, w), *tm = f2((float *)mask, h, w), *tr = f3(ret, 2, h, w);
  lua_State L;
  shim_reset(&L);
  shim_push_float_tensor(&L, tf);
  shim_push_float_tensor(&L, tm);
  shim_push_number(&L, k);
  shim_push_float_tensor(&L, tr);
  const int rc = method_max ? body_pp_fmax(&L) : body_pp_fmed(&L);
  THFloatTensor_free(tf); THFloatTensor_free(tm); THFloatTensor_free(tr);
  return rc;
}

int ref_radial_depth(const float *flow, long h, long w, double mh, double mw, float *ret, float *conf,
                     double infty) {
  THFloatTensor *tf = f3((float *)flow, 2, h, w), *tr = f2(ret, h, w), *tc = f2(conf, h, w);
  lua_State L;
  shim_reset(&L);
  shim_push_float_tensor(&L, tf);
  shim_push_number(&L, mh);
  shim_push_number(&L, mw);
  shim_push_float_tensor(&L, tr);
  shim_push_float_tensor(&L, tc);
  shim_push_number(&L, infty);
  const int rc = body_radial_depth(&L);
  THFloatTensor_free(tf); THFloatTensor_free(tr); THFloatTensor_free(tc);
  return rc;
}

int ref_enlarge_mask(float *mask, long h, long w, int ix, int iy) {
  THFloatTensor *tm = f2(mask, h, w);
  lua_State L;
  shim_reset(&L);
  shim_push_float_tensor(&L, tm);
  shim_push_number(&L, ix);
  shim_push_number(&L, iy);
  const int rc = body_enlarge_mask(&L);
  THFloatTensor_free(tm);
  return rc;
}

/* mask: [2][hdst][wdst] contiguous (the un-padded view of getC2PMask) */
int ref_c2p_mask(float *mask, long hdst, long wdst, double xc, double yc, double kr, double ktheta,
                 double alpha) {
  THFloatTensor *tm = f3(mask, 2, hdst, wdst);
  lua_State L;
  shim_reset(&L);
  shim_push_float_tensor(&L, tm);
  shim_push_number(&L, xc);
  shim_push_number(&L, yc);
  shim_push_number(&L, kr);
  shim_push_number(&L, ktheta);
  shim_push_number(&L, alpha);
  const int rc = body_c2p_mask(&L);
  THFloatTensor_free(tm);
  return rc;
}

int ref_p2c_mask(float *mask, long hdst, long wdst, double xc, double yc, double kx, double ky,
                 double pi2, double invalpha) {
  THFloatTensor *tm = f3(mask, 2, hdst, wdst);
  lua_State L;
  shim_reset(&L);
  shim_push_float_tensor(&L, tm);
  shim_push_number(&L, xc);
  shim_push_number(&L, yc);
  shim_push_number(&L, kx);
  shim_push_number(&L, ky);
  shim_push_number(&L, pi2);
  shim_push_number(&L, invalpha);
  const int rc = body_p2c_mask(&L);
  THFloatTensor_free(tm);
  return rc;
}

int ref_flow2depth(const float *flow, long h, long w, float *depth, float *confs, double xc, double yc,
                   double infty) {
  THFloatTensor *tf = f2((float *)flow, h, w), *td = f2(depth, h, w), *tc = f2(confs, h, w);
  lua_State L;
  shim_reset(&L);
  shim_push_float_tensor(&L, tf);
  shim_push_float_tensor(&L, td);
  shim_push_float_tensor(&L, tc);
  shim_push_number(&L, xc);
  shim_push_number(&L, yc);
  shim_push_number(&L, infty);
  const int rc = body_flow2depth(&L);
  THFloatTensor_free(tf); THFloatTensor_free(td); THFloatTensor_free(tc);
  return rc;
}
