/*
 * ref_extract_tu.cpp -- translation unit that compiles the reference's
 * version2/extract_output.cpp AS IT LIES under /root/reference (path given by
 * -DDM_REF_EXTRACT_SRC=...) against the shim headers, and exports plain-C
 * wrappers around its static Lua C functions.  TEST INFRASTRUCTURE ONLY.
 */
#include DM_REF_EXTRACT_SRC

#include <stdint.h>

extern "C" int ref_extract_output(const float *input, long h, long w, long n, double threshold,
                                  int64_t *ret, float *scores) {
  const long isz[3] = {h, w, n}, osz[2] = {h, w};
  THFloatTensor *tin = shim_float_view(const_cast<float *>(input), 3, isz);
  THFloatTensor *tsc = shim_float_view(scores, 2, osz);
  THLongTensor *tret = shim_long_view(reinterpret_cast<long *>(ret), 2, osz);
  lua_State L;
  shim_reset(&L);
  shim_push_float_tensor(&L, tin);
  shim_push_float_tensor(&L, tsc);
  shim_push_number(&L, threshold);
  shim_push_long_tensor(&L, tret);
  const int rc = ExtractOutput(&L);
  THFloatTensor_free(tin);
  THFloatTensor_free(tsc);
  shim_long_free(tret);
  return rc;
}

extern "C" int ref_extract_output_marginalized(const float *input, long h, long w, long n,
                                               double threshold, double threshold_acc,
                                               int64_t *ret, int64_t *retgd) {
  const long isz[3] = {h, w, n}, osz[2] = {h, w};
  THFloatTensor *tin = shim_float_view(const_cast<float *>(input), 3, isz);
  THLongTensor *tret = shim_long_view(reinterpret_cast<long *>(ret), 2, osz);
  THLongTensor *tgd = shim_long_view(reinterpret_cast<long *>(retgd), 2, osz);
  lua_State L;
  shim_reset(&L);
  shim_push_float_tensor(&L, tin);
  shim_push_number(&L, threshold);
  shim_push_number(&L, threshold_acc);
  shim_push_long_tensor(&L, tret);
  shim_push_long_tensor(&L, tgd);
  const int rc = ExtractOutputMarginalized(&L);
  THFloatTensor_free(tin);
  shim_long_free(tret);
  shim_long_free(tgd);
  return rc;
}
