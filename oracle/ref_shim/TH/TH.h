/*
 * TH/TH.h -- minimal stand-in for Torch7's TH tensor API (see ../luaT.h).
 * Only what version2/extract_output.cpp and x2yxMulti2.c touch.
 * TEST INFRASTRUCTURE ONLY.  Written from scratch for this repo; not Torch code.
 */
#ifndef DM_REF_SHIM_TH_H
#define DM_REF_SHIM_TH_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct THFloatTensor {
  long size[4];
  long stride[4];
  int nDimension;
  float *data;
  int owns;
} THFloatTensor;

typedef struct THLongTensor {
  long size[4];
  long stride[4];
  int nDimension;
  long *data;
  int owns;
} THLongTensor;

THFloatTensor *THFloatTensor_newContiguous(THFloatTensor *t);
THFloatTensor *THFloatTensor_newWithSize4d(long a, long b, long c, long d);
float *THFloatTensor_data(THFloatTensor *t);
int THFloatTensor_isContiguous(THFloatTensor *t);
void THFloatTensor_zero(THFloatTensor *t);
void THFloatTensor_free(THFloatTensor *t);
long *THLongTensor_data(THLongTensor *t);
void THLongTensor_zero(THLongTensor *t);

/* wrapper helpers */
THFloatTensor *shim_float_view(float *data, int nd, const long *size);
THLongTensor *shim_long_view(long *data, int nd, const long *size);
void shim_long_free(THLongTensor *t);

#ifdef __cplusplus
}
#endif
#endif
