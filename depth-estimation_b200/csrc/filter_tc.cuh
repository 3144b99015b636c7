// filter_tc.cuh -- interface of the tensor-core (tcgen05) convolution layers, see filter_tc.cu
#pragma once
#include <vector>

#include "dm_common.cuh"

namespace dm {

struct TcPlan {
  int ok;              // the layer fits the tensor-core kernel
  int KP, Ktot;        // kernel rows padded to 8 per input plane; n_in * KP
  int KtotB;           // K extent of the weight operand: n_in * (KP + 4), one spare chunk per plane
  int G, ngroups;      // output planes per group (per GEMM), number of groups
  int Npad;            // G * kW padded to a multiple of 16
  size_t smem;
};

int tc_plan_layer(const dm_ctx *ctx, int n_in, int n_out, int kh, int kw, TcPlan *p);
void tc_pack_weights(const TcPlan &p, int n_in, int n_out, int kh, int kw, int n_conn, const int *conn, const float *weight,
                     std::vector<float> *out);
int tc_launch_layer(dm_ctx *ctx, const TcPlan &p, const float *B, const float *bias, int n_in, int n_out, int kh, int kw,
                    int tanh_after, const float *in, float *out, int n_img, int h, int w, int pad_l, int pad_r, int pad_t,
                    int pad_b);

}  // namespace dm
