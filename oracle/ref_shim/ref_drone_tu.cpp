/*
 * ref_drone_tu.cpp -- compiles the body of ARdroneAPI::computeDepthMapFromFlow
 * (/root/reference/ardrone/ardrone_api.cpp:99-140), cut out verbatim by extract_inline.py into
 * _ref/drone_depth.inc, against a minimal stand-in for cv::Mat_<float> and the ARdroneAPI members
 * it touches (depthMap, confidenceMap, getIMUTranslation).  TEST INFRASTRUCTURE ONLY: it pins
 * orc_depth_from_xflow / dm_depth_from_xflow to the reference's own statements.
 *
 * The stand-in zero-fills new matrices; cv::Mat leaves them uninitialised, so the reference's
 * depthMap is unspecified where confidenceMap is 0 -- the tests compare depth only where the
 * confidence is 1, and the confidence everywhere.
 */
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
using namespace std;   /* as ardrone_api.cpp:13 */

namespace {
struct SizeShim {
  int width, height;
};
struct matf {   /* common.h:11: typedef cv::Mat_<float> matf */
  int rows, cols;
  float *p;
  bool own;
  matf() : rows(0), cols(0), p(0), own(false) {}
  explicit matf(SizeShim s) : rows(s.height), cols(s.width), p((float *)calloc((size_t)s.height * s.width, 4)), own(true) {}
  matf(int r, int c, float *data) : rows(r), cols(c), p(data), own(false) {}
  matf(const matf &o) : rows(o.rows), cols(o.cols), p(o.p), own(false) {}   /* cv::Mat copies share the data */
  matf &operator=(const matf &o) {
    rows = o.rows; cols = o.cols; p = o.p; own = false;   /* the owner is leaked: test code, a few KB */
    return *this;
  }
  SizeShim size() const { SizeShim s = {cols, rows}; return s; }
  float &operator()(int r, int c) const { return p[(size_t)r * cols + c]; }
};

struct ARdroneAPI {
  matf depthMap, confidenceMap;
  float m_;
  matf getIMUTranslation() const { return matf(1, 1, const_cast<float *>(&m_)); }
  void computeDepthMapFromFlow(const matf &xflow, const matf &mask) {
#include "drone_depth.inc"
  }
};
}  // namespace

extern "C" void ref_depth_from_xflow(const float *xflow, const float *mask, int h, int w, float m, float *depth,
                                     float *conf) {
  ARdroneAPI api;
  api.m_ = m;
  api.computeDepthMapFromFlow(matf(h, w, const_cast<float *>(xflow)), matf(h, w, const_cast<float *>(mask)));
  memcpy(depth, api.depthMap.p, sizeof(float) * (size_t)h * w);
  memcpy(conf, api.confidenceMap.p, sizeof(float) * (size_t)h * w);
  free(api.depthMap.p);
  free(api.confidenceMap.p);
}
