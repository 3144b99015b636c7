/*
 * extract.c -- oracle restatement of extractoutput.extractOutput and
 * extractOutputMarginalized.  TEST INFRASTRUCTURE ONLY (see dm_oracle.h).
 *
 * Follows version2/extract_output.cpp:17-155 and :157-255 (the root
 * extract_output.cpp:17-155 is identical but the file does not compile: it
 * defines ExtractOutputMarginalized twice, :157 and :257).
 * PINNED: tests/test_oracle_ref.py compares these against the reference source
 * compiled as-is into oracle/_ref/ on random and adversarial inputs.
 */
#include "dm_oracle.h"

#include <stddef.h>

/* Compare-exchange pairs of the reference's two sorting networks, in the order
 * the reference applies them (version2/extract_output.cpp:27-33 and :35-61).
 * Each exchange moves the larger value to the lower slot when strictly greater
 * (:17-26), so ties keep their slots. */
static const unsigned char NET4[][2] = {{0, 2}, {1, 3}, {0, 1}, {2, 3}, {1, 2}};
static const unsigned char NET8[][2] = {
    {0, 1}, {2, 3}, {4, 5}, {6, 7}, {0, 2}, {1, 3}, {4, 6}, {5, 7}, {1, 2}, {5, 6},
    {0, 4}, {3, 7}, {1, 5}, {2, 6}, {1, 4}, {3, 6}, {2, 4}, {3, 5}, {3, 4}};

typedef struct {
  float val[8];
  float pos[8]; /* the reference stores the 1-based position as a float (:101) */
  int M;
} highs_t;

static void collect(const float *v, int n, double threshold, highs_t *hs) {
  for (int k = 0; k < hs->M; ++k) hs->val[k] = hs->pos[k] = 0.0f; /* Tensor_(zero), :91 */
  int got = 0;
  for (int i = 0; i < n && got < hs->M; ++i) {
    if (v[i] > threshold) { /* float promoted to double, :98 */
      hs->val[got] = v[i];
      hs->pos[got] = (float)(i + 1);
      ++got;
    }
  }
}

static void sort_desc(highs_t *hs) {
  const unsigned char(*net)[2] = hs->M == 4 ? NET4 : NET8;
  const int nex = hs->M == 4 ? 5 : 19;
  for (int e = 0; e < nex; ++e) {
    const int a = net[e][0], b = net[e][1];
    if (hs->val[b] > hs->val[a]) {
      float t = hs->val[a];
      hs->val[a] = hs->val[b];
      hs->val[b] = t;
      t = hs->pos[a];
      hs->pos[a] = hs->pos[b];
      hs->pos[b] = t;
    }
  }
}

/* :124-129: running prefix in fp32, total in double */
static double prefix_score(highs_t *hs) {
  for (int k = 1; k < hs->M; ++k) hs->val[k] = hs->val[k] + hs->val[k - 1];
  double acc = 0.0;
  for (int k = 0; k < hs->M; ++k) acc += hs->val[k];
  return acc;
}

int64_t orc_extract_output(const float *input, int h, int w, int n, double threshold,
                           int64_t *ret, float *scores) {
  highs_t hs;
  hs.M = threshold < 0.2 ? 8 : 4; /* :82-84 */
  int64_t written = 0;
  for (int64_t p = 0; p < (int64_t)h * w; ++p) {
    collect(input + p * n, n, threshold, &hs);
    if (!(hs.val[0] > 0)) continue; /* :120 / :136: untouched otherwise */
    sort_desc(&hs);
    ret[p] = (int64_t)hs.pos[0];
    scores[p] = (float)prefix_score(&hs);
    ++written;
  }
  return written;
}

int64_t orc_extract_output_marginalized(const float *input, int h, int w, int n,
                                        double threshold, double threshold_acc,
                                        int64_t *ret, int64_t *retgd) {
  highs_t hs;
  hs.M = threshold < 0.2 ? 8 : 4;
  int64_t written = 0;
  for (int64_t p = 0; p < (int64_t)h * w; ++p) retgd[p] = 0; /* THLongTensor_zero(retgd), :167 */
  for (int64_t p = 0; p < (int64_t)h * w; ++p) {
    collect(input + p * n, n, threshold, &hs);
    if (!(hs.val[0] > 0)) continue;
    sort_desc(&hs);
    ret[p] = (int64_t)hs.pos[0];
    if (prefix_score(&hs) >= threshold_acc) retgd[p] = 1; /* :227, :245 */
    ++written;
  }
  return written;
}
