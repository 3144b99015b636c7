/* A C99 host of the drop-in boundary: includes include/depthmatch.h as plain C, links
 * libdepthmatch.so, and walks the error path every reference-side binding relies on (status
 * code + dm_last_error, no longjmp, no abort).  With a GPU (argv[1] == "gpu") it also runs one
 * tiny matching call on host buffers.  Built and run by tests/test_abi.py. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "depthmatch.h"

int main(int argc, char **argv) {
  dm_ctx *ctx = NULL;
  int rc;
  if (dm_version() != 100) return 10;
  rc = dm_create(0, &ctx);
  if (argc < 2 || strcmp(argv[1], "gpu") != 0) {
    /* no device: creation must fail loudly, with a message, and leave ctx NULL */
    if (rc == 0 || ctx != NULL) return 11;
    if (strstr(dm_last_error(), "no CPU fallback") == NULL) return 12;
    printf("no device: %s\n", dm_last_error());
    return 0;
  }
  if (rc != 0) {
    fprintf(stderr, "dm_create: %s\n", dm_last_error());
    return 20;
  }
  {
    enum { C = 3, H1 = 6, W1 = 8, MH = 3, MW = 5, H2 = H1 + MH - 1, W2 = W1 + MW - 1 };
    static float in1[C * H1 * W1], in2[C * H2 * W2], pmax[H1 * W1];
    static int64_t index[H1 * W1];
    dm_pair p;
    dm_extract_out out;
    int i, k, y, x;
    for (i = 0; i < C * H2 * W2; ++i) in2[i] = (float)((i * 2654435761u) % 1000) / 1000.0f;
    for (k = 0; k < C; ++k) /* frame 1 = frame 2 displaced by (dy, dx) = (2, 1) */
      for (y = 0; y < H1; ++y)
        for (x = 0; x < W1; ++x) in1[(k * H1 + y) * W1 + x] = in2[(k * H2 + y + 2) * W2 + x + 1];
    memset(&p, 0, sizeof p);
    memset(&out, 0, sizeof out);
    p.in1 = in1;
    p.in2 = in2;
    p.n_pairs = 1;
    p.channels = C;
    p.h1 = H1;
    p.w1 = W1;
    p.h2 = H2;
    p.w2 = W2;
    out.index = index;
    out.pmax = pmax;
    rc = dm_match_extract(ctx, &p, MH, MW, DM_FLAG_TIE_MIDDLE, 0.11, H1, W1, &out);
    if (rc != 0) {
      fprintf(stderr, "dm_match_extract: %s\n", dm_last_error());
      return 21;
    }
    for (i = 0; i < H1 * W1; ++i)
      if (index[i] != 2 * MW + 1 + 1) return 22;
    /* an argument error comes back as a status + message */
    p.h2 = H1;
    if (dm_match_extract(ctx, &p, MH, MW, 0, 0.11, H1, W1, &out) == 0) return 23;
    if (strstr(dm_last_error(), "smaller than") == NULL) return 24;
    printf("gpu: %d pixels matched, %lld kernel launches\n", H1 * W1, (long long)dm_launch_count(ctx));
  }
  dm_destroy(ctx);
  return 0;
}
