/*
 * ref_x2yx_tu.c -- wraps the reference's x2yxMulti2.c, which is the BODY of a
 * Lua C function as the Torch `inline` package expects it (`%%` is inline's
 * escape for `%`).  oracle/Makefile un-escapes it with sed into
 * oracle/_ref/x2yxMulti2_body.inc (a build artefact, git-ignored) and this file
 * supplies the `int f(lua_State *L) { ... }` frame that `inline.load` would.
 * TEST INFRASTRUCTURE ONLY.
 */
#include <stdint.h>
#include "luaT.h"
#include "TH/TH.h"

static int ref_body(lua_State *L) {
#include "x2yxMulti2_body.inc"
  return 0;
}

int ref_x2yx_multi2(const int64_t *xim, long h, long w, int maxh, int maxw, const double *ratios,
                    int nratios, int64_t *retx, int64_t *rety) {
  const long sz[2] = {h, w};
  double table[32];
  THLongTensor *tx = shim_long_view((long *)xim, 2, sz);
  THLongTensor *trx = shim_long_view((long *)retx, 2, sz);
  THLongTensor *try_ = shim_long_view((long *)rety, 2, sz);
  lua_State L;
  table[0] = 0;
  for (int i = 0; i < nratios && i < 31; ++i) table[i + 1] = ratios[i];
  shim_reset(&L);
  shim_push_long_tensor(&L, tx);
  shim_push_number(&L, maxh);
  shim_push_number(&L, maxw);
  shim_push_table(&L, table, nratios);
  shim_push_long_tensor(&L, trx);
  shim_push_long_tensor(&L, try_);
  const int rc = ref_body(&L);
  shim_long_free(tx);
  shim_long_free(trx);
  shim_long_free(try_);
  return rc;
}
