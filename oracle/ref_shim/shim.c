/*
 * shim.c -- implementation of the luaT / TH stand-ins (see luaT.h, TH/TH.h).
 * TEST INFRASTRUCTURE ONLY.
 */
#include "luaT.h"
#include "TH/TH.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static const char ID_FLOAT[] = "torch.FloatTensor";
static const char ID_LONG[] = "torch.LongTensor";

const void *luaT_checktypename2id(lua_State *L, const char *tname) {
  (void)L;
  if (strcmp(tname, ID_FLOAT) == 0) return ID_FLOAT;
  if (strcmp(tname, ID_LONG) == 0) return ID_LONG;
  fprintf(stderr, "ref shim: unknown tensor type %s\n", tname);
  abort();
}

void *luaT_checkudata(lua_State *L, int idx, const void *id) {
  const shim_slot *s = &L->slot[idx];
  const int want = id == (const void *)ID_FLOAT ? SHIM_FLOAT_TENSOR : SHIM_LONG_TENSOR;
  if (idx < 1 || idx > L->top || s->kind != want) {
    fprintf(stderr, "ref shim: bad argument #%d\n", idx);
    abort();
  }
  return s->ptr;
}

static shim_slot *at(lua_State *L, int idx) {
  if (idx < 0) idx = L->top + 1 + idx;
  return &L->slot[idx];
}

double lua_tonumber(lua_State *L, int idx) {
  const shim_slot *s = at(L, idx);
  return s->kind == SHIM_NUMBER ? s->number : 0.0; /* nil -> 0 like Lua */
}

long lua_tointeger(lua_State *L, int idx) { return (long)lua_tonumber(L, idx); }

int luaL_getn(lua_State *L, int idx) { return at(L, idx)->table_n; }

void lua_pushnumber(lua_State *L, double v) { shim_push_number(L, v); }

/* pops the key, pushes t[key]; only the integer array part 1..n exists */
void lua_gettable(lua_State *L, int idx) {
  const shim_slot *t = at(L, idx);
  shim_slot *key = &L->slot[L->top];
  const int k = (int)key->number;
  if (t->kind == SHIM_TABLE && key->kind == SHIM_NUMBER && (double)k == key->number && k >= 1 &&
      k <= t->table_n) {
    key->kind = SHIM_NUMBER;
    key->number = t->table[k];
  } else {
    key->kind = SHIM_NIL;
    key->number = 0.0;
  }
}

void luaL_openlib(lua_State *L, const char *name, const luaL_reg *l, int nup) {
  (void)L; (void)name; (void)l; (void)nup;
}

void shim_reset(lua_State *L) { memset(L, 0, sizeof(*L)); }

static shim_slot *push(lua_State *L) {
  if (L->top >= 31) abort();
  shim_slot *s = &L->slot[++L->top];
  memset(s, 0, sizeof(*s));
  return s;
}
void shim_push_number(lua_State *L, double v) { shim_slot *s = push(L); s->kind = SHIM_NUMBER; s->number = v; }
void shim_push_float_tensor(lua_State *L, void *t) { shim_slot *s = push(L); s->kind = SHIM_FLOAT_TENSOR; s->ptr = t; }
void shim_push_long_tensor(lua_State *L, void *t) { shim_slot *s = push(L); s->kind = SHIM_LONG_TENSOR; s->ptr = t; }
void shim_push_table(lua_State *L, const double *one_based, int n) {
  shim_slot *s = push(L); s->kind = SHIM_TABLE; s->table = one_based; s->table_n = n;
}

/* ---- TH ---- */
static void set_contig(long *size, long *stride, int nd, const long *sz) {
  long st = 1;
  for (int i = nd - 1; i >= 0; --i) { size[i] = sz[i]; stride[i] = st; st *= sz[i]; }
}
static long numel_f(const THFloatTensor *t) { long n = 1; for (int i = 0; i < t->nDimension; ++i) n *= t->size[i]; return n; }

THFloatTensor *shim_float_view(float *data, int nd, const long *size) {
  THFloatTensor *t = (THFloatTensor *)calloc(1, sizeof(*t));
  t->nDimension = nd; set_contig(t->size, t->stride, nd, size); t->data = data; t->owns = 0;
  return t;
}
THLongTensor *shim_long_view(long *data, int nd, const long *size) {
  THLongTensor *t = (THLongTensor *)calloc(1, sizeof(*t));
  t->nDimension = nd; set_contig(t->size, t->stride, nd, size); t->data = data; t->owns = 0;
  return t;
}
void shim_long_free(THLongTensor *t) { free(t); }

THFloatTensor *THFloatTensor_newContiguous(THFloatTensor *t) {
  /* views made by shim_float_view are always contiguous: return a new handle */
  THFloatTensor *c = (THFloatTensor *)malloc(sizeof(*c));
  *c = *t; c->owns = 0;
  return c;
}
THFloatTensor *THFloatTensor_newWithSize4d(long a, long b, long c, long d) {
  const long sz[4] = {a, b, c, d};
  THFloatTensor *t = (THFloatTensor *)calloc(1, sizeof(*t));
  t->nDimension = 4; set_contig(t->size, t->stride, 4, sz);
  t->data = (float *)malloc(sizeof(float) * (size_t)(a * b * c * d)); t->owns = 1;
  return t;
}
float *THFloatTensor_data(THFloatTensor *t) { return t->data; }
int THFloatTensor_isContiguous(THFloatTensor *t) { (void)t; return 1; }
void THFloatTensor_zero(THFloatTensor *t) { memset(t->data, 0, sizeof(float) * (size_t)numel_f(t)); }
void THFloatTensor_free(THFloatTensor *t) { if (t->owns) free(t->data); free(t); }
long *THLongTensor_data(THLongTensor *t) { return t->data; }
void THLongTensor_zero(THLongTensor *t) {
  long n = 1; for (int i = 0; i < t->nDimension; ++i) n *= t->size[i];
  memset(t->data, 0, sizeof(long) * (size_t)n);
}
