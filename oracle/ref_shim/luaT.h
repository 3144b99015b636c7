/*
 * luaT.h -- minimal stand-in for Torch7's luaT/lua C API, just enough to compile
 * the reference's own native sources (version2/extract_output.cpp, x2yxMulti2.c)
 * UNMODIFIED, from where they lie under /root/reference, into oracle/_ref/.
 * TEST INFRASTRUCTURE ONLY.  Written from scratch for this repo; not Torch code.
 *
 * A lua_State here is a fixed array of argument slots (1-based like the Lua
 * stack) plus a small value stack on top for lua_pushnumber/lua_gettable.
 */
#ifndef DM_REF_SHIM_LUAT_H
#define DM_REF_SHIM_LUAT_H

#ifdef __cplusplus
#define LUA_EXTERNC extern "C"
extern "C" {
#else
#define LUA_EXTERNC extern
#endif

enum { SHIM_NIL = 0, SHIM_NUMBER, SHIM_FLOAT_TENSOR, SHIM_LONG_TENSOR, SHIM_TABLE };

typedef struct shim_slot {
  int kind;
  double number;
  void *ptr;             /* tensor */
  const double *table;   /* SHIM_TABLE: 1-based array part, table[0] unused */
  int table_n;
} shim_slot;

typedef struct lua_State {
  shim_slot slot[32];
  int top; /* number of live slots */
} lua_State;

typedef int (*lua_CFunction)(lua_State *L);
typedef struct luaL_reg {
  const char *name;
  lua_CFunction func;
} luaL_reg;

const void *luaT_checktypename2id(lua_State *L, const char *tname);
void *luaT_checkudata(lua_State *L, int idx, const void *id);
double lua_tonumber(lua_State *L, int idx);
long lua_tointeger(lua_State *L, int idx);
int luaL_getn(lua_State *L, int idx);
void lua_pushnumber(lua_State *L, double v);
void lua_gettable(lua_State *L, int idx);
void luaL_openlib(lua_State *L, const char *name, const luaL_reg *l, int nup);

/* helpers for the wrappers (not part of the Lua API) */
void shim_reset(lua_State *L);
void shim_push_number(lua_State *L, double v);
void shim_push_float_tensor(lua_State *L, void *t);
void shim_push_long_tensor(lua_State *L, void *t);
void shim_push_table(lua_State *L, const double *one_based, int n);

#ifdef __cplusplus
}
#endif
#endif
