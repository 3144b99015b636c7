"""The C-ABI library loads without a GPU and exports every symbol include/depthmatch.h
declares; the host-side index helpers agree with the oracle.  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "depthmatch.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(dm):
    lib = dm.load()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "libdepthmatch.so does not export %s" % n
    from depthmatch import _lib
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"


def test_struct_layout_matches_header(dm):
    from depthmatch import _lib
    assert C.sizeof(_lib.dm_pair) == 2 * 8 + 6 * 4 + 6 * 8
    assert C.sizeof(_lib.dm_extract_out) == 9 * 8


def test_version_and_error_string(dm):
    lib = dm.load()
    assert lib.dm_version() == 100
    assert isinstance(lib.dm_last_error(), bytes)


def test_fails_loudly_without_a_device(dm):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(dm.DepthMatchError) as e:
        dm.Context(0)
    assert "no CPU fallback" in str(e.value)
    # and the public API does not quietly compute on the CPU either
    with pytest.raises(dm.DepthMatchError):
        dm.match_extract(np.zeros((1, 4, 4), np.float32), np.zeros((1, 6, 6), np.float32), 3, 3)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "depth-estimation_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".lua")):
                txt = open(os.path.join(d, f)).read()
                assert "oracle_lib" not in txt and "libdm_oracle" not in txt and "orc_" not in txt, f


@pytest.mark.parametrize("ratios", [[1, 2], [1, 2, 4]])
def test_python_index_helpers_match_oracle(dm, oracle, ratios):
    g = dm.Geometry(maxh=8, maxw=8, ratios=ratios, multiscale=True)
    L = dm.multiscaleLength(g)
    assert L == oracle.multiscale_length(8, 8, ratios)
    for i in range(1, L + 1):
        y, x = dm.x2yxMultiNumber(g, i)
        rc, oy, ox = oracle.x2yx_multi_number(8, 8, ratios, i)
        assert rc == 0 and (y, x) == (oy, ox)
        assert dm.yx2xMulti(g, y, x) == i == oracle.yx2x_multi(8, 8, ratios, y, x)
    assert dm.getMiddleIndex(g) == 28


def test_single_scale_index_helpers(dm):
    g = dm.Geometry(maxh=17, maxw=33)
    assert dm.getMiddleIndex(g) == (9 - 1) * 33 + 17
    assert dm.x2yx(g, dm.yx2x(g, 5, 7)) == (5, 7)
    y, x = dm.x2yx(g, np.array([[1, 33, 34, 17 * 33]]))
    np.testing.assert_array_equal(y, [[1, 1, 2, 17]])
    np.testing.assert_array_equal(x, [[1, 33, 1, 33]])
    assert dm.centered2onebased(g, 0, 0) == (9, 17)


def test_prepare_input_is_a_view_with_the_reference_offsets(dm):
    g = dm.Geometry(maxh=5, maxw=8)
    f1 = np.arange(2 * 20 * 30, dtype=np.float32).reshape(2, 20, 30)
    a, b = dm.prepareInput(g, f1, f1)
    assert a.shape == (2, 16, 23) and b is f1
    assert np.shares_memory(a, f1) and a[0, 0, 0] == f1[0, 2, 3]  # ceil(5/2)-1, ceil(8/2)-1


def _build_c_host(tmp_path):
    import subprocess
    exe = str(tmp_path / "c_abi_host")
    libdir = os.path.join(ROOT, "depth-estimation_b200", "csrc")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi_host.c"), "-o", exe, "-L", libdir, "-ldepthmatch",
                           "-Wl,-rpath," + libdir])
    return exe


def test_header_is_plain_c_and_a_c_host_sees_the_error_path(dm, tmp_path):
    """include/depthmatch.h compiles as pedantic C99, a C program links the library, and without
    a device dm_create fails with a message instead of falling back or aborting."""
    import subprocess
    import torch
    dm.load()
    exe = _build_c_host(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present (the gpu variant runs in test_c_host_runs_the_hot_path)")
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "no CPU fallback" in out.stdout


@pytest.mark.gpu
def test_c_host_runs_the_hot_path(dm, tmp_path):
    import subprocess
    dm.load()
    out = subprocess.run([_build_c_host(tmp_path), "gpu"], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "48 pixels matched" in out.stdout


def _prototypes(text):
    """{name: normalised parameter list} of every `dm_*(...)` declaration in a C fragment."""
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(dm_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S):
        out[m.group(1)] = re.sub(r"\s+", " ", m.group(2)).strip()
    return out


def test_lua_ffi_cdef_matches_the_header():
    """The LuaJIT shim cannot run here; at least its ffi.cdef block must declare the header's
    functions with the same parameter lists (and nothing the header does not have)."""
    header = open(os.path.join(ROOT, "include", "depthmatch.h")).read()
    lua = open(os.path.join(ROOT, "depth-estimation_b200", "lua", "depthmatch_ffi.lua")).read()
    cdef = lua[lua.index("ffi.cdef[["):lua.index("]]", lua.index("ffi.cdef[["))]
    h, l = _prototypes(header), _prototypes(cdef)
    assert set(l) <= set(h), sorted(set(l) - set(h))
    for name, params in l.items():
        assert params == h[name], (name, params, h[name])
    # everything a Lua host needs for the path itself is declared
    for name in ("dm_create", "dm_destroy", "dm_last_error", "dm_match_volume", "dm_match_extract",
                 "dm_extract_output", "dm_match_extract_raw_ssd", "dm_x2yx_multi", "dm_cascade_add", "dm_multiscale_extract",
                 "dm_polar_remap", "dm_flow2depth", "dm_filter_create", "dm_filter_forward",
                 "dm_post_process_image", "dm_enlarge_mask", "dm_warp_homography"):
        assert name in l, name


# ---------------------------------------------------------------------------------------------
# The LuaJIT shims cannot be executed here (no Lua VM in the image).  tests/lua_check.py does what
# is possible without one; the same checker accepts all 55 Lua files of the reference.
# ---------------------------------------------------------------------------------------------
LUA_DIR = os.path.join(ROOT, "depth-estimation_b200", "lua")


@pytest.mark.parametrize("name", ["depthmatch_ffi.lua", "nn_depthmatch.lua", "depth_estimation_api_patch.lua"])
def test_lua_shims_nest_and_close(name):
    import lua_check
    toks = lua_check.check_structure(open(os.path.join(LUA_DIR, name)).read())
    assert len(toks) > 50
    with pytest.raises(lua_check.LuaSyntaxError):       # the checker does reject broken files
        lua_check.check_structure("function f(x) if x then return 1 end")


def test_lua_c_calls_pass_the_headers_argument_counts():
    import lua_check
    protos = lua_check.header_arg_counts(open(os.path.join(ROOT, "include", "depthmatch.h")).read())
    seen = set()
    for name in os.listdir(LUA_DIR):
        toks = lua_check.check_structure(open(os.path.join(LUA_DIR, name)).read())
        for fn, nargs, line in lua_check.calls(toks, "C.dm_"):
            assert fn[2:] in protos, "%s:%d calls %s, which the header does not declare" % (name, line, fn)
            assert nargs == protos[fn[2:]], "%s:%d: %s called with %d arguments, the header has %d" % (
                name, line, fn, nargs, protos[fn[2:]])
            seen.add(fn[2:])
    assert {"dm_match_extract", "dm_match_volume", "dm_extract_output", "dm_x2yx_multi", "dm_cascade_add"} <= seen


def test_lua_api_patch_sets_every_field_the_reference_reads():
    """depth_estimation_api.lua:171-182 reads poutput.full, poutput.y and poutput.full_confidences
    right after the lines the patch replaces (full_confidences was nil -> mask:cmul raised).
    Every field of processOutput's table must be assigned from the module's result, none to nil,
    and nn.DenseMatch must assign each of them on both extraction methods."""
    import lua_check
    patch = lua_check.check_structure(open(os.path.join(LUA_DIR, "depth_estimation_api_patch.lua")).read())
    fields = lua_check.table_fields(patch, "poutput")
    assert fields is not None
    for f in ("index", "y", "x", "confidences", "full", "full_confidences"):
        assert f in fields and fields[f] != "nil", (f, fields)
    mod = open(os.path.join(LUA_DIR, "nn_depthmatch.lua")).read()
    body = mod[mod.index("function DenseMatch:updateOutput"):mod.index("DenseMatch.updateGradInput")]
    for f in ("index", "y", "x", "confidences", "full", "full_confidences"):
        assert re.search(r"ret\.%s\b[^=\n]*=[^=]" % f, body) or re.search(r"\b%s\s*=\s*torch\." % f, body), f
    assert "conf_marginal" in body and "soft_yx" in body and "flow_full" in body
    # the reference lines the patch relies on are still what they were
    ref = os.path.join("/root/reference", "depth_estimation_api.lua")
    if os.path.exists(ref):
        lines = open(ref).read().split("\n")
        assert "processOutput(geometry, moutput, true, nil)" in lines[167]
        assert "poutput.full_confidences" in lines[181]
