"""The C-ABI library loads and exports every symbol include/svr_render.h declares; the ctypes
mirrors agree with the C struct layouts (which static_assert the reference's, include/svr_types.h).
No compute calls: runs without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

from sunvolumerender_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions(header="svr_render.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:[A-Za-z_][\w\s\*]*?)\b([A-Za-z_]\w*)\s*\([^;{]*\)\s*;", src, flags=re.M)
    return sorted(set(names))


def test_header_and_binding_agree():
    declared = _declared_functions()
    bound = sorted(n for n, _, _ in L.SIGNATURES)
    assert declared == bound
    assert _declared_functions("svr_volume_io.h") == sorted(n for n, _, _ in L.SIGNATURES_IO)
    assert _declared_functions("svr_tf_io.h") == sorted(n for n, _, _ in L.SIGNATURES_TF)
    assert _declared_functions("svr_env_io.h") == sorted(n for n, _, _ in L.SIGNATURES_ENV)
    assert _declared_functions("svr_canvas.h") == sorted(n for n, _, _ in L.SIGNATURES_CANVAS)


def test_reference_boundary_symbols_present():
    # pathtracer.h:17-24, raycasting.h:8
    seven = ["render_pathtracer", "setup_volume", "setup_transferfunction", "setup_camera", "setup_env_lights", "setup_area_lights", "render_raycasting"]
    assert [n for n, _, _ in L.SIGNATURES[:7]] == seven


def test_library_exports_every_declared_symbol():
    assert os.path.exists(L.LIB_PATH), "build with `make lib`"
    out = subprocess.check_output(["nm", "-D", "--defined-only", L.LIB_PATH], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [n for n in _declared_functions() + _declared_functions("svr_volume_io.h") + _declared_functions("svr_tf_io.h") + _declared_functions("svr_env_io.h") + _declared_functions("svr_canvas.h") if n not in exported]
    assert not missing, missing


def test_library_loads_and_binds():
    lib = L.load()
    assert lib.svr_version() == 100
    assert lib.svr_get_option(L.OPT_PT_MODE) == 2
    assert lib.svr_get_option(L.OPT_MACROCELL_SIZE) == 0  # automatic
    assert lib.svr_get_option(L.OPT_PT_KERNEL) == 2
    assert lib.svr_get_option(L.OPT_SETUP_SYNC) == 1          # setup_* synchronise like the reference's unless told otherwise
    assert lib.svr_get_option(L.OPT_PT_QUEUE_MIN_DEPTH) == 8
    # option validation is host logic
    assert lib.svr_set_option(L.OPT_PT_MODE, 7) != 0
    assert b"SVR_OPT_PT_MODE" in lib.svr_last_error()
    assert lib.svr_set_option(L.OPT_MACROCELL_SIZE, 12) != 0
    assert lib.svr_set_option(999, 0) != 0


def test_library_is_sm100a_only():
    out = subprocess.check_output(["cuobjdump", "--list-elf", L.LIB_PATH], text=True)
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_struct_layouts_match_c_headers(tmp_path):
    # compile a probe against include/svr_types.h and compare with the ctypes mirrors
    probe = tmp_path / "probe.c"
    probe.write_text(
        '#include <stdio.h>\n#include "svr_types.h"\n'
        "int main(void){printf(\"%zu %zu %zu %zu %zu %zu %zu %zu\\n\", sizeof(svr_volume), sizeof(svr_transfer_function),"
        " sizeof(svr_camera), sizeof(svr_disk), sizeof(svr_area_light), sizeof(svr_env_light), sizeof(svr_render_params), sizeof(svr_bbox));return 0;}\n"
    )
    exe = tmp_path / "probe"
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(probe), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    mirrors = [L.Volume, L.TransferFunction, L.Camera, L.Disk, L.AreaLight, L.EnvLight, L.RenderParams, L.BBox]
    assert sizes == [C.sizeof(m) for m in mirrors]
    assert sizes == [112, 16, 76, 28, 44, 32, 16, 36]  # SURVEY.md section 8b probe of the reference structs


def test_option_enum_and_binding_agree():
    """Every svr_option of include/svr_render.h has the same value in the ctypes binding (OPT_* = SVR_OPT_* without the prefix)."""
    src = open(os.path.join(ROOT, "include", "svr_render.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    body = re.search(r"enum svr_option\s*\{(.*?)\}", src, flags=re.S).group(1)
    entries = re.findall(r"SVR_(OPT_\w+)\s*=\s*(\d+)", body)
    assert len(entries) >= 24
    for name, value in entries:
        assert getattr(L, name) == int(value), name
    assert re.search(r"SVR_OPT_COUNT_", body)
    # defaults a drop-in host gets without calling svr_set_option
    lib = L.load()
    for name, default in (("OPT_PT_PROFILE", 0), ("OPT_PT_LIGHT_CULL", 1), ("OPT_ENV_NEE", 0), ("OPT_PT_BLOCK_SPLIT", 0), ("OPT_PT_PIXEL_CACHE", 1), ("OPT_FUSED_UPLOAD", 1), ("OPT_PT_LOOKAHEAD", 32),
                          ("OPT_PT_WARP_PIXELS", 0), ("OPT_ENV_ENABLED", 0), ("OPT_SHADOW_ESTIMATOR", 0)):
        assert lib.svr_get_option(getattr(L, name)) == default, name
