"""The C-ABI library loads and exports every symbol include/*.h declares; no compute calls (CPU suite)."""
import ctypes as C
import os
import re

import pytest

import tinyraytracing_b200 as trt
from tinyraytracing_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in ("trt.h", "trt_host.h"):
        txt = open(os.path.join(ROOT, "include", h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names |= set(re.findall(r"\b(trt_[a-z0-9_]+)\s*\(", txt))
    return names


def test_library_exports_every_declared_symbol():
    lib = trt.load_library()
    decl = declared_symbols()
    assert decl == set(api.EXPORTS), decl ^ set(api.EXPORTS)
    for s in decl:
        assert hasattr(lib, s), s
    assert lib.trt_version() == 201


def test_sm100a_sass_is_embedded():
    """The shared library must carry sm_100a device code (and nothing the driver would have to JIT)."""
    import shutil
    import subprocess

    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "-lelf", trt.library_path()], stdout=subprocess.PIPE, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_gpu(host_scenes):
    lib = trt.load_library()
    if lib.trt_device_count() > 0:
        pytest.skip("a B200 is present: the no-device error path cannot be exercised")
    with pytest.raises(trt.TrtError) as e:
        trt.DeviceScene(host_scenes["back"], 0)
    assert "(-2)" in str(e.value)  # TRT_ERR_NO_DEVICE: fails loudly, never computes on the CPU


def test_bad_arguments_return_errors_not_exits():
    lib = trt.load_library()
    h = C.c_void_p()
    assert lib.trt_host_scene_load(b"/nonexistent.xml", b"/nonexistent.obj", b"/nonexistent.mtl", b"/tmp", 8, C.byref(h)) == -1
    assert b"xml" in lib.trt_last_error()
    assert lib.trt_scene_create(None, 0, C.byref(h)) == -1
    assert lib.trt_trace_closest(None, None, 0, None, None, 0) == -1
    assert lib.trt_get_stats(None, None) == -1


def test_bench_clock_sampler_degrades_without_a_gpu():
    """bench.py polls NVML for the SM clock inside the timed region; without a device (this container) it must fall
    back and still return the keys of the bench contract instead of raising."""
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    c = bench.ClockSampler(0)
    c.start()
    c.mark()
    out = c.stop()
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(out)


def test_headers_compile_as_plain_c(tmp_path):
    """The boundary is a C ABI: both headers must be valid C99 on their own (no C++ types, no torch types), and every
    function they declare must be the one the library exports (a declaration/definition mismatch would still link)."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "use.c"
    src.write_text('#include "trt.h"\n#include "trt_host.h"\n'
                   "int main(void) { trt_layout_report r; trt_layout_view v; trt_stats s; trt_render_params p;\n"
                   "  return (int)(sizeof r + sizeof v + sizeof s + sizeof p) == 0; }\n")
    out = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only",
                          "-I", os.path.join(root, "include"), str(src)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


def test_inconsistent_descriptions_are_rejected_not_thrown(host_scenes):
    """Negative counts and NULL arrays under a positive count come back as TRT_ERR_INVALID from the host-only entry
    points too (no C++ exception crosses the C ABI; round-1 advisor finding)."""
    import copy

    lib = trt.load_library()
    good = host_scenes["back"].desc
    rep = api.LayoutReport()
    assert lib.trt_layout_check(C.byref(good), C.byref(rep)) == 0
    for field, value in (("n_textures", -1), ("n_lights", -3), ("n_light_tris", -1), ("n_tris", -5), ("n_nodes", -1),
                         ("n_materials", 0), ("width", 1)):
        d = api.SceneDesc.from_buffer_copy(good)
        setattr(d, field, value)
        assert lib.trt_layout_check(C.byref(d), C.byref(rep)) == -1, field
        assert lib.trt_last_error()
    d = api.SceneDesc.from_buffer_copy(good)
    d.v = None
    assert lib.trt_layout_check(C.byref(d), C.byref(rep)) == -1
    h, view = C.c_void_p(), api.LayoutView()
    assert lib.trt_layout_build(C.byref(d), C.byref(h), C.byref(view)) == -1
    h2 = C.c_void_p()
    rc = lib.trt_scene_create(C.byref(d), 0, C.byref(h2))
    assert rc in (-1, -2) and not h2.value
    # the new entry points refuse null / nonsense arguments the same way
    assert lib.trt_render_multi(None, 0, None, None, None) == -1
    assert lib.trt_shade(None, None, None, None, 0, None, None) == -1
    assert lib.trt_accum_save(None, None, 0, 1, 0, 0, b"/tmp/x") == -1
    assert lib.trt_accum_load(None, b"/tmp/x", None, None, None, None, None) == -1
    assert not lib.trt_accum_create(None)


def test_integration_md_shows_the_compiled_bridge():
    """INTEGRATION.md §2 prints integration/trt_bridge.{h,cpp} and main_gpu.sed: the text a maintainer reads is the text
    `make -C oracle bridge` compiles (and tests/test_gpu_bridge.py runs), not a sketch beside it."""
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for name in ("trt_bridge.h", "trt_bridge.cpp", "main_gpu.sed"):
        body = open(os.path.join(ROOT, "integration", name)).read()
        assert body in md, name
