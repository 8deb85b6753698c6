"""The C-ABI shared library loads, exports every symbol include/*.h declares, and fails loudly
(never silently falls back) when no CUDA device is present. No compute calls here."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    names = set()
    for h in ("rt_api.h", "rt_host.h"):
        text = (ROOT / "include" / h).read_text()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(rth?_[a-z0-9_]+)\s*\(", text))
    return names


def test_library_exports_every_declared_symbol(lib, pkg):
    decl = declared_symbols()
    assert {"rt_scene_create", "rt_render", "rt_render_device", "rt_scene_destroy", "rt_last_error", "rth_example_scene"} <= decl
    out = subprocess.run(["nm", "-D", "--defined-only", str(pkg._lib.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
    missing = sorted(decl - exported)
    assert not missing, f"declared in include/*.h but not exported: {missing}"
    for name in decl:
        getattr(lib, name)


def test_library_is_built_for_sm_100a(pkg):
    out = subprocess.run(["cuobjdump", "--list-elf", str(pkg._lib.LIB_PATH)], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_struct_layouts_match_header(lib, pkg, tmp_path):
    """ctypes / numpy mirrors vs. what a C compiler makes of include/rt_api.h."""
    L = pkg._lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "rt_api.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(rt_scene_desc),sizeof(rt_config),sizeof(rt_stats),sizeof(rt_material),sizeof(rt_triangle),sizeof(rt_sphere),"
                   "sizeof(rt_light),offsetof(rt_config,seed),offsetof(rt_config,tile_rows));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), "-o", str(exe), str(src)], check=True)
    got = list(map(int, subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()))
    want = [C.sizeof(L.SceneDesc), C.sizeof(L.Config), C.sizeof(L.Stats), L.MATERIAL_DTYPE.itemsize, L.TRIANGLE_DTYPE.itemsize,
            L.SPHERE_DTYPE.itemsize, L.LIGHT_DTYPE.itemsize, L.Config.seed.offset, L.Config.tile_rows.offset]
    assert got == want
    assert lib.rt_api_version() == 1


def test_argument_validation_without_device(lib, pkg):
    L = pkg._lib
    h = C.c_void_p()
    assert lib.rt_scene_create(None, 0, C.byref(h)) == L.RT_ERR_INVALID_ARGUMENT
    assert b"null" in lib.rt_last_error()
    flat = pkg.example_scene("test_scene", "").flatten()
    flat.triangles["material"][0] = 999  # out-of-range material index
    d = flat.desc()
    assert lib.rt_scene_create(C.byref(d), 0, C.byref(h)) == L.RT_ERR_INVALID_ARGUMENT
    assert b"material" in lib.rt_last_error()
    assert lib.rt_render(None, None, None, None, None) == L.RT_ERR_INVALID_ARGUMENT


def test_no_cpu_fallback(lib, pkg):
    """Without a GPU the product refuses to render; it never routes through the oracle or any CPU path."""
    if lib.rt_device_count() > 0:
        pytest.skip("a CUDA device is present")
    flat = pkg.example_scene("test_scene", "").flatten()
    with pytest.raises(pkg.RtError) as e:
        pkg.DeviceScene(flat)
    assert e.value.code == pkg._lib.RT_ERR_NO_DEVICE
    with pytest.raises(pkg.RtError):
        pkg.RaytracerRenderer("").render(pkg.ImageBuffer(768, 640), flat)
    peak = C.c_double()
    assert lib.rt_measure_fp32_peak(0, C.byref(peak)) == pkg._lib.RT_ERR_NO_DEVICE
    with pytest.raises(pkg.RtError) as e:
        pkg.MultiDeviceScene(flat, n_devices=2)  # rt_multi_create: several GPUs in one process -- none here
    assert e.value.code == pkg._lib.RT_ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    pkg_dir = ROOT / "hslu_i"
    for p in list(pkg_dir.rglob("*.py")) + list(pkg_dir.rglob("*.cu")) + list(pkg_dir.rglob("*.cpp")) + list(pkg_dir.rglob("*.h*")) + list(pkg_dir.rglob("*.cuh")):
        text = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), p
        assert "liboracle" not in text and "rt_oracle" not in text, p
