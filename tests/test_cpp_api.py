"""The header-only C++ mirror (include/idto_b200.hpp) compiles against the C ABI on CPU and — on a GPU —
reproduces the reference's spinner test through the reference-shaped C++ API."""
import os
import subprocess

import pytest

from idto_b200.bake import load_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "test_api")


def _build():
    from idto_b200 import capi
    libdir = os.path.dirname(capi.LIB_PATH)
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "test_api.cc"), "-o", BIN, "-L", libdir, "-lidto_b200",
                    f"-Wl,-rpath,{libdir}", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-lcudart"],
                   check=True)
    return BIN


def test_cpp_api_compiles_and_links():
    assert os.path.exists(_build())


@pytest.mark.gpu
def test_cpp_api_spinner(tmp_path):
    exe = _build()
    txt = tmp_path / "spinner.txt"
    load_model("spinner").save_txt(txt)
    r = subprocess.run([exe, str(txt)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "C++ API test OK" in r.stdout


def test_bake_from_plant_against_the_drake_api_stand_in(tmp_path):
    """include/idto_b200_drake.hpp (BakeFromPlant: MultibodyPlant + SceneGraphInspector -> baked tables) compiled
    against tests/cpp/drake_stub: plants rebuilt from the committed baked models bake back to the same tables, with
    and without a welded body to merge.  No GPU needed (the tables never reach the device here)."""
    from idto_b200 import capi
    libdir = os.path.dirname(capi.LIB_PATH)
    exe = os.path.join(ROOT, "tests", "cpp", "test_bake_from_plant")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                    "-I", os.path.join(ROOT, "tests", "cpp", "drake_stub"),
                    os.path.join(ROOT, "tests", "cpp", "test_bake_from_plant.cc"), "-o", exe, "-L", libdir,
                    "-lidto_b200", f"-Wl,-rpath,{libdir}", "-L/usr/local/cuda/lib64",
                    "-Wl,-rpath,/usr/local/cuda/lib64", "-lcudart"], check=True)
    paths = []
    for name in ("spinner", "mini_cheetah", "allegro_hand", "spinner_capsule", "acrobot"):
        p = tmp_path / f"{name}.txt"
        load_model(name).save_txt(p)
        paths.append(str(p))
    r = subprocess.run([exe] + paths, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "BakeFromPlant test OK" in r.stdout
