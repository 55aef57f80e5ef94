"""CPU-only: libsmrfb.so builds, loads, exports every symbol include/smrfb.h declares, and fails
loudly (no fallback) when there is no GPU.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    from smrf_b200 import build
    return build.build()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "smrfb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smrfb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(libpath):
    lib = ctypes.CDLL(libpath)
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), "libsmrfb.so does not export %s" % n


def test_binding_covers_header(libpath):
    from smrf_b200 import _lib
    assert sorted(_lib._SIGS) == declared_symbols()
    assert _lib.lib().smrfb_version() >= 100


def test_no_cpu_fallback_without_gpu(libpath):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from smrf_b200 import _lib
    from smrf_b200.spatial import IDW
    X, Y = np.meshgrid(np.arange(4.0), np.arange(3.0))
    with pytest.raises(_lib.SmrfbError):
        IDW(np.array([0.5, 2.5]), np.array([0.5, 1.5]), X, Y, mz=np.array([1.0, 2.0]), GridZ=np.ones((3, 4)))
    # the consumers likewise: the device or nothing
    from smrf_b200 import envphys_c
    from smrf_b200.snow import Snow
    with pytest.raises(_lib.SmrfbError):
        Snow.phase_and_density(np.array([-3.0, 1.0]), np.array([1.0, 2.0]), "marks2017")
    with pytest.raises(_lib.SmrfbError):
        envphys_c.cdewpt(np.array([600.0, 900.0]), np.zeros(2), 0.01, 1)


def test_snow_wrapper_argument_errors():
    """snow.py:64-69: an unknown model is a ValueError before anything touches the device."""
    from smrf_b200.snow import Snow
    assert sorted(Snow.MODELS) == ["marks2017", "piecewise_susong1999", "susong1999"]        # snow.py:29-33
    with pytest.raises(ValueError):
        Snow.phase_and_density(np.zeros(2), np.zeros(2), "nope")
    with pytest.raises(ValueError):
        Snow.phase_and_density(np.zeros(3), np.zeros(2), "susong1999")


def test_plain_c_client_compiles_and_links(libpath, tmp_path):
    """include/smrfb.h is plain C99 and examples/c_abi_demo.c -- what a cgo / JNI / Cython binding would wrap -- builds against the
    library with gcc alone; without a GPU it stops at the first call with the library's error text (no fallback)."""
    import subprocess
    import torch
    exe = str(tmp_path / "c_abi_demo")
    hdr = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-fsyntax-only", "-x", "c",
                          os.path.join(ROOT, "include", "smrfb.h")], capture_output=True, text=True)
    assert hdr.returncode == 0 and not hdr.stderr.strip(), hdr.stderr
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "c_abi_demo.c"), "-L", os.path.dirname(libpath), "-lsmrfb",
                        "-Wl,-rpath," + os.path.dirname(libpath), "-lm", "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0 and not r.stderr.strip(), r.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert run.returncode == 0 and "susong1999 at cell 0" in run.stdout, run.stdout + run.stderr
    else:
        assert run.returncode != 0 and "smrfb_device_count" in run.stderr, run.stdout + run.stderr


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "smrf_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
