"""CPU-side checks of the boundary: the shared library loads without a GPU and exports every symbol the header
declares; the ctypes binding covers all of them; product code never imports the oracle."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

from cutdet import _cabi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def handle():
    build.build_native()
    return ctypes.CDLL(build.LIB_PATH)


def test_header_declares_symbols():
    names = _cabi.declared_symbols()
    assert len(names) >= 28
    assert "cutdet_net_forward_frames" in names and "cutdet_glue_orphans" in names


def test_library_exports_every_declared_symbol(handle):
    for name in _cabi.declared_symbols():
        assert hasattr(handle, name), f"{name} declared in include/cutdet_b200.h but not exported"


def test_binding_covers_every_declared_symbol():
    assert sorted(_cabi._SIGNATURES) == _cabi.declared_symbols()


def test_every_declaration_cites_the_reference():
    text = open(_cabi.header_path()).read()
    for section in ("frameID/data.py", "frameID/net.py", "frameID/segmentation.py"):
        assert section in text


def test_no_gpu_calls_needed_for_host_helpers(handle):
    handle.cutdet_abi_version.restype = ctypes.c_int
    assert handle.cutdet_abi_version() == 2
    nw, nh = ctypes.c_int(), ctypes.c_int()
    assert handle.cutdet_target_size(1280, 720, 256, ctypes.byref(nw), ctypes.byref(nh)) == 0
    assert (nw.value, nh.value) == (256, 144)
    assert handle.cutdet_target_size(854, 480, 256, ctypes.byref(nw), ctypes.byref(nh)) == 0
    assert (nw.value, nh.value) == (256, 143)
    assert handle.cutdet_target_size(0, 480, 256, ctypes.byref(nw), ctypes.byref(nh)) == 1
    handle.cutdet_last_error.restype = ctypes.c_char_p
    assert b"cutdet_target_size" in handle.cutdet_last_error()


def test_sass_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", build.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cut-detection_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(dirpath, f)


def test_missing_library_fails_loudly(tmp_path):
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from cutdet import _cabi, build\n"
            "build.LIB_PATH = %r\n"
            "build.find_nvcc = lambda: None\n"
            "try:\n    _cabi.lib()\nexcept RuntimeError as e:\n    print('RAISED', e)\n") % (
        os.path.join(ROOT, "cut-detection_b200"), str(tmp_path / "nope.so"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert "RAISED" in out.stdout and "no fallback" in out.stdout


def test_stale_library_is_not_loaded(tmp_path):
    """A library older than the sources next to it is rebuilt (nvcc here) or refused (no nvcc), never run."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from cutdet import _cabi, build\n"
            "build.STAMP_PATH = %r\n"
            "build.find_nvcc = lambda: None\n"
            "try:\n    _cabi.lib()\nexcept RuntimeError as e:\n    print('RAISED', e)\n") % (
        os.path.join(ROOT, "cut-detection_b200"), str(tmp_path / "no_stamp.txt"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert "RAISED" in out.stdout and "older than the sources" in out.stdout


@pytest.mark.parametrize("h,w", [(720, 1280), (1080, 1920), (360, 640), (288, 512), (2160, 3840), (256, 455), (144, 256),
                                 (480, 854), (1080, 1440), (601, 333)])
def test_host_only_row_list_equals_the_oracle_taps(h, w):
    """cutdet_resize_rows is arithmetic on the host (the CLI calls it before the process has a CUDA context): the rows it
    names are the vertical taps with a non-zero weight of the oracle's coefficient table."""
    import numpy as np
    from cutdet import engine
    from oracle import preprocess as opre
    nw, nh = opre.target_size(w, h, 256)
    rows = engine.resize_rows(h, w, 256)
    if (nh, nw) == (h, w) or (2 * nh, 2 * nw) == (h, w):
        want = np.arange(h)
    else:
        i0, i1, w0, w1 = opre.linear_coeffs(h, nh, clamp_weights=False)
        want = np.unique(np.concatenate([i0, i1[w1 != 0]]))
    assert np.array_equal(rows, want)
