"""CPU: the C-ABI library loads and exports every symbol include/rtda_lovasz.h declares;
argument validation that needs no GPU."""
import ctypes
import os
import re

import pytest

from rtda_semanticsegmentation_b200 import _lib


def _declared_functions():
    text = open(_lib.HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rtda_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.load()


def test_header_and_binding_agree():
    assert _declared_functions() == sorted(_lib.SIGNATURES)


def test_every_declared_symbol_is_exported(lib):
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(raw, name), name


def test_version_and_error_strings(lib):
    assert lib.rtda_version() == 2
    assert lib.rtda_error_string(0) == b"ok"
    for code in (-1, -2, -3, -4):
        assert lib.rtda_error_string(code).startswith(b"rtda:")


def test_profile_stage_table_matches_the_header(lib):
    # the Python reader sizes its arrays with PROFILE_STAGES; the library rejects a smaller count
    n = int(re.search(r"#define\s+RTDA_PROFILE_STAGES\s+(\d+)", open(_lib.HEADER).read()).group(1))
    assert _lib.PROFILE_STAGES == n
    names = [lib.rtda_profile_stage_name(i).decode() for i in range(n)]
    assert len(set(names)) == n and "?" not in names
    for k in ("k_prep", "k_split", "k_part", "k_plan", "k_local", "k_finalize", "k_bwd", "k_confusion"):
        assert k in names
    assert lib.rtda_profile_stage_name(n) == b"?"


def test_workspace_query_and_argument_errors(lib):
    n = ctypes.c_size_t(0)
    assert lib.rtda_lovasz_workspace_bytes(4, 19, 512 * 1024, 0, ctypes.byref(n)) == 0
    pairs = 2 * 19 * 4 * 512 * 1024 * 8  # the two (key, payload) ping-pong buffers dominate
    assert pairs < n.value < pairs * 1.2
    n2 = ctypes.c_size_t(0)
    assert lib.rtda_lovasz_workspace_bytes(4, 19, 512 * 1024, 1, ctypes.byref(n2)) == 0
    assert abs(n2.value - n.value) < 0.05 * n.value
    assert lib.rtda_lovasz_workspace_bytes(0, 19, 10, 0, ctypes.byref(n)) == -1       # RTDA_E_ARG
    assert lib.rtda_lovasz_workspace_bytes(1, 45, 10, 0, ctypes.byref(n)) == -2       # > RTDA_MAX_CLASSES
    assert lib.rtda_lovasz_workspace_bytes(1, 19, 1 << 30, 0, ctypes.byref(n)) == -2  # segment too long
    assert lib.rtda_lovasz_workspace_bytes(1, 19, 10, 0, None) == -1
    # null pointers are rejected before anything is launched
    assert lib.rtda_lovasz_fwd(None, None, 0, 1, 19, 16, 255, 1, 0, None, 0, 1, 1.0, None, None, None, None, None, None, None, None, None, None, None) == -1
    assert lib.rtda_lovasz_bwd_inplace(None, None, None, None, 1, 19, 16, None) == -1
    assert lib.rtda_lovasz_bwd(None, None, None, None, None, 1, 19, 16, 0, None, None) == -1
    assert lib.rtda_confusion_accumulate(None, 0, None, 0, 10, 19, 255, 1, None, None) == -1
    assert lib.rtda_confusion_argmax_accumulate(None, None, 0, 1, 19, 16, 255, 1, None, None) == -1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.NativeLibraryError):
        _lib.load()


def test_shapes_whose_finished_gradient_the_forward_writes(lib):
    """rtda_lovasz_final_grad_supported (host logic only): per image always; global batches when the maps are a multiple
    of 4096 pixels (a level-1 tile lies in one image) and C * B * HW fits the 30-bit pixel index of the sorted pairs."""
    f = lib.rtda_lovasz_final_grad_supported
    assert f(4, 19, 512 * 1024, 1) == 1 and f(16, 19, 720 * 1280, 1) == 1 and f(3, 19, 33 * 33, 1) == 1
    assert f(1, 19, 33 * 47, 0) == 1                      # a batch of one is one image per group
    assert f(4, 19, 512 * 1024, 0) == 1 and f(8, 19, 512 * 1024, 0) == 1 and f(2, 19, 64 * 64, 0) == 1
    assert f(4, 19, 96 * 160, 0) == 0                     # 15360 pixels: not a multiple of 4096
    assert f(4, 19, 33 * 33, 0) == 0
    assert f(64, 19, 1024 * 1024, 0) == 0                 # 19 * 64 Mi elements do not fit the index field
    assert f(0, 19, 4096, 0) == 0 and f(4, 0, 4096, 0) == 0
