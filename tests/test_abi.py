"""The C-ABI library loads on CPU and exports every symbol include/dose_b200.h declares (no compute)."""
import os
import re

import pytest

from conftest import ROOT
from dose_prediction_b200 import _lib


def _declared():
    with open(os.path.join(ROOT, "include", "dose_b200.h")) as f:
        text = f.read()
    return sorted(set(re.findall(r"\b(dp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    handle = _lib.lib()
    names = _declared()
    assert names == _lib.exported_symbols()
    for n in names:
        assert getattr(handle, n) is not None
    assert handle.dp_abi_version() == 1
    assert isinstance(handle.dp_last_error(), bytes)


def test_ctypes_signatures_match_the_header():
    """argument count and kind (pointer / int / long long / float) of every ctypes signature == include/dose_b200.h."""
    import ctypes
    with open(os.path.join(ROOT, "include", "dose_b200.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    seen = 0
    for m in re.finditer(r"\bint\s+(dp_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        name, args = m.group(1), [a.strip() for a in m.group(2).split(",")]
        if name not in _lib._SIGNATURES:
            continue
        sig = _lib._SIGNATURES[name]
        assert len(args) == len(sig), name
        for a, t in zip(args, sig):
            if "*" in a or "cudaStream_t" in a:
                exp = ctypes.c_void_p
            elif a.startswith("long long"):
                exp = ctypes.c_longlong
            elif a.startswith("float"):
                exp = ctypes.c_float
            else:
                exp = ctypes.c_int
            assert exp is t, (name, a)
        seen += 1
    assert seen == len(_lib._SIGNATURES)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    """cuobjdump evidence that the hot kernels are tcgen05/TMA code, not recompiled mma.sync."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    import __graft_entry__
    __graft_entry__.build()
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass
    assert "HMMA." not in sass.replace("UTCHMMA", "")


def test_handle_create_without_a_gpu_reports_an_error():
    """dp_handle_create validates the device; on a box without GPUs it returns NULL and sets dp_last_error (no crash)."""
    import torch
    handle = _lib.lib()
    if torch.cuda.is_available():
        h = handle.dp_handle_create(0)
        assert h and handle.dp_handle_device(h) == 0
        handle.dp_handle_destroy(h)
    bad = handle.dp_handle_create(4096)
    assert not bad and b"dp_handle_create" in handle.dp_last_error()
    handle.dp_handle_destroy(None)
