"""The C-ABI library loads on a CPU-only box and exports every symbol include/esloco_b200.h declares."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from esloco_b200 import build
    build.build()
    from esloco_b200 import _native
    return _native.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "esloco_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(esl_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    from esloco_b200 import _native
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _native.SYMBOLS, f"{n} has no ctypes signature"
    assert sorted(_native.SYMBOLS) == names


def test_abi_version_and_loud_failure_without_gpu(lib):
    assert lib.esl_abi_version() == 2
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ctx = C.c_void_p()
    rc = lib.esl_create(C.byref(ctx), 0)
    assert rc == -3 and not ctx.value  # ESL_ERR_NO_DEVICE: no silent CPU path
    assert b"no CUDA device" in lib.esl_last_error(None)
    from esloco_b200.engine import Engine, EngineError
    with pytest.raises(EngineError):
        Engine(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "esloco_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "libesloco_oracle" not in text, f
