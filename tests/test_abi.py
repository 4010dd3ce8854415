"""The C-ABI library builds, loads and exports what include/mapdamage_b200.h declares."""
import ctypes as C
import re
from pathlib import Path

import pytest

from mapdamage_b200 import _native, build

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    build.build_library()
    return _native.load()


def declared_symbols():
    text = (ROOT / "include" / "mapdamage_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mdg_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(lib):
    names = declared_symbols()
    assert len(names) >= 20
    assert sorted(_native.SYMBOLS) == names
    for name in names:
        assert getattr(lib, name) is not None


def test_abi_version(lib):
    assert lib.mdg_abi_version() == _native.ABI_VERSION


def test_no_silent_cpu_fallback(lib):
    """Without a GPU the engine must refuse to run, not fall back."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mapdamage_b200.engine import DamageEngine

    with pytest.raises(_native.NativeError) as info:
        DamageEngine()
    assert info.value.code == _native.ERR_NO_DEVICE
    assert "no CPU path" in info.value.message


def test_create_rejects_bad_config(lib):
    ctx = C.c_void_p()
    cfg = _native.Config(0, 0, 10, 0, 1, 1024, 2, 0, 1024, 4096, 1 << 16)
    assert lib.mdg_create(C.byref(ctx), C.byref(cfg)) == _native.ERR_ARGUMENT
    assert b"invalid" in lib.mdg_last_error(None)


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under mapdamage_b200/ may reference it."""
    for path in (ROOT / "mapdamage_b200").rglob("*"):
        if path.suffix in (".py", ".cu", ".cuh", ".h"):
            text = path.read_text()
            assert "import oracle" not in text and "mdg_oracle" not in text and "pysam_shim" not in text, path
