"""CPU-side checks of the boundary: the C-ABI library builds/loads and exports every symbol include/*.h declares,
fails loudly without a device, and the product package never touches the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "caretta_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(crt_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_every_declared_symbol():
    from caretta_b200 import build, engine
    lib_path = build.build()
    assert os.path.exists(lib_path)
    lib = ctypes.CDLL(lib_path)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/caretta_b200.h but not exported"
    assert sorted(engine.EXPORTS) == declared, "engine.EXPORTS must list exactly the declared C ABI"
    assert lib.crt_version() >= 100


def test_no_cpu_fallback_without_device():
    import torch
    from caretta_b200 import engine
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(engine.CrtError) as e:
        engine.Engine()
    assert "no CUDA device" in str(e.value)


def test_missing_library_is_loud(monkeypatch, tmp_path):
    from caretta_b200 import engine
    monkeypatch.setattr(engine, "_lib", None)
    monkeypatch.setattr(engine, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(engine.CrtError):
        engine.load_library()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "caretta_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle|libcaretta_oracle|ref_harness|caretta_oracle", flags=re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not pat.search(src), f"{f} references the oracle"
