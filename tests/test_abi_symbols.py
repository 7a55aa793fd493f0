"""The C-ABI library builds, loads without a GPU and exports every symbol that
include/taco_b200.h declares; no compute is attempted here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from tacotron_multispeaker_b200 import _abi
    from tacotron_multispeaker_b200.build import build_library
    build_library()
    return _abi.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "taco_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(taco_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree(lib):
    from tacotron_multispeaker_b200 import _abi
    assert header_symbols() == sorted(_abi.SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in header_symbols():
        assert hasattr(lib, name), name


def test_header_compiles_as_plain_c(tmp_path):
    import subprocess
    c = tmp_path / "t.c"
    c.write_text('#include "taco_b200.h"\nint main(void){ taco_hparams hp; (void)hp; return TACO_OK; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(c),
                           "-o", str(tmp_path / "t.o")])


def test_create_without_gpu_fails_loudly(lib):
    import torch
    from tacotron_multispeaker_b200 import _abi
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    hp = _abi.TacoHParams(80, 1025, 5, 200, 256, 64, 7352, 60)
    h = C.c_void_p()
    assert lib.taco_create(C.byref(hp), 0, C.byref(h)) == _abi.TACO_ERR_UNSUPPORTED
    assert not h.value
    from tacotron_multispeaker_b200.engine import Engine
    from tacotron_multispeaker_b200.hparams import HParams
    with pytest.raises(RuntimeError):      # no CPU fallback in the product path
        Engine(HParams(), 0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "tacotron_multispeaker_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "taco_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f
