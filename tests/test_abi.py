"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol the header declares,
and refuses to run without a GPU (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from dsp_stuff_b200 import build, engine

    build.build()
    return engine.load_library()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "dspb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dspb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    from dsp_stuff_b200 import engine

    declared = header_symbols()
    assert declared, "no declarations found in include/dspb200.h"
    for name in declared:
        assert hasattr(lib, name), f"libdspb200.so does not export {name}"
    assert sorted(engine.ABI_SYMBOLS) == declared


def test_abi_version(lib):
    assert lib.dspb_abi_version() == 1


def test_config_struct_matches_header():
    from dsp_stuff_b200 import engine

    assert ctypes.sizeof(engine.Config) == 6 * 4 + 8 + 2 * 4
    assert [f[0] for f in engine.Config._fields_] == ["channels", "block", "sample_rate", "ref_block", "ring_granule",
                                                     "device", "max_samples", "fir_fft_log2", "fir_mode"]


def test_engine_fails_loudly_without_gpu(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from dsp_stuff_b200.engine import Engine, EngineError

    with pytest.raises(EngineError) as ei:
        Engine(4)
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "dsp_stuff_b200")
    for dp, _, files in os.walk(pkg):
        if "build" in dp.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h")):
                text = open(os.path.join(dp, f), errors="replace").read()
                assert "liboracle" not in text and "import oracle" not in text and "from oracle" not in text, f
