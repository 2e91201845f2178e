"""CPU-only checks of the drop-in boundary: libsgan.so loads, exports every symbol include/sgan.h declares with the
argument counts the ctypes binding assumes, struct layouts agree, and calls fail loudly (no CPU fallback)."""
import ctypes as C
import importlib
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
abi = importlib.import_module("scrabble-gan_b200._abi")


def _header_protos():
    h = open(os.path.join(ROOT, "include", "sgan.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    out = {}
    for name, args in re.findall(r"\b(?:int|size_t|long long|const char\*)\s+(sg_\w+)\s*\(([^;{]*?)\)\s*;", h, flags=re.S):
        args = args.strip()
        out[name] = 0 if args in ("void", "") else len(args.split(","))
    return out


def test_library_is_built_and_loads():
    assert os.path.exists(abi.LIB_PATH), "run `python -c 'import __graft_entry__ as g; g.build()'` first"
    lib = abi.load()
    assert lib.sg_version() >= 100


def test_every_declared_symbol_is_exported_and_bound():
    protos = _header_protos()
    assert len(protos) >= 50
    lib = abi.load()
    for name, nargs in protos.items():
        assert hasattr(lib, name), "libsgan.so does not export " + name
        assert name in abi._PROTOS, "ctypes binding missing for " + name
        assert len(abi._PROTOS[name][1]) == nargs, "argument count mismatch for {}: header {} vs binding {}".format(
            name, nargs, len(abi._PROTOS[name][1]))
    assert set(abi._PROTOS) <= set(protos), "binding declares symbols the header does not: {}".format(set(abi._PROTOS) - set(protos))


def test_conv_desc_layout_matches_c():
    assert abi.load().sg_sizeof_conv_desc() == C.sizeof(abi.ConvDesc)


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    runtime = importlib.import_module("scrabble-gan_b200.runtime")
    with pytest.raises(abi.SganError):
        runtime.Runtime(device=0)
    handle = C.c_void_p()
    rc = abi.load().sg_ctx_create(0, None, C.byref(handle))
    assert rc != abi.SG_OK and len(abi.last_error()) > 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "scrabble-gan_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(base, f)).read()
                assert "sgan_oracle" not in src and "import oracle" not in src and "from oracle" not in src, f
