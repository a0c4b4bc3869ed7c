"""The C-ABI library loads without a GPU and exports every symbol include/kfb.h declares; compute
entry points refuse to run without a device (there is no CPU path)."""

import ctypes
import os
import re

import pytest

from kronfluence_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "kfb.h"), encoding="utf-8").read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kfb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = engine.load_library()
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in kfb.h but not exported by libkfb.so"
    # and the ctypes binding types every one of them
    assert set(names) == set(engine.SIGNATURES), set(names) ^ set(engine.SIGNATURES)
    assert lib.kfb_version() == 100


def test_struct_layouts_match_header():
    assert ctypes.sizeof(engine.KfbLayer) == 18 * 4
    assert ctypes.sizeof(engine.KfbSplit) == 8 * 8
    # kfb_epilogue: int32 kind (+pad), ptr, 2x int64, kfb_split, ptr, int64, 3x int32, float, ptr, int64, 2x int32,
    # int64, int32 (+pad), int64
    assert ctypes.sizeof(engine.KfbEpilogue) == 8 + 8 + 16 + 64 + 8 + 8 + 16 + 8 + 8 + 8 + 8 + 8 + 8
    # ... and the library, compiled from include/kfb.h, agrees
    sizes = [ctypes.c_int() for _ in range(3)]
    engine.load_library().kfb_struct_sizes(*[ctypes.byref(x) for x in sizes])
    assert [x.value for x in sizes] == [ctypes.sizeof(engine.KfbLayer), ctypes.sizeof(engine.KfbSplit),
                                        ctypes.sizeof(engine.KfbEpilogue)]


def test_no_cpu_path():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(engine.KfbError, match="no CPU path"):
        engine.require_device()
    lib = engine.load_library()
    # pure host-side queries still work without a device
    layer = engine.KfbLayer(kind=0, d_in=64, d_out=32, has_bias=1)
    assert lib.kfb_pairwise_workspace_bytes(ctypes.byref(layer), 16, 1) > 0
    assert lib.kfb_eigh_workspace_bytes(128) > 2 * 128 * 128 * 8
    assert lib.kfb_eigh_jacobi_max_dim() == 512


def test_product_never_imports_the_oracle_or_the_tests():
    """The oracle is test infrastructure: nothing under kronfluence_b200/ or examples/ may import `oracle`, `tests` or the
    reference package, statically (AST of every module) or at import time (sys.modules after importing the package)."""
    import ast
    import glob
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    forbidden = {"oracle", "tests", "kronfluence"}
    for path in glob.glob(os.path.join(root, "kronfluence_b200", "**", "*.py"), recursive=True) + glob.glob(
            os.path.join(root, "examples", "*.py")):
        with open(path, encoding="utf-8") as handle:
            tree = ast.parse(handle.read())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [alias.name for alias in node.names]
            elif isinstance(node, ast.ImportFrom) and node.level == 0 and node.module:
                names = [node.module]
            assert not {name.split(".")[0] for name in names} & forbidden, (path, names)
    code = ("import sys; import kronfluence_b200, kronfluence_b200.analyzer, kronfluence_b200.ops, kronfluence_b200.engine, "
            "kronfluence_b200.module, kronfluence_b200.factor.config, kronfluence_b200.utils.common; "
            "bad = [m for m in sys.modules if m.split('.')[0] in ('oracle', 'tests', 'kronfluence')]; "
            "assert not bad, bad")
    subprocess.run([sys.executable, "-c", code], check=True, cwd=root, timeout=300)


def test_integration_md_binding_stub_matches_the_library():
    """The ctypes stub INTEGRATION.md tells a kronfluence maintainer to add (struct kfb_layer / kfb_split field lists,
    the error helper) is executed as written against the built library: struct sizes and field names must be the ones of
    include/kfb.h, and every `kfb_*` entry point the mapping table names must be exported."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "INTEGRATION.md"), encoding="utf-8") as handle:
        text = handle.read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = next(block for block in blocks if "class Layer(ctypes.Structure)" in block)
    library = engine.load_library()
    stub = stub.replace('ctypes.CDLL("libkfb.so")', "_the_library")
    namespace = {"_the_library": library}
    exec(compile(stub, "INTEGRATION.md", "exec"), namespace)  # pylint: disable=exec-used
    sizes = [ctypes.c_int() for _ in range(3)]
    library.kfb_struct_sizes(*[ctypes.byref(x) for x in sizes])
    assert ctypes.sizeof(namespace["Layer"]) == sizes[0].value == ctypes.sizeof(engine.KfbLayer)
    assert ctypes.sizeof(namespace["Split"]) == sizes[1].value == ctypes.sizeof(engine.KfbSplit)
    assert [name for name, _ in namespace["Layer"]._fields_] == [name for name, _ in engine.KfbLayer._fields_]
    assert [name for name, _ in namespace["Split"]._fields_] == [name for name, _ in engine.KfbSplit._fields_]
    for symbol in sorted(set(re.findall(r"\bkfb_[a-z0-9_]+\b", text))):
        if symbol in ("kfb_layer", "kfb_split", "kfb_status", "kfb_comm_"):
            continue  # struct / enum names and the deliberately absent communicator wrappers
        assert hasattr(library, symbol), symbol
