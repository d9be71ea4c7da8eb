"""CPU: the C-ABI library builds (nvcc cross-compiles without a GPU), loads, and exports every symbol that
include/occnerf_b200.h declares.  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "occnerf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(occnerf_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    from occnerf_b200 import _lib, build
    path = build.build()
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/occnerf_b200.h but not exported by {path}"
    assert sorted(_lib.EXPORTED) == names, "occnerf_b200/_lib.py binds a different symbol set than the header declares"


def test_abi_version_and_error_string():
    from occnerf_b200 import _lib
    lib = _lib.load()
    assert lib.occnerf_abi_version() == 3
    assert isinstance(lib.occnerf_last_error(), bytes)


def test_no_cpu_fallback():
    """A CPU tensor must be rejected loudly, not routed anywhere else."""
    import torch
    from occnerf_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.composite_forward(torch.zeros(1, 4, 5), torch.zeros(1, 4), torch.zeros(1, 4), torch.zeros(1, 8), torch.zeros(3))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "occnerf_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f"{f} reaches into oracle/"
