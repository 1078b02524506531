"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/gcl_b200.h declares, and the host layer refuses CPU tensors (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "gcl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gcl_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from gcl_b200 import _cabi
    lib = ctypes.CDLL(_cabi.lib_path())
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gcl_b200.h but not exported"
    assert sorted(_cabi.exported_names()) == names, "ctypes prototypes out of sync with the header"


def test_version_and_error_string():
    from gcl_b200 import _cabi
    lib = _cabi.load()
    assert lib.gcl_version() == _cabi.ABI_VERSION
    # a bad-argument call must return an error code and set the thread-local message (no CUDA needed)
    rc = lib.gcl_spmm_f32(None, None, None, None, None, 1, 1, 1, 4, 4, 4, None, None, None, 0, None)
    assert rc == -1
    assert "null pointer" in _cabi.last_error()
    assert lib.gcl_csr_workspace_bytes(10, 5) > 0


def test_no_cpu_fallback():
    from gcl_b200.nn import GATConv, GCNConv, LayerNorm, SimpleConv
    x = torch.randn(4, 8)
    ei = torch.tensor([[0, 1, 2], [1, 2, 3]])
    for layer in (GCNConv(8, 8), GATConv(8, 8, heads=1, concat=False), SimpleConv(aggr="mean")):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            layer(x, ei)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        LayerNorm(8, mode="node")(x)


def test_state_dict_keys_match_pyg():
    from gcl_b200.nn import GATConv, GCNConv, LayerNorm
    assert sorted(GCNConv(3, 5).state_dict()) == ["bias", "lin.weight"]
    g = GATConv(3, 5, heads=2, concat=False)
    assert sorted(g.state_dict()) == ["att_dst", "att_src", "bias", "lin.weight"]
    assert g.lin.weight.shape == (10, 3) and g.att_src.shape == (1, 2, 5) and g.bias.shape == (5,)
    assert sorted(LayerNorm(4, mode="node").state_dict()) == ["bias", "weight"]
    # PyG <= 2.4 naming is accepted on load
    sd = {"lin_src.weight": torch.ones(10, 3), "lin_dst.weight": torch.ones(10, 3),
          "att_src": torch.zeros(1, 2, 5), "att_dst": torch.zeros(1, 2, 5), "bias": torch.zeros(5)}
    g.load_state_dict(sd)
    assert torch.equal(g.lin.weight, torch.ones(10, 3))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "graphcast-lite_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{fn} imports oracle"
                assert "pyg_shim" not in src and "trimesh_shim" not in src, f"{fn} references the oracle shims"
