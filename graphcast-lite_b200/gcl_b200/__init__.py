"""gcl_b200 -- B200-native (sm_100a) message passing for graphcast-lite.

Host side of the C ABI in include/gcl_b200.h:
  gcl_b200.nn      drop-in GCNConv / GATConv / SimpleConv / LayerNorm / summary  (the PyG import seam of
                   /root/reference/src/models.py:21,25)
  gcl_b200.utils   softmax / scatter / dense_to_sparse                            (models.py:24,220)
  gcl_b200.ops     the torch.autograd.Function custom ops
  gcl_b200.graph   device CSR builder + cache
There is no CPU fallback: CPU tensors or a missing libgcl_b200.so raise.
"""
from . import _cabi, graph, graphs_build, model, nn, ops, train, utils, workloads  # noqa: F401

__version__ = "0.1.0"


def library_path() -> str:
    return _cabi.lib_path()
