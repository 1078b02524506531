"""The reference's own forward, layer by layer, on gcl_b200.nn -- what a graphcast-lite user gets by changing the
import at /root/reference/src/models.py:21 and nothing else.

`reference_forward(model, X)` walks the module tree of a gcl_b200.model.WeatherPrediction exactly the way the
reference's glue does (models.py:106-109 MLP loop, :413-434 GraphLayer loop, :776-874 input assembly with
zeros + cat, slices between the stages): torch.nn.Linear / nn.PReLU for the MLPs (cuBLAS + elementwise kernels, as
in the reference), one conv call per layer through the PyG signature ``conv(x, edge_index) -> [N, C]``, batch 1, no
fusion.  bench.py times it as `dropin_b1` next to the batched, fused mirror; the parity tests run the real
reference glue through the same layers (tests/test_model_gpu.py::test_import_swap_reference_glue_on_gcl_layers).
"""
import torch
import torch.nn as nn

from .nn import GATConv, GCNConv


def _graph_layer(gl, X, edge_index, attention_threshold=0.0, **kwargs):
    if gl.layer_type == "simple_conv":
        return gl.layers(x=X, edge_index=edge_index)
    if gl.layer_type == "sparse_gat":
        for layer in gl.layers:
            if isinstance(layer, GATConv):
                X, (edge_index, _) = layer.forward(X, edge_index, attention_threshold, **kwargs)
            else:
                X = layer(X)
        return X, edge_index
    for layer in gl.layers:
        X = layer(X, edge_index) if isinstance(layer, (GCNConv, GATConv)) else layer(X)
    return X


def _mlp(mlp, X):
    for layer in mlp.MLP:
        X = torch.nn.functional.linear(X, layer.weight, layer.bias) if isinstance(layer, nn.Linear) else layer(X)
    return X


def _model(m, X, edge_index, **kw):
    if m.mlp is not None:
        X = _mlp(m.mlp, X)
    return _graph_layer(m.graph_layer, X, edge_index, **kw)


def reference_forward(model, X: torch.Tensor, attention_threshold: float = 0.0, **kwargs) -> torch.Tensor:
    """X [1, G, T*F] -> [G, F_out], the reference's batch-1 forward (models.py:808-874)."""
    X = X.squeeze(0)
    G, M = model._num_grid_nodes, model._num_mesh_nodes
    grid = torch.cat([X, model.init_grid_features], dim=-1)                       # models.py:786-790
    mesh = torch.cat([torch.zeros((M, X.shape[-1]), device=X.device), model.init_mesh_features], dim=-1)
    enc = _model(model.encoder, torch.cat([grid, mesh], dim=0), model.encoding_graph)
    grid_lat, mesh_lat = enc[:G], enc[G:]
    if model.using_sparse_gat:
        proc, ei = _model(model.processor, mesh_lat, model.processing_graph, attention_threshold=attention_threshold,
                          **kwargs)
        model.processing_graph = ei
    else:
        proc = _model(model.processor, mesh_lat, model.processing_graph)
    dec = _model(model.decoder, torch.cat([grid_lat, proc], dim=0), model.decoding_graph)
    return dec[:G]
