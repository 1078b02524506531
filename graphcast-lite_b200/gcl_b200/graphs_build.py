"""Graph construction for the encode-process-decode model: the icosphere mesh hierarchy, the three
edge lists (grid->mesh, mesh<->mesh, mesh->grid) and the 6 static node features.

Mirrors what WeatherPrediction.__init__ builds (/root/reference/src/models.py:507-570) through
  src/mesh/create_mesh.py:75-223,323-352   hierarchy, level merge, edges from faces
  src/mesh/grid_mesh_connectivity.py:53-184 radius query, containing triangle
  src/create_graphs.py:96-295, src/utils.py:64-245,426-437  edge_index tensors, static features
but vectorised (no per-face Python loops) and with the two searches on the device
(gcl_radius_query_*, gcl_closest_face: brute force in fp64).  Results: edge sets identical to the
reference (mesh and mesh->grid also in the same order; grid->mesh ordered by (sender, receiver) where
the reference has cKDTree's order inside a sender), vertex coordinates bit-identical float32.
"""
from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _cabi


# ------------------------------------------------------------------------------------------ mesh (host)
def _icosahedron() -> Tuple[np.ndarray, np.ndarray]:
    from scipy.spatial.transform import Rotation
    phi = (1 + np.sqrt(5)) / 2
    v = np.array([(c1, c2, 0.0) if k == 0 else (0.0, c1, c2) if k == 1 else (c2, 0.0, c1)
                  for c1 in (1.0, -1.0) for c2 in (phi, -phi) for k in range(3)], dtype=np.float32)
    v /= np.linalg.norm([1.0, phi])
    faces = np.array([(0, 1, 2), (0, 6, 1), (8, 0, 2), (8, 4, 0), (3, 8, 2), (3, 2, 7), (7, 2, 1), (0, 4, 6),
                      (4, 11, 6), (6, 11, 5), (1, 5, 7), (4, 10, 11), (4, 8, 10), (10, 8, 3), (10, 3, 9),
                      (11, 10, 9), (11, 9, 5), (5, 9, 7), (9, 3, 7), (1, 6, 5)], dtype=np.int32)
    tilt = (np.pi - 2 * np.arcsin(phi / np.sqrt(3))) / 2
    v = np.dot(v, Rotation.from_euler(seq="y", angles=tilt).as_matrix())
    return v.astype(np.float32), faces


def _split(verts: np.ndarray, faces: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """4-way split of every face; child vertices numbered in first-use order (face order, edges 12,23,31)."""
    nv = len(verts)
    f = faces.astype(np.int64)
    ends = np.stack([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=1).reshape(-1, 2)  # [3F, 2] in visit order
    key = np.minimum(ends[:, 0], ends[:, 1]) * nv + np.maximum(ends[:, 0], ends[:, 1])
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    creation = np.argsort(first, kind="stable")           # unique keys in order of first use
    new_id = np.empty(len(uniq), dtype=np.int64)
    new_id[creation] = nv + np.arange(len(uniq))
    mid = new_id[inv].reshape(-1, 3)                        # per face: m12, m23, m31
    pa, pb = ends[first[creation], 0], ends[first[creation], 1]
    child = (verts[pa] + verts[pb]) / np.float32(2.0)       # float32 midpoint, like ndarray.mean(0)
    # np.linalg.norm of a float32 3-vector is sqrt(x.dot(x)) with the BLAS sdot kernel; a batched
    # [1,3] @ [3,1] matmul goes through the same kernel, so the float32 bits agree (checked in tests).
    sq = np.matmul(child[:, None, :], child[:, :, None]).reshape(-1)
    child = child / np.sqrt(sq)[:, None]
    out_v = np.concatenate([verts, child.astype(verts.dtype)], axis=0)
    m12, m23, m31 = mid[:, 0], mid[:, 1], mid[:, 2]
    i1, i2, i3 = f[:, 0], f[:, 1], f[:, 2]
    out_f = np.stack([np.stack([i1, m12, m31], 1), np.stack([m12, i2, m23], 1),
                      np.stack([m31, m23, i3], 1), np.stack([m12, m23, m31], 1)], axis=1).reshape(-1, 3)
    return out_v, out_f.astype(np.int32)


def mesh_hierarchy(splits: int) -> List[Tuple[np.ndarray, np.ndarray]]:
    """[(vertices float32 [V,3], faces int32 [F,3])] for levels 0..splits."""
    v, f = _icosahedron()
    out = [(v, f)]
    for _ in range(splits):
        v, f = _split(v, f)
        out.append((v, f))
    return out


def merged_faces(hier, levels: Sequence[int]) -> np.ndarray:
    """Faces of the listed levels only, finest first (create_mesh.py:210-223)."""
    lv = sorted(levels, reverse=True)
    return np.concatenate([hier[l][1] for l in lv], axis=0)


def edges_from_faces(faces: np.ndarray, num_vertices: int) -> np.ndarray:
    """int64 [2, 2U]: unique undirected pairs in lexicographic order, each followed by its reversal."""
    f = faces.astype(np.int64)
    a = np.concatenate([f[:, 0], f[:, 1], f[:, 2]])
    b = np.concatenate([f[:, 1], f[:, 2], f[:, 0]])
    key = np.unique(np.minimum(a, b) * num_vertices + np.maximum(a, b))
    lo, hi = key // num_vertices, key % num_vertices
    out = np.empty((2, 2 * len(key)), dtype=np.int64)
    out[0, 0::2], out[1, 0::2] = lo, hi
    out[0, 1::2], out[1, 1::2] = hi, lo
    return out


def _grid_xyz(lat: np.ndarray, lon: np.ndarray) -> np.ndarray:
    """Unit vectors of the lat-major flattened regular grid, in the dtype of lat/lon."""
    phi, theta = np.meshgrid(np.deg2rad(lon), np.deg2rad(90 - lat))
    return np.stack([np.cos(phi) * np.sin(theta), np.sin(phi) * np.sin(theta), np.cos(theta)], axis=-1).reshape(-1, 3)


def _static_features(lat32: np.ndarray, lon32: np.ndarray) -> np.ndarray:
    phi, theta = np.deg2rad(lon32), np.deg2rad(90 - lat32)
    return np.stack([np.cos(phi) * np.sin(theta), np.sin(phi) * np.sin(theta), np.cos(theta), np.cos(theta),
                     np.cos(phi), np.sin(phi)], axis=-1).astype(np.float32)


def _mesh_lat_lon(verts: np.ndarray):
    phi = np.arctan2(verts[:, 1], verts[:, 0])
    with np.errstate(invalid="ignore"):
        theta = np.arccos(verts[:, 2])
    return (90 - np.rad2deg(theta)).astype(np.float32), np.mod(np.rad2deg(phi), 360).astype(np.float32)


def _max_edge(verts: np.ndarray, faces: np.ndarray):
    s = np.concatenate([faces[:, 0], faces[:, 1], faces[:, 2]])
    r = np.concatenate([faces[:, 1], faces[:, 2], faces[:, 0]])
    return np.linalg.norm(verts[s] - verts[r], axis=-1).max()


def morton_order(xyz: np.ndarray) -> np.ndarray:
    """Permutation that sorts points of the unit sphere along a 3-d Morton (Z-order) curve, int32 [n]."""
    q = np.clip(((np.asarray(xyz, dtype=np.float64) + 1.0) * 0.5 * 1023.0).astype(np.int64), 0, 1023)

    def spread(v):
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    code = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
    return np.argsort(code, kind="stable").astype(np.int32)


def mesh_edge_features(mesh_lat32: np.ndarray, mesh_lon32: np.ndarray, edge_index: np.ndarray) -> np.ndarray:
    """[E, 4] float32 features of the mesh edges for the InteractionNet processor (create_graphs.py:37-91 through
    utils.py:248-418): the sender's position minus the receiver's in the receiver's local frame (rotated so that the
    receiver sits at longitude 0, latitude 0), and its length, both divided by the longest edge."""
    from scipy.spatial.transform import Rotation
    snd, rcv = edge_index[0], edge_index[1]
    phi, theta = np.deg2rad(mesh_lon32), np.deg2rad(90 - mesh_lat32)              # utils.py:212-218 (float32 in, float32 out)
    pos = np.stack([np.cos(phi) * np.sin(theta), np.sin(phi) * np.sin(theta), np.cos(theta)], axis=-1)
    rot = Rotation.from_euler("zy", np.stack([-phi, -theta + np.pi / 2], axis=1)).as_matrix()     # utils.py:389-401
    r_e = rot[rcv]
    rel = np.einsum("bji,bi->bj", r_e, pos[snd]) - np.einsum("bji,bi->bj", r_e, pos[rcv])
    dist = np.linalg.norm(rel, axis=-1, keepdims=True)
    mx = dist.max()
    if mx > 0:
        dist, rel = dist / mx, rel / mx
    return np.concatenate([dist, rel], axis=-1).astype(np.float32)


# ------------------------------------------------------------------------------------------ searches (device)
def _stream():
    return torch.cuda.current_stream().cuda_stream


def radius_query(grid_xyz: np.ndarray, mesh_xyz: np.ndarray, radius: float, mesh_offset: int, device) -> torch.Tensor:
    lib = _cabi.load()
    g = torch.as_tensor(np.ascontiguousarray(grid_xyz, dtype=np.float64), device=device)
    m = torch.as_tensor(np.ascontiguousarray(mesh_xyz, dtype=np.float32), device=device)
    G, M = g.shape[0], m.shape[0]
    offsets = torch.empty(G + 1, dtype=torch.int32, device=device)
    nb = lib.gcl_radius_query_workspace_bytes(G)
    ws = torch.empty(nb, dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        _cabi.check(lib.gcl_radius_query_count(g.data_ptr(), m.data_ptr(), G, M, float(radius), offsets.data_ptr(),
                                               ws.data_ptr(), nb, _stream()), "gcl_radius_query_count")
        E = int(offsets[-1].item())
        ei = torch.empty((2, max(E, 1)), dtype=torch.int64, device=device)
        _cabi.check(lib.gcl_radius_query_fill(g.data_ptr(), m.data_ptr(), G, M, float(radius), offsets.data_ptr(),
                                              ei.data_ptr(), ei.stride(0), int(mesh_offset), _stream()),
                    "gcl_radius_query_fill")
    return ei[:, :E] if E < ei.shape[1] else ei


def closest_face(grid_xyz: np.ndarray, mesh_xyz: np.ndarray, faces: np.ndarray, prefilter_radius: float,
                 device) -> torch.Tensor:
    lib = _cabi.load()
    g = torch.as_tensor(np.ascontiguousarray(grid_xyz, dtype=np.float64), device=device)
    m = torch.as_tensor(np.ascontiguousarray(mesh_xyz, dtype=np.float32), device=device)
    f = torch.as_tensor(np.ascontiguousarray(faces, dtype=np.int32), device=device)
    out = torch.empty(g.shape[0], dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        _cabi.check(lib.gcl_closest_face(g.data_ptr(), m.data_ptr(), f.data_ptr(), g.shape[0], m.shape[0],
                                         f.shape[0], float(prefilter_radius), out.data_ptr(), _stream()),
                    "gcl_closest_face")
    return out


# ------------------------------------------------------------------------------------------ everything
class ModelGraphs:
    """Edge lists (int64 [2, E], device) and static features (fp32, device) of one model."""

    def __init__(self, nlat: int, nlon: int, mesh_levels: Sequence[int], radius_factor: float, device="cuda:0"):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("gcl_b200.ModelGraphs: graph construction runs on a CUDA device; no CPU fallback")
        lat64 = np.linspace(-90, 90, nlat)                       # main.py:45-56
        lon64 = np.linspace(0, 360, nlon, endpoint=False)
        lat32, lon32 = lat64.astype(np.float32), lon64.astype(np.float32)   # models.py:666-667
        hier = mesh_hierarchy(max(mesh_levels))
        verts, faces = hier[-1]
        self.num_grid, self.num_mesh = nlat * nlon, len(verts)
        self.mesh_vertices, self.finest_faces = verts, faces
        G = self.num_grid
        max_edge = _max_edge(verts, faces)
        # grid->mesh: float32 axes (models.py:529-530), radius = max edge * factor (create_graphs.py:131-134)
        self.encoding_graph = radius_query(_grid_xyz(lat32, lon32), verts, float(max_edge * radius_factor), G, device)
        # mesh<->mesh over the listed levels only (create_graphs.py:225-229)
        self.processing_graph = torch.as_tensor(edges_from_faces(merged_faces(hier, mesh_levels), len(verts)),
                                                device=device)
        # mesh->grid: the ORIGINAL float64 axes (models.py:564-565), 3 vertices of the closest finest face
        fid = closest_face(_grid_xyz(lat64, lon64), verts, faces, 3.0 * float(max_edge), device).long()
        tri = torch.as_tensor(faces.astype(np.int64), device=device)[fid].reshape(-1) + G
        gidx = torch.arange(G, device=device).repeat_interleave(3)
        self.decoding_graph = torch.stack([tri, gidx]).contiguous()
        glon, glat = np.meshgrid(lon32, lat32)
        self.init_grid_features = torch.as_tensor(_static_features(glat.reshape(-1), glon.reshape(-1)), device=device)
        mlat, mlon = _mesh_lat_lon(verts)
        self.init_mesh_features = torch.as_tensor(_static_features(mlat, mlon), device=device)
        self._mesh_lat_lon = (mlat, mlon)
        self._edge_features = None
        # scheduling hints for the tiled aggregation kernels: mesh rows along a space-filling curve (grid rows are
        # already lat-major), so that a tile's rows share most of their neighbours
        from . import graph as _graph
        order = morton_order(verts)
        _graph.ORDER_HINTS[self.num_mesh] = order
        _graph.ORDER_HINTS[G + self.num_mesh] = np.concatenate([np.arange(G, dtype=np.int32), G + order]).astype(np.int32)

    @property
    def processing_edge_features(self) -> torch.Tensor:
        """[E_mesh, 4] float32 on the device (built on first use: only the InteractionNet processor reads them)."""
        if self._edge_features is None:
            ei = self.processing_graph.cpu().numpy()
            self._edge_features = torch.as_tensor(mesh_edge_features(*self._mesh_lat_lon, ei),
                                                  device=self.processing_graph.device)
        return self._edge_features
