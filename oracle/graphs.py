"""ORACLE / TEST INFRASTRUCTURE ONLY -- not product code, never imported by gcl_b200.

numpy restatement of the reference's graph construction (the part of the hot path that is the
reference's own code, so it IS pinned: oracle/make_golden.py runs the unmodified reference functions
in the build container and tests/test_oracle.py compares bit-for-bit, plus committed digests in
tests/golden/graphs.json for the GPU box where /root/reference is absent).

Follows:
  icosahedron / 4-way split hierarchy  /root/reference/src/mesh/create_mesh.py:75-207
  filter_mesh, get_edges_from_faces    /root/reference/src/mesh/create_mesh.py:210-223, 323-352
  radius query (G2M)                   /root/reference/src/mesh/grid_mesh_connectivity.py:53-110
  containing triangle (M2G)            /root/reference/src/mesh/grid_mesh_connectivity.py:139-184
  static node features                 /root/reference/src/utils.py:64-245, 426-437
  the three edge_index tensors         /root/reference/src/create_graphs.py:96-295
"""
import os
import sys

import numpy as np
from scipy.spatial import cKDTree
from scipy.spatial.transform import Rotation

_here = os.path.dirname(os.path.abspath(__file__))
if os.path.join(_here, "trimesh_shim") not in sys.path:
    sys.path.insert(0, os.path.join(_here, "trimesh_shim"))
import trimesh  # noqa: E402  (the oracle's restatement, see trimesh_shim/)


def icosahedron():
    """12 unit vertices (float32, rotated so a face pair is polar-aligned) and 20 CCW faces."""
    phi = (1 + np.sqrt(5)) / 2
    verts = []
    for c1 in (1.0, -1.0):
        for c2 in (phi, -phi):
            verts += [(c1, c2, 0.0), (0.0, c1, c2), (c2, 0.0, c1)]
    verts = np.array(verts, dtype=np.float32)
    verts /= np.linalg.norm([1.0, phi])
    faces = [(0, 1, 2), (0, 6, 1), (8, 0, 2), (8, 4, 0), (3, 8, 2), (3, 2, 7), (7, 2, 1),
             (0, 4, 6), (4, 11, 6), (6, 11, 5), (1, 5, 7), (4, 10, 11), (4, 8, 10), (10, 8, 3),
             (10, 3, 9), (11, 10, 9), (11, 9, 5), (5, 9, 7), (9, 3, 7), (1, 6, 5)]
    dihedral = 2 * np.arcsin(phi / np.sqrt(3))
    rot = Rotation.from_euler(seq="y", angles=(np.pi - dihedral) / 2).as_matrix()
    verts = np.dot(verts, rot)
    return verts.astype(np.float32), np.array(faces, dtype=np.int32)


def split_faces(verts: np.ndarray, faces: np.ndarray):
    """One 4-way split; child vertex = normalised midpoint, created in first-use order."""
    child = {}
    out_v = list(verts)

    def mid(i, j):
        key = (i, j) if i < j else (j, i)
        if key not in child:
            p = verts[[i, j]].mean(0)
            p /= np.linalg.norm(p)
            child[key] = len(out_v)
            out_v.append(p)
        return child[key]

    out_f = []
    for i1, i2, i3 in faces:
        i1, i2, i3 = int(i1), int(i2), int(i3)
        m12, m23, m31 = mid(i1, i2), mid(i2, i3), mid(i3, i1)
        out_f += [[i1, m12, m31], [m12, i2, m23], [m31, m23, i3], [m12, m23, m31]]
    return np.array(out_v), np.array(out_f, dtype=np.int32)


def mesh_hierarchy(splits: int):
    v, f = icosahedron()
    out = [(v, f)]
    for _ in range(splits):
        v, f = split_faces(v, f)
        out.append((v, f))
    return out


def merged_faces(hierarchy, levels):
    lv = sorted(levels, reverse=True)
    faces = hierarchy[lv[0]][1]
    for l in lv[1:]:
        faces = np.concatenate((faces, hierarchy[l][1]), axis=0)
    return hierarchy[lv[0]][0], faces


def mesh_edges(faces: np.ndarray) -> np.ndarray:
    """[2, 2*U] : unique undirected pairs (lexicographic), each followed by its reversal."""
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], axis=0).T
    e = np.unique(np.sort(e, axis=0), axis=1)
    out = np.zeros((2, 2 * e.shape[1]), dtype=e.dtype)
    out[:, 0::2] = e
    out[:, 1::2] = e[::-1]
    return out


def grid_xyz(lat: np.ndarray, lon: np.ndarray) -> np.ndarray:
    """[nlat*nlon, 3] float64 unit vectors, lat-major flattening."""
    phi, theta = np.meshgrid(np.deg2rad(lon), np.deg2rad(90 - lat))
    return np.stack([np.cos(phi) * np.sin(theta), np.sin(phi) * np.sin(theta), np.cos(theta)],
                    axis=-1).reshape(-1, 3)


def max_edge_length(verts, faces):
    s = np.concatenate([faces[:, 0], faces[:, 1], faces[:, 2]])
    r = np.concatenate([faces[:, 1], faces[:, 2], faces[:, 0]])
    return np.linalg.norm(verts[s] - verts[r], axis=-1).max()


def g2m_edges(lat, lon, verts, faces, radius_factor: float, num_grid: int) -> np.ndarray:
    radius = max_edge_length(verts, faces) * radius_factor
    hits = cKDTree(verts).query_ball_point(x=grid_xyz(lat, lon), r=radius)
    g = np.concatenate([np.repeat(i, len(h)) for i, h in enumerate(hits)]).astype(int)
    m = np.concatenate(hits).astype(int)
    return np.stack([g, m + num_grid], axis=0).astype(np.int64)


def m2g_edges(lat, lon, verts, faces, num_grid: int) -> np.ndarray:
    pts = grid_xyz(lat, lon)
    _, _, fid = trimesh.proximity.closest_point(trimesh.Trimesh(vertices=verts, faces=faces), pts)
    m = faces[fid].reshape(-1).astype(np.int64)
    g = np.repeat(np.arange(len(pts)), 3)
    return np.stack([m + num_grid, g], axis=0).astype(np.int64)


def mesh_lat_lon(verts):
    phi = np.arctan2(verts[:, 1], verts[:, 0])
    with np.errstate(invalid="ignore"):
        theta = np.arccos(verts[:, 2])
    lon = np.mod(np.rad2deg(phi), 360)
    lat = 90 - np.rad2deg(theta)
    return lat.astype(np.float32), lon.astype(np.float32)


def node_features(lat_f32: np.ndarray, lon_f32: np.ndarray) -> np.ndarray:
    """[n, 6] float32: x, y, z, cos(theta), cos(phi), sin(phi)."""
    phi, theta = np.deg2rad(lon_f32), np.deg2rad(90 - lat_f32)
    cols = [np.cos(phi) * np.sin(theta), np.sin(phi) * np.sin(theta), np.cos(theta),
            np.cos(theta), np.cos(phi), np.sin(phi)]
    return np.stack(cols, axis=-1).astype(np.float32)


def mesh_edge_features(mlat: np.ndarray, mlon: np.ndarray, edge_index: np.ndarray) -> np.ndarray:
    """create_graphs.py:37-91 (_compute_mesh_edge_features): relative position of the sender in the receiver's
    local coordinates (utils.py:248-343, rotation matrices :345-418) and its norm, normalised by the largest norm."""
    senders, receivers = edge_index[0], edge_index[1]
    phi = np.deg2rad(mlon)
    theta = np.deg2rad(90 - mlat)
    node_pos = np.stack((np.cos(phi) * np.sin(theta), np.sin(phi) * np.sin(theta), np.cos(theta)), axis=-1)
    azimuthal_rotation = -phi
    polar_rotation = -theta + np.pi / 2
    mats = Rotation.from_euler("zy", np.stack([azimuthal_rotation, polar_rotation], axis=1)).as_matrix()
    edge_mats = mats[receivers]
    recv_rot = np.einsum("bji,bi->bj", edge_mats, node_pos[receivers])
    send_rot = np.einsum("bji,bi->bj", edge_mats, node_pos[senders])
    relative_position = send_rot - recv_rot
    d = np.linalg.norm(relative_position, axis=-1, keepdims=True)
    max_dist = d.max()
    if max_dist > 0:
        d, relative_position = d / max_dist, relative_position / max_dist
    return np.concatenate([d, relative_position], axis=-1).astype(np.float32)


def build_graphs(nlat: int, nlon: int, mesh_levels, radius_factor: float):
    """Everything WeatherPrediction.__init__ builds (models.py:507-570) for a regular grid."""
    lat64 = np.linspace(-90, 90, nlat)                        # main.py:45-56 (float64)
    lon64 = np.linspace(0, 360, nlon, endpoint=False)
    lat, lon = lat64.astype(np.float32), lon64.astype(np.float32)   # models.py:666-667
    G = nlat * nlon
    hier = mesh_hierarchy(max(mesh_levels))
    fv, ff = hier[-1]
    mlat, mlon = mesh_lat_lon(fv)
    glon, glat = np.meshgrid(lon, lat)
    out = {
        "num_grid": G, "num_mesh": len(fv), "mesh_vertices": fv, "finest_faces": ff,
        "g2m": g2m_edges(lat, lon, fv, ff, radius_factor, G),
        "mesh": mesh_edges(merged_faces(hier, mesh_levels)[1]).astype(np.int64),
        # the decoder graph gets the ORIGINAL float64 axes (models.py:564-565), G2M the float32 ones
        "m2g": m2g_edges(lat64, lon64, fv, ff, G),
        "grid_feats": node_features(glat.reshape(-1).astype(np.float32),
                                    glon.reshape(-1).astype(np.float32)),
        "mesh_feats": node_features(mlat, mlon),
    }
    out["mesh_edge_feats"] = mesh_edge_features(mlat, mlon, out["mesh"])
    return out
