// Graph construction searches on the device (one-time work per model, fp64, exact-order arithmetic):
//   gcl_radius_query_*  grid -> mesh edges: every mesh vertex within `radius` of a grid point
//       (/root/reference/src/mesh/grid_mesh_connectivity.py:53-104, scipy cKDTree.query_ball_point:
//        squared distance ((dx^2 + dy^2) + dz^2) in fp64 compared with radius^2)
//   gcl_closest_face    mesh -> grid edges: the mesh triangle closest to each grid point
//       (/root/reference/src/mesh/grid_mesh_connectivity.py:139-184, trimesh.proximity.closest_point:
//        Ericson closest-point-on-triangle with tol.zero = 1e-13, best two candidates, normal rule when
//        both squared distances exceed tol.merge = 1e-8 and differ by less than it)
// Both are brute force over tiles staged in shared memory (148 SMs make a spatial index pointless at
// these sizes: 131 072 x 40 962 distance tests take about a millisecond).  FP contraction is forbidden
// (explicit __dmul_rn / __dadd_rn) so the results do not depend on how nvcc fuses.
#include "common.cuh"
#include "scan.cuh"

namespace gcl {
namespace {

constexpr int kTile = 512;

__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double dot3(double ax, double ay, double az, double bx, double by, double bz) {
  return dadd(dadd(dmul(ax, bx), dmul(ay, by)), dmul(az, bz));
}

// FILL = false: counts[g] = number of hits.  FILL = true: write hits (ascending mesh id) at offsets[g].
template <bool FILL>
__global__ void __launch_bounds__(256)
    radius_query_kernel(const double* __restrict__ grid, const float* __restrict__ mesh, int64_t G, int64_t M,
                        double r2, int32_t* __restrict__ counts, const int32_t* __restrict__ offsets,
                        int64_t* __restrict__ ei_out, int64_t ei_stride, int64_t mesh_offset) {
  __shared__ double sv[kTile * 3];
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool act = g < G;
  double px = 0, py = 0, pz = 0;
  if (act) { px = grid[g * 3]; py = grid[g * 3 + 1]; pz = grid[g * 3 + 2]; }
  int32_t n = 0;
  int64_t w = (FILL && act) ? offsets[g] : 0;
  for (int64_t base = 0; base < M; base += kTile) {
    const int cnt = (int)min((int64_t)kTile, M - base);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * 3; i += blockDim.x) sv[i] = (double)mesh[base * 3 + i];
    __syncthreads();
    if (!act) continue;
    for (int j = 0; j < cnt; ++j) {
      const double dx = dsub(sv[3 * j], px), dy = dsub(sv[3 * j + 1], py), dz = dsub(sv[3 * j + 2], pz);
      const double d2 = dadd(dadd(dmul(dx, dx), dmul(dy, dy)), dmul(dz, dz));
      if (d2 <= r2) {
        if (FILL) {
          ei_out[w] = g;
          ei_out[ei_stride + w] = mesh_offset + base + j;
          ++w;
        } else {
          ++n;
        }
      }
    }
  }
  if (!FILL && act) counts[g] = n;
}

struct Cand {
  double d2, vx, vy, vz;  // squared distance and (query - closest point)
  int32_t f;
};

// Ericson closest point on triangle (a, b, c) to p, trimesh's vectorised formulation and tolerances.
__device__ __forceinline__ void closest_on_triangle(const double* t, double px, double py, double pz, double& qx,
                                                    double& qy, double& qz) {
  constexpr double Z = 1e-13;  // np.finfo(float64).resolution * 100
  const double ax = t[0], ay = t[1], az = t[2], bx = t[3], by = t[4], bz = t[5], cx = t[6], cy = t[7], cz = t[8];
  const double abx = dsub(bx, ax), aby = dsub(by, ay), abz = dsub(bz, az);
  const double acx = dsub(cx, ax), acy = dsub(cy, ay), acz = dsub(cz, az);
  const double apx = dsub(px, ax), apy = dsub(py, ay), apz = dsub(pz, az);
  const double d1 = dot3(abx, aby, abz, apx, apy, apz), d2 = dot3(acx, acy, acz, apx, apy, apz);
  if (d1 < Z && d2 < Z) { qx = ax; qy = ay; qz = az; return; }
  const double bpx = dsub(px, bx), bpy = dsub(py, by), bpz = dsub(pz, bz);
  const double d3 = dot3(abx, aby, abz, bpx, bpy, bpz), d4 = dot3(acx, acy, acz, bpx, bpy, bpz);
  if (d3 > -Z && d4 <= d3) { qx = bx; qy = by; qz = bz; return; }
  const double vc = dsub(dmul(d1, d4), dmul(d3, d2));
  if (vc < Z && d1 > -Z && d3 < Z) {
    const double v = ddiv(d1, dsub(d1, d3));
    qx = dadd(ax, dmul(v, abx)); qy = dadd(ay, dmul(v, aby)); qz = dadd(az, dmul(v, abz));
    return;
  }
  const double cpx = dsub(px, cx), cpy = dsub(py, cy), cpz = dsub(pz, cz);
  const double d5 = dot3(abx, aby, abz, cpx, cpy, cpz), d6 = dot3(acx, acy, acz, cpx, cpy, cpz);
  if (d6 > -Z && d5 <= d6) { qx = cx; qy = cy; qz = cz; return; }
  const double vb = dsub(dmul(d5, d2), dmul(d1, d6));
  if (vb < Z && d2 > -Z && d6 < Z) {
    const double w = ddiv(d2, dsub(d2, d6));
    qx = dadd(ax, dmul(w, acx)); qy = dadd(ay, dmul(w, acy)); qz = dadd(az, dmul(w, acz));
    return;
  }
  const double va = dsub(dmul(d3, d6), dmul(d5, d4));
  if (va < Z && dsub(d4, d3) > -Z && dsub(d5, d6) > -Z) {
    const double d43 = dsub(d4, d3);
    const double w = ddiv(d43, dadd(d43, dsub(d5, d6)));
    qx = dadd(bx, dmul(w, dsub(cx, bx))); qy = dadd(by, dmul(w, dsub(cy, by))); qz = dadd(bz, dmul(w, dsub(cz, bz)));
    return;
  }
  const double denom = ddiv(1.0, dadd(dadd(va, vb), vc));
  const double v = dmul(vb, denom), w = dmul(vc, denom);
  qx = dadd(dadd(ax, dmul(abx, v)), dmul(acx, w));
  qy = dadd(dadd(ay, dmul(aby, v)), dmul(acy, w));
  qz = dadd(dadd(az, dmul(abz, v)), dmul(acz, w));
}

__device__ __forceinline__ double normal_alignment(const double* t, const Cand& c) {
  const double ux = dsub(t[3], t[0]), uy = dsub(t[4], t[1]), uz = dsub(t[5], t[2]);
  const double vx = dsub(t[6], t[0]), vy = dsub(t[7], t[1]), vz = dsub(t[8], t[2]);
  const double nx = dsub(dmul(uy, vz), dmul(uz, vy)), ny = dsub(dmul(uz, vx), dmul(ux, vz)),
               nz = dsub(dmul(ux, vy), dmul(uy, vx));
  const double nn = sqrt(dadd(dadd(dmul(nx, nx), dmul(ny, ny)), dmul(nz, nz)));
  const double s = sqrt(c.d2);
  return dot3(ddiv(nx, nn), ddiv(ny, nn), ddiv(nz, nn), ddiv(c.vx, s), ddiv(c.vy, s), ddiv(c.vz, s));
}

__global__ void __launch_bounds__(128)
    closest_face_kernel(const double* __restrict__ grid, const float* __restrict__ verts,
                        const int32_t* __restrict__ faces, int64_t G, int64_t F, double prefilter_r2,
                        int32_t* __restrict__ face_out) {
  constexpr int FT = 256;
  __shared__ double st[FT * 9];
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool act = g < G;
  double px = 0, py = 0, pz = 0;
  if (act) { px = grid[g * 3]; py = grid[g * 3 + 1]; pz = grid[g * 3 + 2]; }
  Cand b1{INFINITY, 0, 0, 0, -1}, b2{INFINITY, 0, 0, 0, -1};
  for (int64_t base = 0; base < F; base += FT) {
    const int cnt = (int)min((int64_t)FT, F - base);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * 9; i += blockDim.x) {
      const int f = i / 9, r = i % 9;
      st[i] = (double)verts[(int64_t)faces[(base + f) * 3 + r / 3] * 3 + r % 3];
    }
    __syncthreads();
    if (!act) continue;
    for (int j = 0; j < cnt; ++j) {
      const double* t = st + 9 * j;
      const double ex = t[0] - px, ey = t[1] - py, ez = t[2] - pz;
      if (ex * ex + ey * ey + ez * ez > prefilter_r2) continue;  // cheap reject; never decides a winner
      double qx, qy, qz;
      closest_on_triangle(t, px, py, pz, qx, qy, qz);
      Cand c;
      c.vx = dsub(px, qx); c.vy = dsub(py, qy); c.vz = dsub(pz, qz);
      c.d2 = dot3(c.vx, c.vy, c.vz, c.vx, c.vy, c.vz);
      c.f = (int32_t)(base + j);
      if (c.d2 < b1.d2) { b2 = b1; b1 = c; }
      else if (c.d2 < b2.d2) { b2 = c; }
    }
  }
  if (!act) return;
  int32_t pick = b1.f;
  constexpr double MERGE = 1e-8;
  if (b2.f >= 0 && fabs(b2.d2 - b1.d2) < MERGE && fabs(b1.d2) > MERGE && fabs(b2.d2) > MERGE) {
    double t1[9], t2[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      t1[r] = (double)verts[(int64_t)faces[(int64_t)b1.f * 3 + r / 3] * 3 + r % 3];
      t2[r] = (double)verts[(int64_t)faces[(int64_t)b2.f * 3 + r / 3] * 3 + r % 3];
    }
    if (normal_alignment(t2, b2) > normal_alignment(t1, b1)) pick = b2.f;
  }
  face_out[g] = pick;
}

}  // namespace
}  // namespace gcl

using namespace gcl;

extern "C" size_t gcl_radius_query_workspace_bytes(int64_t n_grid) {
  return n_grid < 0 ? 0 : (size_t)(n_grid + 1) * sizeof(int32_t) + 256;
}

extern "C" int gcl_radius_query_count(const double* grid_xyz, const float* mesh_xyz, int64_t n_grid, int64_t n_mesh,
                                      double radius, int32_t* offsets, void* workspace, size_t workspace_bytes,
                                      void* stream) {
  GCL_CHECK_ARG(grid_xyz && mesh_xyz && offsets && workspace && n_grid > 0 && n_mesh > 0 && radius >= 0,
                "gcl_radius_query_count: bad argument");
  if (workspace_bytes < gcl_radius_query_workspace_bytes(n_grid)) {
    set_error("gcl_radius_query_count: workspace too small");
    return GCL_ERR_WORKSPACE;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int32_t* counts = static_cast<int32_t*>(workspace);
  radius_query_kernel<false><<<(unsigned)ceil_div(n_grid, 256), 256, 0, s>>>(
      grid_xyz, mesh_xyz, n_grid, n_mesh, radius * radius, counts, nullptr, nullptr, 0, 0);
  GCL_CHECK_LAUNCH("gcl_radius_query_count");
  scan_exclusive_kernel<<<1, kScanThreads, 0, s>>>(counts, offsets, n_grid, nullptr);
  GCL_CHECK_LAUNCH("gcl_radius_query_count(scan)");
  return GCL_OK;
}

extern "C" int gcl_radius_query_fill(const double* grid_xyz, const float* mesh_xyz, int64_t n_grid, int64_t n_mesh,
                                     double radius, const int32_t* offsets, int64_t* edge_index_out,
                                     int64_t num_edges, int64_t mesh_index_offset, void* stream) {
  GCL_CHECK_ARG(grid_xyz && mesh_xyz && offsets && edge_index_out && n_grid > 0 && n_mesh > 0 && num_edges >= 0,
                "gcl_radius_query_fill: bad argument");
  radius_query_kernel<true><<<(unsigned)ceil_div(n_grid, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      grid_xyz, mesh_xyz, n_grid, n_mesh, radius * radius, nullptr, offsets, edge_index_out, num_edges,
      mesh_index_offset);
  GCL_CHECK_LAUNCH("gcl_radius_query_fill");
  return GCL_OK;
}

extern "C" int gcl_closest_face(const double* grid_xyz, const float* mesh_xyz, const int32_t* faces, int64_t n_grid,
                                int64_t n_mesh, int64_t n_faces, double prefilter_radius, int32_t* face_out,
                                void* stream) {
  GCL_CHECK_ARG(grid_xyz && mesh_xyz && faces && face_out && n_grid > 0 && n_mesh > 0 && n_faces > 0,
                "gcl_closest_face: bad argument");
  closest_face_kernel<<<(unsigned)ceil_div(n_grid, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      grid_xyz, mesh_xyz, faces, n_grid, n_faces, prefilter_radius * prefilter_radius, face_out);
  GCL_CHECK_LAUNCH("gcl_closest_face");
  return GCL_OK;
}
