// Per-point stencil setup shared by the direct (interp.cu) and tiled (tiles.cu) mesh kernels.
#pragma once
#include "common.cuh"
#include "stencil_weights.cuh"

namespace tpme {

// Fractional mesh coordinate, wrapped stencil origin and 1-D weights of one point.
//   u = r @ r2u                                   (mesh_interpolator.py:326)
//   even n: i0 = floor(u), x = u - (i0 + 1/2);  odd n: i0 = rint(u), x = u - i0   (:329-341)
//   first node index = (i0 + 1 - (n + 1) / 2) mod ns                              (:350-359)
// The modulo is done in floating point on the integer-valued base (exact), followed by a
// compare-and-fix; no integer division.
template <typename T>
struct MeshDims {
  int n[3];
  T inv_n[3];
};

template <typename T>
__device__ __forceinline__ int wrap_base(T base, int n, T inv_n) {
  const T q = floor_t(base * inv_n);
  int i = (int)(base - q * (T)n);
  if (i < 0) i += n;
  if (i >= n) i -= n;
  return i;
}

__device__ __forceinline__ int wrap_add(int i, int n) {   // i in [0, 2n) typically; loop for n < nodes
  while (i >= n) i -= n;
  return i;
}

template <typename T, int METHOD, int N, bool DERIV>
__device__ __forceinline__ void point_stencil(const T* __restrict__ pos, const Mat3<T>& r2u,
                                              const MeshDims<T>& dims, int (&first)[3], T (&w)[3][N],
                                              T (&dw)[3][N]) {
  const T r0 = pos[0], r1 = pos[1], r2 = pos[2];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const T u = r0 * r2u.m[a] + r1 * r2u.m[3 + a] + r2 * r2u.m[6 + a];
    T base, x;
    if (N % 2 == 0) {
      base = floor_t(u);
      x = u - (base + T(0.5));
    } else {
      base = rint_t(u);
      x = u - base;
    }
    first[a] = wrap_base<T>(base + T(1 - (N + 1) / 2), dims.n[a], dims.inv_n[a]);
    Stencil<METHOD, N>::template eval<T, DERIV>(x, w[a], dw[a]);
  }
}


// Optional fused epilogues of the gathers (all pointers may be null):
//   values[i,c] = values[i,c] + scale * val - add_coef[i,c] * self_half - background * dc[c]
//     (the O(N) self / background corrections and the 1/(2V) factor of calculators/pme.py:117-143)
//   grad_positions[i,:] = vjp_scale * (vjp + sum_c coef2[i,c] * dvalues2[i,c,:])
template <typename T>
struct PointEpilogue {
  const T* add_coef;
  const T* dc;
  T scale, self_half, background;
  const T* coef2;
  const T* dvalues2;
  T vjp_scale;
  int enabled;
};

template <typename T>
static MeshDims<T> make_dims(int nx, int ny, int nz) {
  MeshDims<T> d;
  d.n[0] = nx; d.n[1] = ny; d.n[2] = nz;
  d.inv_n[0] = T(1) / T(nx); d.inv_n[1] = T(1) / T(ny); d.inv_n[2] = T(1) / T(nz);
  return d;
}

}  // namespace tpme
