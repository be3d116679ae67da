// Mesh interpolation kernels: charge spreading (points -> mesh) and potential / force
// gathering (mesh -> points) for P3M (n = 1..5) and Lagrange (n = 3..7) stencils.
//
// Replaces the (n^3, N) index/weight tensors and index_put_/fancy-index ops of
// src/torchpme/lib/mesh_interpolator.py:303-457 with in-register stencils.
#include <cstdlib>

#include "common.cuh"
#include "stencil_point.cuh"
#include "../../include/torchpme_b200.h"

namespace tpme {

// do the `nodes` planes starting at `first` reach into the x slab [x0, x0 + nxl) of an nx-periodic axis?
__device__ __forceinline__ bool touches_slab(int first, int nodes, int nx, int x0, int nxl) {
  int ahead = first - x0;      // (first - x0) mod nx < nxl: the first plane lies inside
  if (ahead < 0) ahead += nx;
  int behind = x0 - first;     // (x0 - first) mod nx < nodes: the slab starts inside the stencil
  if (behind < 0) behind += nx;
  return ahead < nxl || behind < nodes;
}

template <typename T, int N>
__device__ __forceinline__ T pick(const T (&arr)[N], int k) {
  T out = arr[0];
#pragma unroll
  for (int i = 1; i < N; ++i) out = (i == k) ? arr[i] : out;
  return out;
}

// A group of G = next_pow2(N) lanes serves one point; lane c owns the z offset c of the stencil
// (the contiguous mesh axis, so a warp-wide access touches runs of N consecutive mesh values per
// (x, y) row) and loops over the N x N (a, b) rows.  Only G lanes repeat the per-point weight
// evaluation.
// marks a stencil plane that lies outside the local x slab (slab-decomposed meshes)
constexpr unsigned kOutside = 0xFFFFFFFFu;

template <int N> struct GroupSize {
  static constexpr int value = N <= 1 ? 1 : N <= 2 ? 2 : N <= 4 ? 4 : 8;
};

// ---------------------------------------------------------------------------------------
// spread
// ---------------------------------------------------------------------------------------
template <typename T, int METHOD, int N>
__global__ void __launch_bounds__(256)
spread_kernel(const T* __restrict__ positions, const T* __restrict__ weights, int64_t n_points,
              int n_channels, Mat3<T> r2u, MeshDims<T> dims, int x0, int nxl, T* __restrict__ mesh,
              const int* __restrict__ point_list, const int* __restrict__ list_count) {
  constexpr int G = GroupSize<N>::value;
  const int64_t total = point_list != nullptr ? (int64_t)*list_count : n_points;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; tid / G < total; tid += stride) {
  const int64_t entry = tid / G;
  const int c = (int)(tid - entry * G);
  if (c >= N) continue;
  const int64_t point = point_list != nullptr ? (int64_t)point_list[entry] : entry;
  const int nx = dims.n[0], ny = dims.n[1], nz = dims.n[2];

  int first[3];
  T w[3][N], dw[3][N];
  point_stencil<T, METHOD, N, false>(positions + 3 * point, r2u, dims, first, w, dw);
  const unsigned plane = (unsigned)ny * nz;
  const int64_t mesh_size = (int64_t)plane * nxl;   // the local slab holds x planes x0 .. x0 + nxl - 1
  unsigned xoff[N], yoff[N];
  {
    int ix = first[0], iy = first[1];
#pragma unroll
    for (int a = 0; a < N; ++a) {
      const unsigned lx = (unsigned)(ix - x0);
      xoff[a] = lx < (unsigned)nxl ? lx * plane : kOutside;
      yoff[a] = (unsigned)iy * nz;
      ix = (ix + 1 >= nx) ? wrap_add(ix + 1, nx) : ix + 1;
      iy = (iy + 1 >= ny) ? wrap_add(iy + 1, ny) : iy + 1;
    }
  }
  const unsigned iz = (unsigned)wrap_add(first[2] + c, nz);
  const T wz = pick<T, N>(w[2], c);
  for (int ch = 0; ch < n_channels; ++ch) {
    const T q = weights[point * n_channels + ch] * wz;
    T* dst = mesh + ch * mesh_size + iz;
#pragma unroll
    for (int a = 0; a < N; ++a) {
      if (xoff[a] == kOutside) continue;
      const T qa = q * w[0][a];
#pragma unroll
      for (int b = 0; b < N; ++b) red_add(dst + (xoff[a] + yoff[b]), qa * w[1][b]);
    }
  }
  }
}

// ---------------------------------------------------------------------------------------
// slab decomposition: list of the points whose stencil reaches into the x slab [x0, x0 + nxl)
// (block-aggregated append; the order inside the list is irrelevant to the results)
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
select_slab_points_kernel(const T* __restrict__ positions, int64_t n_points, Mat3<T> r2u,
                          MeshDims<T> dims, int nodes, int x0, int nxl, int* __restrict__ list,
                          int* __restrict__ count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool keep = false;
  if (i < n_points) {
    const T* pos = positions + 3 * i;
    const T u = pos[0] * r2u.m[0] + pos[1] * r2u.m[3] + pos[2] * r2u.m[6];
    const T base = (nodes % 2 == 0) ? floor_t(u) : rint_t(u);
    const int first = wrap_base<T>(base + T(1 - (nodes + 1) / 2), dims.n[0], dims.inv_n[0]);
    keep = touches_slab(first, nodes, dims.n[0], x0, nxl);
  }
  // one global atomic per thread block: warp counts -> block offsets -> list slots
  __shared__ int warp_count[8];
  __shared__ int block_base;
  const unsigned ballot = __ballot_sync(0xffffffffu, keep);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_count[warp] = __popc(ballot);
  __syncthreads();
  if (threadIdx.x == 0) {
    int total = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) {
      const int c = warp_count[k];
      warp_count[k] = total;
      total += c;
    }
    block_base = total > 0 ? atomicAdd(count, total) : 0;
  }
  __syncthreads();
  if (keep) list[block_base + warp_count[warp] + __popc(ballot & ((1u << lane) - 1))] = (int)i;
}

// ---------------------------------------------------------------------------------------
// gather.  MODE bit 0: values, bit 1: dvalues/dr, bit 2: vjp into grad_positions (+ grad_r2u)
// ---------------------------------------------------------------------------------------
// Sum four per-lane quantities over a group of G lanes (G = 1, 2, 4, 8): the first two steps
// split the quantities between the halves, the rest is a butterfly; every lane gets the totals.
template <typename T, int G>
__device__ __forceinline__ void group_reduce4(T& q0, T& q1, T& q2, T& q3, int lane) {
  if (G == 1) return;
  if (G == 2) {
    q0 += __shfl_xor_sync(0xffffffffu, q0, 1, G);
    q1 += __shfl_xor_sync(0xffffffffu, q1, 1, G);
    q2 += __shfl_xor_sync(0xffffffffu, q2, 1, G);
    q3 += __shfl_xor_sync(0xffffffffu, q3, 1, G);
    return;
  }
  const bool up = (lane & (G / 2)) != 0;
  const T sa = up ? q0 : q2, sb = up ? q1 : q3;
  T ka = (up ? q2 : q0) + __shfl_xor_sync(0xffffffffu, sa, G / 2, G);
  T kb = (up ? q3 : q1) + __shfl_xor_sync(0xffffffffu, sb, G / 2, G);
  const bool up2 = (lane & (G / 4)) != 0;
  T r = (up2 ? kb : ka) + __shfl_xor_sync(0xffffffffu, up2 ? ka : kb, G / 4, G);
#pragma unroll
  for (int off = G / 8; off > 0; off >>= 1) r += __shfl_xor_sync(0xffffffffu, r, off, G);
  // lane 0 of each quarter holds: quarter 0 -> q0, 1 -> q1, 2 -> q2, 3 -> q3
  q0 = __shfl_sync(0xffffffffu, r, 0, G);
  q1 = __shfl_sync(0xffffffffu, r, G / 4, G);
  q2 = __shfl_sync(0xffffffffu, r, G / 2, G);
  q3 = __shfl_sync(0xffffffffu, r, 3 * G / 4, G);
}

template <typename T, int METHOD, int N, int MODE>
__global__ void __launch_bounds__(256, (N <= 4 && sizeof(T) == 4) ? 4 : 2)
gather_kernel(const T* __restrict__ mesh, const T* __restrict__ positions,
              const T* __restrict__ coef, int64_t n_points, int n_channels, Mat3<T> r2u,
              MeshDims<T> dims, int x0, int nxl, T* __restrict__ values, T* __restrict__ dvalues,
              T* __restrict__ grad_positions, int accumulate, T* __restrict__ grad_r2u,
              PointEpilogue<T> epi) {
  constexpr int G = GroupSize<N>::value;
  constexpr bool DERIV = (MODE & 6) != 0;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t point_raw = tid / G;
  const bool valid = point_raw < n_points;
  const int64_t point = valid ? point_raw : n_points - 1;
  const int lane = (int)(tid % G);
  const bool active = lane < N;

  int first[3];
  T w[3][N], dw[3][N];
  point_stencil<T, METHOD, N, DERIV>(positions + 3 * point, r2u, dims, first, w, dw);
  const int nx = dims.n[0], ny = dims.n[1], nz = dims.n[2];
  const unsigned plane = (unsigned)ny * nz;
  const int64_t mesh_size = (int64_t)plane * nxl;   // the local slab holds x planes x0 .. x0 + nxl - 1
  unsigned xoff[N], yoff[N];
  {
    int ix = first[0], iy = first[1];
#pragma unroll
    for (int a = 0; a < N; ++a) {
      const unsigned lx = (unsigned)(ix - x0);
      const bool inside = lx < (unsigned)nxl;   // planes outside the slab: zero weight, loads redirected
      xoff[a] = inside ? lx * plane : 0u;
      if (!inside) { w[0][a] = T(0); if (DERIV) dw[0][a] = T(0); }
      yoff[a] = (unsigned)iy * nz;
      ix = (ix + 1 >= nx) ? wrap_add(ix + 1, nx) : ix + 1;
      iy = (iy + 1 >= ny) ? wrap_add(iy + 1, ny) : iy + 1;
    }
  }
  const unsigned iz = (unsigned)wrap_add(first[2] + (active ? lane : 0), nz);
  const T wz = active ? pick<T, N>(w[2], lane) : T(0);
  const T dwz = (DERIV && active) ? pick<T, N>(dw[2], lane) : T(0);

  T gu[3] = {T(0), T(0), T(0)};  // vjp accumulator in mesh coordinates
  for (int ch = 0; ch < n_channels; ++ch) {
    const T* src = mesh + ch * mesh_size + iz;
    // S0 = sum_ab v wx wy, S1 = sum v dwx wy, S2 = sum v wx dwy   (this lane's z offset)
    T s0 = T(0), s1 = T(0), s2 = T(0);
#pragma unroll
    for (int a = 0; a < N; ++a) {
      T t0 = T(0), t1 = T(0);
#pragma unroll
      for (int b = 0; b < N; ++b) {
        const T v = __ldg(src + (xoff[a] + yoff[b]));
        t0 = fma_t(v, w[1][b], t0);
        if (DERIV) t1 = fma_t(v, dw[1][b], t1);
      }
      s0 = fma_t(w[0][a], t0, s0);
      if (DERIV) {
        s1 = fma_t(dw[0][a], t0, s1);
        s2 = fma_t(w[0][a], t1, s2);
      }
    }
    T val = wz * s0, du0 = wz * s1, du1 = wz * s2, du2 = dwz * s0;
    if (DERIV) {
      group_reduce4<T, G>(val, du0, du1, du2, lane);
    } else {
#pragma unroll
      for (int off = G / 2; off > 0; off >>= 1) val += __shfl_xor_sync(0xffffffffu, val, off, G);
    }
    if (lane == 0 && valid) {
      if (MODE & 1) {
        const int64_t o = point * n_channels + ch;
        if (epi.enabled)
          values[o] = values[o] + epi.scale * val - epi.add_coef[o] * epi.self_half -
                      epi.background * epi.dc[ch];
        else
          values[o] = val;
      }
      if (MODE & 2) {
        T* out = dvalues + (point * n_channels + ch) * 3;
#pragma unroll
        for (int b = 0; b < 3; ++b)  // du_a/dr_b = r2u[b][a]
          out[b] = r2u.m[3 * b] * du0 + r2u.m[3 * b + 1] * du1 + r2u.m[3 * b + 2] * du2;
      }
      if (MODE & 4) {
        const T cf = coef[point * n_channels + ch];
        gu[0] = fma_t(cf, du0, gu[0]);
        gu[1] = fma_t(cf, du1, gu[1]);
        gu[2] = fma_t(cf, du2, gu[2]);
      }
    }
  }
  if (MODE & 4) {
    if (lane == 0 && valid) {
      T* out = grad_positions + 3 * point;
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        T g = r2u.m[3 * b] * gu[0] + r2u.m[3 * b + 1] * gu[1] + r2u.m[3 * b + 2] * gu[2];
        if (epi.enabled && epi.coef2 != nullptr) {
          for (int ch = 0; ch < n_channels; ++ch)
            g = fma_t(epi.coef2[point * n_channels + ch], epi.dvalues2[(point * n_channels + ch) * 3 + b], g);
          g *= epi.vjp_scale;
        }
        out[b] = accumulate ? out[b] + g : g;
      }
    }
    if (grad_r2u != nullptr) {  // block-uniform branch
      __shared__ T red[9][8];
      const bool lead = (lane == 0 && valid);
      const T r[3] = {positions[3 * point], positions[3 * point + 1], positions[3 * point + 2]};
      const int warp = threadIdx.x >> 5, wl = threadIdx.x & 31;
#pragma unroll
      for (int e = 0; e < 9; ++e) {
        T v = lead ? r[e / 3] * gu[e % 3] : T(0);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (wl == 0) red[e][warp] = v;
      }
      __syncthreads();
      if (threadIdx.x < 9) {
        T v = T(0);
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) v += red[threadIdx.x][k];
        red_add(grad_r2u + threadIdx.x, v);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// gather, one thread per point with 16-byte loads along z (nz a multiple of 4 floats / 2 doubles).
//
// The z window of the stencil is widened to whole aligned vectors: NV vectors of VEC elements
// starting at the aligned index below the first node; the 1-D z weights are shifted into that
// window once per point (zeros outside), so a stencil row (a, b) costs NV vector loads and
// NV * VEC fused multiply-adds per accumulated quantity -- about 6x fewer instructions per
// point than the lane-per-z-offset kernel above, with N * N * NV independent loads in flight
// per thread.  Planes outside the local x slab get zero weight (their loads are redirected to
// plane 0); points with no plane inside the slab skip the loads altogether.
// ---------------------------------------------------------------------------------------
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  using type = float4;
  static constexpr int VEC = 4;
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};
template <> struct Vec16<double> {
  using type = double2;
  static constexpr int VEC = 2;
  static __device__ __forceinline__ void load(const double* p, double (&v)[2]) {
    const double2 t = __ldg(reinterpret_cast<const double2*>(p));
    v[0] = t.x; v[1] = t.y;
  }
};

template <typename T, int METHOD, int N, int MODE>
__global__ void __launch_bounds__(128)
gather_point_kernel(const T* __restrict__ mesh, const T* __restrict__ positions,
                    const T* __restrict__ coef, int64_t n_points, int n_channels, Mat3<T> r2u,
                    MeshDims<T> dims, int x0, int nxl, T* __restrict__ values, T* __restrict__ dvalues,
                    T* __restrict__ grad_positions, int accumulate, T* __restrict__ grad_r2u,
                    PointEpilogue<T> epi, const int* __restrict__ point_list,
                    const int* __restrict__ list_count) {
  constexpr int VEC = Vec16<T>::VEC;
  constexpr int NV = (N + VEC - 2) / VEC + 1;     // vectors covering any window of N starting at o < VEC
  constexpr int W = NV * VEC;
  constexpr bool DERIV = (MODE & 6) != 0;
  // with a point list (slab decomposition: the points whose stencil reaches into the local slab)
  // the threads stride over the list, whose length is only known on the device
  const int64_t total = point_list != nullptr ? (int64_t)*list_count : n_points;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  do {
  const bool valid = it < total;
  if (!valid && !((MODE & 4) && grad_r2u != nullptr)) return;
  const int64_t point = valid ? (point_list != nullptr ? (int64_t)point_list[it] : it) : n_points - 1;

  int first[3];
  T w[3][N], dw[3][N];
  point_stencil<T, METHOD, N, DERIV>(positions + 3 * point, r2u, dims, first, w, dw);
  const int nx = dims.n[0], ny = dims.n[1], nz = dims.n[2];
  const unsigned plane = (unsigned)ny * nz;
  const int64_t mesh_size = (int64_t)plane * nxl;
  unsigned xoff[N], yoff[N];
  bool any_inside = false;
  {
    int ix = first[0], iy = first[1];
#pragma unroll
    for (int a = 0; a < N; ++a) {
      const unsigned lx = (unsigned)(ix - x0);
      const bool inside = lx < (unsigned)nxl;
      any_inside |= inside;
      xoff[a] = inside ? lx * plane : 0u;
      if (!inside) { w[0][a] = T(0); if (DERIV) dw[0][a] = T(0); }
      yoff[a] = (unsigned)iy * nz;
      ix = (ix + 1 >= nx) ? wrap_add(ix + 1, nx) : ix + 1;
      iy = (iy + 1 >= ny) ? wrap_add(iy + 1, ny) : iy + 1;
    }
  }
  // z window: aligned vectors, weights shifted by o = first[2] mod VEC
  const int zb = first[2] & ~(VEC - 1);
  const int o = first[2] - zb;
  unsigned zoff[NV];
  T wz[W], dwz[W];
  {
    int z = zb;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      zoff[j] = (unsigned)z;
      z = (z + VEC >= nz) ? wrap_add(z + VEC, nz) : z + VEC;
    }
#pragma unroll
    for (int k = 0; k < W; ++k) {
      T a = T(0), b = T(0);
#pragma unroll
      for (int c = 0; c < N; ++c) {
        if (k - c >= 0 && k - c < VEC) {          // compile-time feasible shifts only
          a = (k - c == o) ? w[2][c] : a;
          if (DERIV) b = (k - c == o) ? dw[2][c] : b;
        }
      }
      wz[k] = a;
      dwz[k] = b;
    }
  }

  T gu[3] = {T(0), T(0), T(0)};  // vjp accumulator in mesh coordinates
  T extra[3] = {T(0), T(0), T(0)};  // sum_c coef2 * dvalues2 of the fused backward epilogue
  const bool with_extra = (MODE & 4) && epi.enabled && epi.coef2 != nullptr;
  for (int ch = 0; ch < n_channels; ++ch) {
    const T* src = mesh + ch * mesh_size;
    // per-point operands of the epilogues: issued before the mesh loads so that their latency
    // hides under the stencil loop
    T cf = T(0), prev = T(0), addc = T(0);
    if (valid) {
      if (MODE & 4) cf = coef[point * n_channels + ch];
      if ((MODE & 1) && epi.enabled) {
        prev = values[point * n_channels + ch];
        addc = epi.add_coef[point * n_channels + ch];
      }
      if (with_extra) {
        const T c2 = epi.coef2[point * n_channels + ch];
        const T* d2 = epi.dvalues2 + (point * n_channels + ch) * 3;
        extra[0] = fma_t(c2, d2[0], extra[0]);
        extra[1] = fma_t(c2, d2[1], extra[1]);
        extra[2] = fma_t(c2, d2[2], extra[2]);
      }
    }
    T val = T(0), du0 = T(0), du1 = T(0), du2 = T(0);
    if (any_inside) {
#pragma unroll(N <= 5 ? N : 1)
      for (int a = 0; a < N; ++a) {
        T ta0 = T(0), ta1 = T(0), ta2 = T(0);
#pragma unroll
        for (int b = 0; b < N; ++b) {
          const T* row = src + (xoff[a] + yoff[b]);
          T t0 = T(0), t1 = T(0);
#pragma unroll
          for (int j = 0; j < NV; ++j) {
            T v[VEC];
            Vec16<T>::load(row + zoff[j], v);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
              t0 = fma_t(v[e], wz[j * VEC + e], t0);
              if (DERIV) t1 = fma_t(v[e], dwz[j * VEC + e], t1);
            }
          }
          ta0 = fma_t(w[1][b], t0, ta0);
          if (DERIV) {
            ta1 = fma_t(dw[1][b], t0, ta1);
            ta2 = fma_t(w[1][b], t1, ta2);
          }
        }
        val = fma_t(w[0][a], ta0, val);
        if (DERIV) {
          du0 = fma_t(dw[0][a], ta0, du0);
          du1 = fma_t(w[0][a], ta1, du1);
          du2 = fma_t(w[0][a], ta2, du2);
        }
      }
    }
    if (valid) {
      if (MODE & 1) {
        const int64_t oidx = point * n_channels + ch;
        if (epi.enabled)
          values[oidx] = prev + epi.scale * val - addc * epi.self_half - epi.background * epi.dc[ch];
        else
          values[oidx] = val;
      }
      if (MODE & 2) {
        T* out = dvalues + (point * n_channels + ch) * 3;
#pragma unroll
        for (int b = 0; b < 3; ++b)  // du_a/dr_b = r2u[b][a]
          out[b] = r2u.m[3 * b] * du0 + r2u.m[3 * b + 1] * du1 + r2u.m[3 * b + 2] * du2;
      }
      if (MODE & 4) {
        gu[0] = fma_t(cf, du0, gu[0]);
        gu[1] = fma_t(cf, du1, gu[1]);
        gu[2] = fma_t(cf, du2, gu[2]);
      }
    }
  }
  if (MODE & 4) {
    if (valid) {
      T* out = grad_positions + 3 * point;
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        T g = r2u.m[3 * b] * gu[0] + r2u.m[3 * b + 1] * gu[1] + r2u.m[3 * b + 2] * gu[2];
        if (with_extra) g = (g + extra[b]) * epi.vjp_scale;
        out[b] = accumulate ? out[b] + g : g;
      }
    }
    if (grad_r2u != nullptr) {  // block-uniform branch
      __shared__ T red[9][4];
      const T r[3] = {positions[3 * point], positions[3 * point + 1], positions[3 * point + 2]};
      const int warp = threadIdx.x >> 5, wl = threadIdx.x & 31;
#pragma unroll
      for (int e = 0; e < 9; ++e) {
        T v = valid ? r[e / 3] * gu[e % 3] : T(0);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (wl == 0) red[e][warp] = v;
      }
      __syncthreads();
      if (threadIdx.x < 9) {
        T v = T(0);
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) v += red[threadIdx.x][k];
        red_add(grad_r2u + threadIdx.x, v);
      }
    }
  }
  it += stride;
  } while (point_list != nullptr && it < total);
}

// ---------------------------------------------------------------------------------------
// host-side dispatch
// ---------------------------------------------------------------------------------------
// local x slab of a decomposed mesh (x0 = 0, nxl = nx: the whole mesh) and, optionally, the device
// list of the points that reach into it
struct SlabArgs {
  int x0, nxl;
  const int* list;
  const int* count;
};

template <typename T, int METHOD, int N>
int launch_spread(const void* positions, const void* weights, int64_t n_points, int n_channels,
                  const double* r2u, int nx, int ny, int nz, SlabArgs sl, void* mesh,
                  cudaStream_t stream) {
  const int64_t threads = n_points * GroupSize<N>::value;  // one lane per (point, z offset)
  const int block = 256;
  int64_t grid = (threads + block - 1) / block;
  if (grid == 0) return 0;
  // list mode: the list length lives on the device, the threads stride over it
  if (sl.list != nullptr && grid > 8ll * num_sms()) grid = 8ll * num_sms();
  spread_kernel<T, METHOD, N><<<(unsigned)grid, block, 0, stream>>>(
      (const T*)positions, (const T*)weights, n_points, n_channels, load_mat3<T>(r2u),
      make_dims<T>(nx, ny, nz), sl.x0, sl.nxl, (T*)mesh, sl.list, sl.count);
  TPME_CUDA_OK(cudaGetLastError());
  return 0;
}

template <typename T, int METHOD, int N, int MODE>
int launch_gather(const void* mesh, const void* positions, const void* coef, int64_t n_points,
                  int n_channels, const double* r2u, int nx, int ny, int nz, SlabArgs sl,
                  void* values, void* dvalues, void* grad_positions, int accumulate, void* grad_r2u,
                  const tpme_point_epilogue* epi_host, cudaStream_t stream) {
  constexpr int G = GroupSize<N>::value;
  const int x0 = sl.x0, nxl = sl.nxl;
  PointEpilogue<T> epi;
  epi.enabled = epi_host != nullptr;
  if (epi_host) {
    epi.add_coef = (const T*)epi_host->add_coef; epi.dc = (const T*)epi_host->dc;
    epi.scale = (T)epi_host->scale; epi.self_half = (T)epi_host->self_half;
    epi.background = (T)epi_host->background;
    epi.coef2 = (const T*)epi_host->coef2; epi.dvalues2 = (const T*)epi_host->dvalues2;
    epi.vjp_scale = (T)epi_host->vjp_scale;
  } else {
    epi.add_coef = epi.dc = epi.coef2 = epi.dvalues2 = nullptr;
    epi.scale = epi.self_half = epi.background = epi.vjp_scale = T(0);
  }
  if (n_points == 0) return 0;
  // one thread per point with 16-byte loads when the z rows allow it (TPME_GATHER=lanes forces the
  // lane-per-z-offset kernel, for A/B tests)
  static const bool force_lanes = [] { const char* e = getenv("TPME_GATHER"); return e && e[0] == 'l'; }();
  if (!force_lanes && nz % Vec16<T>::VEC == 0 && ((uintptr_t)mesh % 16) == 0) {
    const int block = n_points >= 4 * 128 * (int64_t)num_sms() ? 128 : 64;
    int64_t grid = (n_points + block - 1) / block;
    if (sl.list != nullptr) {
      TPME_REQUIRE(grad_r2u == nullptr, "cell gradients are not available with a point list");
      if (grid > 8ll * num_sms()) grid = 8ll * num_sms();
    }
    gather_point_kernel<T, METHOD, N, MODE><<<(unsigned)grid, block, 0, stream>>>(
        (const T*)mesh, (const T*)positions, (const T*)coef, n_points, n_channels,
        load_mat3<T>(r2u), make_dims<T>(nx, ny, nz), x0, nxl, (T*)values, (T*)dvalues, (T*)grad_positions,
        accumulate, (T*)grad_r2u, epi, sl.list, sl.count);
    TPME_CUDA_OK(cudaGetLastError());
    return 0;
  }
  TPME_REQUIRE(sl.list == nullptr, "point lists need mesh rows of whole 16-byte vectors (nz % 4 == 0)");
  const int64_t threads = n_points * G;
  const int block = 256;
  const int64_t grid = (threads + block - 1) / block;
  gather_kernel<T, METHOD, N, MODE><<<(unsigned)grid, block, 0, stream>>>(
      (const T*)mesh, (const T*)positions, (const T*)coef, n_points, n_channels,
      load_mat3<T>(r2u), make_dims<T>(nx, ny, nz), x0, nxl, (T*)values, (T*)dvalues, (T*)grad_positions,
      accumulate, (T*)grad_r2u, epi);
  TPME_CUDA_OK(cudaGetLastError());
  return 0;
}

#define TPME_DISPATCH_STENCIL(CALL)                                         \
  if (method == TPME_P3M) {                                                 \
    switch (nodes) {                                                        \
      case 1: return CALL(TPME_P3M, 1);                                     \
      case 2: return CALL(TPME_P3M, 2);                                     \
      case 3: return CALL(TPME_P3M, 3);                                     \
      case 4: return CALL(TPME_P3M, 4);                                     \
      case 5: return CALL(TPME_P3M, 5);                                     \
    }                                                                       \
  } else if (method == TPME_LAGRANGE) {                                     \
    switch (nodes) {                                                        \
      case 3: return CALL(TPME_LAGRANGE, 3);                                \
      case 4: return CALL(TPME_LAGRANGE, 4);                                \
      case 5: return CALL(TPME_LAGRANGE, 5);                                \
      case 6: return CALL(TPME_LAGRANGE, 6);                                \
      case 7: return CALL(TPME_LAGRANGE, 7);                                \
    }                                                                       \
  }                                                                         \
  set_last_error("stencil", "unsupported (method, interpolation_nodes) pair"); \
  return 2;

template <typename T>
int spread_dispatch(const void* positions, const void* weights, int64_t n_points, int n_channels,
                    const double* r2u, int nx, int ny, int nz, SlabArgs sl, int nodes, int method,
                    void* mesh, cudaStream_t stream) {
#define CALL(M, N) \
  launch_spread<T, M, N>(positions, weights, n_points, n_channels, r2u, nx, ny, nz, sl, mesh, stream)
  TPME_DISPATCH_STENCIL(CALL)
#undef CALL
}

template <typename T, int MODE>
int gather_dispatch(const void* mesh, const void* positions, const void* coef, int64_t n_points,
                    int n_channels, const double* r2u, int nx, int ny, int nz, SlabArgs sl,
                    int nodes, int method, void* values, void* dvalues, void* grad_positions,
                    int accumulate, void* grad_r2u, const tpme_point_epilogue* epi,
                    cudaStream_t stream) {
#define CALL(M, N)                                                                                 \
  launch_gather<T, M, N, MODE>(mesh, positions, coef, n_points, n_channels, r2u, nx, ny, nz, sl,  \
                               values, dvalues, grad_positions, accumulate, grad_r2u, epi, stream)
  TPME_DISPATCH_STENCIL(CALL)
#undef CALL
}

static int check_mesh_args(int dtype, int nx, int ny, int nz, int n_channels, int64_t n_points) {
  TPME_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (float32) or 1 (float64)");
  TPME_REQUIRE(nx > 0 && ny > 0 && nz > 0, "mesh dimensions must be positive");
  TPME_REQUIRE(n_channels >= 0 && n_points >= 0, "negative sizes");
  return 0;
}

}  // namespace tpme

using namespace tpme;

static int check_slab(int nx, int x0, int nxl) {
  TPME_REQUIRE(x0 >= 0 && nxl > 0 && x0 + nxl <= nx, "slab [x0, x0 + nx_local) must lie inside [0, nx)");
  return 0;
}

extern "C" int tpme_spread_slab(int dtype, const void* positions, const void* weights,
                                int64_t n_points, int n_channels, const double* r2u_host, int nx,
                                int ny, int nz, int x0, int nx_local, const int* point_list,
                                const int* list_count, int nodes, int method, void* mesh,
                                int accumulate, void* stream) {
  if (int rc = check_mesh_args(dtype, nx, ny, nz, n_channels, n_points)) return rc;
  if (int rc = check_slab(nx, x0, nx_local)) return rc;
  TPME_REQUIRE((point_list == nullptr) == (list_count == nullptr), "point_list and list_count go together");
  const SlabArgs sl{x0, nx_local, point_list, list_count};
  cudaStream_t s = (cudaStream_t)stream;
  const size_t elem = dtype == 0 ? 4 : 8;
  if (!accumulate)
    TPME_CUDA_OK(cudaMemsetAsync(mesh, 0, elem * (size_t)n_channels * nx_local * ny * nz, s));
  if (n_points == 0 || n_channels == 0) return 0;
  if (dtype == 0)
    return spread_dispatch<float>(positions, weights, n_points, n_channels, r2u_host, nx, ny, nz,
                                  sl, nodes, method, mesh, s);
  return spread_dispatch<double>(positions, weights, n_points, n_channels, r2u_host, nx, ny, nz,
                                 sl, nodes, method, mesh, s);
}

extern "C" int tpme_spread(int dtype, const void* positions, const void* weights,
                           int64_t n_points, int n_channels, const double* r2u_host, int nx,
                           int ny, int nz, int nodes, int method, void* mesh, int accumulate,
                           void* stream) {
  return tpme_spread_slab(dtype, positions, weights, n_points, n_channels, r2u_host, nx, ny, nz, 0,
                          nx, nullptr, nullptr, nodes, method, mesh, accumulate, stream);
}

extern "C" int tpme_gather_slab(int dtype, const void* mesh, const void* positions,
                                int64_t n_points, int n_channels, const double* r2u_host, int nx,
                                int ny, int nz, int x0, int nx_local, const int* point_list,
                                const int* list_count, int nodes, int method, void* values,
                                void* dvalues, const tpme_point_epilogue* epilogue, void* stream) {
  if (int rc = check_mesh_args(dtype, nx, ny, nz, n_channels, n_points)) return rc;
  if (int rc = check_slab(nx, x0, nx_local)) return rc;
  TPME_REQUIRE((point_list == nullptr) == (list_count == nullptr), "point_list and list_count go together");
  const SlabArgs sl{x0, nx_local, point_list, list_count};
  TPME_REQUIRE(values != nullptr || dvalues != nullptr, "nothing to compute");
  if (n_points == 0 || n_channels == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const int mode = (values ? 1 : 0) | (dvalues ? 2 : 0);
#define GO(T, MODE)                                                                          \
  return gather_dispatch<T, MODE>(mesh, positions, nullptr, n_points, n_channels, r2u_host,  \
                                  nx, ny, nz, sl, nodes, method, values, dvalues,            \
                                  nullptr, 0, nullptr, epilogue, s)
  if (dtype == 0) {
    if (mode == 1) GO(float, 1);
    if (mode == 2) GO(float, 2);
    GO(float, 3);
  }
  if (mode == 1) GO(double, 1);
  if (mode == 2) GO(double, 2);
  GO(double, 3);
#undef GO
}

extern "C" int tpme_gather(int dtype, const void* mesh, const void* positions, int64_t n_points,
                           int n_channels, const double* r2u_host, int nx, int ny, int nz,
                           int nodes, int method, void* values, void* dvalues,
                           const tpme_point_epilogue* epilogue, void* stream) {
  return tpme_gather_slab(dtype, mesh, positions, n_points, n_channels, r2u_host, nx, ny, nz, 0, nx,
                          nullptr, nullptr, nodes, method, values, dvalues, epilogue, stream);
}

extern "C" int tpme_gather_vjp_slab(int dtype, const void* mesh, const void* positions,
                                    const void* coef, int64_t n_points, int n_channels,
                                    const double* r2u_host, int nx, int ny, int nz, int x0,
                                    int nx_local, const int* point_list, const int* list_count,
                                    int nodes, int method, void* grad_positions, void* values,
                                    int accumulate, void* grad_r2u,
                                    const tpme_point_epilogue* epilogue, void* stream) {
  if (int rc = check_mesh_args(dtype, nx, ny, nz, n_channels, n_points)) return rc;
  if (int rc = check_slab(nx, x0, nx_local)) return rc;
  TPME_REQUIRE((point_list == nullptr) == (list_count == nullptr), "point_list and list_count go together");
  const SlabArgs sl{x0, nx_local, point_list, list_count};
  TPME_REQUIRE(grad_positions != nullptr && coef != nullptr, "grad_positions / coef missing");
  if (n_points == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (n_channels == 0) {
    if (!accumulate)
      TPME_CUDA_OK(cudaMemsetAsync(grad_positions, 0, (dtype ? 8 : 4) * 3 * (size_t)n_points, s));
    return 0;
  }
#define GO(T, MODE)                                                                           \
  return gather_dispatch<T, MODE>(mesh, positions, coef, n_points, n_channels, r2u_host, nx,  \
                                  ny, nz, sl, nodes, method, values, nullptr,                 \
                                  grad_positions, accumulate, grad_r2u, epilogue, s)
  if (dtype == 0) {
    if (values) GO(float, 5);
    GO(float, 4);
  }
  if (values) GO(double, 5);
  GO(double, 4);
#undef GO
}

extern "C" int tpme_gather_vjp(int dtype, const void* mesh, const void* positions,
                               const void* coef, int64_t n_points, int n_channels,
                               const double* r2u_host, int nx, int ny, int nz, int nodes,
                               int method, void* grad_positions, void* values, int accumulate,
                               void* grad_r2u, const tpme_point_epilogue* epilogue,
                               void* stream) {
  return tpme_gather_vjp_slab(dtype, mesh, positions, coef, n_points, n_channels, r2u_host, nx, ny,
                              nz, 0, nx, nullptr, nullptr, nodes, method, grad_positions, values,
                              accumulate, grad_r2u, epilogue, stream);
}

extern "C" int tpme_slab_select_points(int dtype, const void* positions, int64_t n_points,
                                       const double* r2u_host, int nx, int ny, int nz, int x0,
                                       int nx_local, int nodes, int* point_list, int* list_count,
                                       void* stream) {
  if (int rc = check_mesh_args(dtype, nx, ny, nz, 1, n_points)) return rc;
  if (int rc = check_slab(nx, x0, nx_local)) return rc;
  TPME_REQUIRE(point_list != nullptr && list_count != nullptr, "null output");
  TPME_REQUIRE(n_points < (1ll << 31), "point lists hold 32-bit indices");
  cudaStream_t s = (cudaStream_t)stream;
  TPME_CUDA_OK(cudaMemsetAsync(list_count, 0, sizeof(int), s));
  if (n_points == 0) return 0;
  const int64_t grid = (n_points + 255) / 256;
  if (dtype == 0)
    select_slab_points_kernel<float><<<(unsigned)grid, 256, 0, s>>>(
        (const float*)positions, n_points, load_mat3<float>(r2u_host), make_dims<float>(nx, ny, nz), nodes, x0,
        nx_local, point_list, list_count);
  else
    select_slab_points_kernel<double><<<(unsigned)grid, 256, 0, s>>>(
        (const double*)positions, n_points, load_mat3<double>(r2u_host), make_dims<double>(nx, ny, nz), nodes,
        x0, nx_local, point_list, list_count);
  TPME_CUDA_OK(cudaGetLastError());
  return 0;
}
