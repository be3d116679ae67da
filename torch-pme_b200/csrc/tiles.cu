// Tiled mesh interpolation: atoms binned by mesh tile, the tile of the mesh staged in shared
// memory and moved with the bulk-copy engine (TMA: cp.async.bulk / cp.reduce.async.bulk).
//
// Replaces, like interp.cu, MeshInterpolator.compute_weights / points_to_mesh / mesh_to_points
// (src/torchpme/lib/mesh_interpolator.py:303-457).  interp.cu sends every stencil node of every
// atom to the L2 as its own reduction (64 per atom for 4 nodes), which is bound by the L2
// reduction pipe, and its gather keeps 32 independent 16-byte loads per thread in flight.  Here
//
//   tpme_tile_sort     bins the atoms by (pencil, z chunk) of their FIRST stencil node: a pencil is
//                      a tx x ty footprint of first nodes over the whole z axis, a z chunk is zw
//                      consecutive first nodes.  Output: per-atom records (the three stencil offsets
//                      x in [-1/2, 1/2] and the position of the first node inside the pencil) in bin
//                      order + the permutation + the bin starts.  Done once per set of positions and
//                      shared by the spread, the gather and both backward launches of a step.
//   tpme_tile_spread   one CTA per pencil: the (tx + n - 1) x (ty + n - 1) x nz tile lives in shared
//                      memory; a warp takes one z chunk at a time and adds the n^3 contributions of
//                      its atoms with plain load-add-store (32 lanes = 32 distinct nodes of ONE atom,
//                      bank-conflict free; chunks of equal parity never overlap, so the two parity
//                      phases need no atomics at all); the tile is then flushed row by row with
//                      cp.reduce.async.bulk (add) into the L2-resident mesh: ~3.4 sector reductions
//                      per atom instead of ~15.
//   tpme_tile_gather   one CTA per pencil: rows arrive by cp.async.bulk behind an mbarrier, one
//                      thread per atom reads its n^3 nodes from shared memory; same modes and fused
//                      epilogues as the direct kernels.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "scan.cuh"
#include "stencil_point.cuh"
#include "../../include/torchpme_b200.h"

namespace tpme {

// ---------------------------------------------------------------------------------------
// geometry of the tiling (by value to the kernels)
// ---------------------------------------------------------------------------------------
struct TileGeom {
  int nx, ny, nz;
  int nodes;
  int tx_shift, ty_shift, zw_shift;   // pencil footprint / z chunk width (powers of two)
  int npx, npy, nzc;                  // pencils per axis, z chunks per pencil
  int nzt, tz, nzc_t;                 // z tiles per pencil, first-node z extent of a tile, z chunks per tile
  int rows_x, rows_y;                 // tile rows: tx + nodes - 1, ty + nodes - 1
  int sy, sx;                         // row / plane stride of the shared-memory tile in elements
};

// geometry for a kernel that cuts the pencils into `nzt` z tiles (the spread and the gather choose
// their own: the bins are z chunks of zw planes, any tile extent that is a multiple of zw works)
static TileGeom geom_of(const tpme_tile_plan& p, int nzt) {
  TileGeom g;
  g.nx = p.nx; g.ny = p.ny; g.nz = p.nz; g.nodes = p.nodes;
  auto log2i = [](int v) { int s = 0; while ((1 << s) < v) ++s; return s; };
  g.tx_shift = log2i(p.tx); g.ty_shift = log2i(p.ty); g.zw_shift = log2i(p.zw);
  g.npx = p.npx; g.npy = p.npy; g.nzc = p.nzc;
  g.nzt = nzt; g.tz = p.nz / nzt; g.nzc_t = p.nzc / nzt;
  g.rows_x = p.tx + p.nodes - 1; g.rows_y = p.ty + p.nodes - 1;
  g.sy = g.tz + 4;                                  // = 4 (mod 32): the 4 y rows of a stencil hit different banks
  const int raw = g.rows_y * g.sy;                  // multiple of 4 elements
  g.sx = raw + ((16 - raw % 32) + 32) % 32;         // = 16 (mod 32): the two x planes of a warp access too
  return g;
}

// packed position of the first node: lx | ly << 5 | first_z << 10
__device__ __forceinline__ int pack_first(int lx, int ly, int fz) { return lx | (ly << 5) | (fz << 10); }

template <typename T> struct Rec;   // 4 reals: x offsets of the three axes + packed first node
template <> struct Rec<float> {
  float x[3]; int packed;
  static __device__ __forceinline__ Rec load(const float* p) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    Rec r; r.x[0] = v.x; r.x[1] = v.y; r.x[2] = v.z; r.packed = __float_as_int(v.w); return r;
  }
  static __device__ __forceinline__ Rec load_shared(const float* p) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    Rec r; r.x[0] = v.x; r.x[1] = v.y; r.x[2] = v.z; r.packed = __float_as_int(v.w); return r;
  }
  static __device__ __forceinline__ void store(float* p, float a, float b, float c, int packed) {
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, __int_as_float(packed));
  }
};
template <> struct Rec<double> {
  double x[3]; int packed;
  static __device__ __forceinline__ Rec load(const double* p) {
    const double2 u = __ldg(reinterpret_cast<const double2*>(p));
    const double2 v = __ldg(reinterpret_cast<const double2*>(p) + 1);
    Rec r; r.x[0] = u.x; r.x[1] = u.y; r.x[2] = v.x; r.packed = (int)__double_as_longlong(v.y); return r;
  }
  static __device__ __forceinline__ Rec load_shared(const double* p) {
    const double2 u = reinterpret_cast<const double2*>(p)[0];
    const double2 v = reinterpret_cast<const double2*>(p)[1];
    Rec r; r.x[0] = u.x; r.x[1] = u.y; r.x[2] = v.x; r.packed = (int)__double_as_longlong(v.y); return r;
  }
  static __device__ __forceinline__ void store(double* p, double a, double b, double c, int packed) {
    reinterpret_cast<double2*>(p)[0] = make_double2(a, b);
    reinterpret_cast<double2*>(p)[1] = make_double2(c, __longlong_as_double((long long)packed));
  }
};

// first stencil node (wrapped into the mesh) and stencil offset of one point along the three axes;
// the same arithmetic as point_stencil (stencil_point.cuh), so both kernel families see the same x
template <typename T>
__device__ __forceinline__ void point_first(const T* __restrict__ pos, const Mat3<T>& r2u,
                                            const MeshDims<T>& dims, int nodes, int (&first)[3], T (&x)[3]) {
  const T r0 = pos[0], r1 = pos[1], r2 = pos[2];
  const bool even = (nodes % 2) == 0;
  const T shift = T(1 - (nodes + 1) / 2);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const T u = r0 * r2u.m[a] + r1 * r2u.m[3 + a] + r2 * r2u.m[6 + a];
    T base;
    if (even) {
      base = floor_t(u);
      x[a] = u - (base + T(0.5));
    } else {
      base = rint_t(u);
      x[a] = u - base;
    }
    first[a] = wrap_base<T>(base + shift, dims.n[a], dims.inv_n[a]);
  }
}

__device__ __forceinline__ int bin_of(const TileGeom& g, const int (&first)[3]) {
  return (((first[0] >> g.tx_shift) * g.npy + (first[1] >> g.ty_shift)) * g.nzc) + (first[2] >> g.zw_shift);
}

// ---------------------------------------------------------------------------------------
// sort: count -> scan -> fill
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
tile_count_kernel(const T* __restrict__ positions, int64_t n_points, Mat3<T> r2u, MeshDims<T> dims,
                  TileGeom g, int* __restrict__ counts, int2* __restrict__ key_rank) {
  pdl_trigger();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_points) return;
  int first[3];
  T x[3];
  point_first<T>(positions + 3 * i, r2u, dims, g.nodes, first, x);
  const int key = bin_of(g, first);
  key_rank[i] = make_int2(key, atomicAdd(counts + key, 1));
}

template <typename T>
__global__ void __launch_bounds__(256)
tile_fill_kernel(const T* __restrict__ positions, int64_t n_points, Mat3<T> r2u, MeshDims<T> dims,
                 TileGeom g, const int* __restrict__ start, const int2* __restrict__ key_rank,
                 T* __restrict__ sorted_rec, int* __restrict__ sorted_idx) {
  pdl_trigger();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int first[3] = {0, 0, 0};
  T x[3] = {T(0), T(0), T(0)};
  if (i < n_points) point_first<T>(positions + 3 * i, r2u, dims, g.nodes, first, x);   // not from the chain
  pdl_wait();                                   // bin starts (scan) and ranks (count) are
  if (i >= n_points) return;
  const int2 kr = key_rank[i];
  const int slot = start[kr.x] + kr.y;
  const int lx = first[0] & ((1 << g.tx_shift) - 1), ly = first[1] & ((1 << g.ty_shift) - 1);
  Rec<T>::store(sorted_rec + 4 * (int64_t)slot, x[0], x[1], x[2], pack_first(lx, ly, first[2]));
  sorted_idx[slot] = (int)i;
}

// ---------------------------------------------------------------------------------------
// bulk-copy (TMA) and mbarrier wrappers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <typename T> __device__ __forceinline__ void bulk_reduce_add(T* gdst, const T* ssrc, uint32_t bytes);
template <> __device__ __forceinline__ void bulk_reduce_add<float>(float* gdst, const float* ssrc, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
               :: "l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
template <> __device__ __forceinline__ void bulk_reduce_add<double>(double* gdst, const double* ssrc, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
               :: "l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// ---------------------------------------------------------------------------------------
// spread, 4 nodes.  Shared memory: tile[rows_x][rows_y][nz (+ pad)] | 2 chunk counters.
// Lane l of a warp owns the nodes (a, b, c) = (l >> 4 [+ 2], (l >> 2) & 3, l & 3) of the atom the
// warp is working on: row stride = 4 (mod 32) and plane stride = 16 (mod 32) elements make the 32
// addresses of one load / store hit 32 different banks.
// ---------------------------------------------------------------------------------------
constexpr int kStageWords = 13;   // per staged atom: wx[4], q wy[4], wz[4], tile offset

template <typename T>
__global__ void __launch_bounds__(sizeof(T) == 4 ? 1024 : 512, 1)
tile_spread4_kernel(const T* __restrict__ sorted_rec, const int* __restrict__ sorted_idx,
                    const int* __restrict__ bin_start, const T* __restrict__ weights, int n_channels,
                    int method, TileGeom g, T* __restrict__ mesh, int batch, int debug) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* tile = reinterpret_cast<T*>(smem_raw);
  const int tile_elems = g.rows_x * g.sx;
  int* ctrl = reinterpret_cast<int*>(tile + tile_elems);   // chunk range bounds of the warps (256 bytes reserved)
  // per-warp staging area: `batch` atoms x 13 words (12 one-dimensional weights + tile offset)
  T* st_rec = reinterpret_cast<T*>(reinterpret_cast<unsigned char*>(ctrl) + 256) + (threadIdx.x >> 5) * (kStageWords * batch);

  pdl_trigger();
  const int pencil = blockIdx.x / g.nzt, zt = blockIdx.x - pencil * g.nzt;
  const int bin0 = pencil * g.nzc + zt * g.nzc_t;
  // The chunk starts are parked in the staging areas of the warps when they fit there (else in the not yet
  // used tile): the tile can then be zeroed BEFORE the wait, under the tail of the kernel before.
  const int n_warps_all = blockDim.x >> 5;
  const bool cs_in_stage = (size_t)(g.nzc_t + 1) * sizeof(int) <= (size_t)n_warps_all * kStageWords * batch * sizeof(T);
  if (cs_in_stage) {
    const int n16 = (int)((size_t)tile_elems * sizeof(T) / 16);
    float4* t4 = reinterpret_cast<float4*>(tile);
    for (int k = threadIdx.x; k < n16; k += blockDim.x) t4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  pdl_wait();                                   // sorted atoms / zeroed mesh: written by the kernels before
  const int p_begin = bin_start[bin0], p_end = bin_start[bin0 + g.nzc_t];
  if (p_begin == p_end) return;                 // nothing lands here: the mesh was zeroed by the caller
  const int px = pencil / g.npy, py = pencil - px * g.npy;
  const int z_lo = zt * g.tz;                   // first-node z of this tile: [z_lo, z_lo + tz)
  const int lane = threadIdx.x & 31;
  const int la = lane >> 4, lb = (lane >> 2) & 3, lc = lane & 3;
  T* const lane_ptr = tile + (la * g.sx + lb * g.sy + lc);
  const int plane2 = 2 * g.sx;
  const int64_t mesh_size = (int64_t)g.nx * g.ny * g.nz;

  // ---- z chunks -> warps.  Every warp owns a contiguous range of chunks holding about the same number
  // of atoms (the chunk starts are the prefix sums of the sort).  Only ADJACENT chunks overlap, so a warp
  // first accumulates the first chunk of its range, the CTA synchronises once, and the warp walks
  // through the rest of its range: its last chunk then only meets the neighbour's first chunk, which
  // is complete.  No atomics, one barrier, and the atoms of a warp are one contiguous run of the
  // sorted arrays, so the fetch pipeline never drains.
  const int n_warps = blockDim.x >> 5, warp = threadIdx.x >> 5;
  int s1, mid, e2;
  {
    int* cs = cs_in_stage ? reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(ctrl) + 256)
                          : reinterpret_cast<int*>(tile);   // chunk starts
    for (int k = threadIdx.x; k <= g.nzc_t; k += blockDim.x) cs[k] = bin_start[bin0 + k];
    __syncthreads();
    if (warp == 0) {
      int bound = 0;
      if (lane < n_warps) {                          // first chunk whose start reaches this warp's share
        const int target = p_begin + (int)(((int64_t)(p_end - p_begin) * lane) / n_warps);
        int lo = 0, hi = g.nzc_t;
        while (lo < hi) {
          const int m = (lo + hi) >> 1;
          if (cs[m] < target) lo = m + 1; else hi = m;
        }
        bound = lo;
      }
      // at least two chunks per warp (first chunks of neighbouring warps must not be adjacent)
      for (int w = 1; w < n_warps; ++w) {
        const int prev = __shfl_sync(0xffffffffu, bound, w - 1);
        if (lane == w) bound = min(max(bound, prev + 2), g.nzc_t - 2 * (n_warps - w));
      }
      if (lane < n_warps) ctrl[lane] = bound;
      if (lane == 0) ctrl[n_warps] = g.nzc_t;
    }
    __syncthreads();
    const int c0 = ctrl[warp], c1 = ctrl[warp + 1];
    s1 = cs[c0]; mid = cs[c0 + 1]; e2 = cs[c1];
    __syncthreads();                                 // everybody has read its range: the tile may be zeroed
  }

  for (int ch = 0; ch < n_channels; ++ch) {
    // zero the tile (16-byte stores; the strides are multiples of 4 elements)
    if (ch > 0 || !cs_in_stage) {
      const int n16 = (int)((size_t)tile_elems * sizeof(T) / 16);
      float4* t4 = reinterpret_cast<float4*>(tile);
      for (int k = threadIdx.x; k < n16; k += blockDim.x) t4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    if (!(debug & 1)) {
      // The atoms of a range are fetched 32 at a time, one per lane (coalesced record / index loads, 32
      // independent weight gathers in flight), parked in the warp's staging area and then accumulated
      // one atom after the other by all 32 lanes; the fetch of the next batch is issued before the
      // current one is accumulated.
      Rec<T> pr;            // this lane's atom of the batch in flight
      T pq = T(0);
      auto fetch = [&](int base, int cnt) {
        if (lane < cnt) {
          pr = Rec<T>::load(sorted_rec + 4 * (int64_t)(base + lane));
          pq = __ldg(weights + (int64_t)__ldg(sorted_idx + base + lane) * n_channels + ch);
        }
      };
      int pos = s1;
      int cnt = min(batch, max(mid - s1, 0));
      if (cnt == 0) cnt = min(batch, max(e2 - pos, 0));  // empty first chunk: the batch belongs to the rest
      if (cnt > 0) fetch(pos, cnt);
      // precondition: the batch [pos, pos + cnt) lies inside [.., lim) and is in flight
      auto run_range = [&](int lim, int next_lim) {
        while (cnt > 0) {
          __syncwarp();       // the previous batch is fully consumed
          if (lane < cnt) {
            // The lane that fetched an atom evaluates its 12 one-dimensional weights ONCE and parks them
            // (13 words per atom, odd stride: conflict free) -- the 32 lanes that accumulate the atom then
            // only pick wx[a], wx[a + 2], q wy[b], wz[c] and the tile offset.  Rows hold tz + 4 elements:
            // first_z + c may run into the 3 halo elements behind the tile's own z range (flushed to the
            // next z tile / wrapped to z = 0 by a second bulk reduction): no wrap in the loop.
            T w[3][4], dw[3][4];
            if (method == TPME_P3M) {
#pragma unroll
              for (int a = 0; a < 3; ++a) Stencil<TPME_P3M, 4>::template eval<T, false>(pr.x[a], w[a], dw[a]);
            } else {
#pragma unroll
              for (int a = 0; a < 3; ++a) Stencil<TPME_LAGRANGE, 4>::template eval<T, false>(pr.x[a], w[a], dw[a]);
            }
            T* st = st_rec + kStageWords * lane;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              st[k] = w[0][k];
              st[4 + k] = w[1][k] * pq;
              st[8 + k] = w[2][k];
            }
            const int base = (pr.packed & 31) * g.sx + ((pr.packed >> 5) & 31) * g.sy + ((pr.packed >> 10) - z_lo);
            reinterpret_cast<int*>(st + 12)[0] = base;
          }
          __syncwarp();
          const int n_now = cnt;
          pos += cnt;
          const bool last = pos >= lim;
          cnt = min(batch, max((last ? next_lim : lim) - pos, 0));
          if (cnt > 0) fetch(pos, cnt);
          // software pipeline: the operands of atom j + 1 are read before atom j is accumulated
          const T* sl = st_rec;
          T wa = sl[la], wb = sl[la + 2], wy = sl[4 + lb], wz = sl[8 + lc];
          int off = reinterpret_cast<const int*>(sl + 12)[0];
          for (int j = 0; j < n_now; ++j) {
            const T va = wy * wz * wa, vb = wy * wz * wb;
            T* p = lane_ptr + off;
            if (j + 1 < n_now) {
              sl += kStageWords;
              wa = sl[la]; wb = sl[la + 2]; wy = sl[4 + lb]; wz = sl[8 + lc];
              off = reinterpret_cast<const int*>(sl + 12)[0];
            }
            const T old_a = p[0], old_b = p[plane2];
            p[0] = old_a + va;
            p[plane2] = old_b + vb;
            __syncwarp();   // the next atom of this warp may touch the same nodes from other lanes
          }
          if (last) break;    // the batch now in flight belongs to the next range
        }
      };
      if (mid > s1) run_range(mid, e2);
      __syncthreads();        // all first chunks are complete
      if (e2 > mid) run_range(e2, e2);
    }
    __syncthreads();
    // flush: two bulk reductions (add) per tile row into the global mesh -- the tile's own z range and
    // the 3 halo elements behind it (+ 1 zero pad: 16 / 32 bytes), which belong to the next z tile or,
    // for the last one, wrap to z = 0; periodic wrap in x, y per row
    fence_proxy_async_smem();
    __syncthreads();
    if (!(debug & 2)) {
      T* dst = mesh + ch * mesh_size;
      const int rows = g.rows_x * g.rows_y;
      const uint32_t main_bytes = (uint32_t)(g.tz * sizeof(T)), halo_bytes = (uint32_t)(4 * sizeof(T));
      const int gx0 = px << g.tx_shift, gy0 = py << g.ty_shift;
      int z_hi = z_lo + g.tz;
      z_hi -= z_hi >= g.nz ? g.nz : 0;
      for (int r = threadIdx.x; r < rows; r += blockDim.x) {
        const int rx = r / g.rows_y, ry = r - rx * g.rows_y;
        int gx = gx0 + rx; gx -= gx >= g.nx ? g.nx : 0;
        int gy = gy0 + ry; gy -= gy >= g.ny ? g.ny : 0;
        T* grow = dst + ((int64_t)gx * g.ny + gy) * g.nz;
        const T* srow = tile + rx * g.sx + ry * g.sy;
        if (z_hi != 0) {        // the halo continues the row in the mesh as it does in the tile: one reduction
          bulk_reduce_add<T>(grow + z_lo, srow, main_bytes + halo_bytes);
        } else {                // last z tile: the halo wraps to z = 0
          bulk_reduce_add<T>(grow + z_lo, srow, main_bytes);
          bulk_reduce_add<T>(grow, srow + g.tz, halo_bytes);
        }
      }
      bulk_commit();
      bulk_wait_read_all();   // the tile is re-zeroed (next channel) / freed (exit) only after the reads
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// gather, 4 nodes.  MODE bits as in interp.cu: 1 values, 2 dvalues/dr, 4 vjp (+ grad_r2u).
// ---------------------------------------------------------------------------------------
template <typename T, int METHOD, int MODE>
__global__ void __launch_bounds__(512, 1)
tile_gather4_kernel(const T* __restrict__ mesh, const T* __restrict__ sorted_rec,
                    const int* __restrict__ sorted_idx, const int* __restrict__ bin_start,
                    const T* __restrict__ positions, const T* __restrict__ coef, int n_channels,
                    Mat3<T> r2u, TileGeom g, T* __restrict__ values, T* __restrict__ dvalues,
                    T* __restrict__ grad_positions, int accumulate, T* __restrict__ grad_r2u,
                    PointEpilogue<T> epi, int debug) {
  constexpr int N = 4;
  constexpr bool DERIV = (MODE & 6) != 0;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* tile = reinterpret_cast<T*>(smem_raw);
  const int tile_elems = g.rows_x * g.sx;
  uint64_t* bar = reinterpret_cast<uint64_t*>(tile + tile_elems);
  __shared__ T red[9][16];

  pdl_trigger();
  const int pencil = blockIdx.x / g.nzt, zt = blockIdx.x - pencil * g.nzt;
  const int bin0 = pencil * g.nzc + zt * g.nzc_t;
  if (threadIdx.x == 0) mbar_init(bar, 1);      // set-up that overlaps the tail of the kernel before
  pdl_wait();                                   // the mesh comes from the filter kernel before
  const int p_begin = bin_start[bin0], p_end = bin_start[bin0 + g.nzc_t];
  const bool with_r2u = (MODE & 4) && grad_r2u != nullptr;
  if (p_begin == p_end) return;
  const int px = pencil / g.npy, py = pencil - px * g.npy;
  const int z_lo = zt * g.tz;
  int z_hi = z_lo + g.tz;
  z_hi -= z_hi >= g.nz ? g.nz : 0;
  const int64_t mesh_size = (int64_t)g.nx * g.ny * g.nz;
  const int rows = g.rows_x * g.rows_y;
  const uint32_t main_bytes = (uint32_t)(g.tz * sizeof(T)), halo_bytes = (uint32_t)(4 * sizeof(T));
  const bool with_extra = (MODE & 4) && epi.enabled && epi.coef2 != nullptr;

  __syncthreads();                              // the mbarrier (initialised above) is visible to all threads
  T cellsum[9];
#pragma unroll
  for (int e = 0; e < 9; ++e) cellsum[e] = T(0);

  for (int ch = 0; ch < n_channels; ++ch) {
    if (ch > 0) __syncthreads();   // everybody is done reading the previous channel's tile
    if (threadIdx.x == 0 && !(debug & 2)) mbar_expect_tx(bar, (main_bytes + halo_bytes) * (uint32_t)rows);
    __syncthreads();
    if (!(debug & 2)) {
      const T* src = mesh + ch * mesh_size;
      const int gx0 = px << g.tx_shift, gy0 = py << g.ty_shift;
      for (int r = threadIdx.x; r < rows; r += blockDim.x) {
        const int rx = r / g.rows_y, ry = r - rx * g.rows_y;
        int gx = gx0 + rx; gx -= gx >= g.nx ? g.nx : 0;
        int gy = gy0 + ry; gy -= gy >= g.ny ? g.ny : 0;
        const T* grow = src + ((int64_t)gx * g.ny + gy) * g.nz;
        T* srow = tile + rx * g.sx + ry * g.sy;
        if (z_hi != 0) {        // the tile's own z range and the halo behind it are contiguous in the mesh
          bulk_load(srow, grow + z_lo, main_bytes + halo_bytes, bar);
        } else {                // last z tile: the halo wraps to z = 0
          bulk_load(srow, grow + z_lo, main_bytes, bar);
          bulk_load(srow + g.tz, grow, halo_bytes, bar);
        }
      }
    }
    // first atom of this thread: its record travels while the tile is still arriving
    Rec<T> r_next;
    int i_next = 0;
    if (p_begin + (int)threadIdx.x < p_end) {
      r_next = Rec<T>::load(sorted_rec + 4 * (int64_t)(p_begin + threadIdx.x));
      i_next = __ldg(sorted_idx + p_begin + threadIdx.x);
    }
    if (!(debug & 2)) mbar_wait(bar, (uint32_t)(ch & 1));
    for (int j = p_begin + threadIdx.x; j < p_end && !(debug & 1); j += blockDim.x) {
      // the record / index of this thread's NEXT atom are fetched before the current one is gathered
      const Rec<T> r = r_next;
      const int64_t point = i_next;
      if (j + (int)blockDim.x < p_end) {
        r_next = Rec<T>::load(sorted_rec + 4 * (int64_t)(j + blockDim.x));
        i_next = __ldg(sorted_idx + j + blockDim.x);
      }
      const int64_t o = point * n_channels + ch;
      // per-point operands of the epilogues first: their latency hides under the stencil loop
      T cf = T(0), prev = T(0), addc = T(0);
      if (MODE & 4) cf = coef[o];
      if ((MODE & 1) && epi.enabled) { prev = values[o]; addc = epi.add_coef[o]; }
      T w[3][N], dw[3][N];
#pragma unroll
      for (int a = 0; a < 3; ++a) Stencil<METHOD, N>::template eval<T, DERIV>(r.x[a], w[a], dw[a]);
      const int lx = r.packed & 31, ly = (r.packed >> 5) & 31, fz = r.packed >> 10;
      const T* base = tile + lx * g.sx + ly * g.sy + (fz - z_lo);
      T val = T(0), du0 = T(0), du1 = T(0), du2 = T(0);
#pragma unroll
      for (int a = 0; a < N; ++a) {
        T ta0 = T(0), ta1 = T(0), ta2 = T(0);
#pragma unroll
        for (int b = 0; b < N; ++b) {
          const T* row = base + a * g.sx + b * g.sy;
          T t0 = T(0), t1 = T(0);
#pragma unroll
          for (int c = 0; c < N; ++c) {
            const T v = row[c];
            t0 = fma_t(v, w[2][c], t0);
            if (DERIV) t1 = fma_t(v, dw[2][c], t1);
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
      if (MODE & 1) {
        if (epi.enabled)
          values[o] = prev + epi.scale * val - addc * epi.self_half - epi.background * epi.dc[ch];
        else
          values[o] = val;
      }
      if (MODE & 2) {
        T* out = dvalues + o * 3;
#pragma unroll
        for (int b = 0; b < 3; ++b)  // du_a/dr_b = r2u[b][a]
          out[b] = r2u.m[3 * b] * du0 + r2u.m[3 * b + 1] * du1 + r2u.m[3 * b + 2] * du2;
      }
      if (MODE & 4) {
        const T gu[3] = {cf * du0, cf * du1, cf * du2};
        T* out = grad_positions + 3 * point;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          T gb = r2u.m[3 * b] * gu[0] + r2u.m[3 * b + 1] * gu[1] + r2u.m[3 * b + 2] * gu[2];
          if (with_extra) {
            if (ch == 0)   // the saved forward derivative enters once, summed over the channels
              for (int c2 = 0; c2 < n_channels; ++c2)
                gb = fma_t(epi.coef2[point * n_channels + c2], epi.dvalues2[(point * n_channels + c2) * 3 + b], gb);
            gb *= epi.vjp_scale;
          }
          out[b] = (accumulate || ch > 0) ? out[b] + gb : gb;
        }
        if (with_r2u) {
          const T rr[3] = {positions[3 * point], positions[3 * point + 1], positions[3 * point + 2]};
#pragma unroll
          for (int e = 0; e < 9; ++e) cellsum[e] = fma_t(rr[e / 3], gu[e % 3], cellsum[e]);
        }
      }
    }
  }
  if (with_r2u) {   // block-uniform branch
    const int warp = threadIdx.x >> 5, wl = threadIdx.x & 31;
#pragma unroll
    for (int e = 0; e < 9; ++e) {
      T v = cellsum[e];
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

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

static int max_smem_optin() {
  static int cached = 0;
  if (!cached) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&cached, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (cached <= 0) cached = 227 * 1024;
  }
  return cached;
}

template <typename T, int METHOD, int MODE>
static int launch_tile_gather(const tpme_tile_plan* plan, const void* mesh, const void* rec, const int* idx,
                              const int* bin_start, const void* positions, const void* coef, int n_channels,
                              const double* r2u, void* values, void* dvalues, void* grad_positions,
                              int accumulate, void* grad_r2u, const tpme_point_epilogue* epi_host,
                              cudaStream_t stream) {
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
  const TileGeom g = geom_of(*plan, plan->gather_nzt);
  const size_t smem = (size_t)g.rows_x * g.sx * sizeof(T) + 16;
  auto kernel = tile_gather4_kernel<T, METHOD, MODE>;
  static thread_local bool configured = false;
  if (!configured) {   // dynamic limit = opt-in maximum minus the kernel's static shared memory
    cudaFuncAttributes attr;
    TPME_CUDA_OK(cudaFuncGetAttributes(&attr, kernel));
    TPME_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      max_smem_optin() - (int)attr.sharedSizeBytes));
    configured = true;
  }
  TPME_CUDA_OK(launch_pdl(kernel, dim3((unsigned)(g.npx * g.npy * g.nzt)), dim3(plan->gather_threads), smem, stream, pdl_for<T>(),
      (const T*)mesh, (const T*)rec, idx, bin_start, (const T*)positions, (const T*)coef, n_channels,
      load_mat3<T>(r2u), g, (T*)values, (T*)dvalues, (T*)grad_positions, accumulate, (T*)grad_r2u, epi,
      getenv("TPME_TILE_DEBUG") ? atoi(getenv("TPME_TILE_DEBUG")) : 0));
  TPME_CUDA_OK(cudaGetLastError());
  return 0;
}

template <typename T, int MODE>
static int tile_gather_method(int method, const tpme_tile_plan* plan, const void* mesh, const void* rec,
                              const int* idx, const int* bin_start, const void* positions, const void* coef,
                              int n_channels, const double* r2u, void* values, void* dvalues,
                              void* grad_positions, int accumulate, void* grad_r2u,
                              const tpme_point_epilogue* epi, cudaStream_t stream) {
  if (method == TPME_P3M)
    return launch_tile_gather<T, TPME_P3M, MODE>(plan, mesh, rec, idx, bin_start, positions, coef, n_channels,
                                                 r2u, values, dvalues, grad_positions, accumulate, grad_r2u, epi, stream);
  return launch_tile_gather<T, TPME_LAGRANGE, MODE>(plan, mesh, rec, idx, bin_start, positions, coef, n_channels,
                                                    r2u, values, dvalues, grad_positions, accumulate, grad_r2u, epi, stream);
}

}  // namespace tpme

using namespace tpme;

extern "C" int tpme_tile_plan_make(int dtype, int nx, int ny, int nz, int nodes, int method,
                                   int64_t n_points, tpme_tile_plan* plan) {
  TPME_REQUIRE(plan != nullptr, "null plan");
  TPME_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (float32) or 1 (float64)");
  memset(plan, 0, sizeof(*plan));
  const int elem = dtype == 0 ? 4 : 8;
  // what the tiled kernels cover; everything else stays on the direct kernels of interp.cu
  if (nodes != 4 || (method != TPME_P3M && method != TPME_LAGRANGE)) return 3;
  if (!is_pow2(nx) || !is_pow2(ny) || !is_pow2(nz) || nz < 8 || nz > 32768 || nx < 4 || ny < 4) return 3;
  if (n_points <= 0 || n_points >= (1ll << 31)) return 3;
  plan->nx = nx; plan->ny = ny; plan->nz = nz; plan->nodes = nodes;
  // The tiled spread is the faster spread on every measured workload once the pencils are cut into z
  // tiles (profiles/r02_summary.md: c3 52 -> 33 us, c5 38 -> 20 us, c4 95 -> 77 us); TPME_TILE_SPREAD = off
  // keeps the direct kernel (the plan then has one bin per gather tile and serves the gathers only).
  plan->spread_tiled = 1;
  if (const char* env = getenv("TPME_TILE_SPREAD")) {
    if (strcmp(env, "off") == 0) plan->spread_tiled = 0;
  }
  // footprint of a pencil in first stencil nodes.  (8, 8) is the measured optimum of the B200 sweeps
  // (profiles/r02_summary.md); smaller meshes take the largest footprint that still gives every SM a pencil.
  static const int cand[][2] = {{8, 8}, {4, 8}, {4, 4}, {2, 4}, {2, 2}};
  int tx = 0, ty = 0;
  if (const char* env = getenv("TPME_TILE")) {
    int a = 0, b = 0;
    if (sscanf(env, "%d,%d", &a, &b) == 2 && is_pow2(a) && is_pow2(b) && a <= 16 && b <= 16) { tx = a; ty = b; }
  }
  const int smem_cap = max_smem_optin() - 2048;   // static shared memory of the gather + slack
  const int sms = num_sms();
  for (int k = 0; k < 5 && tx == 0; ++k) {
    const int a = cand[k][0], b = cand[k][1];
    if (a > nx || b > ny) continue;
    if ((int64_t)(nx / a) * (ny / b) >= sms || k == 4) { tx = a; ty = b; }
  }
  if (tx > nx || ty > ny) return 3;
  plan->tx = tx; plan->ty = ty;
  plan->npx = nx / tx; plan->npy = ny / ty;
  // z tiles: the pencil is cut along z until a tile (+ the staging areas of the spread) is small enough for
  // several CTAs per SM -- then one CTA accumulates while another one zeroes / flushes / loads its tile.
  // A tile keeps >= 16 first-node planes (>= 4 z chunks).  TPME_TILE_NZT overrides.
  const int rows = (tx + nodes - 1) * (ty + nodes - 1);
  auto stage_bytes_of = [&](int w, int batch) { return (size_t)w * kStageWords * batch * elem; };
  auto tile_bytes = [&](int nzt) {
    const size_t raw = (size_t)(ty + nodes - 1) * (nz / nzt + 4);
    return (size_t)(tx + nodes - 1) * (raw + ((16 - raw % 32) + 32) % 32) * elem + 256;
  };
  auto warps_of = [&](int nzt) {
    int max_warps = dtype == 0 ? 32 : 16;
    if (const char* env = getenv("TPME_TILE_WARPS")) {
      const int v = atoi(env);
      if (v >= 1 && v <= max_warps) max_warps = v;
    }
    int w = (nz / nzt / 4) / 2;                  // every warp needs at least two z chunks of 4 planes
    if (w > max_warps) w = max_warps;
    if (w > 16 && nzt > 1) w = 16;
    const double per_tile = (double)n_points / ((double)plan->npx * plan->npy * nzt);
    while (w > 2 && per_tile < 16.0 * w) w >>= 1;
    return w < 1 ? 1 : w;
  };
  // spread: tiles of 32 first-node planes (8 z chunks): 4 .. 8 tiles per SM; gather: the whole pencil when it
  // fits (its tile load is two bulk copies per row and z tile: fewer, longer copies win there)
  int nzt = nz / 32 > 0 ? nz / 32 : 1;
  if (const char* env = getenv("TPME_TILE_NZT")) {
    const int v = atoi(env);
    if (is_pow2(v) && nz / v >= 8) nzt = v;
  }
  int gather_nzt = 1;
  while (nz / (2 * gather_nzt) >= 8 && (int64_t)tile_bytes(gather_nzt) > smem_cap) gather_nzt *= 2;
  if (const char* env = getenv("TPME_TILE_GATHER_NZT")) {
    const int v = atoi(env);
    if (is_pow2(v) && nz / v >= 8) gather_nzt = v;
  }
  if ((int64_t)tile_bytes(gather_nzt) > smem_cap) return 3;
  plan->nzt = nzt;
  plan->gather_nzt = gather_nzt;
  plan->zw = plan->spread_tiled ? 4 : nz / gather_nzt;   // >= nodes - 1: only adjacent chunks overlap
  plan->nzc = nz / plan->zw;
  plan->n_bins = plan->npx * plan->npy * plan->nzc;
  plan->row_stride = nz / nzt + 4;
  plan->plane_stride = 0;
  int warps = warps_of(nzt);
  while (warps > 1 && (int64_t)(tile_bytes(nzt) + stage_bytes_of(warps, 16)) > smem_cap) warps >>= 1;
  plan->spread_threads = 32 * warps;
  plan->spread_batch = (int64_t)(tile_bytes(nzt) + stage_bytes_of(warps, 32)) <= smem_cap ? 32 : 16;
  plan->smem_bytes = (int)(tile_bytes(nzt) + stage_bytes_of(warps, plan->spread_batch));
  if (plan->smem_bytes > smem_cap) return 3;
  const double per_tile = (double)n_points / ((double)plan->npx * plan->npy * gather_nzt);
  int gt = 64;
  while (gt < 512 && gt < per_tile) gt <<= 1;
  plan->gather_threads = gt;
  return 0;
}

extern "C" int64_t tpme_tile_bin_count_ints(const tpme_tile_plan* plan) {
  if (plan == nullptr || plan->n_bins <= 0) return 0;
  return ((plan->n_bins + 1) & ~1) + 2 * scan_state_words(plan->n_bins);
}

extern "C" int tpme_tile_sort(int dtype, const tpme_tile_plan* plan, const void* positions, int64_t n_points,
                              const double* r2u_host, int* bin_count, int* bin_start,
                              int* key_rank, void* sorted_rec, int* sorted_idx, void* stream) {
  TPME_REQUIRE(plan != nullptr && plan->n_bins > 0, "invalid tile plan");
  TPME_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (float32) or 1 (float64)");
  TPME_REQUIRE(n_points >= 0 && n_points < (1ll << 31), "the tiled kernels hold 32-bit point indices");
  TPME_REQUIRE(((uintptr_t)sorted_rec % 16) == 0 && ((uintptr_t)key_rank % 8) == 0, "workspace alignment");
  cudaStream_t s = (cudaStream_t)stream;
  const TileGeom g = geom_of(*plan, plan->nzt);
  TPME_REQUIRE(((uintptr_t)bin_count % 8) == 0, "workspace alignment");
  // the counters and, behind them (8-byte aligned, tpme_tile_bin_count_ints), the scratch of the scan: one memset
  void* scan_state = bin_count + ((plan->n_bins + 1) & ~1);
  TPME_CUDA_OK(cudaMemsetAsync(bin_count, 0, sizeof(int) * (size_t)tpme_tile_bin_count_ints(plan), s));
  const unsigned grid = (unsigned)((n_points + 255) / 256);
  if (n_points > 0) {
    if (dtype == 0)
      tile_count_kernel<float><<<grid, 256, 0, s>>>((const float*)positions, n_points, load_mat3<float>(r2u_host),
                                                    make_dims<float>(g.nx, g.ny, g.nz), g, bin_count, (int2*)key_rank);
    else
      tile_count_kernel<double><<<grid, 256, 0, s>>>((const double*)positions, n_points, load_mat3<double>(r2u_host),
                                                     make_dims<double>(g.nx, g.ny, g.nz), g, bin_count, (int2*)key_rank);
  }
  TPME_CUDA_OK(launch_exclusive_scan<int>(bin_count, bin_start, plan->n_bins, scan_state, s, true,
                                          dtype == 1 || pdl_everywhere()));
  if (n_points > 0) {
    if (dtype == 0)
      TPME_CUDA_OK(launch_pdl(tile_fill_kernel<float>, dim3(grid), dim3(256), 0, s, pdl_for<float>(), (const float*)positions, n_points,
                              load_mat3<float>(r2u_host), make_dims<float>(g.nx, g.ny, g.nz), g, bin_start,
                              (const int2*)key_rank, (float*)sorted_rec, sorted_idx));
    else
      TPME_CUDA_OK(launch_pdl(tile_fill_kernel<double>, dim3(grid), dim3(256), 0, s, pdl_for<double>(), (const double*)positions, n_points,
                              load_mat3<double>(r2u_host), make_dims<double>(g.nx, g.ny, g.nz), g, bin_start,
                              (const int2*)key_rank, (double*)sorted_rec, sorted_idx));
  }
  TPME_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int tpme_tile_spread(int dtype, const tpme_tile_plan* plan, const void* sorted_rec,
                                const int* sorted_idx, const int* bin_start, const void* weights,
                                int64_t n_points, int n_channels, int method, void* mesh, int accumulate,
                                void* stream) {
  TPME_REQUIRE(plan != nullptr && plan->n_bins > 0 && plan->nodes == 4, "invalid tile plan");
  TPME_REQUIRE(plan->spread_tiled && plan->zw == 4 && plan->nzc / plan->nzt >= 2,
               "this tile plan was made without z chunks (spread_tiled = 0)");
  TPME_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (float32) or 1 (float64)");
  TPME_REQUIRE(method == TPME_P3M || method == TPME_LAGRANGE, "unknown interpolation method");
  TPME_REQUIRE(((uintptr_t)mesh % 16) == 0, "mesh must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t elem = dtype == 0 ? 4 : 8;
  if (!accumulate)
    TPME_CUDA_OK(cudaMemsetAsync(mesh, 0, elem * (size_t)n_channels * plan->nx * plan->ny * plan->nz, s));
  if (n_points == 0 || n_channels == 0) return 0;
  const TileGeom g = geom_of(*plan, plan->nzt);
  const size_t smem = (size_t)g.rows_x * g.sx * elem + 256 +
                      (size_t)(plan->spread_threads / 32) * kStageWords * plan->spread_batch * elem;
  const unsigned grid = (unsigned)(g.npx * g.npy * g.nzt);
  const int dbg = getenv("TPME_TILE_DEBUG") ? atoi(getenv("TPME_TILE_DEBUG")) : 0;
  if (dtype == 0) {
    static thread_local bool configured = false;
    if (!configured) {
      TPME_CUDA_OK(cudaFuncSetAttribute(tile_spread4_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        max_smem_optin()));
      configured = true;
    }
    TPME_CUDA_OK(launch_pdl(tile_spread4_kernel<float>, dim3(grid), dim3(plan->spread_threads), smem, s, pdl_for<float>(),
        (const float*)sorted_rec, sorted_idx, bin_start, (const float*)weights, n_channels, method, g, (float*)mesh, plan->spread_batch, dbg));
  } else {
    static thread_local bool configured = false;
    if (!configured) {
      TPME_CUDA_OK(cudaFuncSetAttribute(tile_spread4_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        max_smem_optin()));
      configured = true;
    }
    TPME_CUDA_OK(launch_pdl(tile_spread4_kernel<double>, dim3(grid), dim3(plan->spread_threads), smem, s, pdl_for<double>(),
        (const double*)sorted_rec, sorted_idx, bin_start, (const double*)weights, n_channels, method, g, (double*)mesh, plan->spread_batch, dbg));
  }
  TPME_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int tpme_tile_gather(int dtype, const tpme_tile_plan* plan, const void* mesh, const void* sorted_rec,
                                const int* sorted_idx, const int* bin_start, const void* positions,
                                const void* coef, int64_t n_points, int n_channels, const double* r2u_host,
                                int method, void* values, void* dvalues, void* grad_positions, int accumulate,
                                void* grad_r2u, const tpme_point_epilogue* epilogue, void* stream) {
  TPME_REQUIRE(plan != nullptr && plan->n_bins > 0 && plan->nodes == 4, "invalid tile plan");
  TPME_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (float32) or 1 (float64)");
  TPME_REQUIRE(method == TPME_P3M || method == TPME_LAGRANGE, "unknown interpolation method");
  TPME_REQUIRE(((uintptr_t)mesh % 16) == 0, "mesh must be 16-byte aligned");
  TPME_REQUIRE(values != nullptr || dvalues != nullptr || grad_positions != nullptr, "nothing to compute");
  TPME_REQUIRE(grad_positions == nullptr || coef != nullptr, "coef missing");
  TPME_REQUIRE(grad_positions == nullptr || dvalues == nullptr, "dvalues and grad_positions are exclusive");
  TPME_REQUIRE(grad_r2u == nullptr || positions != nullptr, "grad_r2u needs the positions");
  if (n_points == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (n_channels == 0) {
    if (grad_positions != nullptr && !accumulate)
      TPME_CUDA_OK(cudaMemsetAsync(grad_positions, 0, (dtype ? 8 : 4) * 3 * (size_t)n_points, s));
    return 0;
  }
  const int mode = (values ? 1 : 0) | (dvalues ? 2 : 0) | (grad_positions ? 4 : 0);
#define GO(T, MODE)                                                                                       \
  return tile_gather_method<T, MODE>(method, plan, mesh, sorted_rec, sorted_idx, bin_start, positions,   \
                                     coef, n_channels, r2u_host, values, dvalues, grad_positions,         \
                                     accumulate, grad_r2u, epilogue, s)
  if (dtype == 0) {
    switch (mode) {
      case 1: GO(float, 1);
      case 2: GO(float, 2);
      case 3: GO(float, 3);
      case 4: GO(float, 4);
      case 5: GO(float, 5);
    }
  } else {
    switch (mode) {
      case 1: GO(double, 1);
      case 2: GO(double, 2);
      case 3: GO(double, 3);
      case 4: GO(double, 4);
      case 5: GO(double, 5);
    }
  }
#undef GO
  set_last_error("tpme_tile_gather", "unsupported output combination");
  return 1;
}
