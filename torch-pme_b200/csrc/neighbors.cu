// Cell-list neighbor list on the GPU (SURVEY.md section 8f rank 1 -- the step before the hot path;
// the reference delegates it to the external `vesin` package, tests/helpers.py:240-275) and the
// differentiable pair distances that connect its output to the calculators.
//
//   tpme_nl_sort    wrap + bin the atoms, counting sort by bin (count -> single-pass scan -> place)
//   tpme_nl_pairs   one thread per sorted atom walks the neighbouring bins ONCE: the partners it finds
//                   are parked in shared memory, every warp reserves a contiguous range of the output with
//                   one atomic on the pair counter and then writes (i, j), |r_ij| and the image shift S of
//                   its pairs, neighbouring lanes storing neighbouring pairs.  The pair count stays on the device (no host
//                   synchronisation is needed to go on); pairs beyond `capacity` are dropped (the
//                   caller compares the count with the capacity) -- so a step with a fixed-capacity
//                   list is one CUDA graph.  Half lists visit half of the bins (neighbors_core.h).
//   tpme_pair_distances / _backward   d_p = |r_j + S_p . cell - r_i| and its vector-Jacobian
//                   product with respect to positions (and the cell)
// The sorted records are 16 / 32 bytes (wrapped position + atom index): neighbouring threads sit in
// the same bin and read the same candidates, so the candidate loads are L1 broadcasts.  The per-atom
// pieces live in neighbors_core.h and are run on the CPU by tests/test_neighbors.py.
#include "common.cuh"
#include "neighbors_core.h"
#include "scan.cuh"
#include "../../include/torchpme_b200.h"

namespace tpme {

template <typename T>
__global__ void __launch_bounds__(256)
nl_bin_kernel(const T* __restrict__ positions, int64_t n, NeighborGeometry g, int* __restrict__ bin_count,
              int2* __restrict__ key_rank) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  T w[3];
  int k[3], bin;
  nl_locate<T>(positions + 3 * i, g, w, k, bin);
  key_rank[i] = make_int2(bin, atomicAdd(bin_count + bin, 1));
}

template <typename T>
__global__ void __launch_bounds__(256)
nl_place_kernel(const T* __restrict__ positions, int64_t n, NeighborGeometry g, const int* __restrict__ bin_start,
                const int2* __restrict__ key_rank, NlRecord<T>* __restrict__ sorted, NlShift* __restrict__ sshift) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  T w[3];
  NlShift sh;
  nl_locate<T>(positions + 3 * i, g, w, sh.k, sh.bin);
  const int2 kr = key_rank[i];
  const int slot = bin_start[kr.x] + kr.y;
  NlRecord<T> rec;
  rec.x = w[0]; rec.y = w[1]; rec.z = w[2];
  rec.index = (decltype(rec.index))i;
  sorted[slot] = rec;
  sshift[slot] = sh;
}

constexpr int kNlThreads = 128;
constexpr int kNlHits = 32;        // partners parked per thread; atoms with more search a second time

struct HitRecorder {               // hit list of one thread: hits[k][thread], conflict-free
  uint2* mine;
  int kept;
  __device__ __forceinline__ void operator()(int s, unsigned image) {
    if (kept < kNlHits) mine[kept * kNlThreads] = make_uint2((unsigned)s, image);
    ++kept;
  }
};
struct HitCounter {
  __device__ __forceinline__ void operator()(int, unsigned) {}
};
template <typename T, typename I>
struct HitWriter {                 // second search of a crowded atom: writes the partners number >= kNlHits
  int64_t slot, base, capacity;
  const NlRecord<T>* sorted;
  const NlShift* sshift;
  const NeighborGeometry* g;
  I* indices; T* distances; int* shifts;
  int ordinal;
  __device__ __forceinline__ void operator()(int s, unsigned image) {
    const int64_t o = base + ordinal;
    if (ordinal >= kNlHits && o < capacity) nl_emit<T, I>(slot, s, image, sorted, sshift, *g, o, indices, distances, shifts);
    ++ordinal;
  }
};

template <typename T, typename I, bool FILL>
__global__ void __launch_bounds__(kNlThreads, 6)
nl_pairs_kernel(const NlRecord<T>* __restrict__ sorted, const NlShift* __restrict__ sshift,
                const int* __restrict__ bin_start, int64_t n, NeighborGeometry g, int64_t capacity,
                I* __restrict__ indices, T* __restrict__ distances, int* __restrict__ shifts,
                unsigned long long* __restrict__ n_pairs) {
  __shared__ uint2 hits[FILL ? kNlHits * kNlThreads : 1];
  const int64_t slot = (int64_t)blockIdx.x * kNlThreads + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int found = 0;
  if (slot < n) {
    if (FILL) {
      HitRecorder rec{hits + threadIdx.x, 0};
      found = nl_visit_slot<T>(slot, sorted, sshift, bin_start, g, rec);
    } else {
      HitCounter cnt;
      found = nl_visit_slot<T>(slot, sorted, sshift, bin_start, g, cnt);
    }
  }
  // Every WARP reserves its own range of the output (prefix sums of the counts by shuffles, one atomic on the
  // pair counter): no CTA barrier, so a warp whose atoms had few partners does not wait for the others.
  int incl = found;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += y;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  if (total == 0) return;
  unsigned long long reserved = 0;
  if (lane == 0) reserved = atomicAdd(n_pairs, (unsigned long long)total);
  const int64_t base = (int64_t)__shfl_sync(0xffffffffu, reserved, 0);
  if (!FILL) return;
  const int local = incl - found;                      // first pair of this lane inside the warp's range
  __syncwarp();                                        // the hit lists of the warp are complete
  // Write-out, coalesced: pair p of the warp's range is written by lane p mod 32 -- its owner (the lane that
  // found it) is looked up in the prefix sums, its partner in the owner's hit list -- so that neighbouring
  // lanes store neighbouring pairs (one thread writing its own run would touch 32 sectors per store).
  const int64_t slot0 = slot - lane;
  const int tid0 = threadIdx.x - lane;
  for (int p0 = 0; p0 < total; p0 += 32) {
    const int p = p0 + lane;
    int lo = 0, hi = 31;                               // last lane whose first pair is <= p
#pragma unroll
    for (int step = 0; step < 5; ++step) {
      const int mid = (lo + hi + 1) >> 1;
      const int first_mid = __shfl_sync(0xffffffffu, local, mid);
      if (first_mid <= p) lo = mid; else hi = mid - 1;
    }
    const int k = p - __shfl_sync(0xffffffffu, local, lo);
    if (p < total && k < kNlHits && base + p < capacity) {   // beyond the parked ones: written by the owner below
      const uint2 h = hits[k * kNlThreads + tid0 + lo];
      nl_emit<T, I>(slot0 + lo, (int)h.x, h.y, sorted, sshift, g, base + p, indices, distances, shifts);
    }
  }
  if (found > kNlHits) {                                // crowded atom: search again for the rest
    HitWriter<T, I> w{slot, base + local, capacity, sorted, sshift, &g, indices, distances, shifts, 0};
    nl_visit_slot<T>(slot, sorted, sshift, bin_start, g, w);
  }
}

// ---- differentiable pair distances ------------------------------------------------------------
template <typename I>
__device__ __forceinline__ void load_pair(const I* idx, int64_t p, int64_t& i, int64_t& j);
template <>
__device__ __forceinline__ void load_pair<int64_t>(const int64_t* idx, int64_t p, int64_t& i, int64_t& j) {
  const longlong2 v = *reinterpret_cast<const longlong2*>(idx + 2 * p);
  i = v.x; j = v.y;
}
template <>
__device__ __forceinline__ void load_pair<int32_t>(const int32_t* idx, int64_t p, int64_t& i, int64_t& j) {
  const int2 v = *reinterpret_cast<const int2*>(idx + 2 * p);
  i = v.x; j = v.y;
}

template <typename T, typename I>
__device__ __forceinline__ void pair_vector(const T* __restrict__ pos, const I* __restrict__ idx,
                                            const int* __restrict__ shifts, const Mat3<T>& cell, int64_t p,
                                            int64_t& i, int64_t& j, T (&s)[3], T (&v)[3]) {
  load_pair<I>(idx, p, i, j);
  s[0] = (T)shifts[3 * p]; s[1] = (T)shifts[3 * p + 1]; s[2] = (T)shifts[3 * p + 2];
#pragma unroll
  for (int c = 0; c < 3; ++c)
    v[c] = (__ldg(pos + 3 * j + c) - __ldg(pos + 3 * i + c)) + (s[0] * cell.m[c] + s[1] * cell.m[3 + c] + s[2] * cell.m[6 + c]);
}

template <typename T, typename I>
__global__ void __launch_bounds__(256)
pair_distance_kernel(const T* __restrict__ pos, const I* __restrict__ idx, const int* __restrict__ shifts,
                     Mat3<T> cell, int64_t n_pairs, const int64_t* __restrict__ n_pairs_dev, T* __restrict__ out) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs || (n_pairs_dev != nullptr && p >= *n_pairs_dev)) return;
  int64_t i, j;
  T s[3], v[3];
  pair_vector<T, I>(pos, idx, shifts, cell, p, i, j, s, v);
  out[p] = nl_sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
}

// one reduction per atom and pair: the three components travel as ONE 16-byte vector reduction into a
// gradient buffer padded to 4 reals per atom (fp32: red.global.add.v4.f32, sm_90+); fp64 has no vector
// form and issues three scalar reductions into the same 32-byte sector
__device__ __forceinline__ void red_add3(float* row4, float x, float y, float z) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(row4), "f"(x), "f"(y), "f"(z), "f"(0.0f) : "memory");
}
__device__ __forceinline__ void red_add3(double* row4, double x, double y, double z) {
  red_add(row4, x); red_add(row4 + 1, y); red_add(row4 + 2, z);
}

// grad_pos4[j] += g_p v_p / d_p, grad_pos4[i] -= ..., grad_cell[a][c] += g_p S_a v_c / d_p
template <typename T, typename I>
__global__ void __launch_bounds__(256)
pair_distance_backward_kernel(const T* __restrict__ pos, const I* __restrict__ idx, const int* __restrict__ shifts,
                              Mat3<T> cell, const T* __restrict__ grad_d, int64_t n_pairs,
                              const int64_t* __restrict__ n_pairs_dev, T* __restrict__ grad_pos4,
                              T* __restrict__ grad_cell) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = p < n_pairs && (n_pairs_dev == nullptr || p < *n_pairs_dev);
  T gc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) gc[k] = T(0);
  if (active) {
    int64_t i, j;
    T s[3], v[3];
    pair_vector<T, I>(pos, idx, shifts, cell, p, i, j, s, v);
    const T d2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const T scale = d2 > T(0) ? grad_d[p] / nl_sqrt(d2) : T(0);
    const T f0 = scale * v[0], f1 = scale * v[1], f2 = scale * v[2];
    red_add3(grad_pos4 + 4 * j, f0, f1, f2);
    red_add3(grad_pos4 + 4 * i, -f0, -f1, -f2);
#pragma unroll
    for (int a = 0; a < 3; ++a) { gc[3 * a] = s[a] * f0; gc[3 * a + 1] = s[a] * f1; gc[3 * a + 2] = s[a] * f2; }
  }
  if (grad_cell != nullptr) {   // block-uniform
    __shared__ T red[9][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      T x = gc[k];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
      if (lane == 0) red[k][warp] = x;
    }
    __syncthreads();
    if (threadIdx.x < 9) {
      T x = T(0);
      for (int w = 0; w < 8; ++w) x += red[threadIdx.x][w];
      if (x != T(0)) red_add(grad_cell + threadIdx.x, x);
    }
  }
}

static int make_geometry(const tpme_neighbor_search* s, NeighborGeometry* g) {
  TPME_REQUIRE(s != nullptr, "search parameters missing");
  TPME_REQUIRE(s->cutoff > 0, "cutoff must be positive");
  int64_t total = 1;
  for (int a = 0; a < 3; ++a) {
    TPME_REQUIRE(s->n_bins[a] >= 1 && s->reach[a] >= 0, "bad bin layout");
    TPME_REQUIRE(s->periodic[a] || s->n_bins[a] == 1, "non-periodic directions take one bin");
    TPME_REQUIRE(s->reach[a] / s->n_bins[a] + 2 <= 127, "cutoff spans more than 125 images of the cell");
    g->n_bins[a] = s->n_bins[a];
    g->reach[a] = s->reach[a];
    g->periodic[a] = s->periodic[a] != 0;
    total *= s->n_bins[a];
  }
  TPME_REQUIRE(total < (1ll << 30), "too many bins");
  for (int k = 0; k < 9; ++k) g->cell[k] = s->cell[k];
  TPME_REQUIRE(invert3(g->cell, g->inv_cell), "singular cell");
  g->cutoff_sq = s->cutoff * s->cutoff;
  g->full_list = s->full_list != 0;
  return 0;
}

static int64_t total_bins(const tpme_neighbor_search* s) {
  return (int64_t)s->n_bins[0] * s->n_bins[1] * s->n_bins[2];
}

}  // namespace tpme

using namespace tpme;

extern "C" int64_t tpme_nl_scratch_ints(int64_t n_atoms, const tpme_neighbor_search* search) {
  if (search == nullptr || n_atoms < 0) return 0;
  const int64_t bins = total_bins(search);
  return ((bins + 3) & ~3ll) + 2 * ((scan_state_words(bins) + 1) & ~1ll) + 2 * n_atoms + 4;
}

extern "C" int tpme_nl_sort(int dtype, const void* positions, int64_t n_atoms,
                            const tpme_neighbor_search* search, int* scratch, int* bin_start,
                            void* sorted_rec, int* sorted_shift, void* stream) {
  TPME_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 or 1");
  NeighborGeometry g;
  if (int rc = make_geometry(search, &g)) return rc;
  TPME_REQUIRE(n_atoms >= 0 && n_atoms < (1ll << 31), "the neighbor list holds 32-bit atom slots");
  TPME_REQUIRE(((uintptr_t)scratch % 16) == 0 && ((uintptr_t)sorted_rec % 16) == 0 && ((uintptr_t)sorted_shift % 16) == 0,
               "workspace alignment");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t bins = total_bins(search);
  int* bin_count = scratch;
  void* scan_state = scratch + ((bins + 3) & ~3ll);
  int2* key_rank = reinterpret_cast<int2*>(scratch + ((bins + 3) & ~3ll) + 2 * ((scan_state_words(bins) + 1) & ~1ll));
  // counters + scan scratch behind them: one memset
  TPME_CUDA_OK(cudaMemsetAsync(bin_count, 0, sizeof(int) * (size_t)(((bins + 3) & ~3ll) + 2 * ((scan_state_words(bins) + 1) & ~1ll)), s));
  const unsigned grid = (unsigned)((n_atoms + 255) / 256);
  if (n_atoms > 0) {
    if (dtype == 0) nl_bin_kernel<float><<<grid, 256, 0, s>>>((const float*)positions, n_atoms, g, bin_count, key_rank);
    else nl_bin_kernel<double><<<grid, 256, 0, s>>>((const double*)positions, n_atoms, g, bin_count, key_rank);
  }
  TPME_CUDA_OK(launch_exclusive_scan<int>(bin_count, bin_start, bins, scan_state, s, true));
  if (n_atoms > 0) {
    if (dtype == 0)
      nl_place_kernel<float><<<grid, 256, 0, s>>>((const float*)positions, n_atoms, g, bin_start, key_rank,
                                                  (NlRecord<float>*)sorted_rec, (NlShift*)sorted_shift);
    else
      nl_place_kernel<double><<<grid, 256, 0, s>>>((const double*)positions, n_atoms, g, bin_start, key_rank,
                                                   (NlRecord<double>*)sorted_rec, (NlShift*)sorted_shift);
  }
  TPME_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int tpme_nl_pairs(int dtype, const void* sorted_rec, const int* sorted_shift, const int* bin_start,
                             int64_t n_atoms, const tpme_neighbor_search* search, int64_t capacity,
                             int index_is_int64, void* indices, void* distances, int* shifts,
                             int64_t* n_pairs_dev, void* stream) {
  TPME_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 or 1");
  TPME_REQUIRE(n_pairs_dev != nullptr, "pair counter missing");
  NeighborGeometry g;
  if (int rc = make_geometry(search, &g)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  TPME_CUDA_OK(cudaMemsetAsync(n_pairs_dev, 0, sizeof(int64_t), s));
  if (n_atoms == 0) return 0;
  const unsigned grid = (unsigned)((n_atoms + kNlThreads - 1) / kNlThreads);
  unsigned long long* counter = reinterpret_cast<unsigned long long*>(n_pairs_dev);
#define GO(T, I, FILL)                                                                                       \
  nl_pairs_kernel<T, I, FILL><<<grid, kNlThreads, 0, s>>>((const NlRecord<T>*)sorted_rec, (const NlShift*)sorted_shift, \
                                                          bin_start, n_atoms, g, capacity, (I*)indices,      \
                                                          (T*)distances, shifts, counter)
  if (indices == nullptr) {            // count only
    if (dtype == 0) GO(float, int32_t, false); else GO(double, int32_t, false);
  } else {
    TPME_REQUIRE(distances != nullptr && shifts != nullptr && capacity >= 0, "output buffers missing");
    if (dtype == 0) { if (index_is_int64) GO(float, int64_t, true); else GO(float, int32_t, true); }
    else            { if (index_is_int64) GO(double, int64_t, true); else GO(double, int32_t, true); }
  }
#undef GO
  TPME_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int tpme_pair_distances(int dtype, const void* positions, const double* cell_host,
                                   const void* neighbor_indices, int index_is_int64, const int* shifts,
                                   int64_t n_pairs, const int64_t* n_pairs_dev, void* distances, void* stream) {
  TPME_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 or 1");
  TPME_REQUIRE(cell_host != nullptr, "cell missing");
  if (n_pairs == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((n_pairs + 255) / 256);
#define GO(T, I)                                                                                            \
  pair_distance_kernel<T, I><<<grid, 256, 0, s>>>((const T*)positions, (const I*)neighbor_indices, shifts, \
                                                  load_mat3<T>(cell_host), n_pairs, n_pairs_dev, (T*)distances)
  if (dtype == 0) { if (index_is_int64) GO(float, int64_t); else GO(float, int32_t); }
  else            { if (index_is_int64) GO(double, int64_t); else GO(double, int32_t); }
#undef GO
  TPME_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int tpme_pair_distances_backward(int dtype, const void* positions, const double* cell_host,
                                            const void* neighbor_indices, int index_is_int64, const int* shifts,
                                            const void* grad_distances, int64_t n_pairs,
                                            const int64_t* n_pairs_dev, void* grad_positions, void* grad_cell,
                                            void* stream) {
  TPME_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 or 1");
  TPME_REQUIRE(cell_host != nullptr && grad_positions != nullptr, "cell / grad_positions missing");
  TPME_REQUIRE(((uintptr_t)grad_positions % 16) == 0, "grad_positions (N,4) must be 16-byte aligned");
  if (n_pairs == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((n_pairs + 255) / 256);
#define GO(T, I)                                                                                      \
  pair_distance_backward_kernel<T, I><<<grid, 256, 0, s>>>((const T*)positions, (const I*)neighbor_indices, \
      shifts, load_mat3<T>(cell_host), (const T*)grad_distances, n_pairs, n_pairs_dev, (T*)grad_positions,   \
      (T*)grad_cell)
  if (dtype == 0) { if (index_is_int64) GO(float, int64_t); else GO(float, int32_t); }
  else            { if (index_is_int64) GO(double, int64_t); else GO(double, int32_t); }
#undef GO
  TPME_CUDA_OK(cudaGetLastError());
  return 0;
}
