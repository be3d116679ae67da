// EXPERIMENTAL: cell-list neighbor list on the GPU (SURVEY.md section 8f rank 1 -- the step before
// the hot path; the reference delegates it to the external `vesin` package).  Two passes over the
// spatially sorted atoms with the search loop of neighbors_core.h: count, (exclusive scan by the
// caller), fill.  Binning / sorting of the atoms is cheap plumbing done by the caller.
#include "common.cuh"
#include "neighbors_core.h"
#include "../../include/torchpme_b200.h"

namespace tpme {

template <typename T>
__global__ void __launch_bounds__(128)
neighbor_count_kernel(const T* __restrict__ wrapped, const int* __restrict__ wrap_shift,
                      const int* __restrict__ atom_bins, const int* __restrict__ order,
                      const int* __restrict__ bin_start, int64_t n, NeighborGeometry g,
                      int* __restrict__ counts) {
  const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n) return;
  counts[slot] = neighbor_search_atom<T, false>(slot, wrapped, wrap_shift, atom_bins, order, bin_start, g, 0,
                                                nullptr, nullptr, nullptr);
}

template <typename T>
__global__ void __launch_bounds__(128)
neighbor_fill_kernel(const T* __restrict__ wrapped, const int* __restrict__ wrap_shift,
                     const int* __restrict__ atom_bins, const int* __restrict__ order,
                     const int* __restrict__ bin_start, int64_t n, NeighborGeometry g,
                     const int64_t* __restrict__ offsets, int64_t* __restrict__ indices,
                     T* __restrict__ distances_sq, int* __restrict__ shifts) {
  const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n) return;
  neighbor_search_atom<T, true>(slot, wrapped, wrap_shift, atom_bins, order, bin_start, g, offsets[slot],
                                indices, distances_sq, shifts);
}

static int make_geometry(const tpme_neighbor_search* s, NeighborGeometry* g) {
  TPME_REQUIRE(s != nullptr, "search parameters missing");
  TPME_REQUIRE(s->cutoff > 0, "cutoff must be positive");
  for (int a = 0; a < 3; ++a) {
    TPME_REQUIRE(s->n_bins[a] >= 1 && s->reach[a] >= 0, "bad bin layout");
    g->n_bins[a] = s->n_bins[a];
    g->reach[a] = s->reach[a];
    g->periodic[a] = s->periodic[a] != 0;
  }
  for (int k = 0; k < 9; ++k) g->cell[k] = s->cell[k];
  g->cutoff_sq = s->cutoff * s->cutoff;
  g->full_list = s->full_list != 0;
  return 0;
}

}  // namespace tpme

using namespace tpme;

extern "C" int tpme_neighbor_count(int dtype, const void* wrapped, const int* wrap_shift,
                                   const int* atom_bins, const int* order, const int* bin_start,
                                   int64_t n_atoms, const tpme_neighbor_search* search, int* counts,
                                   void* stream) {
  TPME_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 or 1");
  NeighborGeometry g;
  if (int rc = make_geometry(search, &g)) return rc;
  if (n_atoms == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((n_atoms + 127) / 128);
  if (dtype == 0)
    neighbor_count_kernel<float><<<grid, 128, 0, s>>>((const float*)wrapped, wrap_shift, atom_bins, order,
                                                      bin_start, n_atoms, g, counts);
  else
    neighbor_count_kernel<double><<<grid, 128, 0, s>>>((const double*)wrapped, wrap_shift, atom_bins, order,
                                                       bin_start, n_atoms, g, counts);
  TPME_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int tpme_neighbor_fill(int dtype, const void* wrapped, const int* wrap_shift,
                                  const int* atom_bins, const int* order, const int* bin_start,
                                  int64_t n_atoms, const tpme_neighbor_search* search,
                                  const int64_t* offsets, int64_t* indices, void* distances_sq,
                                  int* shifts, void* stream) {
  TPME_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 or 1");
  NeighborGeometry g;
  if (int rc = make_geometry(search, &g)) return rc;
  if (n_atoms == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((n_atoms + 127) / 128);
  if (dtype == 0)
    neighbor_fill_kernel<float><<<grid, 128, 0, s>>>((const float*)wrapped, wrap_shift, atom_bins, order,
                                                     bin_start, n_atoms, g, offsets, indices,
                                                     (float*)distances_sq, shifts);
  else
    neighbor_fill_kernel<double><<<grid, 128, 0, s>>>((const double*)wrapped, wrap_shift, atom_bins, order,
                                                      bin_start, n_atoms, g, offsets, indices,
                                                      (double*)distances_sq, shifts);
  TPME_CUDA_OK(cudaGetLastError());
  return 0;
}
