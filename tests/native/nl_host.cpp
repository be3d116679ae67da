// TEST INFRASTRUCTURE: runs the neighbor-search loop of torch-pme_b200/csrc/neighbors_core.h -- the
// code the CUDA kernels execute per thread -- on the CPU, one "thread" after the other, so that
// tests/test_neighbors.py can check it against the brute-force oracle without a GPU.
//   g++ -O2 -std=c++17 -shared -fPIC -o nl_host.so nl_host.cpp
#include "../../torch-pme_b200/csrc/neighbors_core.h"

using namespace tpme;

static NeighborGeometry geometry(const double* cell, const int* n_bins, const int* reach, const int* periodic,
                                 int full_list, double cutoff) {
  NeighborGeometry g;
  for (int k = 0; k < 9; ++k) g.cell[k] = cell[k];
  for (int a = 0; a < 3; ++a) {
    g.n_bins[a] = n_bins[a];
    g.reach[a] = reach[a];
    g.periodic[a] = periodic[a];
  }
  g.cutoff_sq = cutoff * cutoff;
  g.full_list = full_list;
  return g;
}

template <typename T>
static void count(const T* wrapped, const int* wrap_shift, const int* atom_bins, const int* order,
                  const int* bin_start, int64_t n, const NeighborGeometry& g, int* counts) {
  for (int64_t slot = 0; slot < n; ++slot)
    counts[slot] = neighbor_search_atom<T, false>(slot, wrapped, wrap_shift, atom_bins, order, bin_start, g, 0,
                                                  nullptr, nullptr, nullptr);
}

template <typename T>
static void fill(const T* wrapped, const int* wrap_shift, const int* atom_bins, const int* order,
                 const int* bin_start, int64_t n, const NeighborGeometry& g, const int64_t* offsets,
                 int64_t* indices, T* dist_sq, int* shifts) {
  for (int64_t slot = 0; slot < n; ++slot)
    neighbor_search_atom<T, true>(slot, wrapped, wrap_shift, atom_bins, order, bin_start, g, offsets[slot],
                                  indices, dist_sq, shifts);
}

extern "C" int nl_host_count(int dtype, const void* wrapped, const int* wrap_shift, const int* atom_bins,
                             const int* order, const int* bin_start, int64_t n, const double* cell,
                             const int* n_bins, const int* reach, const int* periodic, int full_list,
                             double cutoff, int* counts) {
  const NeighborGeometry g = geometry(cell, n_bins, reach, periodic, full_list, cutoff);
  if (dtype == 0) count<float>((const float*)wrapped, wrap_shift, atom_bins, order, bin_start, n, g, counts);
  else count<double>((const double*)wrapped, wrap_shift, atom_bins, order, bin_start, n, g, counts);
  return 0;
}

extern "C" int nl_host_fill(int dtype, const void* wrapped, const int* wrap_shift, const int* atom_bins,
                            const int* order, const int* bin_start, int64_t n, const double* cell,
                            const int* n_bins, const int* reach, const int* periodic, int full_list,
                            double cutoff, const int64_t* offsets, int64_t* indices, void* dist_sq,
                            int* shifts) {
  const NeighborGeometry g = geometry(cell, n_bins, reach, periodic, full_list, cutoff);
  if (dtype == 0)
    fill<float>((const float*)wrapped, wrap_shift, atom_bins, order, bin_start, n, g, offsets, indices,
                (float*)dist_sq, shifts);
  else
    fill<double>((const double*)wrapped, wrap_shift, atom_bins, order, bin_start, n, g, offsets, indices,
                 (double*)dist_sq, shifts);
  return 0;
}
