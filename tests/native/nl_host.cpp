// TEST INFRASTRUCTURE: runs the per-atom pieces of torch-pme_b200/csrc/neighbors_core.h -- the code the
// CUDA kernels of neighbors.cu execute per thread -- on the CPU, one "thread" after the other (wrap + bin,
// counting sort, search, write-out), so that tests/test_neighbors.py can check them against the brute-force
// oracle without a GPU.
//   g++ -O2 -std=c++17 -shared -fPIC -o nl_host.so nl_host.cpp
#include <vector>

#include "../../torch-pme_b200/csrc/neighbors_core.h"

using namespace tpme;

static bool geometry(const double* cell, const int* n_bins, const int* reach, const int* periodic,
                     int full_list, double cutoff, NeighborGeometry* g) {
  for (int k = 0; k < 9; ++k) g->cell[k] = cell[k];
  if (!invert3(g->cell, g->inv_cell)) return false;
  for (int a = 0; a < 3; ++a) {
    g->n_bins[a] = n_bins[a];
    g->reach[a] = reach[a];
    g->periodic[a] = periodic[a];
  }
  g->cutoff_sq = cutoff * cutoff;
  g->full_list = full_list;
  return true;
}

template <typename T> struct Sorted {
  std::vector<NlRecord<T>> rec;
  std::vector<NlShift> shift;
  std::vector<int> bin_start;
};

struct Collect {   // stands in for the shared-memory hit list of a CUDA thread
  std::vector<int> s;
  std::vector<unsigned> image;
  void operator()(int slot, unsigned im) { s.push_back(slot); image.push_back(im); }
};

template <typename T>
static void sort_atoms(const T* pos, int64_t n, const NeighborGeometry& g, Sorted<T>& s) {
  const int64_t bins = (int64_t)g.n_bins[0] * g.n_bins[1] * g.n_bins[2];
  std::vector<int> count(bins, 0), rank(n), key(n);
  for (int64_t i = 0; i < n; ++i) {   // nl_bin_kernel
    T w[3];
    int k[3], bin;
    nl_locate<T>(pos + 3 * i, g, w, k, bin);
    key[i] = bin;
    rank[i] = count[bin]++;
  }
  s.bin_start.assign(bins + 1, 0);
  for (int64_t b = 0; b < bins; ++b) s.bin_start[b + 1] = s.bin_start[b] + count[b];
  s.rec.resize(n);
  s.shift.resize(n);
  for (int64_t i = 0; i < n; ++i) {   // nl_place_kernel
    T w[3];
    NlShift sh;
    nl_locate<T>(pos + 3 * i, g, w, sh.k, sh.bin);
    const int slot = s.bin_start[key[i]] + rank[i];
    s.rec[slot].x = w[0]; s.rec[slot].y = w[1]; s.rec[slot].z = w[2];
    s.rec[slot].index = (decltype(s.rec[slot].index))i;
    s.shift[slot] = sh;
  }
}

// nl_pairs_kernel, one "thread" after the other; indices == NULL: count only
template <typename T, typename I>
static int64_t build(const T* pos, int64_t n, const NeighborGeometry& g, int64_t capacity, I* indices,
                     T* distances, int* shifts) {
  Sorted<T> s;
  sort_atoms<T>(pos, n, g, s);
  int64_t total = 0;
  for (int64_t slot = 0; slot < n; ++slot) {
    Collect hits;
    const int found = nl_visit_slot<T>(slot, s.rec.data(), s.shift.data(), s.bin_start.data(), g, hits);
    if (indices != nullptr)
      for (int k = 0; k < found; ++k)
        if (total + k < capacity)
          nl_emit<T, I>(slot, hits.s[k], hits.image[k], s.rec.data(), s.shift.data(), g, total + k, indices, distances, shifts);
    total += found;
  }
  return total;
}

extern "C" int64_t nl_host_build(int dtype, const void* positions, int64_t n, const double* cell, const int* n_bins,
                                 const int* reach, const int* periodic, int full_list, double cutoff,
                                 int64_t capacity, int index_is_int64, void* indices, void* distances,
                                 int* shifts) {
  NeighborGeometry g;
  if (!geometry(cell, n_bins, reach, periodic, full_list, cutoff, &g)) return -1;
  if (dtype == 0) {
    if (index_is_int64) return build<float, int64_t>((const float*)positions, n, g, capacity, (int64_t*)indices, (float*)distances, shifts);
    return build<float, int32_t>((const float*)positions, n, g, capacity, (int32_t*)indices, (float*)distances, shifts);
  }
  if (index_is_int64) return build<double, int64_t>((const double*)positions, n, g, capacity, (int64_t*)indices, (double*)distances, shifts);
  return build<double, int32_t>((const double*)positions, n, g, capacity, (int32_t*)indices, (double*)distances, shifts);
}
