// Cell-list neighbor search, one atom at a time -- shared by the CUDA kernels (neighbors.cu) and by
// the host harness that validates the very same code on a CPU (tests/native/nl_host.cpp).
//
// EXPERIMENTAL (the step *before* the hot path; the reference relies
// on the external `vesin` package, tests/helpers.py:240-275, examples/basic-usage.py:166-169).
//
// Conventions: cell rows are lattice vectors; a pair (i, j, S) means the image r_j + S . cell of
// atom j seen from atom i.  Atoms are binned by their fractional coordinates wrapped into the
// cell (periodic directions); bins are slabs between lattice planes, `n_bins[a]` per direction,
// and `reach[a]` = ceil(cutoff / slab thickness) bins are visited on both sides.  Walking the
// *unwrapped* bin coordinate b + db and splitting it into (wrapped bin, image count) visits every
// (bin, image) combination exactly once, also when the cell is smaller than the cutoff.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TPME_HD __host__ __device__ __forceinline__
#else
#define TPME_HD inline
#endif

namespace tpme {

struct NeighborGeometry {
  double cell[9];        // row-major, rows = lattice vectors
  int n_bins[3];
  int reach[3];
  int periodic[3];
  double cutoff_sq;
  int full_list;
};

TPME_HD int floor_div(int a, int b) {   // b > 0
  int q = a / b;
  return (a % b != 0 && a < 0) ? q - 1 : q;
}

// Visits all neighbors of the atom in sorted slot `slot`.  With FILL == false only counts them.
//   wrapped    (N,3) positions wrapped into the cell, original atom order
//   wrap_shift (N,3) k_i with  wrapped_i = r_i - k_i . cell
//   atom_bins  (N,3) bin coordinates of every atom
//   order      (N)   sorted slot -> atom (atoms sorted by linear bin index)
//   bin_start  (n_bins_total + 1) first sorted slot of every bin
// Half lists keep (i, j, S) with i < j, and self images (i, i, S) with S lexicographically positive.
template <typename T, bool FILL>
TPME_HD int neighbor_search_atom(int64_t slot, const T* wrapped, const int* wrap_shift, const int* atom_bins,
                                 const int* order, const int* bin_start, const NeighborGeometry& g,
                                 int64_t out_offset, int64_t* indices, T* distances_sq, int* shifts) {
  const int i = order[slot];
  const T xi = wrapped[3 * i], yi = wrapped[3 * i + 1], zi = wrapped[3 * i + 2];
  const int bx = atom_bins[3 * i], by = atom_bins[3 * i + 1], bz = atom_bins[3 * i + 2];
  int found = 0;
  for (int dx = -g.reach[0]; dx <= g.reach[0]; ++dx) {
    const int ux = bx + dx;
    if (!g.periodic[0] && (ux < 0 || ux >= g.n_bins[0])) continue;
    const int sx = g.periodic[0] ? floor_div(ux, g.n_bins[0]) : 0;
    const int wx = ux - sx * g.n_bins[0];
    for (int dy = -g.reach[1]; dy <= g.reach[1]; ++dy) {
      const int uy = by + dy;
      if (!g.periodic[1] && (uy < 0 || uy >= g.n_bins[1])) continue;
      const int sy = g.periodic[1] ? floor_div(uy, g.n_bins[1]) : 0;
      const int wy = uy - sy * g.n_bins[1];
      for (int dz = -g.reach[2]; dz <= g.reach[2]; ++dz) {
        const int uz = bz + dz;
        if (!g.periodic[2] && (uz < 0 || uz >= g.n_bins[2])) continue;
        const int sz = g.periodic[2] ? floor_div(uz, g.n_bins[2]) : 0;
        const int wz = uz - sz * g.n_bins[2];
        // image translation S . cell
        const T tx = (T)(sx * g.cell[0] + sy * g.cell[3] + sz * g.cell[6]);
        const T ty = (T)(sx * g.cell[1] + sy * g.cell[4] + sz * g.cell[7]);
        const T tz = (T)(sx * g.cell[2] + sy * g.cell[5] + sz * g.cell[8]);
        const bool zero_shift = (sx == 0 && sy == 0 && sz == 0);
        const bool positive_shift = sx > 0 || (sx == 0 && (sy > 0 || (sy == 0 && sz > 0)));
        const int bin = (wx * g.n_bins[1] + wy) * g.n_bins[2] + wz;
        for (int s = bin_start[bin]; s < bin_start[bin + 1]; ++s) {
          const int j = order[s];
          if (i == j) {
            if (zero_shift) continue;
            if (!g.full_list && !positive_shift) continue;
          } else if (!g.full_list && i > j) {
            continue;
          }
          const T ddx = wrapped[3 * j] + tx - xi;
          const T ddy = wrapped[3 * j + 1] + ty - yi;
          const T ddz = wrapped[3 * j + 2] + tz - zi;
          const T r2 = ddx * ddx + ddy * ddy + ddz * ddz;
          if (!(r2 < (T)g.cutoff_sq)) continue;
          if (FILL) {
            const int64_t o = out_offset + found;
            indices[2 * o] = i;
            indices[2 * o + 1] = j;
            distances_sq[o] = r2;
            // shift with respect to the original (unwrapped) positions
            shifts[3 * o] = sx + wrap_shift[3 * i] - wrap_shift[3 * j];
            shifts[3 * o + 1] = sy + wrap_shift[3 * i + 1] - wrap_shift[3 * j + 1];
            shifts[3 * o + 2] = sz + wrap_shift[3 * i + 2] - wrap_shift[3 * j + 2];
          }
          ++found;
        }
      }
    }
  }
  return found;
}

}  // namespace tpme
