// Cell-list neighbor search, the per-atom pieces -- shared by the CUDA kernels (neighbors.cu) and by
// the host harness that runs the very same code on a CPU (tests/native/nl_host.cpp).
//
// the step *before* the hot path; the reference relies on the external
// `vesin` package (tests/helpers.py:240-275, examples/basic-usage.py:166-169).
//
// Conventions: cell rows are lattice vectors; a pair (i, j, S) means the image r_j + S . cell of
// atom j seen from atom i.  Atoms are wrapped into the cell along the periodic directions
// (wrapped = r - k . cell, k integer) and binned by their fractional coordinates; bins are slabs
// between lattice planes, `n_bins[a]` per direction, and `reach[a]` = ceil(cutoff / slab thickness)
// bins are visited on both sides.  The atoms are sorted by linear bin index ((bx * ny + by) * nz + bz),
// so the z bins of one (x, y) column are one contiguous run of sorted records.  Walking the
// *unwrapped* bin coordinate b + db and splitting it into (wrapped bin, image count) visits every
// (bin, image) combination exactly once, also when the cell is smaller than the cutoff.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define TPME_HD __host__ __device__ __forceinline__
#else
#define TPME_HD inline
#endif

namespace tpme {

struct NeighborGeometry {
  double cell[9];        // row-major, rows = lattice vectors
  double inv_cell[9];    // cell^-1: fractional = r (row vector) . inv_cell
  int n_bins[3];
  int reach[3];
  int periodic[3];
  double cutoff_sq;
  int full_list;
};

// sorted per-atom record: wrapped position + original atom index (one 16 / 32 byte load per candidate)
template <typename T> struct NlRecord;
template <> struct alignas(16) NlRecord<float> { float x, y, z; int32_t index; };
template <> struct alignas(16) NlRecord<double> { double x, y, z; int64_t index; };
// wrap shift k (wrapped = r - k . cell) and linear bin of the atom in the same sorted slot
struct alignas(16) NlShift { int k[3]; int bin; };

TPME_HD float nl_sqrt(float x) { return sqrtf(x); }
TPME_HD double nl_sqrt(double x) { return sqrt(x); }

TPME_HD int floor_div(int a, int b) {   // b > 0
  int q = a / b;
  return (a % b != 0 && a < 0) ? q - 1 : q;
}

inline bool invert3(const double* m, double* inv) {
  const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
  const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
  if (det == 0.0) return false;
  const double r = 1.0 / det;
  inv[0] = c00 * r; inv[1] = (m[2] * m[7] - m[1] * m[8]) * r; inv[2] = (m[1] * m[5] - m[2] * m[4]) * r;
  inv[3] = c01 * r; inv[4] = (m[0] * m[8] - m[2] * m[6]) * r; inv[5] = (m[2] * m[3] - m[0] * m[5]) * r;
  inv[6] = c02 * r; inv[7] = (m[1] * m[6] - m[0] * m[7]) * r; inv[8] = (m[0] * m[4] - m[1] * m[3]) * r;
  return true;
}

// wrap one atom into the cell and bin it (fp64 arithmetic for either position type)
template <typename T>
TPME_HD void nl_locate(const T* pos, const NeighborGeometry& g, T (&wrapped)[3], int (&k)[3], int& bin) {
  const double r[3] = {(double)pos[0], (double)pos[1], (double)pos[2]};
  double w[3] = {r[0], r[1], r[2]};
  int b[3];
  for (int a = 0; a < 3; ++a) {
    if (!g.periodic[a]) { k[a] = 0; b[a] = 0; continue; }
    const double f = r[0] * g.inv_cell[a] + r[1] * g.inv_cell[3 + a] + r[2] * g.inv_cell[6 + a];
    const double fl = floor(f);
    k[a] = (int)fl;
    double fw = f - fl;
    fw = fw < 0.0 ? 0.0 : (fw > 1.0 ? 1.0 : fw);
    int bb = (int)(fw * g.n_bins[a]);
    b[a] = bb < g.n_bins[a] - 1 ? bb : g.n_bins[a] - 1;
    for (int c = 0; c < 3; ++c) w[c] -= fl * g.cell[3 * a + c];
  }
  for (int c = 0; c < 3; ++c) wrapped[c] = (T)w[c];
  bin = (b[0] * g.n_bins[1] + b[1]) * g.n_bins[2] + b[2];
}

// image count of the unwrapped bin coordinate u (n bins per period): floor(u / n) without a division
// for the usual case of a cell larger than the reach
TPME_HD int image_of(int u, int n) {
  if (u >= 0 && u < n) return 0;
  if (u < 0 && u >= -n) return -1;
  if (u >= n && u < 2 * n) return 1;
  return floor_div(u, n);
}

// image shifts travel as one word: (s + 128) per axis, 8 bits each
TPME_HD unsigned pack_image(int sx, int sy, int sz) {
  return (unsigned)(sx + 128) | ((unsigned)(sy + 128) << 8) | ((unsigned)(sz + 128) << 16);
}
TPME_HD void unpack_image(unsigned w, int& sx, int& sy, int& sz) {
  sx = (int)(w & 255u) - 128; sy = (int)((w >> 8) & 255u) - 128; sz = (int)((w >> 16) & 255u) - 128;
}

// squared distance between the atom `me` and the image (sx, sy, sz) of `other` -- ONE expression shared by
// the search and by the code that writes the pair out, so both see the same bits
template <typename T>
TPME_HD T nl_dist_sq(const NlRecord<T>& me, const NlRecord<T>& other, const NeighborGeometry& g, int sx, int sy, int sz) {
  const T tx = (T)(sx * g.cell[0] + sy * g.cell[3] + sz * g.cell[6]);
  const T ty = (T)(sx * g.cell[1] + sy * g.cell[4] + sz * g.cell[7]);
  const T tz = (T)(sx * g.cell[2] + sy * g.cell[5] + sz * g.cell[8]);
  const T ddx = other.x + tx - me.x;
  const T ddy = other.y + ty - me.y;
  const T ddz = other.z + tz - me.z;
  return ddx * ddx + ddy * ddy + ddz * ddz;
}

// Visits all neighbors of the atom in sorted slot `slot`: hit(s, image) is called for every pair found with
// the sorted slot `s` of the partner and the packed image shift (with respect to the WRAPPED positions);
// returns their number.
//   full lists: all bins within reach, every pair from both ends (only the zero-shift self pair is skipped);
//   half lists: every unordered pair exactly once -- only the bins whose unwrapped offset (dx, dy, dz) is
//     lexicographically positive are visited, plus the later atoms of the own bin.  The offset of a pair seen
//     from its other end is the negative one, so exactly one end finds it; a self image (i, i, S) has the
//     offset S . n_bins and is found for the lexicographically positive S.  Half the distance tests.
template <typename T, typename F>
TPME_HD int nl_visit_slot(int64_t slot, const NlRecord<T>* __restrict__ sorted, const NlShift* __restrict__ sshift,
                          const int* __restrict__ bin_start, const NeighborGeometry& g, F& hit) {
  const NlRecord<T> me = sorted[slot];
  const int my_bin = sshift[slot].bin;
  const int bz = my_bin % g.n_bins[2];
  const int bxy = my_bin / g.n_bins[2];
  const int by = bxy % g.n_bins[1], bx = bxy / g.n_bins[1];
  const T cutoff_sq = (T)g.cutoff_sq;
  const bool full = g.full_list != 0;
  int found = 0;
  for (int dx = full ? -g.reach[0] : 0; dx <= g.reach[0]; ++dx) {
    const int ux = bx + dx;
    if (!g.periodic[0] && (ux < 0 || ux >= g.n_bins[0])) continue;
    const int sx = g.periodic[0] ? image_of(ux, g.n_bins[0]) : 0;
    const int wx = ux - sx * g.n_bins[0];
    for (int dy = (full || dx > 0) ? -g.reach[1] : 0; dy <= g.reach[1]; ++dy) {
      const int uy = by + dy;
      if (!g.periodic[1] && (uy < 0 || uy >= g.n_bins[1])) continue;
      const int sy = g.periodic[1] ? image_of(uy, g.n_bins[1]) : 0;
      const int wy = uy - sy * g.n_bins[1];
      const int column = (wx * g.n_bins[1] + wy) * g.n_bins[2];
      const bool own_column = !full && dx == 0 && dy == 0;     // half lists: dz >= 0 only
      // the z bins within reach, cut into runs that share one image count sz (contiguous sorted records)
      int uz = own_column ? bz : bz - g.reach[2];
      const int uz_end = bz + g.reach[2];
      while (uz <= uz_end) {
        const bool own_bin = own_column && uz == bz;           // starts in the atom's own bin: later slots only
        int sz, wz0, wz1;
        if (!g.periodic[2]) {
          sz = 0;
          wz0 = uz < 0 ? 0 : uz;
          wz1 = uz_end < g.n_bins[2] - 1 ? uz_end : g.n_bins[2] - 1;
          uz = uz_end + 1;
          if (wz0 > wz1) break;
        } else {
          sz = image_of(uz, g.n_bins[2]);
          wz0 = uz - sz * g.n_bins[2];
          const int len = (uz_end - uz) < (g.n_bins[2] - 1 - wz0) ? (uz_end - uz) : (g.n_bins[2] - 1 - wz0);
          wz1 = wz0 + len;
          uz += len + 1;
        }
        const unsigned image = pack_image(sx, sy, sz);
        const bool zero_shift = (sx == 0 && sy == 0 && sz == 0);
        const int s_end = bin_start[column + wz1 + 1];
        int s = bin_start[column + wz0];
        if (own_bin) s = (int)slot + 1;
        // four candidates are loaded before the first one is tested (independent loads in flight)
        for (; s < s_end; s += 4) {
          NlRecord<T> cand[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) cand[k] = sorted[s + k < s_end ? s + k : s_end - 1];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (s + k >= s_end) break;
            if (full && zero_shift && s + k == slot) continue;
            const T r2 = nl_dist_sq<T>(me, cand[k], g, sx, sy, sz);
            if (!(r2 < cutoff_sq)) continue;
            hit(s + k, image);
            ++found;
          }
        }
      }
    }
  }
  return found;
}

// writes pair number `o` of the list: the atom in `slot` and the image `image` of the atom in slot `s`.
// Half lists are oriented (i < j, or i == j with S lexicographically positive -- guaranteed by the visit).
template <typename T, typename I>
TPME_HD void nl_emit(int64_t slot, int s, unsigned image, const NlRecord<T>* __restrict__ sorted,
                     const NlShift* __restrict__ sshift, const NeighborGeometry& g, int64_t o,
                     I* __restrict__ indices, T* __restrict__ distances, int* __restrict__ shifts) {
  const NlRecord<T> me = sorted[slot], other = sorted[s];
  const NlShift mine = sshift[slot], theirs = sshift[s];
  int sx, sy, sz;
  unpack_image(image, sx, sy, sz);
  distances[o] = nl_sqrt(nl_dist_sq<T>(me, other, g, sx, sy, sz));
  // shift with respect to the original (unwrapped) positions
  int S[3] = {sx + mine.k[0] - theirs.k[0], sy + mine.k[1] - theirs.k[1], sz + mine.k[2] - theirs.k[2]};
  const bool flip = !g.full_list && me.index > other.index;
  indices[2 * o] = (I)(flip ? other.index : me.index);
  indices[2 * o + 1] = (I)(flip ? me.index : other.index);
  for (int c = 0; c < 3; ++c) shifts[3 * o + c] = flip ? -S[c] : S[c];
}

}  // namespace tpme
