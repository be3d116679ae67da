// Hand-written real 3-D FFT for power-of-two meshes with the Green's-function multiply fused
// into the x pass:   out = iFFT3( G(k) * FFT3(in) )   in five passes
//
//     rows R2C (z)  ->  lines FFT (y)  ->  lines FFT . G . iFFT (x)  ->  lines iFFT (y)  ->  rows C2R (z)
//
// (replaces cuFFT R2C + multiply kernel + cuFFT C2R = 7 launches and 2R + 12K bytes by 5 launches
// and 2R + 8K bytes, R / K = real / half-complex mesh bytes; reference ops:
// lib/kspace_filter.py:169-187).
//
// Each transform of length N = R_a R_b (R_c) is done as 2 (3) radix groups of up to 16 points held
// in registers; only the exchange between groups goes through shared memory:
//   * the first group loads straight from global memory into registers, the last one stores
//     straight from registers to global memory;
//   * forward transforms are decimation-in-frequency (natural in, bit-reversed out), inverse
//     ones decimation-in-time (bit-reversed in, natural out); along the strided axes (y, x) the
//     bit reversal is folded into the line index of the global access (free: coalescing comes
//     from the contiguous z axis), so global memory always holds the standard layout
//     (C, nx, ny, nz/2+1);
//   * in the x pass the last forward group, the multiply by G(k) and the first inverse group
//     work on the same registers:  DIF -> G at the bit-reversed frequency -> DIT;
//   * G(k) is evaluated from per-CTA axis tables (k = a[ix] + b[iy,iz], sines by angle addition),
//     one exp and one division per k-point.
#pragma once
#include <cooperative_groups.h>
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "green.cuh"

namespace tpme {
namespace fft {

template <typename T> struct alignas(2 * sizeof(T)) C2 { T x, y; };

template <typename T> __device__ __forceinline__ C2<T> operator+(C2<T> a, C2<T> b) { return {a.x + b.x, a.y + b.y}; }
template <typename T> __device__ __forceinline__ C2<T> operator-(C2<T> a, C2<T> b) { return {a.x - b.x, a.y - b.y}; }
template <typename T> __device__ __forceinline__ C2<T> cmul(C2<T> a, C2<T> b) {
  return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
template <typename T> __device__ __forceinline__ C2<T> conj(C2<T> a) { return {a.x, -a.y}; }

__host__ __device__ constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n / 2); }
// radix of the decimation-in-frequency group that splits a block of size M
__host__ __device__ constexpr int group_radix(int M) {
  return ilog2(M) <= 4 ? M : (ilog2(M) <= 8 ? (1 << ((ilog2(M) + 1) / 2)) : 8);
}
__host__ __device__ constexpr int num_groups(int N) { return N <= 1 ? 0 : 1 + num_groups(N / group_radix(N)); }
template <int N> __device__ __forceinline__ int bitrev(int i) {
  constexpr int shift = (32 - ilog2(N)) & 31;
  return N <= 1 ? 0 : (int)(__brev((unsigned)i) >> shift);
}

// forward twiddles e^{-2 pi i k / N}, k < N/2, computed in double
template <typename T, int N>
__device__ __forceinline__ void fill_twiddles(C2<T>* tw, int tid, int nthreads) {
  for (int k = tid; k < N / 2; k += nthreads) {
    double s, c;
    sincospi(-2.0 * (double)k / (double)N, &s, &c);
    tw[k] = {(T)c, (T)s};
  }
}

// b * e^{-2 pi i k16 / 16} (k16 is a compile-time constant after unrolling, 0 <= k16 < 8)
template <typename T>
__device__ __forceinline__ C2<T> mul_root16(C2<T> b, int k16) {
  const T c1 = T(0.92387953251128673848), s1 = T(0.38268343236508978178), h = T(0.70710678118654752440);
  switch (k16) {
    case 0: return b;
    case 1: return cmul(b, C2<T>{c1, -s1});
    case 2: return {h * (b.x + b.y), h * (b.y - b.x)};
    case 3: return cmul(b, C2<T>{s1, -c1});
    case 4: return {b.y, -b.x};
    case 5: return cmul(b, C2<T>{-s1, -c1});
    case 6: return {h * (b.y - b.x), -h * (b.x + b.y)};
    default: return cmul(b, C2<T>{-c1, -s1});
  }
}

// R-point decimation-in-frequency group on registers: v[i] = element (j0 + i * M / R) of a block
// of size M of a length-N transform.  SIGN -1 forward, +1 inverse (conjugated twiddles).
template <typename T, int N, int M, int R, int SIGN, int TWMUL>
__device__ __forceinline__ void dif_regs(C2<T> (&v)[R], int j0, const C2<T>* __restrict__ tw) {
#pragma unroll
  for (int hs = R / 2, s = 0; hs >= 1; hs >>= 1, ++s) {
    const C2<T> b = (M == R) ? C2<T>{T(1), T(0)} : tw[((j0 * (N / M)) << s) * TWMUL];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      if (i & hs) continue;
      const int k16 = ((i & (hs - 1)) << s) * (16 / R);
      C2<T> w = mul_root16(b, k16);
      if (SIGN > 0) w = conj(w);
      const C2<T> lo = v[i], hi = v[i + hs];
      v[i] = lo + hi;
      v[i + hs] = (M == R && k16 == 0) ? lo - hi : cmul(lo - hi, w);
    }
  }
}

// R-point decimation-in-time group on registers: v[i] = element (j0 + i * Q) of a block of size
// Q * R (Q = product of the radices already applied).
template <typename T, int N, int Q, int R, int SIGN, int TWMUL>
__device__ __forceinline__ void dit_regs(C2<T> (&v)[R], int j0, const C2<T>* __restrict__ tw) {
#pragma unroll
  for (int hs = 1; hs < R; hs <<= 1) {
    const C2<T> b = (Q == 1) ? C2<T>{T(1), T(0)} : tw[(j0 * (N / (2 * Q * hs))) * TWMUL];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      if (i & hs) continue;
      const int k16 = (i & (hs - 1)) * (8 / hs);
      C2<T> w = mul_root16(b, k16);
      if (SIGN > 0) w = conj(w);
      const C2<T> t = (Q == 1 && k16 == 0) ? v[i + hs] : cmul(v[i + hs], w);
      const C2<T> lo = v[i];
      v[i] = lo + t;
      v[i + hs] = lo - t;
    }
  }
}

// Radix chain of a length-N transform: DIF groups RA (block N), RB (block N/RA), RC (block
// N/(RA RB)); the DIT chain is its mirror (RC/RB first at spacing 1, RA last at spacing N/RA).
template <int N> struct Chain {
  static constexpr int NG = num_groups(N);
  static constexpr int RA = group_radix(N);
  static constexpr int MB = N / RA;                       // block size seen by the second group
  static constexpr int RB = NG >= 2 ? group_radix(MB) : 1;
  static constexpr int MC = MB / RB;
  static constexpr int RC = NG >= 3 ? group_radix(MC) : 1;
  static_assert(NG >= 1 && NG <= 3 && RA * RB * RC == N, "unsupported transform length");
  static constexpr int RL = NG == 1 ? RA : (NG == 2 ? RB : RC);   // last DIF / first DIT radix
};

// ---------------------------------------------------------------------------------------
// Green's function in the x pass, from per-CTA axis tables.  The CTA has a fixed iy and a chunk
// of iz; table A is indexed by ix, table B by the column (iy, iz).  Variants (template GV):
//   GV_ORTHO    Coulomb-type G (exp(-c k^2)/k^2), reciprocal cell diagonal: fully separable,
//               k^2 = A.k2 + B.k2,  G = amp A.f B.f / k^2  (the P3M factor is folded into f)
//   GV_TRI      Coulomb-type G, triclinic cell: k = A.k + B.k, one exp per k-point
//   GV_TRI_P3M  same with the P3M influence function, sines by angle addition
//   GV_GENERIC  anything else (table, inverse power law p >= 2): green_value() per k-point
// ---------------------------------------------------------------------------------------
enum { GV_ORTHO = 0, GV_TRI = 1, GV_TRI_P3M = 2, GV_GENERIC = 3 };

template <typename GT> struct AxisEntry { GT k[3], s[3], c[3]; };   // largest table entry

template <typename GT> struct FastMath;
template <> struct FastMath<float> {
  static __device__ __forceinline__ float exp(float x) { return __expf(x); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdividef(a, b); }
};
template <> struct FastMath<double> {
  static __device__ __forceinline__ double exp(double x) { return ::exp(x); }
  static __device__ __forceinline__ double div(double a, double b) { return a / b; }
};

// h / sin(h) raised to 2n (1 at h = 0, 0 where sin(h) = 0): one axis of 1 / U^2
template <typename GT>
__device__ __forceinline__ GT inv_sinc_pow(GT h, int nodes) {
  if (nodes <= 0 || h == GT(0)) return GT(1);
  const GT s = MathFn<GT>::sin(h);
  if (s == GT(0)) return GT(0);
  const GT r = h / s, r2 = r * r;
  GT out = GT(1);
  for (int i = 0; i < nodes; ++i) out *= r2;
  return out;
}

// fill table entry `e` for frequency vector k = f0 * B[row0] (+ f1 * B[row1])
template <typename GT, int GV>
__device__ __forceinline__ void make_axis_entry(AxisEntry<GT>& e, const GreenDev<GT>& g, GT f0, int row0,
                                                GT f1, int row1) {
  GT k[3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
    k[a] = f0 * g.recip[3 * row0 + a] + (row1 >= 0 ? f1 * g.recip[3 * row1 + a] : GT(0));
  if (GV == GV_ORTHO) {
    const GT k2 = k[0] * k[0] + k[1] * k[1] + k[2] * k[2];
    GT f = MathFn<GT>::exp(-g.half_s2 * k2);
#pragma unroll
    for (int a = 0; a < 3; ++a) f *= inv_sinc_pow<GT>(GT(0.5) * g.spacing[a] * k[a], g.p3m_nodes);
    e.k[0] = k2;
    e.k[1] = f;
  } else {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      e.k[a] = k[a];
      if (GV == GV_TRI_P3M) MathFn<GT>::sincos(GT(0.5) * g.spacing[a] * k[a], &e.s[a], &e.c[a]);
    }
  }
}

template <typename GT, int GV>
__device__ __forceinline__ GT green_from_axes(const GreenDev<GT>& g, const AxisEntry<GT>& ea,
                                              const AxisEntry<GT>& eb) {
  using F = FastMath<GT>;
  if (GV == GV_ORTHO) {
    const GT k_sq = ea.k[0] + eb.k[0];
    const GT val = F::div(g.amplitude * ea.k[1] * eb.k[1], k_sq);
    return k_sq == GT(0) ? g.k0_value : val;
  }
  const GT kk[3] = {ea.k[0] + eb.k[0], ea.k[1] + eb.k[1], ea.k[2] + eb.k[2]};
  const GT k_sq = kk[0] * kk[0] + kk[1] * kk[1] + kk[2] * kk[2];
  GT num = g.amplitude * F::exp(-g.half_s2 * k_sq);
  GT den = k_sq;
  if (GV == GV_TRI_P3M) {
    // 1/U^2 = (hx hy hz / (sin hx sin hy sin hz))^(2n) with sinc(0) = 1
    GT sn = GT(1), hp = GT(1);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const GT h = GT(0.5) * g.spacing[a] * kk[a];
      const GT s = ea.s[a] * eb.c[a] + ea.c[a] * eb.s[a];
      const bool zero = (h == GT(0));
      sn *= zero ? GT(1) : s;
      hp *= zero ? GT(1) : h;
    }
    const GT ratio = F::div(hp, sn);     // 1 / prod sinc (a ratio: the plain products underflow in fp32)
    const GT r2 = ratio * ratio;
    GT ru = GT(1);
    for (int i = 0; i < g.p3m_nodes; ++i) ru *= r2;
    // sinc -> 0 far outside the first Brillouin zone of a skewed cell: keep 0 * huge = 0
    ru = ru < GT(sizeof(GT) == 4 ? 1e30 : 1e300) ? ru : GT(sizeof(GT) == 4 ? 1e30 : 1e300);
    num = (sn == GT(0) || num == GT(0)) ? GT(0) : num * ru;
  }
  const GT val = F::div(num, den);
  return k_sq == GT(0) ? g.k0_value : val;
}

__device__ __forceinline__ int freq_index(int i, int n) { return i < (n + 1) / 2 ? i : i - n; }

// ---------------------------------------------------------------------------------------
// strided passes (y or x): a CTA owns all N points of the transform axis for `zc` contiguous
// inner (z) elements.  MODE 0: forward, 1: inverse, 2: forward . G . inverse (x pass only).
// Tile origin of work item `o`: (o / d1) * s1 + (o % d1) * s0 + chunk * zc; line stride `ls`.
// ---------------------------------------------------------------------------------------
// skewed position inside a row buffer: one pad element every 16, so that the stride-RL accesses of
// the second radix group fall into distinct shared-memory banks
__device__ __forceinline__ int skew(int p) { return p + (p >> 4); }

template <typename T> struct MaxThreads { static constexpr int value = sizeof(T) == 8 ? 128 : 256; };

template <typename T, typename GT, int N, int MODE, int GV, bool REMOTE>
__global__ void __launch_bounds__(MaxThreads<T>::value)
lines_fft_kernel(C2<T>* __restrict__ data, int n_inner, int zc, int n_chunks, int64_t ls, int d1,
                 int64_t s1, int64_t s0, GreenDev<GT> green, int nx, int ny, int nz, T* __restrict__ dc_out,
                 int y_off, RemoteStore rs) {
  using CH = Chain<N>;
  constexpr int NG = CH::NG, RA = CH::RA, RB = CH::RB, RL = CH::RL;
  constexpr int QA = N / RA;                 // element spacing inside the first DIF group
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C2<T>* tile = reinterpret_cast<C2<T>*>(smem_raw);
  C2<T>* tw = tile + (NG > 1 ? (size_t)N * zc : 0);
  AxisEntry<GT>* ax_a = reinterpret_cast<AxisEntry<GT>*>(tw + N / 2);   // [N], MODE 2 only
  AxisEntry<GT>* ax_b = ax_a + N;                                       // [zc]
  const int tid = threadIdx.x, nt = blockDim.x;
  const int o = blockIdx.x / n_chunks, chunk = blockIdx.x - o * n_chunks;
  const int z0 = chunk * zc;
  const int cols = min(zc, n_inner - z0);
  const int o_hi = o / d1, o_lo = o - o_hi * d1;
  C2<T>* base = data + o_hi * s1 + o_lo * s0 + z0;
  // final stores: in place, or into the peers' buffers
  const int64_t r_base = REMOTE ? (int64_t)(o / rs.dA) * rs.sA + (int64_t)(o % rs.dA) * rs.sB + rs.off + z0 : 0;
  const int r_mask = (1 << rs.shift) - 1;
  auto store_line = [&](int line, int c, C2<T> v) {
    if (REMOTE)
      (reinterpret_cast<C2<T>*>(rs.p[line >> rs.shift]) + r_base)[(int64_t)(line & r_mask) * rs.sL + c] = v;
    else
      base[(int64_t)line * ls + c] = v;
  };

  pdl_trigger();
  if (NG > 1) fill_twiddles<T, N>(tw, tid, nt);
  if (MODE == 2 && GV != GV_GENERIC) {
    for (int i = tid; i < N + cols; i += nt) {
      if (i < N) make_axis_entry<GT, GV>(ax_a[i], green, (GT)freq_index(i, nx), 0, GT(0), -1);
      else make_axis_entry<GT, GV>(ax_b[i - N], green, (GT)freq_index(o_lo + y_off, ny), 1, (GT)(z0 + i - N), 2);
    }
  }
  if (NG > 1 || MODE == 2) __syncthreads();
  pdl_wait();      // twiddles and Green axis tables above overlap the tail of the pass before

  // ---- first group: global -> registers ------------------------------------------------
  if (MODE == 0 || MODE == 2) {
    // DIF group A on lines j0 + i * QA
    for (int w = tid; w < cols * QA; w += nt) {
      const int j0 = w / cols, c = w - j0 * cols;
      C2<T> v[RA];
#pragma unroll
      for (int i = 0; i < RA; ++i) v[i] = base[(int64_t)(j0 + i * QA) * ls + c];
      dif_regs<T, N, N, RA, -1, 1>(v, j0, tw);
      if (NG == 1) {
        if (MODE == 2) {
          // single-group transform: multiply and invert in the same registers
          AxisEntry<GT> eb;
          if (GV != GV_GENERIC) eb = ax_b[c];
          if (dc_out != nullptr && o_lo + y_off == 0 && z0 + c == 0) dc_out[o_hi] = v[0].x;   // ix = 0
#pragma unroll
          for (int i = 0; i < RA; ++i) {
            const int ix = bitrev<N>(i);
            T gv;
            if (GV == GV_GENERIC) {
              const int64_t flat = ((int64_t)ix * ny + o_lo + y_off) * (nz / 2 + 1) + z0 + c;
              gv = (T)green_value<GT, T>(green, ix, o_lo + y_off, z0 + c, nx, ny, nz, flat);
            } else {
              gv = (T)green_from_axes<GT, GV>(green, ax_a[ix], eb);
            }
            v[i] = {v[i].x * gv, v[i].y * gv};
          }
          dit_regs<T, N, 1, RA, +1, 1>(v, 0, tw);
#pragma unroll
          for (int i = 0; i < RA; ++i) store_line(i, c, v[i]);
        } else {
#pragma unroll
          for (int i = 0; i < RA; ++i) store_line(bitrev<N>(i), c, v[i]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < RA; ++i) tile[(j0 + i * QA) * zc + c] = v[i];
      }
    }
  } else {
    // DIT first group (spacing 1, radix RL) on bit-reversed positions blk * RL + i
    for (int w = tid; w < cols * (N / RL); w += nt) {
      const int blk = w / cols, c = w - blk * cols;
      C2<T> v[RL];
#pragma unroll
      for (int i = 0; i < RL; ++i) v[i] = base[(int64_t)bitrev<N>(blk * RL + i) * ls + c];
      dit_regs<T, N, 1, RL, +1, 1>(v, 0, tw);
      if (NG == 1) {
#pragma unroll
        for (int i = 0; i < RL; ++i) base[(int64_t)i * ls + c] = v[i];
      } else {
#pragma unroll
        for (int i = 0; i < RL; ++i) tile[(blk * RL + i) * zc + c] = v[i];
      }
    }
  }
  if (NG == 1) return;
  __syncthreads();

  // ---- middle of the forward chain (N = 512 only) ------------------------------------------
  if (NG == 3 && (MODE == 0 || MODE == 2)) {
    constexpr int MB = CH::MB, QB = MB / RB;
    for (int w = tid; w < cols * (N / RB); w += nt) {
      const int g = w / cols, c = w - g * cols;
      const int blk = g / QB, j0 = g - blk * QB;
      C2<T>* p = tile + (blk * MB + j0) * zc + c;
      C2<T> v[RB];
#pragma unroll
      for (int i = 0; i < RB; ++i) v[i] = p[i * QB * zc];
      dif_regs<T, N, MB, RB, -1, 1>(v, j0, tw);
#pragma unroll
      for (int i = 0; i < RB; ++i) p[i * QB * zc] = v[i];
    }
    __syncthreads();
  }

  // ---- last forward group (+ G + first inverse group in the x pass) ------------------------
  if (MODE == 0 || MODE == 2) {
    for (int w = tid; w < cols * (N / RL); w += nt) {
      const int blk = w / cols, c = w - blk * cols;
      C2<T> v[RL];
#pragma unroll
      for (int i = 0; i < RL; ++i) v[i] = tile[(blk * RL + i) * zc + c];
      dif_regs<T, N, RL, RL, -1, 1>(v, 0, tw);
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < RL; ++i) store_line(bitrev<N>(blk * RL + i), c, v[i]);
      } else {
        AxisEntry<GT> eb;
        if (GV != GV_GENERIC) eb = ax_b[c];
        if (dc_out != nullptr && blk == 0 && o_lo + y_off == 0 && z0 + c == 0) dc_out[o_hi] = v[0].x;   // ix = 0
#pragma unroll
        for (int i = 0; i < RL; ++i) {
          const int ix = bitrev<N>(blk * RL + i);
          T gv;
          if (GV == GV_GENERIC) {
            const int64_t flat = ((int64_t)ix * ny + o_lo + y_off) * (nz / 2 + 1) + z0 + c;
            gv = (T)green_value<GT, T>(green, ix, o_lo + y_off, z0 + c, nx, ny, nz, flat);
          } else {
            gv = (T)green_from_axes<GT, GV>(green, ax_a[ix], eb);
          }
          v[i] = {v[i].x * gv, v[i].y * gv};
        }
        dit_regs<T, N, 1, RL, +1, 1>(v, 0, tw);
#pragma unroll
        for (int i = 0; i < RL; ++i) tile[(blk * RL + i) * zc + c] = v[i];
      }
    }
    if (MODE == 0) return;
    __syncthreads();
  }

  // ---- middle of the inverse chain (N = 512 only) ------------------------------------------
  if (NG == 3) {
    constexpr int QB = RL;   // spacing after the first DIT group
    for (int w = tid; w < cols * (N / RB); w += nt) {
      const int g = w / cols, c = w - g * cols;
      const int blk = g / QB, j0 = g - blk * QB;
      C2<T>* p = tile + (blk * QB * RB + j0) * zc + c;
      C2<T> v[RB];
#pragma unroll
      for (int i = 0; i < RB; ++i) v[i] = p[i * QB * zc];
      dit_regs<T, N, QB, RB, +1, 1>(v, j0, tw);
#pragma unroll
      for (int i = 0; i < RB; ++i) p[i * QB * zc] = v[i];
    }
    __syncthreads();
  }

  // ---- last inverse group: registers -> global (natural order) ------------------------------
  for (int w = tid; w < cols * QA; w += nt) {
    const int j0 = w / cols, c = w - j0 * cols;
    C2<T> v[RA];
#pragma unroll
    for (int i = 0; i < RA; ++i) v[i] = tile[(j0 + i * QA) * zc + c];
    dit_regs<T, N, QA, RA, +1, 1>(v, j0, tw);
#pragma unroll
    for (int i = 0; i < RA; ++i) store_line(j0 + i * QA, c, v[i]);
  }
}

// ---------------------------------------------------------------------------------------
// z pass, forward: rows of NZ reals -> NZ/2+1 complex (packed half-length DIF + split).
// The last radix group writes its outputs to their natural (bit-reversed back) positions of a
// second buffer, so the split reads and the global stores are conflict-free / coalesced.
// ---------------------------------------------------------------------------------------
template <typename T, int NZ>
__global__ void __launch_bounds__(256)
rows_r2c_kernel(const T* __restrict__ in, C2<T>* __restrict__ out, int64_t n_rows, int rows_per_cta) {
  constexpr int H = NZ / 2, P = H + 1;
  using CH = Chain<H>;
  constexpr int NG = CH::NG, RA = CH::RA, RL = CH::RL, QA = H / RA;
  static_assert(NG <= 2, "row transforms use at most two radix groups");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int P1 = H + (H >> 4) + 1;   // pitch of the skewed exchange buffer
  C2<T>* buf1 = reinterpret_cast<C2<T>*>(smem_raw);
  C2<T>* buf2 = buf1 + (NG == 2 ? (size_t)rows_per_cta * P1 : 0);
  C2<T>* tw = buf2 + (size_t)rows_per_cta * P;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int64_t row0 = (int64_t)blockIdx.x * rows_per_cta;
  const int rows = (int)min((int64_t)rows_per_cta, n_rows - row0);
  pdl_trigger();
  fill_twiddles<T, NZ>(tw, tid, nt);
  __syncthreads();
  pdl_wait();      // the twiddles above overlap the tail of the kernel before
  const C2<T>* in2 = reinterpret_cast<const C2<T>*>(in) + row0 * H;   // z[m] = x[2m] + i x[2m+1]
  for (int w = tid; w < rows * QA; w += nt) {
    const int r = w / QA, j0 = w - r * QA;
    C2<T> v[RA];
#pragma unroll
    for (int i = 0; i < RA; ++i) v[i] = in2[r * H + j0 + i * QA];
    dif_regs<T, H, H, RA, -1, 2>(v, j0, tw);
    if (NG == 1) {
#pragma unroll
      for (int i = 0; i < RA; ++i) buf2[r * P + bitrev<H>(i)] = v[i];
    } else {
#pragma unroll
      for (int i = 0; i < RA; ++i) buf1[r * P1 + skew(j0 + i * QA)] = v[i];
    }
  }
  __syncthreads();
  if (NG == 2) {
    for (int w = tid; w < rows * (H / RL); w += nt) {
      const int r = w / (H / RL), blk = w - r * (H / RL);
      C2<T> v[RL];
#pragma unroll
      for (int i = 0; i < RL; ++i) v[i] = buf1[r * P1 + skew(blk * RL + i)];
      dif_regs<T, H, RL, RL, -1, 2>(v, 0, tw);
#pragma unroll
      for (int i = 0; i < RL; ++i) buf2[r * P + bitrev<H>(blk * RL + i)] = v[i];
    }
    __syncthreads();
  }
  // split:  X[k] = (A + B)/2 - i/2 w^k (A - B),  A = Z[k], B = conj(Z[H-k]);  straight to global
  C2<T>* o = out + row0 * P;
  for (int i = tid; i < rows * (H / 2 + 1); i += nt) {
    const int r = i / (H / 2 + 1), k = i - r * (H / 2 + 1);
    const C2<T>* row = buf2 + r * P;
    C2<T>* orow = o + (int64_t)r * P;
    if (k == 0) {
      const C2<T> z0 = row[0];
      orow[0] = {z0.x + z0.y, T(0)};
      orow[H] = {z0.x - z0.y, T(0)};
    } else {
      const int kk = H - k;
      const C2<T> zk = row[k], zkk = row[kk];
      const C2<T> w = tw[k];                       // e^{-2 pi i k / NZ}
      {
        const C2<T> s = {T(0.5) * (zk.x + zkk.x), T(0.5) * (zk.y - zkk.y)};     // (A + B)/2
        const C2<T> d = {T(0.5) * (zk.x - zkk.x), T(0.5) * (zk.y + zkk.y)};     // (A - B)/2
        const C2<T> wd = cmul(d, w);
        orow[k] = {s.x + wd.y, s.y - wd.x};                                     // s - i wd
      }
      if (kk != k) {
        // X[H-k]: A' = Z[H-k], B' = conj(Z[k]), w^{H-k} = -conj(w^k)
        const C2<T> s = {T(0.5) * (zkk.x + zk.x), T(0.5) * (zkk.y - zk.y)};
        const C2<T> d = {T(0.5) * (zkk.x - zk.x), T(0.5) * (zkk.y + zk.y)};
        const C2<T> wd = {-(d.x * w.x + d.y * w.y), -(d.y * w.x - d.x * w.y)};
        orow[kk] = {s.x + wd.y, s.y - wd.x};
      }
    }
  }
}

// z pass, inverse: NZ/2+1 complex -> NZ reals (unnormalised sum over the Hermitian extension)
template <typename T, int NZ>
__global__ void __launch_bounds__(256)
rows_c2r_kernel(const C2<T>* __restrict__ in, T* __restrict__ out, int64_t n_rows, int rows_per_cta) {
  constexpr int H = NZ / 2, P = H + 1;
  using CH = Chain<H>;
  constexpr int NG = CH::NG, RA = CH::RA, RL = CH::RL, QA = H / RA;
  static_assert(NG <= 2, "row transforms use at most two radix groups");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int P1 = H + (H >> 4) + 1;   // pitch of the skewed exchange buffer
  C2<T>* buf1 = reinterpret_cast<C2<T>*>(smem_raw);
  C2<T>* buf2 = buf1 + (NG == 2 ? (size_t)rows_per_cta * P1 : 0);
  C2<T>* tw = buf2 + (size_t)rows_per_cta * P;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int64_t row0 = (int64_t)blockIdx.x * rows_per_cta;
  const int rows = (int)min((int64_t)rows_per_cta, n_rows - row0);
  pdl_trigger();
  fill_twiddles<T, NZ>(tw, tid, nt);
  __syncthreads();
  pdl_wait();      // the twiddles above overlap the tail of the kernel before
  // un-split straight from global memory (natural order in buf2):
  //   Z'[k] = (X[k] + conj X[H-k]) + i conj(w^k) (X[k] - conj X[H-k])
  const C2<T>* src = in + row0 * P;
  {
    // four (X[k], X[H-k]) pairs per thread are in flight before the first one is consumed (the loop with
    // one pair per iteration stalled on its two loads: long_scoreboard 8.4 in profiles/r02_ncu_full_c4.csv)
    constexpr int U = 4, KH = H / 2 + 1;
    const int total = rows * KH;
    for (int i0 = tid; i0 < total; i0 += U * nt) {
      C2<T> xa[U], xb[U];
      int rr[U], kq[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * nt;
        rr[u] = i / KH;
        kq[u] = i - rr[u] * KH;
        if (i < total) {
          const C2<T>* irow = src + (int64_t)rr[u] * P;
          xa[u] = irow[kq[u]];
          xb[u] = irow[H - kq[u]];
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (i0 + u * nt >= total) break;
        const int k = kq[u], kk = H - k;
        C2<T>* row = buf2 + rr[u] * P;
        const C2<T> xk = xa[u], xkk = xb[u];
        if (k == 0) {
          row[0] = {xk.x + xkk.x, xk.x - xkk.x};                 // X[0], X[H] are real
          continue;
        }
        const C2<T> w = tw[k];
        {
          const C2<T> s = {xk.x + xkk.x, xk.y - xkk.y};          // X[k] + conj X[H-k]
          const C2<T> d = {xk.x - xkk.x, xk.y + xkk.y};          // X[k] - conj X[H-k]
          const C2<T> wd = {d.x * w.x + d.y * w.y, d.y * w.x - d.x * w.y};   // conj(w) d
          row[k] = {s.x - wd.y, s.y + wd.x};                     // s + i wd
        }
        if (kk != k) {
          const C2<T> s = {xkk.x + xk.x, xkk.y - xk.y};
          const C2<T> d = {xkk.x - xk.x, xkk.y + xk.y};
          const C2<T> wd = {-(d.x * w.x - d.y * w.y), -(d.x * w.y + d.y * w.x)};   // conj(w^{H-k}) = -w^k
          row[kk] = {s.x - wd.y, s.y + wd.x};
        }
      }
    }
  }
  __syncthreads();
  C2<T>* o = reinterpret_cast<C2<T>*>(out) + row0 * H;
  if (NG == 1) {
    for (int r = tid; r < rows; r += nt) {
      C2<T> v[RA];
#pragma unroll
      for (int i = 0; i < RA; ++i) v[i] = buf2[r * P + bitrev<H>(i)];
      dit_regs<T, H, 1, RA, +1, 2>(v, 0, tw);
#pragma unroll
      for (int i = 0; i < RA; ++i) o[(int64_t)r * H + i] = v[i];
    }
    return;
  }
  // first DIT group: position blk * RL + i of the bit-reversed sequence = natural index bitrev(..)
  for (int w = tid; w < rows * (H / RL); w += nt) {
    const int r = w / (H / RL), blk = w - r * (H / RL);
    C2<T> v[RL];
#pragma unroll
    for (int i = 0; i < RL; ++i) v[i] = buf2[r * P + bitrev<H>(blk * RL + i)];
    dit_regs<T, H, 1, RL, +1, 2>(v, 0, tw);
#pragma unroll
    for (int i = 0; i < RL; ++i) buf1[r * P1 + skew(blk * RL + i)] = v[i];
  }
  __syncthreads();
  for (int w = tid; w < rows * QA; w += nt) {
    const int r = w / QA, j0 = w - r * QA;
    C2<T> v[RA];
#pragma unroll
    for (int i = 0; i < RA; ++i) v[i] = buf1[r * P1 + skew(j0 + i * QA)];
    dit_regs<T, H, QA, RA, +1, 2>(v, j0, tw);
#pragma unroll
    for (int i = 0; i < RA; ++i) o[(int64_t)r * H + j0 + i * QA] = v[i];   // (y[2m], y[2m+1])
  }
}

// ---------------------------------------------------------------------------------------
// fused (y, z) plane transforms: one CTA per (channel, x) plane keeps the whole half-complex
// plane [NY][NZ/2+1] in shared memory, so the z and y passes cost one global read and one global
// write in total (used when the plane fits: <= 128 x 128 here).
// ---------------------------------------------------------------------------------------
// the plane kernels hold at most one radix-16 group of complex values per thread (<= 92 registers in
// fp64): up to 512 threads per CTA, so that the single CTA an SM holds (the plane fills its shared
// memory) has enough warps to hide the global and shared-memory latency
constexpr int kPlaneMaxThreads = 512;

template <typename T, int NY, int NZ>
__global__ void __launch_bounds__(kPlaneMaxThreads)
plane_r2c_kernel(const T* __restrict__ in, C2<T>* __restrict__ out, int rows_per_chunk) {
  constexpr int H = NZ / 2, P = H + 1, P1 = H + (H >> 4) + 1;
  using CZ = Chain<H>;
  using CY = Chain<NY>;
  static_assert(CZ::NG <= 2 && CY::NG <= 2, "plane kernels use at most two radix groups per axis");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C2<T>* plane = reinterpret_cast<C2<T>*>(smem_raw);
  C2<T>* buf1 = plane + (size_t)NY * P;
  C2<T>* twz = buf1 + (CZ::NG == 2 ? (size_t)rows_per_chunk * P1 : 0);
  C2<T>* twy = twz + NZ / 2;
  const int tid = threadIdx.x, nt = blockDim.x;
  pdl_trigger();
  fill_twiddles<T, NZ>(twz, tid, nt);
  fill_twiddles<T, NY>(twy, tid, nt);
  __syncthreads();
  pdl_wait();      // the twiddles above overlap the tail of the kernel before
  const C2<T>* in2 = reinterpret_cast<const C2<T>*>(in) + (int64_t)blockIdx.x * NY * H;
  // ---- z pass (packed half-length DIF, natural order into the plane) -----------------------
  constexpr int QAz = H / CZ::RA;
  for (int row0 = 0; row0 < NY; row0 += rows_per_chunk) {
    const int rows = min(rows_per_chunk, NY - row0);
    for (int w = tid; w < rows * QAz; w += nt) {
      const int r = w / QAz, j0 = w - r * QAz;
      C2<T> v[CZ::RA];
#pragma unroll
      for (int i = 0; i < CZ::RA; ++i) v[i] = in2[(row0 + r) * H + j0 + i * QAz];
      dif_regs<T, H, H, CZ::RA, -1, 2>(v, j0, twz);
      if (CZ::NG == 1) {
#pragma unroll
        for (int i = 0; i < CZ::RA; ++i) plane[(row0 + r) * P + bitrev<H>(i)] = v[i];
      } else {
#pragma unroll
        for (int i = 0; i < CZ::RA; ++i) buf1[r * P1 + skew(j0 + i * QAz)] = v[i];
      }
    }
    __syncthreads();
    if (CZ::NG == 2) {
      constexpr int RL = CZ::RL;
      for (int w = tid; w < rows * (H / RL); w += nt) {
        const int r = w / (H / RL), blk = w - r * (H / RL);
        C2<T> v[RL];
#pragma unroll
        for (int i = 0; i < RL; ++i) v[i] = buf1[r * P1 + skew(blk * RL + i)];
        dif_regs<T, H, RL, RL, -1, 2>(v, 0, twz);
#pragma unroll
        for (int i = 0; i < RL; ++i) plane[(row0 + r) * P + bitrev<H>(blk * RL + i)] = v[i];
      }
      __syncthreads();
    }
  }
  // ---- split in place -------------------------------------------------------------------
  for (int i = tid; i < NY * (H / 2 + 1); i += nt) {
    const int r = i / (H / 2 + 1), k = i - r * (H / 2 + 1);
    C2<T>* row = plane + r * P;
    if (k == 0) {
      const C2<T> z0 = row[0];
      row[0] = {z0.x + z0.y, T(0)};
      row[H] = {z0.x - z0.y, T(0)};
    } else {
      const int kk = H - k;
      const C2<T> zk = row[k], zkk = row[kk];
      const C2<T> w = twz[k];
      {
        const C2<T> s = {T(0.5) * (zk.x + zkk.x), T(0.5) * (zk.y - zkk.y)};
        const C2<T> d = {T(0.5) * (zk.x - zkk.x), T(0.5) * (zk.y + zkk.y)};
        const C2<T> wd = cmul(d, w);
        row[k] = {s.x + wd.y, s.y - wd.x};
      }
      if (kk != k) {
        const C2<T> s = {T(0.5) * (zkk.x + zk.x), T(0.5) * (zkk.y - zk.y)};
        const C2<T> d = {T(0.5) * (zkk.x - zk.x), T(0.5) * (zkk.y + zk.y)};
        const C2<T> wd = {-(d.x * w.x + d.y * w.y), -(d.y * w.x - d.x * w.y)};
        row[kk] = {s.x + wd.y, s.y - wd.x};
      }
    }
  }
  __syncthreads();
  // ---- y pass (DIF along the columns, bit reversal folded into the global line index) -------
  C2<T>* o = out + (int64_t)blockIdx.x * NY * P;
  constexpr int QAy = NY / CY::RA;
  for (int w = tid; w < P * QAy; w += nt) {
    const int j0 = w / P, c = w - j0 * P;
    C2<T> v[CY::RA];
#pragma unroll
    for (int i = 0; i < CY::RA; ++i) v[i] = plane[(j0 + i * QAy) * P + c];
    dif_regs<T, NY, NY, CY::RA, -1, 1>(v, j0, twy);
    if (CY::NG == 1) {
#pragma unroll
      for (int i = 0; i < CY::RA; ++i) o[bitrev<NY>(i) * P + c] = v[i];
    } else {
#pragma unroll
      for (int i = 0; i < CY::RA; ++i) plane[(j0 + i * QAy) * P + c] = v[i];
    }
  }
  if (CY::NG == 1) return;
  __syncthreads();
  constexpr int RLy = CY::RL;
  for (int w = tid; w < P * (NY / RLy); w += nt) {
    const int blk = w / P, c = w - blk * P;
    C2<T> v[RLy];
#pragma unroll
    for (int i = 0; i < RLy; ++i) v[i] = plane[(blk * RLy + i) * P + c];
    dif_regs<T, NY, RLy, RLy, -1, 1>(v, 0, twy);
#pragma unroll
    for (int i = 0; i < RLy; ++i) o[bitrev<NY>(blk * RLy + i) * P + c] = v[i];
  }
}

template <typename T, int NY, int NZ>
__global__ void __launch_bounds__(kPlaneMaxThreads)
plane_c2r_kernel(const C2<T>* __restrict__ in, T* __restrict__ out, int rows_per_chunk) {
  constexpr int H = NZ / 2, P = H + 1, P1 = H + (H >> 4) + 1;
  using CZ = Chain<H>;
  using CY = Chain<NY>;
  static_assert(CZ::NG <= 2 && CY::NG <= 2, "plane kernels use at most two radix groups per axis");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C2<T>* plane = reinterpret_cast<C2<T>*>(smem_raw);
  C2<T>* buf1 = plane + (size_t)NY * P;
  C2<T>* twz = buf1 + (CZ::NG == 2 ? (size_t)rows_per_chunk * P1 : 0);
  C2<T>* twy = twz + NZ / 2;
  const int tid = threadIdx.x, nt = blockDim.x;
  pdl_trigger();
  fill_twiddles<T, NZ>(twz, tid, nt);
  fill_twiddles<T, NY>(twy, tid, nt);
  __syncthreads();
  pdl_wait();      // the twiddles above overlap the tail of the kernel before
  const C2<T>* src = in + (int64_t)blockIdx.x * NY * P;
  // ---- y pass (DIT along the columns; bit-reversed line order read straight from global) ----
  constexpr int RLy = CY::RL, QAy = NY / CY::RA;
  for (int w = tid; w < P * (NY / RLy); w += nt) {
    const int blk = w / P, c = w - blk * P;
    C2<T> v[RLy];
#pragma unroll
    for (int i = 0; i < RLy; ++i) v[i] = src[bitrev<NY>(blk * RLy + i) * P + c];
    dit_regs<T, NY, 1, RLy, +1, 1>(v, 0, twy);
#pragma unroll
    for (int i = 0; i < RLy; ++i) plane[(blk * RLy + i) * P + c] = v[i];
  }
  __syncthreads();
  if (CY::NG == 2) {
    for (int w = tid; w < P * QAy; w += nt) {
      const int j0 = w / P, c = w - j0 * P;
      C2<T> v[CY::RA];
#pragma unroll
      for (int i = 0; i < CY::RA; ++i) v[i] = plane[(j0 + i * QAy) * P + c];
      dit_regs<T, NY, QAy, CY::RA, +1, 1>(v, j0, twy);
#pragma unroll
      for (int i = 0; i < CY::RA; ++i) plane[(j0 + i * QAy) * P + c] = v[i];
    }
    __syncthreads();
  }
  // ---- un-split in place:  Z'[k] = (X[k] + conj X[H-k]) + i conj(w^k) (X[k] - conj X[H-k]) ----
  for (int i = tid; i < NY * (H / 2 + 1); i += nt) {
    const int r = i / (H / 2 + 1), k = i - r * (H / 2 + 1);
    C2<T>* row = plane + r * P;
    if (k == 0) {
      const T x0 = row[0].x, xh = row[H].x;
      row[0] = {x0 + xh, x0 - xh};
    } else {
      const int kk = H - k;
      const C2<T> xk = row[k], xkk = row[kk];
      const C2<T> w = twz[k];
      {
        const C2<T> s = {xk.x + xkk.x, xk.y - xkk.y};
        const C2<T> d = {xk.x - xkk.x, xk.y + xkk.y};
        const C2<T> wd = {d.x * w.x + d.y * w.y, d.y * w.x - d.x * w.y};
        row[k] = {s.x - wd.y, s.y + wd.x};
      }
      if (kk != k) {
        const C2<T> s = {xkk.x + xk.x, xkk.y - xk.y};
        const C2<T> d = {xkk.x - xk.x, xkk.y + xk.y};
        const C2<T> wd = {-(d.x * w.x - d.y * w.y), -(d.x * w.y + d.y * w.x)};
        row[kk] = {s.x - wd.y, s.y + wd.x};
      }
    }
  }
  __syncthreads();
  // ---- z pass (packed half-length DIT) -------------------------------------------------------
  C2<T>* o = reinterpret_cast<C2<T>*>(out) + (int64_t)blockIdx.x * NY * H;
  constexpr int RLz = CZ::RL, QAz = H / CZ::RA;
  if (CZ::NG == 1) {
    for (int r = tid; r < NY; r += nt) {
      C2<T> v[CZ::RA];
#pragma unroll
      for (int i = 0; i < CZ::RA; ++i) v[i] = plane[r * P + bitrev<H>(i)];
      dit_regs<T, H, 1, CZ::RA, +1, 2>(v, 0, twz);
#pragma unroll
      for (int i = 0; i < CZ::RA; ++i) o[r * H + i] = v[i];
    }
    return;
  }
  for (int row0 = 0; row0 < NY; row0 += rows_per_chunk) {
    const int rows = min(rows_per_chunk, NY - row0);
    for (int w = tid; w < rows * (H / RLz); w += nt) {
      const int r = w / (H / RLz), blk = w - r * (H / RLz);
      C2<T> v[RLz];
#pragma unroll
      for (int i = 0; i < RLz; ++i) v[i] = plane[(row0 + r) * P + bitrev<H>(blk * RLz + i)];
      dit_regs<T, H, 1, RLz, +1, 2>(v, 0, twz);
#pragma unroll
      for (int i = 0; i < RLz; ++i) buf1[r * P1 + skew(blk * RLz + i)] = v[i];
    }
    __syncthreads();
    for (int w = tid; w < rows * QAz; w += nt) {
      const int r = w / QAz, j0 = w - r * QAz;
      C2<T> v[CZ::RA];
#pragma unroll
      for (int i = 0; i < CZ::RA; ++i) v[i] = buf1[r * P1 + skew(j0 + i * QAz)];
      dit_regs<T, H, QAz, CZ::RA, +1, 2>(v, j0, twz);
#pragma unroll
      for (int i = 0; i < CZ::RA; ++i) o[(row0 + r) * H + j0 + i * QAz] = v[i];
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// fused (y, z) plane transforms for planes that do not fit one SM (256 x 256 in fp32: the half-complex
// plane is 264 KB): a thread-block CLUSTER of two CTAs shares the plane, CTA h holding the rows
// [h NY/2, (h+1) NY/2) in its shared memory.  The z pass, the split and the contiguous-row radix group
// of the y pass work on local rows; the strided radix group of the y pass needs rows of both halves:
// the CTAs split the columns between them and read / write the partner's rows through DISTRIBUTED
// SHARED MEMORY (8 of the 16 operands of every item), bracketed by cluster barriers.  One global read
// and one global write per plane instead of the two round trips of the separate z / y passes.
// ---------------------------------------------------------------------------------------
template <typename T, int NY, int NZ>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPlaneMaxThreads)
plane2_r2c_kernel(const T* __restrict__ in, C2<T>* __restrict__ out, int rows_per_chunk) {
  namespace cg = cooperative_groups;
  constexpr int H = NZ / 2, P = H + 1, P1 = H + (H >> 4) + 1, RLOC = NY / 2;
  using CZ = Chain<H>;
  using CY = Chain<NY>;
  static_assert(CZ::NG == 2 && CY::NG == 2, "the cluster plane kernels use two radix groups per axis");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C2<T>* plane = reinterpret_cast<C2<T>*>(smem_raw);      // [RLOC][P]: the rows of this CTA
  C2<T>* buf1 = plane + (size_t)RLOC * P;
  C2<T>* twz = buf1 + (size_t)rows_per_chunk * P1;
  C2<T>* twy = twz + NZ / 2;
  cg::cluster_group cluster = cg::this_cluster();
  const int half = (int)cluster.block_rank();
  const int plane_id = blockIdx.x >> 1;
  C2<T>* other = cluster.map_shared_rank(plane, half ^ 1);
  const int tid = threadIdx.x, nt = blockDim.x;
  pdl_trigger();
  fill_twiddles<T, NZ>(twz, tid, nt);
  fill_twiddles<T, NY>(twy, tid, nt);
  __syncthreads();
  pdl_wait();      // the twiddles above overlap the tail of the kernel before
  const C2<T>* in2 = reinterpret_cast<const C2<T>*>(in) + ((int64_t)plane_id * NY + half * RLOC) * H;
  // ---- z pass of the local rows (packed half-length DIF, natural order into the plane) --------
  constexpr int QAz = H / CZ::RA, RL = CZ::RL;
  for (int row0 = 0; row0 < RLOC; row0 += rows_per_chunk) {
    const int rows = min(rows_per_chunk, RLOC - row0);
    for (int w = tid; w < rows * QAz; w += nt) {
      const int r = w / QAz, j0 = w - r * QAz;
      C2<T> v[CZ::RA];
#pragma unroll
      for (int i = 0; i < CZ::RA; ++i) v[i] = in2[(row0 + r) * H + j0 + i * QAz];
      dif_regs<T, H, H, CZ::RA, -1, 2>(v, j0, twz);
#pragma unroll
      for (int i = 0; i < CZ::RA; ++i) buf1[r * P1 + skew(j0 + i * QAz)] = v[i];
    }
    __syncthreads();
    for (int w = tid; w < rows * (H / RL); w += nt) {
      const int r = w / (H / RL), blk = w - r * (H / RL);
      C2<T> v[RL];
#pragma unroll
      for (int i = 0; i < RL; ++i) v[i] = buf1[r * P1 + skew(blk * RL + i)];
      dif_regs<T, H, RL, RL, -1, 2>(v, 0, twz);
#pragma unroll
      for (int i = 0; i < RL; ++i) plane[(row0 + r) * P + bitrev<H>(blk * RL + i)] = v[i];
    }
    __syncthreads();
  }
  // ---- split in place (local rows) --------------------------------------------------------------
  for (int i = tid; i < RLOC * (H / 2 + 1); i += nt) {
    const int r = i / (H / 2 + 1), k = i - r * (H / 2 + 1);
    C2<T>* row = plane + r * P;
    if (k == 0) {
      const C2<T> z0 = row[0];
      row[0] = {z0.x + z0.y, T(0)};
      row[H] = {z0.x - z0.y, T(0)};
    } else {
      const int kk = H - k;
      const C2<T> zk = row[k], zkk = row[kk];
      const C2<T> w = twz[k];
      {
        const C2<T> s = {T(0.5) * (zk.x + zkk.x), T(0.5) * (zk.y - zkk.y)};
        const C2<T> d = {T(0.5) * (zk.x - zkk.x), T(0.5) * (zk.y + zkk.y)};
        const C2<T> wd = cmul(d, w);
        row[k] = {s.x + wd.y, s.y - wd.x};
      }
      if (kk != k) {
        const C2<T> s = {T(0.5) * (zkk.x + zk.x), T(0.5) * (zkk.y - zk.y)};
        const C2<T> d = {T(0.5) * (zkk.x - zk.x), T(0.5) * (zkk.y + zk.y)};
        const C2<T> wd = {-(d.x * w.x + d.y * w.y), -(d.y * w.x - d.x * w.y)};
        row[kk] = {s.x + wd.y, s.y - wd.x};
      }
    }
  }
  cluster.sync();     // both halves of the plane are complete
  // ---- y pass, strided radix group: rows j0 + i QAy live in both CTAs; this CTA takes half of the columns
  constexpr int QAy = NY / CY::RA, RA = CY::RA;
  const int c0 = half == 0 ? 0 : (P + 1) / 2, c1 = half == 0 ? (P + 1) / 2 : P, ncol = c1 - c0;
  for (int w = tid; w < ncol * QAy; w += nt) {
    const int j0 = w / ncol, c = c0 + (w - j0 * ncol);
    C2<T> v[RA];
#pragma unroll
    for (int i = 0; i < RA; ++i) {
      const int r = j0 + i * QAy;                       // compile-time owner: rows < RLOC belong to CTA 0
      C2<T>* base = ((i * QAy >= RLOC) == (half == 1)) ? plane : other;
      v[i] = base[(r & (RLOC - 1)) * P + c];
    }
    dif_regs<T, NY, NY, RA, -1, 1>(v, j0, twy);
#pragma unroll
    for (int i = 0; i < RA; ++i) {
      const int r = j0 + i * QAy;
      C2<T>* base = ((i * QAy >= RLOC) == (half == 1)) ? plane : other;
      base[(r & (RLOC - 1)) * P + c] = v[i];
    }
  }
  cluster.sync();     // the partner has finished reading / writing this CTA's rows
  // ---- y pass, contiguous radix group on the local rows; bit reversal folded into the global row index
  C2<T>* o = out + (int64_t)plane_id * NY * P;
  constexpr int RLy = CY::RL;
  for (int w = tid; w < P * (RLOC / RLy); w += nt) {
    const int blk = w / P, c = w - blk * P;
    C2<T> v[RLy];
#pragma unroll
    for (int i = 0; i < RLy; ++i) v[i] = plane[(blk * RLy + i) * P + c];
    dif_regs<T, NY, RLy, RLy, -1, 1>(v, 0, twy);
#pragma unroll
    for (int i = 0; i < RLy; ++i) o[bitrev<NY>(half * RLOC + blk * RLy + i) * P + c] = v[i];
  }
}

template <typename T, int NY, int NZ>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPlaneMaxThreads)
plane2_c2r_kernel(const C2<T>* __restrict__ in, T* __restrict__ out, int rows_per_chunk) {
  namespace cg = cooperative_groups;
  constexpr int H = NZ / 2, P = H + 1, P1 = H + (H >> 4) + 1, RLOC = NY / 2;
  using CZ = Chain<H>;
  using CY = Chain<NY>;
  static_assert(CZ::NG == 2 && CY::NG == 2, "the cluster plane kernels use two radix groups per axis");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C2<T>* plane = reinterpret_cast<C2<T>*>(smem_raw);
  C2<T>* buf1 = plane + (size_t)RLOC * P;
  C2<T>* twz = buf1 + (size_t)rows_per_chunk * P1;
  C2<T>* twy = twz + NZ / 2;
  cg::cluster_group cluster = cg::this_cluster();
  const int half = (int)cluster.block_rank();
  const int plane_id = blockIdx.x >> 1;
  C2<T>* other = cluster.map_shared_rank(plane, half ^ 1);
  const int tid = threadIdx.x, nt = blockDim.x;
  pdl_trigger();
  fill_twiddles<T, NZ>(twz, tid, nt);
  fill_twiddles<T, NY>(twy, tid, nt);
  __syncthreads();
  pdl_wait();      // the twiddles above overlap the tail of the kernel before
  const C2<T>* src = in + (int64_t)plane_id * NY * P;
  // ---- y pass, contiguous DIT group into the local rows (bit-reversed row order read from global) ----
  constexpr int RLy = CY::RL, QAy = NY / CY::RA, RA = CY::RA;
  for (int w = tid; w < P * (RLOC / RLy); w += nt) {
    const int blk = w / P, c = w - blk * P;
    C2<T> v[RLy];
#pragma unroll
    for (int i = 0; i < RLy; ++i) v[i] = src[bitrev<NY>(half * RLOC + blk * RLy + i) * P + c];
    dit_regs<T, NY, 1, RLy, +1, 1>(v, 0, twy);
#pragma unroll
    for (int i = 0; i < RLy; ++i) plane[(blk * RLy + i) * P + c] = v[i];
  }
  cluster.sync();
  // ---- y pass, strided DIT group over the rows of both CTAs (this CTA: half of the columns) ----------
  const int c0 = half == 0 ? 0 : (P + 1) / 2, c1 = half == 0 ? (P + 1) / 2 : P, ncol = c1 - c0;
  for (int w = tid; w < ncol * QAy; w += nt) {
    const int j0 = w / ncol, c = c0 + (w - j0 * ncol);
    C2<T> v[RA];
#pragma unroll
    for (int i = 0; i < RA; ++i) {
      const int r = j0 + i * QAy;
      C2<T>* base = ((i * QAy >= RLOC) == (half == 1)) ? plane : other;
      v[i] = base[(r & (RLOC - 1)) * P + c];
    }
    dit_regs<T, NY, QAy, RA, +1, 1>(v, j0, twy);
#pragma unroll
    for (int i = 0; i < RA; ++i) {
      const int r = j0 + i * QAy;
      C2<T>* base = ((i * QAy >= RLOC) == (half == 1)) ? plane : other;
      base[(r & (RLOC - 1)) * P + c] = v[i];
    }
  }
  cluster.sync();
  // ---- un-split in place (local rows) ---------------------------------------------------------------
  for (int i = tid; i < RLOC * (H / 2 + 1); i += nt) {
    const int r = i / (H / 2 + 1), k = i - r * (H / 2 + 1);
    C2<T>* row = plane + r * P;
    if (k == 0) {
      const T x0 = row[0].x, xh = row[H].x;
      row[0] = {x0 + xh, x0 - xh};
    } else {
      const int kk = H - k;
      const C2<T> xk = row[k], xkk = row[kk];
      const C2<T> w = twz[k];
      {
        const C2<T> s = {xk.x + xkk.x, xk.y - xkk.y};
        const C2<T> d = {xk.x - xkk.x, xk.y + xkk.y};
        const C2<T> wd = {d.x * w.x + d.y * w.y, d.y * w.x - d.x * w.y};
        row[k] = {s.x - wd.y, s.y + wd.x};
      }
      if (kk != k) {
        const C2<T> s = {xkk.x + xk.x, xkk.y - xk.y};
        const C2<T> d = {xkk.x - xk.x, xkk.y + xk.y};
        const C2<T> wd = {-(d.x * w.x - d.y * w.y), -(d.x * w.y + d.y * w.x)};
        row[kk] = {s.x - wd.y, s.y + wd.x};
      }
    }
  }
  __syncthreads();
  // ---- z pass of the local rows (packed half-length DIT) -----------------------------------------------
  C2<T>* o = reinterpret_cast<C2<T>*>(out) + ((int64_t)plane_id * NY + half * RLOC) * H;
  constexpr int RLz = CZ::RL, QAz = H / CZ::RA;
  for (int row0 = 0; row0 < RLOC; row0 += rows_per_chunk) {
    const int rows = min(rows_per_chunk, RLOC - row0);
    for (int w = tid; w < rows * (H / RLz); w += nt) {
      const int r = w / (H / RLz), blk = w - r * (H / RLz);
      C2<T> v[RLz];
#pragma unroll
      for (int i = 0; i < RLz; ++i) v[i] = plane[(row0 + r) * P + bitrev<H>(blk * RLz + i)];
      dit_regs<T, H, 1, RLz, +1, 2>(v, 0, twz);
#pragma unroll
      for (int i = 0; i < RLz; ++i) buf1[r * P1 + skew(blk * RLz + i)] = v[i];
    }
    __syncthreads();
    for (int w = tid; w < rows * QAz; w += nt) {
      const int r = w / QAz, j0 = w - r * QAz;
      C2<T> v[CZ::RA];
#pragma unroll
      for (int i = 0; i < CZ::RA; ++i) v[i] = buf1[r * P1 + skew(j0 + i * QAz)];
      dit_regs<T, H, QAz, CZ::RA, +1, 2>(v, j0, twz);
#pragma unroll
      for (int i = 0; i < CZ::RA; ++i) o[(row0 + r) * H + j0 + i * QAz] = v[i];
    }
    __syncthreads();
  }
  cluster.sync();     // no CTA exits while the partner may still touch its shared memory
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
inline bool supported_dim(int n) { return n >= 8 && n <= 512 && (n & (n - 1)) == 0; }

template <typename K>
inline int allow_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024)
    TPME_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

template <typename T, int NZ>
int launch_rows(bool forward, const void* in, void* out, int64_t n_rows, cudaStream_t s) {
  constexpr int H = NZ / 2, P = H + 1;
  constexpr int items_per_row = H / Chain<H>::RA > 0 ? H / Chain<H>::RA : 1;
  // one radix group of work per thread (256 items per CTA), >= 2 CTAs per SM when possible
  int rows = 256 / items_per_row;
  if (rows < 1) rows = 1;
  const int64_t want_ctas = 2 * (int64_t)num_sms();
  while (rows > 2 && (n_rows + rows - 1) / rows < want_ctas) rows /= 2;
  const int64_t grid = (n_rows + rows - 1) / rows;
  int threads = rows * items_per_row;
  threads = threads < 64 ? 64 : (threads > 256 ? 256 : (threads + 31) / 32 * 32);
  const size_t smem = ((size_t)rows * (P + (Chain<H>::NG == 2 ? H + (H >> 4) + 1 : 0)) + NZ / 2) * sizeof(C2<T>);
  if (forward) {
    if (int rc = allow_smem(rows_r2c_kernel<T, NZ>, smem)) return rc;
    TPME_CUDA_OK(launch_pdl(rows_r2c_kernel<T, NZ>, dim3((unsigned)grid), dim3(threads), smem, s, pdl_for<T>(), (const T*)in, (C2<T>*)out, n_rows, rows));
  } else {
    if (int rc = allow_smem(rows_c2r_kernel<T, NZ>, smem)) return rc;
    TPME_CUDA_OK(launch_pdl(rows_c2r_kernel<T, NZ>, dim3((unsigned)grid), dim3(threads), smem, s, pdl_for<T>(), (const C2<T>*)in, (T*)out, n_rows, rows));
  }
  TPME_CUDA_OK(cudaGetLastError());
  return 0;
}

template <typename T>
int dispatch_rows(int nz, bool forward, const void* in, void* out, int64_t n_rows, cudaStream_t s) {
  switch (nz) {
    case 8: return launch_rows<T, 8>(forward, in, out, n_rows, s);
    case 16: return launch_rows<T, 16>(forward, in, out, n_rows, s);
    case 32: return launch_rows<T, 32>(forward, in, out, n_rows, s);
    case 64: return launch_rows<T, 64>(forward, in, out, n_rows, s);
    case 128: return launch_rows<T, 128>(forward, in, out, n_rows, s);
    case 256: return launch_rows<T, 256>(forward, in, out, n_rows, s);
    case 512: return launch_rows<T, 512>(forward, in, out, n_rows, s);
  }
  set_last_error("fft", "unsupported mesh size for the hand-written FFT");
  return 3;
}

template <typename T, typename GT, int N, int MODE, int GV>
int launch_lines(void* data, int n_outer, int n_inner, int64_t ls, int d1, int64_t s1, int64_t s0,
                 const GreenDev<GT>& green, int nx, int ny, int nz, void* dc_out, cudaStream_t s, int y_off = 0,
                 const RemoteStore* rs = nullptr) {
  // columns per CTA: one first-group radix item per thread, at least ~2 CTAs per SM when the
  // mesh is small, never narrower than 8 columns (64-byte runs)
  constexpr int items_per_col = N / Chain<N>::RA > 0 ? N / Chain<N>::RA : 1;
  constexpr int max_threads = MaxThreads<T>::value;
  int max_zc = max_threads / items_per_col;
  if (max_zc < 8) max_zc = 8;
  if (max_zc > n_inner) max_zc = n_inner;
  int n_chunks = (n_inner + max_zc - 1) / max_zc;
  const int want = 2 * num_sms();
  while (n_outer * n_chunks < want && (n_inner + n_chunks) / (n_chunks + 1) >= 8) ++n_chunks;
  const int zc = (n_inner + n_chunks - 1) / n_chunks;
  n_chunks = (n_inner + zc - 1) / zc;
  int threads = zc * items_per_col;
  threads = threads < 64 ? 64 : (threads > max_threads ? max_threads : (threads + 31) / 32 * 32);
  size_t smem = (Chain<N>::NG > 1 ? ((size_t)N * zc) : 0) * sizeof(C2<T>) + (N / 2) * sizeof(C2<T>);
  if (MODE == 2) smem += ((size_t)N + zc) * sizeof(AxisEntry<GT>);
  RemoteStore remote;
  if (rs != nullptr) remote = *rs;
  else memset(&remote, 0, sizeof(remote));
  // the variant whose final stores go to peer GPUs exists for the forward and the fused x pass only
  auto kernel = lines_fft_kernel<T, GT, N, MODE, GV, false>;
  if (rs != nullptr) {
    if (MODE == 1) { set_last_error("fft", "no remote-store variant of the inverse pass"); return 3; }
    kernel = lines_fft_kernel<T, GT, N, MODE, GV, (MODE != 1)>;
  }
  if (int rc = allow_smem(kernel, smem)) return rc;
  TPME_CUDA_OK(launch_pdl(kernel, dim3((unsigned)(n_outer * n_chunks)), dim3(threads), smem, s, pdl_for<T>(),
      (C2<T>*)data, n_inner, zc, n_chunks, ls, d1, s1, s0, green, nx, ny, nz, (T*)dc_out, y_off, remote));
  return 0;
}

template <typename T, typename GT, int MODE, int GV>
int dispatch_lines(int n, void* data, int n_outer, int n_inner, int64_t ls, int d1, int64_t s1, int64_t s0,
                   const GreenDev<GT>& green, int nx, int ny, int nz, void* dc_out, cudaStream_t s, int y_off = 0,
                   const RemoteStore* rs = nullptr) {
#define TPME_LINES(NN) \
  case NN: return launch_lines<T, GT, NN, MODE, GV>(data, n_outer, n_inner, ls, d1, s1, s0, green, nx, ny, nz, dc_out, s, y_off, rs);
  switch (n) {
    TPME_LINES(8) TPME_LINES(16) TPME_LINES(32) TPME_LINES(64) TPME_LINES(128) TPME_LINES(256) TPME_LINES(512)
  }
#undef TPME_LINES
  set_last_error("fft", "unsupported mesh size for the hand-written FFT");
  return 3;
}

// fused (y, z) plane passes; returns -1 when the plane does not fit (caller falls back)
template <typename T, int NY, int NZ>
int launch_plane(bool forward, const void* in, void* out, int n_planes, cudaStream_t s) {
  constexpr int H = NZ / 2, P = H + 1, P1 = H + (H >> 4) + 1;
  constexpr int items_per_row = H / Chain<H>::RA > 0 ? H / Chain<H>::RA : 1;
  // threads: one radix item of the y pass each, up to the launch bound (TPME_PLANE_THREADS overrides)
  static const int env_threads = [] { const char* e = getenv("TPME_PLANE_THREADS"); return e ? atoi(e) : 0; }();
  int threads = (P * (NY / Chain<NY>::RA) + 31) / 32 * 32;
  if (threads > kPlaneMaxThreads) threads = kPlaneMaxThreads;
  if (env_threads >= 32 && env_threads <= kPlaneMaxThreads) threads = env_threads / 32 * 32;
  if (threads < 64) threads = 64;
  int rows = threads / items_per_row;
  if (rows < 1) rows = 1;
  if (rows > NY) rows = NY;
  auto smem_of = [&](int r) {
    return ((size_t)NY * P + (Chain<H>::NG == 2 ? (size_t)r * P1 : 0) + NZ / 2 + NY / 2) * sizeof(C2<T>);
  };
  while (rows > 1 && smem_of(rows) > 216 * 1024) rows >>= 1;   // the exchange buffer of the z pass shrinks first
  const size_t smem = smem_of(rows);
  if (smem > 216 * 1024) return -1;
  if (forward) {
    if (int rc = allow_smem(plane_r2c_kernel<T, NY, NZ>, smem)) return rc;
    TPME_CUDA_OK(launch_pdl(plane_r2c_kernel<T, NY, NZ>, dim3(n_planes), dim3(threads), smem, s, pdl_for<T>(), (const T*)in, (C2<T>*)out, rows));
  } else {
    if (int rc = allow_smem(plane_c2r_kernel<T, NY, NZ>, smem)) return rc;
    TPME_CUDA_OK(launch_pdl(plane_c2r_kernel<T, NY, NZ>, dim3(n_planes), dim3(threads), smem, s, pdl_for<T>(), (const C2<T>*)in, (T*)out, rows));
  }
  TPME_CUDA_OK(cudaGetLastError());
  return 0;
}

// two-CTA cluster variant: half of the plane per CTA
template <typename T, int NY, int NZ>
int launch_plane2(bool forward, const void* in, void* out, int n_planes, cudaStream_t s) {
  constexpr int H = NZ / 2, P = H + 1, P1 = H + (H >> 4) + 1, RLOC = NY / 2;
  constexpr int items_per_row = H / Chain<H>::RA > 0 ? H / Chain<H>::RA : 1;
  const int threads = kPlaneMaxThreads;
  int rows = threads / items_per_row;
  if (rows > RLOC) rows = RLOC;
  auto smem_of = [&](int r) { return ((size_t)RLOC * P + (size_t)r * P1 + NZ / 2 + NY / 2) * sizeof(C2<T>); };
  while (rows > 1 && smem_of(rows) > 216 * 1024) rows >>= 1;
  const size_t smem = smem_of(rows);
  if (smem > 216 * 1024) return -1;
  if (forward) {
    if (int rc = allow_smem(plane2_r2c_kernel<T, NY, NZ>, smem)) return rc;
    plane2_r2c_kernel<T, NY, NZ><<<2 * n_planes, threads, smem, s>>>((const T*)in, (C2<T>*)out, rows);
  } else {
    if (int rc = allow_smem(plane2_c2r_kernel<T, NY, NZ>, smem)) return rc;
    plane2_c2r_kernel<T, NY, NZ><<<2 * n_planes, threads, smem, s>>>((const C2<T>*)in, (T*)out, rows);
  }
  TPME_CUDA_OK(cudaGetLastError());
  return 0;
}

template <typename T> struct ClusterPlane {   // which planes go to the two-CTA kernels
  static int run(int, int, bool, const void*, void*, int, cudaStream_t) { return -1; }
};
template <> struct ClusterPlane<float> {
  static int run(int ny, int nz, bool forward, const void* in, void* out, int n_planes, cudaStream_t s) {
    // opt-in (TPME_FFT_CLUSTER=1): measured on B200 at 256^3 (profiles/r02_summary.md) the two cluster
    // kernels take 77 us each against 69 / 87 us of the separate z / y passes they replace -- no gain yet
    static const bool off = [] { const char* e = getenv("TPME_FFT_CLUSTER"); return !(e && e[0] == '1'); }();
    if (off) return -1;
    if (ny == 256 && nz == 256) return launch_plane2<float, 256, 256>(forward, in, out, n_planes, s);
    return -1;
  }
};

template <typename T>
int dispatch_plane(int ny, int nz, bool forward, const void* in, void* out, int n_planes, cudaStream_t s) {
  {
    const int rc = ClusterPlane<T>::run(ny, nz, forward, in, out, n_planes, s);
    if (rc >= 0) return rc;
  }
#define TPME_PLANE(A, B) if (ny == A && nz == B) return launch_plane<T, A, B>(forward, in, out, n_planes, s);
  TPME_PLANE(16, 16) TPME_PLANE(32, 32) TPME_PLANE(64, 64) TPME_PLANE(128, 128)
  TPME_PLANE(32, 16) TPME_PLANE(16, 32) TPME_PLANE(64, 32) TPME_PLANE(32, 64) TPME_PLANE(128, 64) TPME_PLANE(64, 128)
#undef TPME_PLANE
  return -1;
}

// which Green variant the x pass can use for this filter
template <typename GT>
int green_variant(const GreenDev<GT>& g) {
  const bool coulomb_form = (g.kind == 1) || (g.kind == 2 && g.exponent == 1);
  if (!coulomb_form) return GV_GENERIC;
  const bool ortho = g.recip[1] == 0 && g.recip[2] == 0 && g.recip[3] == 0 && g.recip[5] == 0 &&
                     g.recip[6] == 0 && g.recip[7] == 0;
  if (ortho) return GV_ORTHO;
  return g.p3m_nodes > 0 ? GV_TRI_P3M : GV_TRI;
}

// x pass fused with G on a (channels, nx, ny_local, nz/2+1) half-complex array holding the rows
// y0 .. y0 + ny_local - 1 of the global mesh (ny_local == ny, y0 == 0: the whole mesh)
template <typename T, typename GT, bool FAST>
int x_pass_green(void* hat, int channels, int nx, int ny, int nz, int y0, int ny_local,
                 const GreenDev<GT>& green_in, void* dc_out, cudaStream_t s, const RemoteStore* rs = nullptr) {
  const int nzh = nz / 2 + 1;
  GreenDev<GT> green = green_in;
  const int gv = FAST ? green_variant(green) : (int)GV_GENERIC;
  // the fast variants evaluate amp * exp(-c k^2) / k^2; the p = 1 power law carries 1 / (c k^2)
  if (gv != GV_GENERIC && green.kind == 2) green.amplitude = green.amplitude / green.half_s2;
  // outer = (c, y), line stride ny_local * nzh
  int rc = 0;
#define TPME_XPASS(GV)                                                                                   \
  rc = dispatch_lines<T, GT, 2, GV>(nx, hat, channels * ny_local, nzh, (int64_t)ny_local * nzh, ny_local, \
                                    (int64_t)nx * ny_local * nzh, nzh, green, nx, ny, nz, dc_out, s, y0, rs)
  if (FAST && gv == GV_ORTHO) TPME_XPASS(GV_ORTHO);
  else if (FAST && gv == GV_TRI) TPME_XPASS(GV_TRI);
  else if (FAST && gv == GV_TRI_P3M) TPME_XPASS(GV_TRI_P3M);
  else TPME_XPASS(GV_GENERIC);
#undef TPME_XPASS
  return rc;
}

// (y, z) passes of `planes` (channel, x) planes: real (planes, ny, nz) -> half-complex
// (planes, ny, nz/2+1) when `forward`, the reverse otherwise (the half-complex input is destroyed).
// Fused per-plane kernel when the plane fits in shared memory, separate z / y passes otherwise.
template <typename T, typename GT>
int yz_passes(bool forward, void* real, void* hat, int planes, int ny, int nz, cudaStream_t s,
              const RemoteStore* rs = nullptr) {
  const int nzh = nz / 2 + 1;
  const int64_t rows = (int64_t)planes * ny;
  GreenDev<GT> unused{};
  if (forward) {
    // with a remote destination the separate z / y passes are used: only their y pass can push
    const int prc = rs != nullptr ? -1 : dispatch_plane<T>(ny, nz, true, real, hat, planes, s);
    if (prc >= 0) return prc;
    if (int rc = dispatch_rows<T>(nz, true, real, hat, rows, s)) return rc;
    // y pass: outer = (c, x), line stride nzh
    return dispatch_lines<T, GT, 0, GV_GENERIC>(ny, hat, planes, nzh, nzh, 1 << 30, 0, (int64_t)ny * nzh,
                                                unused, 0, ny, nz, nullptr, s, 0, rs);
  }
  const int prc = dispatch_plane<T>(ny, nz, false, hat, real, planes, s);
  if (prc >= 0) return prc;
  if (int rc = dispatch_lines<T, GT, 1, GV_GENERIC>(ny, hat, planes, nzh, nzh, 1 << 30, 0, (int64_t)ny * nzh,
                                                    unused, 0, ny, nz, nullptr, s)) return rc;
  return dispatch_rows<T>(nz, false, hat, real, rows, s);
}

// out = iFFT3(G * FFT3(in)), unnormalised; `hat` is (C, nx, ny, nz/2+1) complex scratch.
// FAST = false compiles only the generic Green variant (used for float meshes with double G).
template <typename T, typename GT, bool FAST>
int filter_pow2(const void* in, void* out, void* hat, int channels, int nx, int ny, int nz,
                const GreenDev<GT>& green, void* dc_out, cudaStream_t s) {
  if (int rc = yz_passes<T, GT>(true, const_cast<void*>(in), hat, channels * nx, ny, nz, s)) return rc;
  if (int rc = x_pass_green<T, GT, FAST>(hat, channels, nx, ny, nz, 0, ny, green, dc_out, s)) return rc;
  return yz_passes<T, GT>(false, out, hat, channels * nx, ny, nz, s);
}

}  // namespace fft
}  // namespace tpme
