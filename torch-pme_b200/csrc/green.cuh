// Green's / influence functions evaluated per k-point from the reciprocal cell (shared by the
// stand-alone multiply kernel and the fused x-pass of the hand-written FFT).
//
// Reference: generate_kvectors_for_mesh (lib/kvectors.py:24-102), Potential.lr_from_k_sq
// (potentials/coulomb.py:122-142, potentials/inversepowerlaw.py:108-141, lib/math.py:16-104),
// P3MKSpaceFilter._compute_influence (lib/kspace_filter.py:307-316,349-361).
#pragma once
#include <cmath>
#include <cstdlib>

#include "common.cuh"
#include "../../include/torchpme_b200.h"

namespace tpme {

// ---- special functions ---------------------------------------------------------------
template <typename T> struct MathFn;
template <> struct MathFn<float> {
  static __device__ __forceinline__ float exp(float x) { return expf(x); }
  static __device__ __forceinline__ float log(float x) { return logf(x); }
  static __device__ __forceinline__ float sqrt(float x) { return sqrtf(x); }
  static __device__ __forceinline__ float erfc(float x) { return erfcf(x); }
  static __device__ __forceinline__ float sin(float x) { return sinf(x); }
  static __device__ __forceinline__ float abs(float x) { return fabsf(x); }
  static __device__ __forceinline__ void sincos(float x, float* s, float* c) { sincosf(x, s, c); }
};
template <> struct MathFn<double> {
  static __device__ __forceinline__ double exp(double x) { return ::exp(x); }
  static __device__ __forceinline__ double log(double x) { return ::log(x); }
  static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
  static __device__ __forceinline__ double erfc(double x) { return ::erfc(x); }
  static __device__ __forceinline__ double sin(double x) { return ::sin(x); }
  static __device__ __forceinline__ double abs(double x) { return ::fabs(x); }
  static __device__ __forceinline__ void sincos(double x, double* s, double* c) { ::sincos(x, s, c); }
};

// Exponential integral E1, same algorithm as lib/math.py:16-60 (series for x <= 1,
// continued fraction with 20 + floor(80/x) levels above).
template <typename T>
__device__ T exp1_dev(T x) {
  using M = MathFn<T>;
  if (!(x > T(0))) return T(INFINITY);
  if (x <= T(1)) {
    T e1 = T(1), r = T(1);
    for (int k = 1; k < 26; ++k) {
      const T kp = T(k + 1);
      r = -r * T(k) * x / (kp * kp);
      e1 += r;
      if (M::abs(r) <= M::abs(e1) * T(1e-15)) break;
    }
    return T(-0.577215664901532860606512090082402431) - M::log(x) + x * e1;
  }
  const int m = 20 + (int)(T(80) / x);
  T t0 = T(0);
  for (int k = m; k > 0; --k) t0 = T(k) / (T(1) + T(k) / (x + t0));
  return M::exp(-x) / (x + t0);
}

// f_p(z) = Gamma((3-p)/2, z) / z^((3-p)/2)   (lib/math.py:79-104)
template <typename T>
__device__ T gammaincc_over_powerlaw_dev(int p, T z) {
  using M = MathFn<T>;
  const T pi = T(3.14159265358979323846);
  switch (p) {
    case 1: return M::exp(-z) / z;
    case 2: return M::sqrt(pi / z) * M::erfc(M::sqrt(z));
    case 3: return exp1_dev<T>(z);
    case 4: return T(2) * (M::exp(-z) - M::sqrt(pi * z) * M::erfc(M::sqrt(z)));
    case 5: return M::exp(-z) - z * exp1_dev<T>(z);
    default:
      return ((T(2) - T(4) * z) * M::exp(-z) + T(4) * M::sqrt(pi * z * z * z) * M::erfc(M::sqrt(z))) / T(3);
  }
}

// Natural cubic spline through (x, y) with second derivatives m, n knots, evaluated like
// lib/splines.py:CubicSpline.forward (interval = searchsorted(x, q, right=True) - 1 clamped to [0, n-2],
// so the end intervals extrapolate):  a y_i + b y_{i+1} + ((a^3 - a) m_i + (b^3 - b) m_{i+1}) h^2 / 6
__device__ __forceinline__ double cubic_spline_dev(const double* __restrict__ x, const double* __restrict__ y,
                                                   const double* __restrict__ m, int n, double q) {
  int lo = 0, hi = n;                       // first knot > q
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(x + mid) <= q) lo = mid + 1; else hi = mid;
  }
  int i = lo - 1;
  i = i < 0 ? 0 : (i > n - 2 ? n - 2 : i);
  const double x0 = __ldg(x + i), x1 = __ldg(x + i + 1);
  const double h = x1 - x0;
  const double a = (x1 - q) / h, b = 1.0 - a;
  const double curvature = (a * (a * a - 1.0)) * __ldg(m + i) + (b * (b * b - 1.0)) * __ldg(m + i + 1);
  return a * __ldg(y + i) + b * __ldg(y + i + 1) + curvature * (h * h / 6.0);
}

// SplinePotential's reciprocal-space kernel (potentials/spline.py:108-121,151-157) at k^2 = q.
// kind 3: one spline over k^2, table = [x(n), y(n), m(n)].
// kind 4: spline on the 1/k^2 axis (lib/splines.py:CubicSplineReciprocal): table = [xi(n), yi(n), mi(n)] of
//         the inverse-axis spline (its first knot is 1/x = 0) followed by the three-knot head spline
//         [xh(3), yh(3), mh(3)] that takes over below the first grid point xh[1].
__device__ __forceinline__ double spline_kernel_dev(int kind, const double* __restrict__ t, int n, double q) {
  if (kind == 3) return cubic_spline_dev(t, t + n, t + 2 * n, n, q);
  const double* head = t + 3 * n;
  const double first = __ldg(head + 1);
  if (q < first) return cubic_spline_dev(head, head + 3, head + 6, 3, q);
  return cubic_spline_dev(t, t + n, t + 2 * n, n, 1.0 / q);
}

// coefficients of the finite-difference approximations to ik (kspace_filter.py:282-293), row = order - 1
__device__ __constant__ const double kP3MDiffCoeff[6][6] = {
    {1.0, 0.0, 0.0, 0.0, 0.0, 0.0},
    {4.0 / 3, -1.0 / 3, 0.0, 0.0, 0.0, 0.0},
    {3.0 / 2, -3.0 / 5, 1.0 / 10, 0.0, 0.0, 0.0},
    {8.0 / 5, -4.0 / 5, 8.0 / 35, -1.0 / 35, 0.0, 0.0},
    {5.0 / 3, -20.0 / 21, 5.0 / 14, -5.0 / 63, 1.0 / 126, 0.0},
    {12.0 / 7, -15.0 / 14, 10.0 / 21, -1.0 / 7, 2.0 / 77, -1.0 / 465},
};

template <typename T>
struct GreenDev {
  int kind, exponent, p3m_nodes;
  int p3m_mode, diff_order;   // P3M influence function of modes 1..3 with a differential operator of order 1..6
  T recip[9];
  T spacing[3];
  T half_s2;      // smearing^2 / 2
  T amplitude;    // scale * prefactor * (4 pi | ipl prefactor)
  T k0_value;     // value at k = 0 (already scaled)
  const void* table;
};

// scale * G(k) at integer mesh frequency (ix, iy, iz) of the rFFT layout.  EXT compiles the extended kinds in
// (spline kernels, P3M influence modes 1..3): the stand-alone table / multiply kernels do, the fused x pass of
// the hand-written FFT does not (its register budget is tuned for the closed forms; extended kinds reach it
// as a table written by tpme_green_table, see is_extended_green).
template <typename T, typename S, bool EXT = false>
__device__ __forceinline__ T green_value(const GreenDev<T>& g, int ix, int iy, int iz, int nx,
                                         int ny, int nz, int64_t flat) {
  using M = MathFn<T>;
  if (g.kind == 0) return (T) reinterpret_cast<const S*>(g.table)[flat] * g.amplitude;
  // fftfreq(n) * n  (lib/kvectors.py:56-70)
  const T fx = (T)(ix < (nx + 1) / 2 ? ix : ix - nx);
  const T fy = (T)(iy < (ny + 1) / 2 ? iy : iy - ny);
  const T fz = (T)iz;
  const T kx = fx * g.recip[0] + fy * g.recip[3] + fz * g.recip[6];
  const T ky = fx * g.recip[1] + fy * g.recip[4] + fz * g.recip[7];
  const T kz = fx * g.recip[2] + fy * g.recip[5] + fz * g.recip[8];
  const T k_sq = kx * kx + ky * ky + kz * kz;
  T val;
  if (EXT && g.kind >= 3) {
    // tabulated kernel: cubic spline in k^2 (double arithmetic, the tables are double)
    val = g.amplitude * (T)spline_kernel_dev(g.kind, reinterpret_cast<const double*>(g.table), g.exponent, (double)k_sq);
  } else if (k_sq == T(0)) {
    val = g.k0_value;
  } else if (g.kind == 1 || g.exponent == 1) {
    // 4 pi exp(-s^2 k^2 / 2) / k^2   (coulomb.py:137-142); IPL p=1 is identical
    val = g.amplitude * M::exp(-g.half_s2 * k_sq) / (g.kind == 1 ? k_sq : g.half_s2 * k_sq);
  } else {
    val = g.amplitude * gammaincc_over_powerlaw_dev<T>(g.exponent, g.half_s2 * k_sq);
  }
  if (g.p3m_nodes > 0) {
    // 1 / U^2, U^2 = [prod_a sinc(k_a h_a / 2 pi)]^(2n)   (kspace_filter.py:307-316,349-361)
    const T hx = T(0.5) * kx * g.spacing[0], hy = T(0.5) * ky * g.spacing[1],
            hz = T(0.5) * kz * g.spacing[2];
    const T sx = hx == T(0) ? T(1) : M::sin(hx) / hx;
    const T sy = hy == T(0) ? T(1) : M::sin(hy) / hy;
    const T sz = hz == T(0) ? T(1) : M::sin(hz) / hz;
    const T s = sx * sy * sz;
    T u2 = T(1);
    const T s2 = s * s;
    for (int i = 0; i < g.p3m_nodes; ++i) u2 *= s2;
    if (!EXT || g.p3m_mode == 0) {
      val = (u2 == T(0)) ? T(0) : val / u2;
    } else {
      // (k . D)^mode / (U^2 |D|^(4 mode)),  D_a = sum_i c_i / (i + 1) sin((i + 1) k_a h_a) / h_a
      // (kspace_filter.py:318-347; eq. 30 of doi:10.1063/1.3000389)
      const T kk[3] = {kx, ky, kz};
      T d_sq = T(0), k_dot_d = T(0);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const T kh = kk[a] * g.spacing[a];
        T acc = T(0);
        for (int i = 0; i < g.diff_order; ++i)
          acc += (T)(kP3MDiffCoeff[g.diff_order - 1][i] / (i + 1)) * M::sin(kh * T(i + 1));
        const T d = acc / g.spacing[a];
        d_sq += d * d;
        k_dot_d += kk[a] * d;
      }
      T numer = T(1), d4 = T(1);
      for (int i = 0; i < g.p3m_mode; ++i) { numer *= k_dot_d; d4 *= d_sq * d_sq; }
      const T denom = u2 * d4;
      val = (denom == T(0)) ? T(0) : val * numer / denom;
    }
  }
  return val;
}

template <typename T>
inline GreenDev<T> make_green(const tpme_green* h) {
  GreenDev<T> g;
  g.kind = h->kind;
  g.exponent = h->exponent;
  g.p3m_nodes = h->p3m_nodes;
  g.p3m_mode = h->p3m_mode & 255;
  g.diff_order = (h->p3m_mode >> 8) & 255;
  for (int i = 0; i < 9; ++i) g.recip[i] = (T)h->recip[i];
  for (int i = 0; i < 3; ++i) g.spacing[i] = (T)h->spacing[i];
  const double s2 = h->smearing * h->smearing;
  g.half_s2 = (T)(0.5 * s2);
  g.table = h->table;
  const double pi = 3.14159265358979323846;
  double amp = h->scale, k0 = 0.0;
  if (h->kind == 1) {
    amp *= h->prefactor * 4.0 * pi;
  } else if (h->kind >= 3) {
    amp *= h->prefactor;
  } else if (h->kind == 2) {
    // prefac = pi^1.5 / Gamma(p/2) (2 s^2)^((3-p)/2)   (inversepowerlaw.py:121-125)
    const double p = h->exponent;
    const double peff = (3.0 - p) / 2.0;
    const double pre = pow(pi, 1.5) / tgamma(p / 2.0) * pow(2.0 * s2, peff);
    amp *= h->prefactor * pre;
    if (h->exponent > 3) k0 = h->scale * h->prefactor * (-pre / peff);  // :134-137
  }
  g.amplitude = (T)amp;
  g.k0_value = (T)k0;
  return g;
}

// kinds that only the EXT instantiation of green_value evaluates
inline bool is_extended_green(const tpme_green* g) {
  return g->kind >= 3 || (g->p3m_nodes > 0 && (g->p3m_mode & 255) != 0);
}

inline int check_green(const tpme_green* g) {
  TPME_REQUIRE(g != nullptr, "green parameters missing");
  TPME_REQUIRE(g->kind >= 0 && g->kind <= 4,
               "green kind must be 0 (table), 1 (coulomb), 2 (ipl), 3 (spline) or 4 (reciprocal-axis spline)");
  TPME_REQUIRE((g->kind != 0 && g->kind < 3) || g->table != nullptr, "table / spline kinds need a table pointer");
  TPME_REQUIRE(g->kind < 3 || g->exponent >= 2, "spline kinds carry the number of knots (>= 2) in `exponent`");
  TPME_REQUIRE(g->kind != 2 || (g->exponent >= 1 && g->exponent <= 6), "Unsupported exponent");
  TPME_REQUIRE(g->p3m_nodes >= 0 && g->p3m_nodes <= 7, "bad p3m_nodes");
  TPME_REQUIRE((g->p3m_mode & 255) <= 3 && ((g->p3m_mode & 255) == 0 || (((g->p3m_mode >> 8) & 255) >= 1 && ((g->p3m_mode >> 8) & 255) <= 6)),
               "P3M mode must be 0..3 with a differential order 1..6");
  return 0;
}

// IPL with p >= 2 has cancellations (erfc / E1 differences) -> evaluate G in double even
// for float meshes; everything else uses the storage precision like the reference.
// Evaluate G in double for float meshes?  Only on request (TPME_GREEN_FP64=1): the closed forms
// of the inverse power laws lose relative accuracy in fp32 only where exp(-z) has already made
// G negligible (z >~ 20), and the reference evaluates them in the mesh precision too.
inline bool needs_double_math(const tpme_green* g) {
  static const bool forced = [] { const char* e = getenv("TPME_GREEN_FP64"); return e && e[0] == '1'; }();
  return forced && g->kind == 2 && g->exponent >= 2;
}


}  // namespace tpme
