// Real-space pair sum over the neighbor list, forward and analytic backward.
//
// Replaces Calculator._compute_rspace (src/torchpme/calculators/calculator.py:43-87) and
// Potential.sr_from_dist (potentials/potential.py:106-138) with one pass over the pairs.
// The short-range kernel is evaluated in its cancellation-free closed form
// Q(p/2, d^2 / 2 s^2) / d^p instead of "full - long range"
// (coulomb.py:80-120, inversepowerlaw.py:54-106).
#include <cmath>

#include "common.cuh"
#include "../../include/torchpme_b200.h"

namespace tpme {

template <typename T>
struct PairPot {
  int kind, exponent, exclusion_degree;
  T inv_2s2;      // 1 / (2 s^2)
  T inv_s2;       // 1 / s^2
  T prefactor;
  T inv_gamma;    // 1 / Gamma(p/2)
  T exclusion_radius;  // <= 0: unset
};

__device__ __forceinline__ float exp_t(float x) { return expf(x); }
__device__ __forceinline__ double exp_t(double x) { return exp(x); }
__device__ __forceinline__ float erfc_t(float x) { return erfcf(x); }
__device__ __forceinline__ double erfc_t(double x) { return erfc(x); }
__device__ __forceinline__ float erf_t(float x) { return erff(x); }
__device__ __forceinline__ double erf_t(double x) { return erf(x); }
__device__ __forceinline__ float sqrt_t(float x) { return sqrtf(x); }
__device__ __forceinline__ double sqrt_t(double x) { return sqrt(x); }
__device__ __forceinline__ void sincospi_t(float x, float* s, float* c) { sincospif(x, s, c); }
__device__ __forceinline__ void sincospi_t(double x, double* s, double* c) { sincospi(x, s, c); }

template <typename T>
__device__ __forceinline__ T ipow(T x, int n) {
  T r = T(1);
  for (int i = 0; i < n; ++i) r *= x;
  return r;
}

// v_SR(d) and dv_SR/dd (both times prefactor).  p = 1 for Coulomb.
template <typename T, bool DERIV>
__device__ __forceinline__ void short_range(const PairPot<T>& pp, T d, T& v, T& dv) {
  const int p = pp.kind == 1 ? 1 : pp.exponent;
  const T x = d * d * pp.inv_2s2;
  const T ex = exp_t(-x);
  const T inv_sqrt_pi = T(0.56418958354775628695);
  T q;          // regularised upper incomplete gamma Q(p/2, x)
  T xpow;       // x^(p/2 - 1)
  const T sx = sqrt_t(x);
  switch (p) {
    case 1: q = erfc_t(sx); xpow = T(1) / sx; break;
    case 2: q = ex; xpow = T(1); break;
    case 3: q = erfc_t(sx) + T(2) * inv_sqrt_pi * sx * ex; xpow = sx; break;
    case 4: q = ex * (T(1) + x); xpow = x; break;
    case 5: q = erfc_t(sx) + T(2) * inv_sqrt_pi * sx * ex * (T(1) + T(2) * x / T(3)); xpow = x * sx; break;
    default: q = ex * (T(1) + x + T(0.5) * x * x); xpow = x * x; break;
  }
  const T inv_d = T(1) / d;
  const T inv_dp = ipow(inv_d, p);
  T sr = q * inv_dp;
  T dsr = T(0);
  if (DERIV) {
    const T dq = -xpow * ex * pp.inv_gamma * d * pp.inv_s2;
    dsr = (dq - T(p) * q * inv_d) * inv_dp;
  }
  if (pp.exclusion_radius > T(0)) {
    // v = -v_LR f_cut,  v_LR = d^-p - v_SR,  f_cut = 1 - ((1 - cos(pi d / rc)) / 2)^deg  (potential.py:78-88,136-138)
    T lr, dlr = T(0);
    if (p == 1) {
      const T a = sqrt_t(pp.inv_2s2);
      lr = erf_t(a * d) * inv_d;
      if (DERIV) dlr = T(2) * inv_sqrt_pi * a * ex * inv_d - lr * inv_d;
    } else {
      lr = inv_dp - sr;
      if (DERIV) dlr = -T(p) * inv_dp * inv_d - dsr;
    }
    T f = T(0), df = T(0);
    if (d < pp.exclusion_radius) {
      T s, c;
      sincospi_t(d / pp.exclusion_radius, &s, &c);
      const T h = T(0.5) * (T(1) - c);
      f = T(1) - ipow(h, pp.exclusion_degree);
      if (DERIV)
        df = -T(pp.exclusion_degree) * ipow(h, pp.exclusion_degree - 1) * T(0.5) * s *
             T(3.14159265358979323846) / pp.exclusion_radius;
    }
    sr = -lr * f;
    if (DERIV) dsr = -(dlr * f + lr * df);
  }
  v = pp.prefactor * sr;
  dv = pp.prefactor * dsr;
}

// Fast path: plain 1/r (Coulomb or p = 1 power law), no exclusion zone.
//   v = pref erfc(a d) / d,   dv/dd = -(v + pref 2a/sqrt(pi) exp(-a^2 d^2)) / d,   a = 1/(s sqrt 2)
template <typename T> struct PairFast;
template <> struct PairFast<float> {
  static __device__ __forceinline__ float rcp(float x) { return __frcp_rn(x); }
  static __device__ __forceinline__ float exp(float x) { return __expf(x); }
};
template <> struct PairFast<double> {
  static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
  static __device__ __forceinline__ double exp(double x) { return ::exp(x); }
};

template <typename T, bool DERIV>
__device__ __forceinline__ void coulomb_short_range(T a, T pref, T d, T& v, T& dv) {
  const T ad = a * d;
  const T inv_d = PairFast<T>::rcp(d);
  v = pref * erfc_t(ad) * inv_d;
  if (DERIV)
    dv = -(v + pref * T(1.1283791670955125739) * a * PairFast<T>::exp(-ad * ad)) * inv_d;
}

template <typename I> struct IndexPair;
template <> struct IndexPair<int64_t> {
  static __device__ __forceinline__ void load(const int64_t* idx, int64_t p, int64_t& i, int64_t& j) {
    const longlong2 v = *reinterpret_cast<const longlong2*>(idx + 2 * p);
    i = v.x; j = v.y;
  }
};
template <> struct IndexPair<int32_t> {
  static __device__ __forceinline__ void load(const int32_t* idx, int64_t p, int64_t& i, int64_t& j) {
    const int2 v = *reinterpret_cast<const int2*>(idx + 2 * p);
    i = v.x; j = v.y;
  }
};

// single-channel Coulomb-type potential: the common case, no channel loop, no branches on kind
template <typename T, typename I, bool HALF>
__global__ void __launch_bounds__(256)
pair_forward_coulomb_kernel(const T* __restrict__ charges, const I* __restrict__ idx,
                            const T* __restrict__ dist, const uint8_t* __restrict__ mask,
                            int64_t n_pairs, const int64_t* __restrict__ n_dev, T a, T half_pref,
                            T* __restrict__ out) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs || (n_dev != nullptr && p >= *n_dev)) return;
  if (mask != nullptr && mask[p] == 0) return;
  int64_t i, j;
  IndexPair<I>::load(idx, p, i, j);
  T v, dv;
  coulomb_short_range<T, false>(a, half_pref, dist[p], v, dv);
  red_add(out + i, __ldg(charges + j) * v);
  if (HALF) red_add(out + j, __ldg(charges + i) * v);
}

template <typename T, typename I, bool HALF>
__global__ void __launch_bounds__(256)
pair_backward_coulomb_kernel(const T* __restrict__ charges, const I* __restrict__ idx,
                             const T* __restrict__ dist, const uint8_t* __restrict__ mask,
                             const T* __restrict__ grad_out, int64_t n_pairs,
                             const int64_t* __restrict__ n_dev, T a, T half_pref,
                             T* __restrict__ grad_charges, T* __restrict__ grad_pairs) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs || (n_dev != nullptr && p >= *n_dev)) return;
  if (mask != nullptr && mask[p] == 0) {
    if (grad_pairs) grad_pairs[p] = T(0);
    return;
  }
  int64_t i, j;
  IndexPair<I>::load(idx, p, i, j);
  T v, dv;
  coulomb_short_range<T, true>(a, half_pref, dist[p], v, dv);
  const T gi = __ldg(grad_out + i);
  T acc = gi * __ldg(charges + j);
  if (grad_charges) red_add(grad_charges + j, gi * v);
  if (HALF) {
    const T gj = __ldg(grad_out + j);
    acc = fma_t(gj, __ldg(charges + i), acc);
    if (grad_charges) red_add(grad_charges + i, gj * v);
  }
  if (grad_pairs) grad_pairs[p] = acc * dv;
}

template <typename T, typename I>
__global__ void __launch_bounds__(256)
pair_forward_kernel(const T* __restrict__ charges, const I* __restrict__ idx,
                    const T* __restrict__ dist, const T* __restrict__ pair_values,
                    const uint8_t* __restrict__ mask, int64_t n_pairs,
                    const int64_t* __restrict__ n_dev, int n_channels,
                    int full_list, PairPot<T> pp, T* __restrict__ out) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs || (n_dev != nullptr && p >= *n_dev)) return;
  if (mask != nullptr && mask[p] == 0) return;
  const int64_t i = (int64_t)idx[2 * p], j = (int64_t)idx[2 * p + 1];
  T v, dv;
  if (pp.kind == 0) v = pair_values[p];
  else short_range<T, false>(pp, dist[p], v, dv);
  v *= T(0.5);  // the final "/ 2" of calculator.py:87
  for (int c = 0; c < n_channels; ++c) {
    red_add(out + i * n_channels + c, charges[j * n_channels + c] * v);
    if (!full_list) red_add(out + j * n_channels + c, charges[i * n_channels + c] * v);
  }
}

template <typename T, typename I>
__global__ void __launch_bounds__(256)
pair_backward_kernel(const T* __restrict__ charges, const I* __restrict__ idx,
                     const T* __restrict__ dist, const T* __restrict__ pair_values,
                     const uint8_t* __restrict__ mask, const T* __restrict__ grad_out,
                     int64_t n_pairs, const int64_t* __restrict__ n_dev, int n_channels, int full_list,
                     PairPot<T> pp, T* __restrict__ grad_charges, T* __restrict__ grad_pairs) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs || (n_dev != nullptr && p >= *n_dev)) return;
  if (mask != nullptr && mask[p] == 0) {
    if (grad_pairs) grad_pairs[p] = T(0);
    return;
  }
  const int64_t i = (int64_t)idx[2 * p], j = (int64_t)idx[2 * p + 1];
  T v, dv = T(1);
  if (pp.kind == 0) v = pair_values[p];
  else short_range<T, true>(pp, dist[p], v, dv);
  T acc = T(0);
  for (int c = 0; c < n_channels; ++c) {
    const T gi = grad_out[i * n_channels + c];
    const T qj = charges[j * n_channels + c];
    acc = fma_t(gi, qj, acc);
    if (grad_charges) red_add(grad_charges + j * n_channels + c, T(0.5) * gi * v);
    if (!full_list) {
      const T gj = grad_out[j * n_channels + c];
      const T qi = charges[i * n_channels + c];
      acc = fma_t(gj, qi, acc);
      if (grad_charges) red_add(grad_charges + i * n_channels + c, T(0.5) * gj * v);
    }
  }
  if (grad_pairs) grad_pairs[p] = T(0.5) * acc * dv;
}

template <typename T>
static PairPot<T> make_pair_pot(const tpme_pair_potential* h) {
  PairPot<T> pp;
  pp.kind = h->kind;
  pp.exponent = h->kind == 1 ? 1 : h->exponent;
  pp.exclusion_degree = h->exclusion_degree;
  const double s2 = h->smearing * h->smearing;
  pp.inv_2s2 = (T)(h->kind == 0 ? 0.0 : 0.5 / s2);
  pp.inv_s2 = (T)(h->kind == 0 ? 0.0 : 1.0 / s2);
  pp.prefactor = (T)h->prefactor;
  pp.inv_gamma = (T)(h->kind == 0 ? 1.0 : 1.0 / tgamma(0.5 * pp.exponent));
  pp.exclusion_radius = (T)h->exclusion_radius;
  return pp;
}

static bool is_plain_coulomb(const tpme_pair_potential* h) {
  return (h->kind == 1 || (h->kind == 2 && h->exponent == 1)) && !(h->exclusion_radius > 0);
}

static int check_pair(const tpme_pair_potential* h, const void* dist, const void* values) {
  TPME_REQUIRE(h != nullptr, "pair potential missing");
  TPME_REQUIRE(h->kind >= 0 && h->kind <= 2, "pair kind must be 0, 1 or 2");
  if (h->kind == 0) {
    TPME_REQUIRE(values != nullptr, "kind 0 needs pair_values");
  } else {
    TPME_REQUIRE(dist != nullptr, "distances missing");
    TPME_REQUIRE(h->smearing > 0, "smearing must be positive");
    TPME_REQUIRE(h->kind == 1 || (h->exponent >= 1 && h->exponent <= 6), "Unsupported exponent");
  }
  return 0;
}

}  // namespace tpme

using namespace tpme;

extern "C" int tpme_pair_forward(int dtype, const void* charges, const void* neighbor_indices,
                                 int index_is_int64, const void* distances,
                                 const void* pair_values, const uint8_t* pair_mask,
                                 int64_t n_pairs, const int64_t* n_pairs_dev, int64_t n_atoms, int n_channels,
                                 int full_neighbor_list, const tpme_pair_potential* pot,
                                 void* out, void* stream) {
  (void)n_atoms;
  if (int rc = check_pair(pot, distances, pair_values)) return rc;
  TPME_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 or 1");
  if (n_pairs == 0 || n_channels == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((n_pairs + 255) / 256);
  if (is_plain_coulomb(pot) && n_channels == 1) {
    const double a = 1.0 / (pot->smearing * sqrt(2.0)), hp = 0.5 * pot->prefactor;
#define GOC(T, I, H)                                                                                  \
  pair_forward_coulomb_kernel<T, I, H><<<grid, 256, 0, s>>>((const T*)charges, (const I*)neighbor_indices, \
      (const T*)distances, pair_mask, n_pairs, n_pairs_dev, (T)a, (T)hp, (T*)out)
#define GOCI(T, I) do { if (full_neighbor_list) GOC(T, I, false); else GOC(T, I, true); } while (0)
    if (dtype == 0) { if (index_is_int64) GOCI(float, int64_t); else GOCI(float, int32_t); }
    else            { if (index_is_int64) GOCI(double, int64_t); else GOCI(double, int32_t); }
#undef GOCI
#undef GOC
    TPME_CUDA_OK(cudaGetLastError());
    return 0;
  }
#define GO(T, I)                                                                              \
  pair_forward_kernel<T, I><<<grid, 256, 0, s>>>((const T*)charges, (const I*)neighbor_indices, \
      (const T*)distances, (const T*)pair_values, pair_mask, n_pairs, n_pairs_dev, n_channels, \
      full_neighbor_list, make_pair_pot<T>(pot), (T*)out)
  if (dtype == 0) { if (index_is_int64) GO(float, int64_t); else GO(float, int32_t); }
  else            { if (index_is_int64) GO(double, int64_t); else GO(double, int32_t); }
#undef GO
  TPME_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int tpme_pair_backward(int dtype, const void* charges, const void* neighbor_indices,
                                  int index_is_int64, const void* distances,
                                  const void* pair_values, const uint8_t* pair_mask,
                                  const void* grad_out, int64_t n_pairs, const int64_t* n_pairs_dev,
                                  int64_t n_atoms,
                                  int n_channels, int full_neighbor_list,
                                  const tpme_pair_potential* pot, void* grad_charges,
                                  void* grad_pairs, void* stream) {
  (void)n_atoms;
  if (int rc = check_pair(pot, distances, pair_values)) return rc;
  TPME_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 or 1");
  if (n_pairs == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((n_pairs + 255) / 256);
  if (is_plain_coulomb(pot) && n_channels == 1) {
    const double a = 1.0 / (pot->smearing * sqrt(2.0)), hp = 0.5 * pot->prefactor;
#define GOC(T, I, H)                                                                                   \
  pair_backward_coulomb_kernel<T, I, H><<<grid, 256, 0, s>>>((const T*)charges, (const I*)neighbor_indices, \
      (const T*)distances, pair_mask, (const T*)grad_out, n_pairs, n_pairs_dev, (T)a, (T)hp, (T*)grad_charges, \
      (T*)grad_pairs)
#define GOCI(T, I) do { if (full_neighbor_list) GOC(T, I, false); else GOC(T, I, true); } while (0)
    if (dtype == 0) { if (index_is_int64) GOCI(float, int64_t); else GOCI(float, int32_t); }
    else            { if (index_is_int64) GOCI(double, int64_t); else GOCI(double, int32_t); }
#undef GOCI
#undef GOC
    TPME_CUDA_OK(cudaGetLastError());
    return 0;
  }
#define GO(T, I)                                                                               \
  pair_backward_kernel<T, I><<<grid, 256, 0, s>>>((const T*)charges, (const I*)neighbor_indices, \
      (const T*)distances, (const T*)pair_values, pair_mask, (const T*)grad_out, n_pairs,       \
      n_pairs_dev, n_channels, full_neighbor_list, make_pair_pot<T>(pot), (T*)grad_charges, (T*)grad_pairs)
  if (dtype == 0) { if (index_is_int64) GO(float, int64_t); else GO(float, int32_t); }
  else            { if (index_is_int64) GO(double, int64_t); else GO(double, int32_t); }
#undef GO
  TPME_CUDA_OK(cudaGetLastError());
  return 0;
}
