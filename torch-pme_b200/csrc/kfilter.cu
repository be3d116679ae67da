// Reciprocal-space stage: real 3-D FFT (cuFFT), Green's / influence function evaluated on
// the fly from the reciprocal cell (no k-vector tensors), inverse FFT.
//
// Replaces KSpaceFilter / P3MKSpaceFilter (src/torchpme/lib/kspace_filter.py:97-197,293-361),
// generate_kvectors_for_mesh (lib/kvectors.py:24-102) and Potential.lr_from_k_sq
// (potentials/coulomb.py:122-142, potentials/inversepowerlaw.py:108-141, lib/math.py:16-104).
#include <cufft.h>

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
};
template <> struct MathFn<double> {
  static __device__ __forceinline__ double exp(double x) { return ::exp(x); }
  static __device__ __forceinline__ double log(double x) { return ::log(x); }
  static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
  static __device__ __forceinline__ double erfc(double x) { return ::erfc(x); }
  static __device__ __forceinline__ double sin(double x) { return ::sin(x); }
  static __device__ __forceinline__ double abs(double x) { return ::fabs(x); }
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

template <typename T>
struct GreenDev {
  int kind, exponent, p3m_nodes;
  T recip[9];
  T spacing[3];
  T half_s2;      // smearing^2 / 2
  T amplitude;    // scale * prefactor * (4 pi | ipl prefactor)
  T k0_value;     // value at k = 0 (already scaled)
  const void* table;
};

// scale * G(k) at integer mesh frequency (ix, iy, iz) of the rFFT layout
template <typename T, typename S>
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
  if (k_sq == T(0)) {
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
    val = (u2 == T(0)) ? T(0) : val / u2;
  }
  return val;
}

// S = storage type of the mesh, T = arithmetic type of the Green's function
template <typename S, typename T>
__global__ void __launch_bounds__(256)
green_multiply_kernel(S* __restrict__ hat, int n_channels, int nx, int ny, int nz, GreenDev<T> g,
                      S* __restrict__ dc_out) {
  const int nzh = nz / 2 + 1;
  const int64_t total = (int64_t)nx * ny * nzh;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int iz = (int)(i % nzh);
    const int64_t t = i / nzh;
    const int iy = (int)(t % ny);
    const int ix = (int)(t / ny);
    const S gv = (S)green_value<T, S>(g, ix, iy, iz, nx, ny, nz, i);
    for (int c = 0; c < n_channels; ++c) {
      S* p = hat + 2 * (c * total + i);
      if (i == 0 && dc_out != nullptr) dc_out[c] = p[0];
      if (sizeof(S) == 4) {
        float2 v = *reinterpret_cast<float2*>(p);
        v.x *= gv; v.y *= gv;
        *reinterpret_cast<float2*>(p) = v;
      } else {
        double2 v = *reinterpret_cast<double2*>(p);
        v.x *= gv; v.y *= gv;
        *reinterpret_cast<double2*>(p) = v;
      }
    }
  }
}

template <typename S, typename T>
__global__ void __launch_bounds__(256)
green_table_kernel(S* __restrict__ out, int nx, int ny, int nz, GreenDev<T> g) {
  const int nzh = nz / 2 + 1;
  const int64_t total = (int64_t)nx * ny * nzh;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int iz = (int)(i % nzh);
    const int64_t t = i / nzh;
    out[i] = (S)green_value<T, S>(g, (int)(t / ny), (int)(t % ny), iz, nx, ny, nz, i);
  }
}

template <typename S>
__global__ void __launch_bounds__(256)
table_vjp_kernel(const S* __restrict__ x_hat, const S* __restrict__ y_hat, int n_channels, int nx,
                 int ny, int nz, S scale, S* __restrict__ grad_table) {
  const int nzh = nz / 2 + 1;
  const int64_t total = (int64_t)nx * ny * nzh;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int iz = (int)(i % nzh);
    const bool self_conj = (iz == 0) || (nz % 2 == 0 && iz == nz / 2);
    S acc = S(0);
    for (int c = 0; c < n_channels; ++c) {
      const S* x = x_hat + 2 * (c * total + i);
      const S* y = y_hat + 2 * (c * total + i);
      acc += x[0] * y[0] + x[1] * y[1];
    }
    grad_table[i] = acc * scale * (self_conj ? S(1) : S(2));
  }
}

template <typename T>
static GreenDev<T> make_green(const tpme_green* h) {
  GreenDev<T> g;
  g.kind = h->kind;
  g.exponent = h->exponent;
  g.p3m_nodes = h->p3m_nodes;
  for (int i = 0; i < 9; ++i) g.recip[i] = (T)h->recip[i];
  for (int i = 0; i < 3; ++i) g.spacing[i] = (T)h->spacing[i];
  const double s2 = h->smearing * h->smearing;
  g.half_s2 = (T)(0.5 * s2);
  g.table = h->table;
  const double pi = 3.14159265358979323846;
  double amp = h->scale, k0 = 0.0;
  if (h->kind == 1) {
    amp *= h->prefactor * 4.0 * pi;
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

static int green_grid(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

static int check_green(const tpme_green* g) {
  TPME_REQUIRE(g != nullptr, "green parameters missing");
  TPME_REQUIRE(g->kind >= 0 && g->kind <= 2, "green kind must be 0 (table), 1 (coulomb) or 2 (ipl)");
  TPME_REQUIRE(g->kind != 0 || g->table != nullptr, "table kind needs a table pointer");
  TPME_REQUIRE(g->kind != 2 || (g->exponent >= 1 && g->exponent <= 6), "Unsupported exponent");
  TPME_REQUIRE(g->p3m_nodes >= 0 && g->p3m_nodes <= 7, "bad p3m_nodes");
  return 0;
}

// IPL with p >= 2 has cancellations (erfc / E1 differences) -> evaluate G in double even
// for float meshes; everything else uses the storage precision like the reference.
static bool needs_double_math(const tpme_green* g) { return g->kind == 2 && g->exponent >= 2; }

static const char* cufft_err(cufftResult r) {
  switch (r) {
    case CUFFT_SUCCESS: return "CUFFT_SUCCESS";
    case CUFFT_INVALID_PLAN: return "CUFFT_INVALID_PLAN";
    case CUFFT_ALLOC_FAILED: return "CUFFT_ALLOC_FAILED";
    case CUFFT_INVALID_VALUE: return "CUFFT_INVALID_VALUE";
    case CUFFT_INTERNAL_ERROR: return "CUFFT_INTERNAL_ERROR";
    case CUFFT_EXEC_FAILED: return "CUFFT_EXEC_FAILED";
    case CUFFT_SETUP_FAILED: return "CUFFT_SETUP_FAILED";
    case CUFFT_INVALID_SIZE: return "CUFFT_INVALID_SIZE";
    default: return "CUFFT error";
  }
}
#define TPME_CUFFT_OK(expr)                                   \
  do {                                                        \
    cufftResult r__ = (expr);                                 \
    if (r__ != CUFFT_SUCCESS) {                               \
      ::tpme::set_last_error(#expr, cufft_err(r__));          \
      return 200 + (int)r__;                                  \
    }                                                         \
  } while (0)

}  // namespace tpme

struct tpme_fft_plan_s {
  cufftHandle fwd = 0, inv = 0;
  int dtype = 0, nx = 0, ny = 0, nz = 0, batch = 0;
};

using namespace tpme;

extern "C" int tpme_fft_plan_create(tpme_fft_plan* plan, int dtype, int nx, int ny, int nz,
                                    int batch) {
  TPME_REQUIRE(plan != nullptr, "null plan pointer");
  TPME_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 or 1");
  TPME_REQUIRE(nx > 0 && ny > 0 && nz > 0 && batch > 0, "bad plan dimensions");
  tpme_fft_plan_s* p = new tpme_fft_plan_s();
  p->dtype = dtype; p->nx = nx; p->ny = ny; p->nz = nz; p->batch = batch;
  int n[3] = {nx, ny, nz};
  const int nzh = nz / 2 + 1;
  const int rdist = nx * ny * nz, cdist = nx * ny * nzh;
  int rembed[3] = {nx, ny, nz}, cembed[3] = {nx, ny, nzh};
  cufftResult r1 = cufftPlanMany(&p->fwd, 3, n, rembed, 1, rdist, cembed, 1, cdist,
                                 dtype == 0 ? CUFFT_R2C : CUFFT_D2Z, batch);
  if (r1 != CUFFT_SUCCESS) { delete p; set_last_error("cufftPlanMany(fwd)", cufft_err(r1)); return 200 + (int)r1; }
  cufftResult r2 = cufftPlanMany(&p->inv, 3, n, cembed, 1, cdist, rembed, 1, rdist,
                                 dtype == 0 ? CUFFT_C2R : CUFFT_Z2D, batch);
  if (r2 != CUFFT_SUCCESS) { cufftDestroy(p->fwd); delete p; set_last_error("cufftPlanMany(inv)", cufft_err(r2)); return 200 + (int)r2; }
  *plan = p;
  return 0;
}

extern "C" int tpme_fft_plan_destroy(tpme_fft_plan plan) {
  if (!plan) return 0;
  cufftDestroy(plan->fwd);
  cufftDestroy(plan->inv);
  delete plan;
  return 0;
}

extern "C" int tpme_rfft3(tpme_fft_plan plan, const void* mesh, void* mesh_hat, void* stream) {
  TPME_REQUIRE(plan != nullptr, "null plan");
  TPME_CUFFT_OK(cufftSetStream(plan->fwd, (cudaStream_t)stream));
  if (plan->dtype == 0)
    TPME_CUFFT_OK(cufftExecR2C(plan->fwd, (cufftReal*)mesh, (cufftComplex*)mesh_hat));
  else
    TPME_CUFFT_OK(cufftExecD2Z(plan->fwd, (cufftDoubleReal*)mesh, (cufftDoubleComplex*)mesh_hat));
  return 0;
}

extern "C" int tpme_irfft3(tpme_fft_plan plan, void* mesh_hat, void* mesh, void* stream) {
  TPME_REQUIRE(plan != nullptr, "null plan");
  TPME_CUFFT_OK(cufftSetStream(plan->inv, (cudaStream_t)stream));
  if (plan->dtype == 0)
    TPME_CUFFT_OK(cufftExecC2R(plan->inv, (cufftComplex*)mesh_hat, (cufftReal*)mesh));
  else
    TPME_CUFFT_OK(cufftExecZ2D(plan->inv, (cufftDoubleComplex*)mesh_hat, (cufftDoubleReal*)mesh));
  return 0;
}

extern "C" int tpme_green_multiply(int dtype, void* mesh_hat, int n_channels, int nx, int ny,
                                   int nz, const tpme_green* green, void* dc_out, void* stream) {
  if (int rc = check_green(green)) return rc;
  TPME_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 or 1");
  const int64_t total = (int64_t)nx * ny * (nz / 2 + 1);
  if (total == 0 || n_channels == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = green_grid(total);
  if (dtype == 1)
    green_multiply_kernel<double, double><<<grid, 256, 0, s>>>((double*)mesh_hat, n_channels, nx, ny, nz, make_green<double>(green), (double*)dc_out);
  else if (needs_double_math(green))
    green_multiply_kernel<float, double><<<grid, 256, 0, s>>>((float*)mesh_hat, n_channels, nx, ny, nz, make_green<double>(green), (float*)dc_out);
  else
    green_multiply_kernel<float, float><<<grid, 256, 0, s>>>((float*)mesh_hat, n_channels, nx, ny, nz, make_green<float>(green), (float*)dc_out);
  TPME_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int tpme_green_table(int dtype, void* table_out, int nx, int ny, int nz,
                                const tpme_green* green, void* stream) {
  if (int rc = check_green(green)) return rc;
  TPME_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 or 1");
  const int64_t total = (int64_t)nx * ny * (nz / 2 + 1);
  if (total == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = green_grid(total);
  if (dtype == 1)
    green_table_kernel<double, double><<<grid, 256, 0, s>>>((double*)table_out, nx, ny, nz, make_green<double>(green));
  else if (needs_double_math(green))
    green_table_kernel<float, double><<<grid, 256, 0, s>>>((float*)table_out, nx, ny, nz, make_green<double>(green));
  else
    green_table_kernel<float, float><<<grid, 256, 0, s>>>((float*)table_out, nx, ny, nz, make_green<float>(green));
  TPME_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int tpme_green_table_vjp(int dtype, const void* x_hat, const void* y_hat,
                                    int n_channels, int nx, int ny, int nz, double scale,
                                    void* grad_table, void* stream) {
  TPME_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 or 1");
  const int64_t total = (int64_t)nx * ny * (nz / 2 + 1);
  if (total == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = green_grid(total);
  if (dtype == 1)
    table_vjp_kernel<double><<<grid, 256, 0, s>>>((const double*)x_hat, (const double*)y_hat, n_channels, nx, ny, nz, scale, (double*)grad_table);
  else
    table_vjp_kernel<float><<<grid, 256, 0, s>>>((const float*)x_hat, (const float*)y_hat, n_channels, nx, ny, nz, (float)scale, (float*)grad_table);
  TPME_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int tpme_kfilter_apply(tpme_fft_plan plan, const void* mesh_in, void* mesh_out,
                                  void* work_hat, void* keep_hat, const tpme_green* green,
                                  void* dc_out, void* stream) {
  TPME_REQUIRE(plan != nullptr, "null plan");
  if (int rc = tpme_rfft3(plan, mesh_in, work_hat, stream)) return rc;
  if (keep_hat != nullptr) {
    const size_t bytes = (size_t)(plan->dtype ? 16 : 8) * plan->batch * plan->nx * plan->ny * (plan->nz / 2 + 1);
    TPME_CUDA_OK(cudaMemcpyAsync(keep_hat, work_hat, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  }
  if (int rc = tpme_green_multiply(plan->dtype, work_hat, plan->batch, plan->nx, plan->ny, plan->nz, green, dc_out, stream)) return rc;
  return tpme_irfft3(plan, work_hat, mesh_out, stream);
}
