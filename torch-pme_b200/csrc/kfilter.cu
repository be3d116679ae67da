// Reciprocal-space stage: real 3-D FFT (cuFFT), Green's / influence function evaluated on
// the fly from the reciprocal cell (no k-vector tensors), inverse FFT.
//
// Replaces KSpaceFilter / P3MKSpaceFilter (src/torchpme/lib/kspace_filter.py:97-197,293-361),
// generate_kvectors_for_mesh (lib/kvectors.py:24-102) and Potential.lr_from_k_sq
// (potentials/coulomb.py:122-142, potentials/inversepowerlaw.py:108-141, lib/math.py:16-104).
#include <cufft.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "green.cuh"
#include "../../include/torchpme_b200.h"

namespace tpme {

// hand-written FFT . G . iFFT for power-of-two meshes (fft3d.cuh, instantiated in fft_*.cu)
int filter_pow2_f32(const void*, void*, void*, int, int, int, int, const GreenDev<float>&, void*, cudaStream_t);
int filter_pow2_f64(const void*, void*, void*, int, int, int, int, const GreenDev<double>&, void*, cudaStream_t);
int filter_pow2_f32d(const void*, void*, void*, int, int, int, int, const GreenDev<double>&, void*, cudaStream_t);
static bool fft_supported_dim(int n) { return n >= 8 && n <= 512 && (n & (n - 1)) == 0; }

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
    const S gv = (S)green_value<T, S, true>(g, ix, iy, iz, nx, ny, nz, i);
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
    out[i] = (S)green_value<T, S, true>(g, (int)(t / ny), (int)(t % ny), iz, nx, ny, nz, i);
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

static int green_grid(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

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
  bool own_fft = false;  // power-of-two mesh: hand-written fused FFT . G . iFFT (fft3d.cuh)
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
  const char* force = getenv("TPME_FFT");  // "cufft" forces the library path (A/B testing)
  p->own_fft = fft_supported_dim(nx) && fft_supported_dim(ny) && fft_supported_dim(nz) &&
               !(force != nullptr && strcmp(force, "cufft") == 0);
  *plan = p;
  return 0;
}

// number of our kernels one tpme_kfilter_apply launches for this plan: 0 = cuFFT path (+ the
// multiply kernel), 3 = fused (y,z)-plane passes + x pass, 5 = separate z / y / x passes
extern "C" int tpme_fft_plan_uses_own_fft(tpme_fft_plan plan) {
  if (plan == nullptr || !plan->own_fft) return 0;
  auto listed = [](int n) { return n == 16 || n == 32 || n == 64 || n == 128; };
  const int ny = plan->ny, nz = plan->nz;
  bool fused = listed(ny) && listed(nz) && (ny == nz || ny == 2 * nz || nz == 2 * ny);
  // 256 x 256 fp32 planes: the two-CTA cluster kernels (half a plane per CTA, exchange through DSMEM)
  static const bool cluster_off = [] { const char* e = getenv("TPME_FFT_CLUSTER"); return !(e && e[0] == '1'); }();
  if (plan->dtype == 0 && ny == 256 && nz == 256 && !cluster_off) fused = true;
  return fused ? 3 : 5;
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
  if (plan->own_fft && keep_hat == nullptr) {
    if (int rc = check_green(green)) return rc;
    TPME_REQUIRE(!is_extended_green(green),
                 "spline kernels and P3M modes 1-3 reach the fused FFT as a table: tpme_green_table, then kind 0");
    cudaStream_t s = (cudaStream_t)stream;
    if (plan->dtype == 1)
      return filter_pow2_f64(mesh_in, mesh_out, work_hat, plan->batch, plan->nx, plan->ny, plan->nz,
                             make_green<double>(green), dc_out, s);
    if (needs_double_math(green))
      return filter_pow2_f32d(mesh_in, mesh_out, work_hat, plan->batch, plan->nx, plan->ny, plan->nz,
                              make_green<double>(green), dc_out, s);
    return filter_pow2_f32(mesh_in, mesh_out, work_hat, plan->batch, plan->nx, plan->ny, plan->nz,
                           make_green<float>(green), dc_out, s);
  }
  if (int rc = tpme_rfft3(plan, mesh_in, work_hat, stream)) return rc;
  if (keep_hat != nullptr) {
    const size_t bytes = (size_t)(plan->dtype ? 16 : 8) * plan->batch * plan->nx * plan->ny * (plan->nz / 2 + 1);
    TPME_CUDA_OK(cudaMemcpyAsync(keep_hat, work_hat, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  }
  if (int rc = tpme_green_multiply(plan->dtype, work_hat, plan->batch, plan->nx, plan->ny, plan->nz, green, dc_out, stream)) return rc;
  return tpme_irfft3(plan, work_hat, mesh_out, stream);
}
