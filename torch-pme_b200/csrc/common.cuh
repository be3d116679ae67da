// Shared device/host helpers for the sm_100a PME/P3M kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define TPME_P3M 0
#define TPME_LAGRANGE 1

namespace tpme {

// ---- error plumbing (C ABI returns int; the message is kept per thread) ---------------
void set_last_error(const char* where, const char* what);

#define TPME_CUDA_OK(expr)                                              \
  do {                                                                  \
    cudaError_t err__ = (expr);                                         \
    if (err__ != cudaSuccess) {                                         \
      ::tpme::set_last_error(#expr, cudaGetErrorString(err__));         \
      return 100 + (int)err__;                                          \
    }                                                                   \
  } while (0)

#define TPME_REQUIRE(cond, msg)                                         \
  do {                                                                  \
    if (!(cond)) {                                                      \
      ::tpme::set_last_error(#cond, msg);                               \
      return 1;                                                         \
    }                                                                   \
  } while (0)

// ---- small math helpers --------------------------------------------------------------
__device__ __forceinline__ float fma_t(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }

template <typename T> struct Mat3 { T m[9]; };  // row-major, passed by value to kernels

template <typename T>
__host__ inline Mat3<T> load_mat3(const double* host9) {
  Mat3<T> out;
  for (int i = 0; i < 9; ++i) out.m[i] = (T)host9[i];
  return out;
}

// non-negative modulo for possibly negative i (atoms may sit outside the cell;
// reference: ``(idx + i) % ns`` in lib/mesh_interpolator.py:352)
__device__ __forceinline__ int wrap_index(int i, int n) {
  int r = i % n;
  return r < 0 ? r + n : r;
}

__device__ __forceinline__ float floor_t(float x) { return floorf(x); }
__device__ __forceinline__ double floor_t(double x) { return floor(x); }
__device__ __forceinline__ float rint_t(float x) { return rintf(x); }   // ties-to-even == torch.round
__device__ __forceinline__ double rint_t(double x) { return rint(x); }

// native reduction atomics (no return value -> RED.E.ADD in SASS)
__device__ __forceinline__ void red_add(float* p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void red_add(double* p, double v) { atomicAdd(p, v); }

// Optional redirection of the final stores of a strided FFT pass into the buffers of other GPUs
// (slab decomposition: the transposing all-to-all is fused into the FFT pass, the results travel
// as NVLink peer stores while other CTAs still compute).  Line `l` of work item `o` goes to
//   p[l >> shift] + (o / dA) * sA + (o % dA) * sB + (l & ((1 << shift) - 1)) * sL + off + z
constexpr int kMaxRanks = 16;
struct RemoteStore {
  void* p[kMaxRanks];
  int enabled, shift, dA;
  int64_t sA, sB, sL, off;
};

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// The mesh pipeline is a chain of 6-8 short dependent kernels per pass.  Launched with
// cudaLaunchAttributeProgrammaticStreamSerialization a kernel may become resident while its predecessor is
// still running: everything it does before pdl_wait() (shared-memory zeroing, twiddle / Green axis tables,
// mbarrier set-up -- nothing that reads memory written by a kernel of the chain) overlaps the predecessor's
// tail; pdl_wait() returns once the predecessor has completed and its writes are visible.  Every CTA of every
// kernel of the chain executes pdl_wait() before it exits, so "predecessor complete" implies "everything
// upstream complete".  pdl_trigger() at the top of a kernel lets its successor be scheduled as soon as all
// CTAs of this kernel have started.  Without the launch attribute both are no-ops.  TPME_PDL=0 switches
// the attribute off, TPME_PDL=all asks for it in fp32 too.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();
bool pdl_everywhere();
template <typename T> inline bool pdl_for() { return sizeof(T) == 8 || pdl_everywhere(); }

// `overlap`: ask for the programmatic dependency.  Measured on B200 (profiles/r02_summary.md section 10): it pays
// for the fp64 kernels, whose prologues (double-precision twiddles, Green axis tables) are expensive -- c3
// 0.308 -> 0.297 ms -- and costs a little for the fp32 ones (c2 +10 %, c4 +2 %), so the callers pass
// sizeof(T) == 8.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                              bool overlap, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (overlap && pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int num_sms() {
  static int cached = 0;
  if (!cached) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    if (cached <= 0) cached = 148;
  }
  return cached;
}

}  // namespace tpme
