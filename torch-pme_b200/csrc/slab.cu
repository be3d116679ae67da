// Slab-decomposed reciprocal-space stage (one x slab of the real mesh per GPU):
//
//   (y,z) passes on the local x planes -> exchange (x slabs -> y slabs) -> x pass . G . inverse x
//   pass on the local y rows -> exchange back -> inverse (y,z) passes
//
// (the single-GPU reference ops are lib/kspace_filter.py:169-187).
// The exchange is a strided block copy whose destinations are either blocks of a local send
// buffer (then NCCL all-to-all moves them) or the receive buffers of the peer GPUs mapped into
// this process (NVLink peer stores: pack + transfer + unpack in one kernel), followed by a
// device-side barrier over flags in peer memory.
#include <cstring>

#include "common.cuh"
#include "green.cuh"
#include "../../include/torchpme_b200.h"

namespace tpme {

int slab_yz_f32(bool, void*, void*, int, int, int, cudaStream_t, const RemoteStore*);
int slab_yz_f64(bool, void*, void*, int, int, int, cudaStream_t, const RemoteStore*);
int slab_x_f32(void*, int, int, int, int, int, int, const GreenDev<float>&, void*, cudaStream_t, const RemoteStore*);
int slab_x_f64(void*, int, int, int, int, int, int, const GreenDev<double>&, void*, cudaStream_t, const RemoteStore*);
int slab_x_f32d(void*, int, int, int, int, int, int, const GreenDev<double>&, void*, cudaStream_t, const RemoteStore*);
static_assert(kMaxRanks == TPME_MAX_RANKS, "rank limit of the kernels and of the ABI differ");

static int ilog2_exact(int n) {
  int s = 0;
  while ((1 << s) < n) ++s;
  return (1 << s) == n ? s : -1;
}

static bool pow2_dim(int n) { return n >= 8 && n <= 512 && (n & (n - 1)) == 0; }

// ---------------------------------------------------------------------------------------
// exchange copy: for every (c, p, a) copy `run` contiguous 16-byte (or 8-byte) words
//   dst[p][c * dst_c + a * dst_a + i] = src[c * src_c + p * src_p + a * src_a + i]
// One CTA row per (c, p, a) chunk, grid.y splits long runs.
// ---------------------------------------------------------------------------------------
struct PeerPointers { void* p[TPME_MAX_RANKS]; };

template <typename W>
__global__ void __launch_bounds__(256)
exchange_copy_kernel(const W* __restrict__ src, PeerPointers dst, int n_p, int n_a, int64_t run,
                     int64_t src_c, int64_t src_p, int64_t src_a, int64_t dst_c, int64_t dst_a) {
  const int chunk = blockIdx.x;
  const int a = chunk % n_a;
  const int p = (chunk / n_a) % n_p;
  const int c = chunk / (n_a * n_p);
  const W* s = src + c * src_c + p * src_p + a * src_a;
  W* d = reinterpret_cast<W*>(dst.p[p]) + c * dst_c + a * dst_a;
  // four independent 16-byte loads in flight per thread before the (possibly remote) stores
  const int64_t stride = (int64_t)gridDim.y * blockDim.x;
  int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < run; i += 4 * stride) {
    const W v0 = s[i], v1 = s[i + stride], v2 = s[i + 2 * stride], v3 = s[i + 3 * stride];
    d[i] = v0; d[i + stride] = v1; d[i + 2 * stride] = v2; d[i + 3 * stride] = v3;
  }
  for (; i < run; i += stride) d[i] = s[i];
}

// ---------------------------------------------------------------------------------------
// device-side barrier over flags in peer memory.  Every rank owns `flags[n_ranks]` (uint32,
// zero-initialised) plus a local epoch counter; arrival adds 1 to flags[rank] of every peer, the
// wait spins until all of its own flags reached the epoch.  No kernel argument changes between
// calls, so the barrier can be replayed from a CUDA graph.  A rank that waits longer than
// `timeout_ns` raises `*error` instead of hanging the GPU.
// ---------------------------------------------------------------------------------------
__global__ void peer_barrier_kernel(PeerPointers flags, int n_ranks, int rank, unsigned* epoch,
                                    unsigned long long timeout_ns, int* error) {
  __shared__ unsigned target;
  if (threadIdx.x == 0) {
    target = *epoch + 1;
    *epoch = target;
  }
  __syncthreads();
  const int peer = threadIdx.x;
  if (peer < n_ranks) {
    __threadfence_system();   // order this GPU's earlier peer stores before the arrival
    unsigned* remote = reinterpret_cast<unsigned*>(flags.p[peer]) + rank;
    atomicAdd_system(remote, 1u);
    volatile unsigned* mine = reinterpret_cast<unsigned*>(flags.p[rank]) + peer;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int)(*mine - target) < 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t - t0 > timeout_ns) {
        atomicExch(error, 1 + peer);
        break;
      }
      __nanosleep(64);
    }
    __threadfence_system();
  }
}

// ---------------------------------------------------------------------------------------
// all-reduce (sum) over peer memory in one kernel: rank r pulls slice r of every rank's input
// region (W loads in flight per thread), sums them in rank order and pushes the result into slice
// r of every rank's output region.  Every element is reduced by exactly one rank, so all ranks
// end up with bitwise identical sums.  The caller brackets the kernel with peer barriers.
// ---------------------------------------------------------------------------------------
template <typename V, typename T, int VEC>
__global__ void __launch_bounds__(256)
peer_allreduce_kernel(PeerPointers in, PeerPointers out, int n_ranks, int64_t begin, int64_t end) {
  // [begin, end) in units of V (16 bytes)
  for (int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end;
       i += (int64_t)gridDim.x * blockDim.x) {
    V acc = reinterpret_cast<const V*>(in.p[0])[i];
    T* a = reinterpret_cast<T*>(&acc);
#pragma unroll 4
    for (int p = 1; p < n_ranks; ++p) {
      const V v = reinterpret_cast<const V*>(in.p[p])[i];
      const T* b = reinterpret_cast<const T*>(&v);
#pragma unroll
      for (int e = 0; e < VEC; ++e) a[e] += b[e];
    }
    for (int p = 0; p < n_ranks; ++p) reinterpret_cast<V*>(out.p[p])[i] = acc;
  }
}

// ---------------------------------------------------------------------------------------
// all-reduce (sum) through the NVSwitch (NVLS): `mc` is the MULTICAST address of a buffer that every
// rank of the node has bound (one multicast object over W symmetric allocations).  Rank r reduces
// slice r with multimem.ld_reduce -- one load that the switch answers with the sum over all W
// copies -- and broadcasts the result with multimem.st, which the switch stores into all W copies,
// in place.  Per element 1 load + 1 store leave the GPU instead of W loads + W stores of the
// peer-pointer version above, and every rank ends up with the bitwise identical in-switch sum.
// The caller brackets the kernel with peer barriers.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float4 multimem_ld_reduce(const float4* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float4* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ double multimem_ld_reduce(const double* mc) {
  double v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f64 %0, [%1];" : "=d"(v) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(double* mc, double v) {
  asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" :: "l"(mc), "d"(v) : "memory");
}

template <typename V>
__global__ void __launch_bounds__(256)
multimem_allreduce_kernel(V* mc, int64_t begin, int64_t end) {
  // [begin, end) in units of V; two independent reductions in flight per thread
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + stride < end; i += 2 * stride) {
    const V a = multimem_ld_reduce(mc + i);
    const V b = multimem_ld_reduce(mc + i + stride);
    multimem_st(mc + i, a);
    multimem_st(mc + i + stride, b);
  }
  if (i < end) multimem_st(mc + i, multimem_ld_reduce(mc + i));
}

}  // namespace tpme

using namespace tpme;

extern "C" int tpme_multimem_allreduce(int dtype, void* multicast_ptr, int n_ranks, int rank, int64_t n,
                                       void* stream) {
  TPME_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 or 1");
  TPME_REQUIRE(n_ranks > 0 && n_ranks <= TPME_MAX_RANKS && rank >= 0 && rank < n_ranks, "bad rank layout");
  TPME_REQUIRE(multicast_ptr != nullptr && ((uintptr_t)multicast_ptr % 16) == 0, "multicast address missing / unaligned");
  TPME_REQUIRE(n >= 0, "negative size");
  if (n == 0) return 0;
  const int vec = dtype == 0 ? 4 : 1;                     // float4 / double per multimem operation
  const int64_t words = (n + vec - 1) / vec;               // the buffer is padded to whole 16-byte words
  const int64_t per = (words + n_ranks - 1) / n_ranks;
  const int64_t begin = per * rank < words ? per * rank : words;
  const int64_t end = begin + per < words ? begin + per : words;
  if (end <= begin) return 0;
  int64_t grid = (end - begin + 511) / 512;
  const int64_t cap = 2ll * num_sms();
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == 0)
    multimem_allreduce_kernel<float4><<<(unsigned)grid, 256, 0, s>>>((float4*)multicast_ptr, begin, end);
  else
    multimem_allreduce_kernel<double><<<(unsigned)grid, 256, 0, s>>>((double*)multicast_ptr, begin, end);
  TPME_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int tpme_slab_fft_yz(int dtype, int forward, void* real_mesh, void* mesh_hat,
                                int n_planes, int ny, int nz, void* stream) {
  TPME_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 or 1");
  TPME_REQUIRE(pow2_dim(ny) && pow2_dim(nz),
               "the slab-decomposed FFT needs power-of-two mesh dimensions in 8..512");
  if (n_planes <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  return dtype == 0 ? slab_yz_f32(forward != 0, real_mesh, mesh_hat, n_planes, ny, nz, s, nullptr)
                    : slab_yz_f64(forward != 0, real_mesh, mesh_hat, n_planes, ny, nz, s, nullptr);
}

static int check_peers(const tpme_slab_peers* peers, int nx, int ny) {
  TPME_REQUIRE(peers != nullptr, "peer table missing");
  TPME_REQUIRE(peers->n_ranks > 0 && peers->n_ranks <= TPME_MAX_RANKS && peers->rank >= 0 &&
               peers->rank < peers->n_ranks, "bad rank layout");
  TPME_REQUIRE(nx % peers->n_ranks == 0 && ny % peers->n_ranks == 0, "the world size has to divide nx and ny");
  for (int p = 0; p < peers->n_ranks; ++p)
    TPME_REQUIRE(peers->hat[p] != nullptr && peers->hat_t[p] != nullptr, "null peer buffer");
  return 0;
}

// forward (y,z) passes of the local x planes; the y pass stores its output straight into the
// y-slab arrays of all ranks (rank p receives the rows [p ny/W, (p+1) ny/W))
extern "C" int tpme_slab_fft_yz_push(int dtype, void* real_mesh, int n_channels, int nx, int ny,
                                     int nz, const tpme_slab_peers* peers, void* stream) {
  TPME_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 or 1");
  TPME_REQUIRE(pow2_dim(nx) && pow2_dim(ny) && pow2_dim(nz),
               "the slab-decomposed FFT needs power-of-two mesh dimensions in 8..512");
  if (int rc = check_peers(peers, nx, ny)) return rc;
  if (n_channels <= 0) return 0;
  const int w = peers->n_ranks, rank = peers->rank;
  const int nxl = nx / w, nyl = ny / w, nzh = nz / 2 + 1;
  RemoteStore rs;
  memset(&rs, 0, sizeof(rs));
  for (int p = 0; p < w; ++p) rs.p[p] = peers->hat_t[p];
  rs.enabled = 1;
  rs.shift = ilog2_exact(nyl);
  TPME_REQUIRE(rs.shift >= 0, "rows per rank must be a power of two");
  // work item o = c * nxl + xl  ->  (c * nx + rank * nxl + xl) * nyl * nzh + yl * nzh + z
  rs.dA = nxl;
  rs.sA = (int64_t)nx * nyl * nzh;
  rs.sB = (int64_t)nyl * nzh;
  rs.sL = nzh;
  rs.off = (int64_t)rank * nxl * nyl * nzh;
  cudaStream_t s = (cudaStream_t)stream;
  void* scratch = peers->hat[rank];
  return dtype == 0 ? slab_yz_f32(true, real_mesh, scratch, n_channels * nxl, ny, nz, s, &rs)
                    : slab_yz_f64(true, real_mesh, scratch, n_channels * nxl, ny, nz, s, &rs);
}

// x pass . G . inverse x pass on the local y rows; the results are stored straight into the
// x-slab arrays of all ranks (rank p receives the planes [p nx/W, (p+1) nx/W))
extern "C" int tpme_slab_fft_x_green_push(int dtype, int n_channels, int nx, int ny, int nz,
                                          const tpme_green* green, const tpme_slab_peers* peers,
                                          void* stream) {
  TPME_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 or 1");
  TPME_REQUIRE(pow2_dim(nx), "the slab-decomposed FFT needs power-of-two mesh dimensions in 8..512");
  if (int rc = check_peers(peers, nx, ny)) return rc;
  if (int rc = check_green(green)) return rc;
  TPME_REQUIRE(!is_extended_green(green), "the slab x pass evaluates closed-form kernels and tables only");
  if (n_channels <= 0) return 0;
  const int w = peers->n_ranks, rank = peers->rank;
  const int nxl = nx / w, nyl = ny / w, nzh = nz / 2 + 1;
  RemoteStore rs;
  memset(&rs, 0, sizeof(rs));
  for (int p = 0; p < w; ++p) rs.p[p] = peers->hat[p];
  rs.enabled = 1;
  rs.shift = ilog2_exact(nxl);
  TPME_REQUIRE(rs.shift >= 0, "planes per rank must be a power of two");
  // work item o = c * nyl + yl  ->  (c * nxl + xl) * ny * nzh + (rank * nyl + yl) * nzh + z
  rs.dA = nyl;
  rs.sA = (int64_t)nxl * ny * nzh;
  rs.sB = nzh;
  rs.sL = (int64_t)ny * nzh;
  rs.off = (int64_t)rank * nyl * nzh;
  cudaStream_t s = (cudaStream_t)stream;
  void* hat_t = peers->hat_t[rank];
  const int y0 = rank * nyl;
  if (dtype == 1)
    return slab_x_f64(hat_t, n_channels, nx, ny, nz, y0, nyl, make_green<double>(green), nullptr, s, &rs);
  if (needs_double_math(green))
    return slab_x_f32d(hat_t, n_channels, nx, ny, nz, y0, nyl, make_green<double>(green), nullptr, s, &rs);
  return slab_x_f32(hat_t, n_channels, nx, ny, nz, y0, nyl, make_green<float>(green), nullptr, s, &rs);
}

extern "C" int tpme_slab_fft_x_green(int dtype, void* mesh_hat_t, int n_channels, int nx, int ny,
                                     int nz, int y0, int ny_local, const tpme_green* green,
                                     void* stream) {
  TPME_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 or 1");
  TPME_REQUIRE(pow2_dim(nx), "the slab-decomposed FFT needs power-of-two mesh dimensions in 8..512");
  TPME_REQUIRE(y0 >= 0 && ny_local > 0 && y0 + ny_local <= ny, "y slab must lie inside [0, ny)");
  if (int rc = check_green(green)) return rc;
  TPME_REQUIRE(!is_extended_green(green), "the slab x pass evaluates closed-form kernels and tables only");
  if (n_channels <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == 1)
    return slab_x_f64(mesh_hat_t, n_channels, nx, ny, nz, y0, ny_local, make_green<double>(green), nullptr, s, nullptr);
  if (needs_double_math(green))
    return slab_x_f32d(mesh_hat_t, n_channels, nx, ny, nz, y0, ny_local, make_green<double>(green), nullptr, s, nullptr);
  return slab_x_f32(mesh_hat_t, n_channels, nx, ny, nz, y0, ny_local, make_green<float>(green), nullptr, s, nullptr);
}

extern "C" int tpme_slab_exchange_copy(int elem_bytes, const void* src, void* const* dst_host,
                                       int n_c, int n_p, int n_a, int64_t run, int64_t src_c,
                                       int64_t src_p, int64_t src_a, int64_t dst_c, int64_t dst_a,
                                       void* stream) {
  TPME_REQUIRE(elem_bytes == 8 || elem_bytes == 16, "elements are complex float (8) or complex double (16)");
  TPME_REQUIRE(n_p > 0 && n_p <= TPME_MAX_RANKS, "number of destinations out of range");
  TPME_REQUIRE(n_c >= 0 && n_a >= 0 && run >= 0, "negative sizes");
  if (n_c == 0 || n_a == 0 || run == 0) return 0;
  PeerPointers dst;
  memset(&dst, 0, sizeof(dst));
  for (int p = 0; p < n_p; ++p) dst.p[p] = dst_host[p];
  const int64_t chunks = (int64_t)n_c * n_p * n_a;
  TPME_REQUIRE(chunks < (1ll << 31), "too many exchange chunks");
  cudaStream_t s = (cudaStream_t)stream;
  // 16-byte words when everything is 16-byte aligned, else 8-byte words
  bool wide = elem_bytes == 16;
  if (!wide) {
    wide = (run % 2 == 0) && (src_c % 2 == 0) && (src_p % 2 == 0) && (src_a % 2 == 0) &&
           (dst_c % 2 == 0) && (dst_a % 2 == 0) && ((uintptr_t)src % 16 == 0);
    for (int p = 0; p < n_p && wide; ++p) wide = ((uintptr_t)dst.p[p] % 16 == 0);
  }
  const int64_t words = (wide && elem_bytes == 8) ? run / 2 : run;
  const int div = (wide && elem_bytes == 8) ? 2 : 1;
  // enough CTAs to fill the machine (~8 per SM), at most one CTA per 256 words of a chunk
  int64_t split = (4ll * num_sms() + chunks - 1) / chunks;
  const int64_t max_split = (words + 1023) / 1024;
  if (split > max_split) split = max_split;
  if (split < 1) split = 1;
  if (split > 65535) split = 65535;
  dim3 grid((unsigned)chunks, (unsigned)split);
  if (wide)
    exchange_copy_kernel<uint4><<<grid, 256, 0, s>>>((const uint4*)src, dst, n_p, n_a, words, src_c / div,
                                                     src_p / div, src_a / div, dst_c / div, dst_a / div);
  else
    exchange_copy_kernel<uint2><<<grid, 256, 0, s>>>((const uint2*)src, dst, n_p, n_a, words, src_c, src_p,
                                                     src_a, dst_c, dst_a);
  TPME_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- peer memory ------------------------------------------------------------------------
extern "C" int tpme_peer_buffer_create(int64_t bytes, void** dev_ptr, unsigned char* handle_out) {
  TPME_REQUIRE(bytes > 0 && dev_ptr != nullptr && handle_out != nullptr, "bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == TPME_IPC_HANDLE_BYTES, "IPC handle size");
  void* p = nullptr;
  TPME_CUDA_OK(cudaMalloc(&p, (size_t)bytes));
  TPME_CUDA_OK(cudaMemset(p, 0, (size_t)bytes));
  cudaIpcMemHandle_t h;
  cudaError_t err = cudaIpcGetMemHandle(&h, p);
  if (err != cudaSuccess) {
    cudaFree(p);
    set_last_error("cudaIpcGetMemHandle", cudaGetErrorString(err));
    return 100 + (int)err;
  }
  memcpy(handle_out, &h, sizeof(h));
  *dev_ptr = p;
  return 0;
}

extern "C" int tpme_peer_buffer_open(const unsigned char* handle, void** dev_ptr) {
  TPME_REQUIRE(handle != nullptr && dev_ptr != nullptr, "bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  TPME_CUDA_OK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

extern "C" int tpme_peer_buffer_close(void* dev_ptr) {
  if (dev_ptr) TPME_CUDA_OK(cudaIpcCloseMemHandle(dev_ptr));
  return 0;
}

extern "C" int tpme_peer_buffer_destroy(void* dev_ptr) {
  if (dev_ptr) TPME_CUDA_OK(cudaFree(dev_ptr));
  return 0;
}

extern "C" int tpme_peer_barrier(void* const* flags_host, int n_ranks, int rank, void* epoch,
                                 double timeout_seconds, void* error_flag, void* stream) {
  TPME_REQUIRE(n_ranks > 0 && n_ranks <= TPME_MAX_RANKS && rank >= 0 && rank < n_ranks, "bad rank layout");
  TPME_REQUIRE(flags_host != nullptr && epoch != nullptr && error_flag != nullptr, "null pointers");
  PeerPointers flags;
  memset(&flags, 0, sizeof(flags));
  for (int p = 0; p < n_ranks; ++p) flags.p[p] = flags_host[p];
  const unsigned long long ns = (unsigned long long)(timeout_seconds * 1e9);
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags, n_ranks, rank, (unsigned*)epoch, ns,
                                                          (int*)error_flag);
  TPME_CUDA_OK(cudaGetLastError());
  return 0;
}

// all-reduce (sum) of `n` reals held in peer-mapped regions: `in_host[p]` / `out_host[p]` are rank
// p's input / output regions as mapped in this process (16-byte aligned, padded to a multiple of
// 16 bytes * n_ranks).  Call between two tpme_peer_barrier calls.
extern "C" int tpme_peer_allreduce(int dtype, void* const* in_host, void* const* out_host, int n_ranks,
                                   int rank, int64_t n, void* stream) {
  TPME_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 or 1");
  TPME_REQUIRE(n_ranks > 0 && n_ranks <= TPME_MAX_RANKS && rank >= 0 && rank < n_ranks, "bad rank layout");
  TPME_REQUIRE(n >= 0, "negative size");
  if (n == 0) return 0;
  PeerPointers in, out;
  memset(&in, 0, sizeof(in));
  memset(&out, 0, sizeof(out));
  for (int p = 0; p < n_ranks; ++p) {
    TPME_REQUIRE(in_host[p] != nullptr && out_host[p] != nullptr, "null peer region");
    TPME_REQUIRE(((uintptr_t)in_host[p] % 16) == 0 && ((uintptr_t)out_host[p] % 16) == 0, "regions must be 16-byte aligned");
    in.p[p] = in_host[p];
    out.p[p] = out_host[p];
  }
  const int vec = dtype == 0 ? 4 : 2;
  const int64_t words = (n + vec - 1) / vec;               // the regions are padded, see above
  const int64_t per = (words + n_ranks - 1) / n_ranks;
  const int64_t begin = per * rank < words ? per * rank : words;
  const int64_t end = begin + per < words ? begin + per : words;
  if (end <= begin) return 0;
  int64_t grid = (end - begin + 255) / 256;
  const int64_t cap = 4ll * num_sms();
  if (grid > cap) grid = cap;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == 0)
    peer_allreduce_kernel<float4, float, 4><<<(unsigned)grid, 256, 0, s>>>(in, out, n_ranks, begin, end);
  else
    peer_allreduce_kernel<double2, double, 2><<<(unsigned)grid, 256, 0, s>>>(in, out, n_ranks, begin, end);
  TPME_CUDA_OK(cudaGetLastError());
  return 0;
}
