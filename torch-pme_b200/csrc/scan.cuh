// Single-pass exclusive prefix sum of int32 counters (chained scan with decoupled look-back).
//
// Used for the bin tables of the tile sort (tiles.cu) and of the neighbor list (neighbors.cu) and for
// the per-atom pair counts of the latter.  One CTA per tile of 4096 counters: a tile takes its number
// from a ticket counter (so every lower-numbered tile is already running: the look-back below cannot
// wait for a CTA that has not been scheduled), publishes its aggregate, adds up the published
// aggregates / inclusive prefixes of the tiles before it (one warp, 32 tiles per look) and publishes
// its own inclusive prefix.  Flag and value travel in ONE 64-bit word, so no fence is needed between
// them.  One launch, n reads + n writes; the single-CTA scan this replaces took 9 us for 64 k and
// 25 us for 256 k counters.
//
// `state` is caller-provided scratch of scan_state_words(n) 64-bit words, zeroed by the launcher.
#pragma once
#include "common.cuh"

namespace tpme {

constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

inline int64_t scan_tiles(int64_t n) { return (n + kScanTile - 1) / kScanTile; }
inline int64_t scan_state_words(int64_t n) { return scan_tiles(n) + 1; }   // ticket counter + one status word per tile

constexpr unsigned long long kScanAggregate = 1ull << 62, kScanPrefix = 2ull << 62, kScanValueMask = (1ull << 62) - 1;

template <typename OutT>
__global__ void __launch_bounds__(kScanThreads)
exclusive_scan_kernel(const int* __restrict__ in, OutT* __restrict__ out, int64_t n,
                      unsigned long long* __restrict__ state) {
  __shared__ int s_tile;
  __shared__ long long s_warp[kScanThreads / 32];
  __shared__ long long s_prefix;
  pdl_trigger();
  pdl_wait();                                  // the counters come from the kernel before
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(&state[0], 1ull);
  __syncthreads();
  const int tile = s_tile;
  volatile unsigned long long* status = state + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t base = (int64_t)tile * kScanTile + (int64_t)threadIdx.x * kScanItems;

  int v[kScanItems];
  long long sum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = base + k < n ? in[base + k] : 0;
    sum += v[k];
  }
  long long incl = sum;                       // inclusive scan of the thread sums inside the warp
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const long long y = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += y;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    long long w = lane < kScanThreads / 32 ? s_warp[lane] : 0;
    long long wi = w;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, wi, off);
      if (lane >= off) wi += y;
    }
    if (lane < kScanThreads / 32) s_warp[lane] = wi - w;   // exclusive offsets of the warps
    const long long aggregate = __shfl_sync(0xffffffffu, wi, 31);
    long long exclusive = 0;
    if (tile > 0) {
      if (lane == 0) status[tile] = kScanAggregate | (unsigned long long)aggregate;
      int look = tile - 1;
      while (true) {
        const int t = look - lane;
        unsigned long long word = kScanPrefix;               // "tile -1": inclusive prefix 0
        if (t >= 0) {
          do { word = status[t]; } while ((word >> 62) == 0);
        }
        const long long value = (long long)(word & kScanValueMask);
        const unsigned has_prefix = __ballot_sync(0xffffffffu, (word >> 62) == 2);
        long long part = value;
        if (has_prefix) part = lane <= __ffs(has_prefix) - 1 ? value : 0;   // nearest prefix and what lies between
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
        exclusive += part;
        if (has_prefix) break;
        look -= 32;
      }
    }
    if (lane == 0) {
      status[tile] = kScanPrefix | (unsigned long long)(exclusive + aggregate);
      s_prefix = exclusive;
      if ((int64_t)(tile + 1) * kScanTile >= n) out[n] = (OutT)(exclusive + aggregate);   // last tile: the total
    }
  }
  __syncthreads();
  long long run = s_prefix + s_warp[warp] + (incl - sum);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < n) out[base + k] = (OutT)run;
    run += v[k];
  }
}

// out[0 .. n] = exclusive prefix sums of in[0 .. n-1] (out[n] = total)
// `state_is_zero`: the caller has zeroed the scratch already (e.g. together with the counters, one memset)
template <typename OutT>
inline cudaError_t launch_exclusive_scan(const int* in, OutT* out, int64_t n, void* state, cudaStream_t s,
                                         bool state_is_zero = false, bool overlap = false) {
  if (n <= 0) return cudaMemsetAsync(out, 0, sizeof(OutT), s);
  if (!state_is_zero) {
    cudaError_t err = cudaMemsetAsync(state, 0, sizeof(unsigned long long) * (size_t)scan_state_words(n), s);
    if (err != cudaSuccess) return err;
  }
  return launch_pdl(exclusive_scan_kernel<OutT>, dim3((unsigned)scan_tiles(n)), dim3(kScanThreads), 0, s, overlap, in, out, n,
                    (unsigned long long*)state);
}

}  // namespace tpme
