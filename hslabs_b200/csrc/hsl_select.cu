// Selection over the (all-gathered) per-candidate costs: argmin and the k cheapest candidates.
//
// Order everywhere: ascending (cost, index) -- what a stable sort of the costs gives; NaN costs (failed candidates) are
// never selected.  The costs of a search iteration are small (8 B per candidate: 512 KiB for BASELINE config 3), so the
// kernels are latency bound, not bandwidth bound:
//
//   hsl_argmin_kernel   one block, eight independent loads per thread and iteration (the round-1 kernel issued one
//                       dependent load per iteration and grew with the number of gathered shards)
//   top-k               a bitonic sort of (sortable 64-bit image of the cost, index) pairs over all n costs in a
//                       stream-ordered workspace, then the first k are emitted.  Tiles of 4096 pairs are sorted / merged
//                       in shared memory; only the compare-exchange steps whose partner lies in another tile are separate
//                       launches (none for n <= 4096, 10 + 5 for n = 65536).  The round-1 kernel made k passes over n in
//                       ONE block (2.7e8 loads for the elite 4096 of 65536).
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "hsl_gather_dev.cuh"
#include "hsl_internal.h"

namespace {

__device__ __forceinline__ bool hsl_better(double v, long long i, double bv, long long bidx) {
  return i >= 0 && (bidx < 0 || v < bv || (v == bv && i < bidx));
}

// GATHERED: the costs are a peer-memory gather buffer (hsl_gather.cu): wait for every rank's flag first (acquire), and read
// the costs past L1 (they were written by other GPUs).
template <bool GATHERED>
__global__ void hsl_argmin_kernel(const double* __restrict__ cost, int64_t n, int64_t* __restrict__ out_index, double* __restrict__ out_value,
                                  const unsigned long long* flags, int nranks, unsigned long long epoch) {
  __shared__ double sv[32];
  __shared__ long long si[32];
  HSL_GRID_DEP_WAIT();   // launched with programmatic dependent launch behind the kernel that wrote the costs
  if (GATHERED) {
    if (threadIdx.x < nranks) hsl_wait_flag(flags + threadIdx.x, epoch, (unsigned long long*)(flags + HSL_MAX_PEERS + 1));
    __syncthreads();
  }
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  double best = inf;
  long long bi = -1;
  constexpr int U = 8;
  const int64_t stride = (int64_t)blockDim.x * U;
  for (int64_t base = threadIdx.x; base < n; base += stride) {
    double c[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int64_t i = base + (int64_t)u * blockDim.x;
      c[u] = (i < n) ? (GATHERED ? __ldcg(cost + i) : cost[i]) : inf;   // independent loads: one round trip per U candidates
    }
#pragma unroll
    for (int u = 0; u < U; u++)
      if (c[u] < best) { best = c[u]; bi = base + (int64_t)u * blockDim.x; }  // NaN compares false; ascending indices per thread
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
    const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (hsl_better(ov, oi, best, bi)) { best = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = (blockDim.x + 31) / 32;
    best = (threadIdx.x < nw) ? sv[threadIdx.x] : inf;
    bi = (threadIdx.x < nw) ? si[threadIdx.x] : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (hsl_better(ov, oi, best, bi)) { best = ov; bi = oi; }
    }
    if (threadIdx.x == 0) {
      if (out_index) *out_index = bi;
      if (out_value) *out_value = (bi >= 0) ? best : __longlong_as_double(0x7ff8000000000000LL);
    }
  }
}

// ---------------------------------------------------------------- bitonic sort of (key, index) pairs
constexpr int TILE = 4096, SORT_THREADS = 1024;
constexpr uint64_t KEY_INVALID = ~0ull;

// order-preserving image of a double in an unsigned 64-bit integer; NaN -> KEY_INVALID (sorts last, never emitted)
__device__ __forceinline__ uint64_t sort_key(double c) {
  if (c != c) return KEY_INVALID;
  const uint64_t b = (uint64_t)__double_as_longlong(c);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_value(uint64_t k) {
  const uint64_t b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}
__device__ __forceinline__ bool pair_less(uint64_t ka, uint32_t ia, uint64_t kb, uint32_t ib) { return ka < kb || (ka == kb && ia < ib); }

// compare-exchange steps j = j0, j0/2, ..., 1 of stage s on one tile held in shared memory
__device__ __forceinline__ void tile_steps(uint64_t* sk, uint32_t* si, int64_t tile_base, int64_t s, int j0) {
  for (int j = j0; j > 0; j >>= 1) {
    for (int t = threadIdx.x; t < TILE / 2; t += SORT_THREADS) {
      const int lo = ((t / j) * 2 * j) + (t % j), hi = lo + j;
      const bool up = (((tile_base + lo) & s) == 0);
      const uint64_t ka = sk[lo], kb = sk[hi];
      const uint32_t ia = si[lo], ib = si[hi];
      if (pair_less(kb, ib, ka, ia) == up) { sk[lo] = kb; sk[hi] = ka; si[lo] = ib; si[hi] = ia; }
    }
    __syncthreads();
  }
}

// keys / indices of the padded array [P] from the costs, each tile sorted through stage s = TILE
__global__ void __launch_bounds__(SORT_THREADS) hsl_sort_tiles_kernel(const double* __restrict__ cost, int64_t n, uint64_t* __restrict__ keys,
                                                                      uint32_t* __restrict__ idxs) {
  __shared__ uint64_t sk[TILE];
  __shared__ uint32_t si[TILE];
  const int64_t base = (int64_t)blockIdx.x * TILE;
  for (int t = threadIdx.x; t < TILE; t += SORT_THREADS) {
    const int64_t i = base + t;
    sk[t] = (i < n) ? sort_key(cost[i]) : KEY_INVALID;
    si[t] = (i < n) ? (uint32_t)i : 0xffffffffu;
  }
  __syncthreads();
  for (int64_t s = 2; s <= TILE; s <<= 1) tile_steps(sk, si, base, s, (int)(s >> 1));
  for (int t = threadIdx.x; t < TILE; t += SORT_THREADS) { keys[base + t] = sk[t]; idxs[base + t] = si[t]; }
}
// one compare-exchange step (stage s, distance j >= TILE) across tiles
__global__ void hsl_sort_global_step_kernel(uint64_t* __restrict__ keys, uint32_t* __restrict__ idxs, int64_t P, int64_t s, int64_t j) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P / 2) return;
  const int64_t lo = ((t / j) * 2 * j) + (t % j), hi = lo + j;
  const bool up = ((lo & s) == 0);
  const uint64_t ka = keys[lo], kb = keys[hi];
  const uint32_t ia = idxs[lo], ib = idxs[hi];
  if (pair_less(kb, ib, ka, ia) == up) { keys[lo] = kb; keys[hi] = ka; idxs[lo] = ib; idxs[hi] = ia; }
}
// the remaining steps j = TILE/2 .. 1 of stage s, tile by tile in shared memory
__global__ void __launch_bounds__(SORT_THREADS) hsl_sort_merge_kernel(uint64_t* __restrict__ keys, uint32_t* __restrict__ idxs, int64_t s) {
  __shared__ uint64_t sk[TILE];
  __shared__ uint32_t si[TILE];
  const int64_t base = (int64_t)blockIdx.x * TILE;
  for (int t = threadIdx.x; t < TILE; t += SORT_THREADS) { sk[t] = keys[base + t]; si[t] = idxs[base + t]; }
  __syncthreads();
  tile_steps(sk, si, base, s, TILE / 2);
  for (int t = threadIdx.x; t < TILE; t += SORT_THREADS) { keys[base + t] = sk[t]; idxs[base + t] = si[t]; }
}
// The whole sort and the emission in ONE cooperative launch when every tile can be resident at once (P / TILE <= the
// number of co-resident blocks: up to 148 tiles = 606 208 costs on a B200): grid-wide barriers replace the ~16 dependent
// launches of the multi-kernel path, whose launch latencies dominated (168 us -> see profiles/ for top-4096 of 65536).
__global__ void __launch_bounds__(SORT_THREADS) hsl_topk_coop_kernel(const double* __restrict__ cost, int64_t n, uint64_t* __restrict__ keys,
                                                                     uint32_t* __restrict__ idxs, int64_t P, int k,
                                                                     int64_t* __restrict__ out_index, double* __restrict__ out_value) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ uint64_t sk[TILE];
  __shared__ uint32_t si[TILE];
  const int64_t base = (int64_t)blockIdx.x * TILE;
  for (int t = threadIdx.x; t < TILE; t += SORT_THREADS) {
    const int64_t i = base + t;
    sk[t] = (i < n) ? sort_key(cost[i]) : KEY_INVALID;
    si[t] = (i < n) ? (uint32_t)i : 0xffffffffu;
  }
  __syncthreads();
  for (int64_t s = 2; s <= TILE; s <<= 1) tile_steps(sk, si, base, s, (int)(s >> 1));
  for (int64_t s = 2 * (int64_t)TILE; s <= P; s <<= 1) {
    for (int t = threadIdx.x; t < TILE; t += SORT_THREADS) { keys[base + t] = sk[t]; idxs[base + t] = si[t]; }
    grid.sync();
    for (int64_t j = s >> 1; j >= TILE; j >>= 1) {
      // this block's share of the P/2 compare-exchange pairs of the step
      for (int64_t t = (int64_t)blockIdx.x * (TILE / 2) + threadIdx.x; t < (int64_t)(blockIdx.x + 1) * (TILE / 2); t += SORT_THREADS) {
        const int64_t lo = ((t / j) * 2 * j) + (t % j), hi = lo + j;
        const bool up = ((lo & s) == 0);
        const uint64_t ka = keys[lo], kb = keys[hi];
        const uint32_t ia = idxs[lo], ib = idxs[hi];
        if (pair_less(kb, ib, ka, ia) == up) { keys[lo] = kb; keys[hi] = ka; idxs[lo] = ib; idxs[hi] = ia; }
      }
      grid.sync();
    }
    for (int t = threadIdx.x; t < TILE; t += SORT_THREADS) { sk[t] = keys[base + t]; si[t] = idxs[base + t]; }
    __syncthreads();
    tile_steps(sk, si, base, s, TILE / 2);
  }
  // emit: the first k pairs of the sorted array live in the first tiles
  for (int t = threadIdx.x; t < TILE; t += SORT_THREADS) {
    const int64_t i = base + t;
    if (i >= k) continue;
    const bool ok = sk[t] != KEY_INVALID;
    if (out_index) out_index[i] = ok ? (int64_t)si[t] : -1;
    if (out_value) out_value[i] = ok ? key_value(sk[t]) : __longlong_as_double(0x7ff8000000000000LL);
  }
}

// ---------------------------------------------------------------- k <= TILE: sort the tiles, then prune pairwise
// For the k <= TILE cheapest of n costs the full sort is not needed: once every tile is sorted ascending, two tiles A, B
// contribute only the TILE smallest of their union, and min(A[i], B[TILE-1-i]) is exactly that set as a bitonic sequence --
// one merge (log2 TILE steps) sorts it again.  log2(tiles) rounds of that leave the answer in tile 0: 4 grid barriers for
// 65536 costs instead of the 14 of the full sort, and no global exchange steps.
// Inside a tile the steps with distance < 8 run in registers: a thread owns 8 consecutive elements, so the three last steps
// of every stage (and the whole stages 2, 4, 8) cost one block barrier instead of three (95 barriers for 65536 costs
// instead of 126).
constexpr int PRUNE_THREADS = TILE / 8;

__device__ __forceinline__ void cswap(uint64_t& ka, uint32_t& ia, uint64_t& kb, uint32_t& ib, bool up) {
  if (pair_less(kb, ib, ka, ia) == up) { const uint64_t tk = ka; ka = kb; kb = tk; const uint32_t ti = ia; ia = ib; ib = ti; }
}
// steps j = min(4, s/2) .. 1 of stages s_lo .. s_hi (s_hi <= 8, or s_lo == s_hi) on the thread's own 8 elements
__device__ __forceinline__ void reg_steps(uint64_t* sk, uint32_t* si, int s_lo, int s_hi) {
  const int e0 = threadIdx.x * 8;
  uint64_t k[8];
  uint32_t ix[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { k[i] = sk[e0 + i]; ix[i] = si[e0 + i]; }
  for (int s = s_lo; s <= s_hi; s <<= 1) {
#pragma unroll
    for (int j = 4; j > 0; j >>= 1) {
      if (j > (s >> 1)) continue;
#pragma unroll
      for (int t = 0; t < 4; t++) {
        const int lo = ((t / j) * 2 * j) + (t % j), hi = lo + j;
        cswap(k[lo], ix[lo], k[hi], ix[hi], ((e0 + lo) & s) == 0);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; i++) { sk[e0 + i] = k[i]; si[e0 + i] = ix[i]; }
  __syncthreads();
}
// steps j = j0 .. 1 of stage s on a tile sorted ascending as a whole (s == TILE) or by the bitonic pattern (s < TILE)
__device__ __forceinline__ void prune_tile_steps(uint64_t* sk, uint32_t* si, int s, int j0) {
  for (int j = j0; j >= 8; j >>= 1) {
    for (int t = threadIdx.x; t < TILE / 2; t += PRUNE_THREADS) {
      const int lo = ((t / j) * 2 * j) + (t % j), hi = lo + j;
      const bool up = ((lo & s) == 0);
      const uint64_t ka = sk[lo], kb = sk[hi];
      const uint32_t ia = si[lo], ib = si[hi];
      if (pair_less(kb, ib, ka, ia) == up) { sk[lo] = kb; sk[hi] = ka; si[lo] = ib; si[hi] = ia; }
    }
    __syncthreads();
  }
  reg_steps(sk, si, s, s);
}
__global__ void __launch_bounds__(PRUNE_THREADS) hsl_topk_prune_kernel(const double* __restrict__ cost, int64_t n, uint64_t* __restrict__ keys,
                                                                       uint32_t* __restrict__ idxs, int tiles, int k,
                                                                       int64_t* __restrict__ out_index, double* __restrict__ out_value) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ __align__(16) uint64_t sk[TILE];
  __shared__ __align__(16) uint32_t si[TILE];
  const int64_t base = (int64_t)blockIdx.x * TILE;
  for (int t = threadIdx.x; t < TILE; t += PRUNE_THREADS) {
    const int64_t i = base + t;
    sk[t] = (i < n) ? sort_key(cost[i]) : KEY_INVALID;
    si[t] = (i < n) ? (uint32_t)i : 0xffffffffu;
  }
  __syncthreads();
  reg_steps(sk, si, 2, 8);
  for (int s = 16; s <= TILE; s <<= 1) prune_tile_steps(sk, si, s, s >> 1);   // every tile ascending (lo & TILE == 0 for all lo)
  for (int stride = 1; stride < tiles; stride <<= 1) {
    const bool giver = (blockIdx.x & (2 * stride - 1)) == stride, taker = (blockIdx.x & (2 * stride - 1)) == 0;
    if (giver) for (int t = threadIdx.x; t < TILE; t += PRUNE_THREADS) { keys[base + t] = sk[t]; idxs[base + t] = si[t]; }
    grid.sync();
    if (taker && (int)blockIdx.x + stride < tiles) {
      const int64_t pb = base + (int64_t)stride * TILE;
      for (int t = threadIdx.x; t < TILE; t += PRUNE_THREADS) {
        const uint64_t kb = keys[pb + TILE - 1 - t];
        const uint32_t ib = idxs[pb + TILE - 1 - t];
        if (pair_less(kb, ib, sk[t], si[t])) { sk[t] = kb; si[t] = ib; }
      }
      __syncthreads();
      prune_tile_steps(sk, si, TILE, TILE / 2);
    }
  }
  if (blockIdx.x == 0)
    for (int t = threadIdx.x; t < k; t += PRUNE_THREADS) {
      const bool ok = sk[t] != KEY_INVALID;
      if (out_index) out_index[t] = ok ? (int64_t)si[t] : -1;
      if (out_value) out_value[t] = ok ? key_value(sk[t]) : __longlong_as_double(0x7ff8000000000000LL);
    }
}

__global__ void hsl_topk_emit_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ idxs, int k, int64_t* __restrict__ out_index,
                                     double* __restrict__ out_value) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k) return;
  const uint64_t key = keys[i];
  const bool ok = key != KEY_INVALID;
  if (out_index) out_index[i] = ok ? (int64_t)idxs[i] : -1;
  if (out_value) out_value[i] = ok ? key_value(key) : __longlong_as_double(0x7ff8000000000000LL);
}

}  // namespace

cudaError_t hsl_launch_argmin(const double* cost, int64_t n, int64_t* out_index, double* out_value, cudaStream_t st) {
  HslPdlConfig pc(dim3(1), dim3(1024), 0, st);
  return cudaLaunchKernelEx(&pc.cfg, hsl_argmin_kernel<false>, cost, n, out_index, out_value, (const unsigned long long*)nullptr, 0, 0ull);
}
cudaError_t hsl_launch_argmin_gathered(const double* cost, int64_t n, int64_t* out_index, double* out_value, const unsigned long long* flags,
                                       int nranks, unsigned long long epoch, cudaStream_t st) {
  HslPdlConfig pc(dim3(1), dim3(1024), 0, st);
  return cudaLaunchKernelEx(&pc.cfg, hsl_argmin_kernel<true>, cost, n, out_index, out_value, flags, nranks, epoch);
}

cudaError_t hsl_launch_topk(const double* cost, int64_t n, int k, int64_t* out_index, double* out_value, cudaStream_t st) {
  int64_t P = TILE;
  while (P < n) P <<= 1;
  uint64_t* keys = nullptr;
  cudaError_t e = cudaMallocAsync((void**)&keys, (size_t)P * (sizeof(uint64_t) + sizeof(uint32_t)), st);  // stream-ordered workspace
  if (e != cudaSuccess) return e;
  uint32_t* idxs = (uint32_t*)(keys + P);
  const unsigned tiles = (unsigned)(P / TILE);
  {
    static int coop_blocks[64];  // co-resident blocks of the cooperative kernel, per device (0 = not queried yet, -1 = unsupported)
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64) {
      if (coop_blocks[dev] == 0) {
        int coop = 0, sms = 0, per_sm = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hsl_topk_coop_kernel, SORT_THREADS, 0);
        coop_blocks[dev] = (coop && per_sm > 0) ? sms * per_sm : -1;
      }
      if (coop_blocks[dev] > 0 && (int)tiles <= coop_blocks[dev] && k <= TILE) {   // the common case: an elite set of at most one tile
        int64_t n_ = n;
        int k_ = k, tiles_ = (int)tiles;
        void* args[] = {(void*)&cost, (void*)&n_, (void*)&keys, (void*)&idxs, (void*)&tiles_, (void*)&k_, (void*)&out_index, (void*)&out_value};
        e = cudaLaunchCooperativeKernel((const void*)hsl_topk_prune_kernel, dim3(tiles), dim3(PRUNE_THREADS), args, 0, st);
        const cudaError_t ef = cudaFreeAsync(keys, st);
        return e != cudaSuccess ? e : ef;
      }
      if (coop_blocks[dev] > 0 && (int)tiles <= coop_blocks[dev]) {
        int64_t n_ = n, P_ = P;
        int k_ = k;
        void* args[] = {(void*)&cost, (void*)&n_, (void*)&keys, (void*)&idxs, (void*)&P_, (void*)&k_, (void*)&out_index, (void*)&out_value};
        e = cudaLaunchCooperativeKernel((const void*)hsl_topk_coop_kernel, dim3(tiles), dim3(SORT_THREADS), args, 0, st);
        const cudaError_t ef = cudaFreeAsync(keys, st);
        return e != cudaSuccess ? e : ef;
      }
    }
  }
  hsl_sort_tiles_kernel<<<tiles, SORT_THREADS, 0, st>>>(cost, n, keys, idxs);
  for (int64_t s = 2 * (int64_t)TILE; s <= P; s <<= 1) {
    for (int64_t j = s >> 1; j >= TILE; j >>= 1)
      hsl_sort_global_step_kernel<<<(unsigned)((P / 2 + 255) / 256), 256, 0, st>>>(keys, idxs, P, s, j);
    hsl_sort_merge_kernel<<<tiles, SORT_THREADS, 0, st>>>(keys, idxs, s);
  }
  hsl_topk_emit_kernel<<<(k + 255) / 256, 256, 0, st>>>(keys, idxs, k, out_index, out_value);
  e = cudaGetLastError();
  const cudaError_t e2 = cudaFreeAsync(keys, st);
  return e != cudaSuccess ? e : e2;
}

int hsl_topk_launches(int64_t n) {  // kernels the multi-launch path of hsl_launch_topk issues for n costs (the cooperative path: 1)
  int64_t P = TILE;
  while (P < n) P <<= 1;
  int c = 2;
  for (int64_t s = 2 * (int64_t)TILE; s <= P; s <<= 1) {
    for (int64_t j = s >> 1; j >= TILE; j >>= 1) c++;
    c++;
  }
  return c;
}
