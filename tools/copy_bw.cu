// Ceiling probe for the row-copy phase: how fast can 206 MB be moved HBM -> HBM on one B200 with
// (a) cudaMemcpyAsync D2D, (b) a plain uint4 grid-stride copy at several launch shapes, (c) per-warp TMA bulk
// copies through shared memory (the mechanism of insert.cu phase 5) at several slot sizes / stages.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o copy_bw copy_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int U>
__global__ void copy_plain(const uint4* __restrict__ s, uint4* __restrict__ d, int64_t n) {
  const int64_t T = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += U * T) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) if (i + u * T < n) v[u] = __ldcs(s + i + u * T);
#pragma unroll
    for (int u = 0; u < U; ++u) if (i + u * T < n) __stcs(d + i + u * T, v[u]);
  }
}

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// each warp: STAGES slots of SLOT bytes; lane moves `piece` bytes per round
template <int STAGES>
__global__ void copy_tma(const char* __restrict__ s, char* __restrict__ d, int64_t rows, int piece, int slot_bytes, int warps_used) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint64_t bar[32 * STAGES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int k = 0; k < warps_used * STAGES; ++k) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bar[k])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp >= warps_used) return;
  const int L = slot_bytes / piece < 32 ? slot_bytes / piece : 32;
  const int64_t gw = (int64_t)blockIdx.x * warps_used + warp, ngw = (int64_t)gridDim.x * warps_used;
  unsigned char* base = sm + (size_t)warp * STAGES * slot_bytes;
  uint32_t par[STAGES];
  int64_t it[STAGES];
  for (int st = 0; st < STAGES; ++st) par[st] = 0;
  // prologue: fill all stages
  int64_t next = gw * L;
  auto issue = [&](int st) {
    it[st] = next;
    if (next < rows) {
      const int64_t item = next + lane;
      const bool act = lane < L && item < rows;
      uint32_t len = act ? piece : 0;
      uint32_t tot = __reduce_add_sync(0xffffffffu, len);
      if (lane == 0) asm volatile("{\n\t.reg .b64 t;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1;\n\t}" ::"r"(su32(&bar[warp * STAGES + st])), "r"(tot) : "memory");
      __syncwarp();
      if (act) asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(su32(base + st * slot_bytes + lane * piece)), "l"(s + item * piece), "r"(len), "r"(su32(&bar[warp * STAGES + st])) : "memory");
    }
    next += ngw * L;
  };
  for (int st = 0; st < STAGES; ++st) issue(st);
  for (int st = 0;; st = (st + 1) % STAGES) {
    if (it[st] >= rows) break;
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(su32(&bar[warp * STAGES + st])), "r"(par[st]) : "memory");
    par[st] ^= 1;
    const int64_t item = it[st] + lane;
    if (lane < L && item < rows) asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d + item * piece), "r"(su32(base + st * slot_bytes + lane * piece)), "r"(piece) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
    issue(st);
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  const int64_t rows = 250000; const int piece = 800; const int64_t bytes = rows * piece;  // 200 MB
  char *s, *d; CK(cudaMalloc(&s, bytes)); CK(cudaMalloc(&d, bytes)); CK(cudaMemset(s, 1, bytes));
  char* flush; CK(cudaMalloc(&flush, 512 << 20));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](const char* name, auto fn) {
    float best = 1e9;
    for (int r = 0; r < 5; ++r) {
      cudaMemsetAsync(flush, r, 512 << 20);
      cudaEventRecord(e0); fn(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    printf("%-52s %7.1f us  %6.0f GB/s %s\n", name, best * 1e3, 2.0 * bytes / best / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
  };
  timeit("cudaMemcpyAsync D2D", [&] { cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice); });
  const int64_t n16 = bytes / 16;
  timeit("plain uint4 U=4, 148*8 x 256", [&] { copy_plain<4><<<148 * 8, 256>>>((const uint4*)s, (uint4*)d, n16); });
  timeit("plain uint4 U=4, 296 x 512", [&] { copy_plain<4><<<296, 512>>>((const uint4*)s, (uint4*)d, n16); });
  timeit("plain uint4 U=8, 296 x 512", [&] { copy_plain<8><<<296, 512>>>((const uint4*)s, (uint4*)d, n16); });
  timeit("plain uint4 U=8, 148*8 x 256", [&] { copy_plain<8><<<148 * 8, 256>>>((const uint4*)s, (uint4*)d, n16); });
  timeit("plain uint4 U=2, 148*16 x 128", [&] { copy_plain<2><<<148 * 16, 128>>>((const uint4*)s, (uint4*)d, n16); });
  cudaFuncSetAttribute(copy_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
  cudaFuncSetAttribute(copy_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
  cudaFuncSetAttribute(copy_tma<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
  timeit("tma 1 stage, 4 warps x 24 KB, 296 x 512", [&] { copy_tma<1><<<296, 512, 98304>>>(s, d, rows, piece, 24576, 4); });
  timeit("tma 2 stages, 4 warps x 12 KB, 296 x 512", [&] { copy_tma<2><<<296, 512, 98304>>>(s, d, rows, piece, 12288, 4); });
  timeit("tma 3 stages, 4 warps x 8 KB, 296 x 512", [&] { copy_tma<3><<<296, 512, 98304>>>(s, d, rows, piece, 8192, 4); });
  timeit("tma 2 stages, 8 warps x 6 KB, 296 x 512", [&] { copy_tma<2><<<296, 512, 98304>>>(s, d, rows, piece, 6144, 8); });
  timeit("tma 1 stage, 16 warps x 6 KB, 296 x 512", [&] { copy_tma<1><<<296, 512, 98304>>>(s, d, rows, piece, 6144, 16); });
  timeit("tma 2 stages, 2 warps x 24 KB, 296 x 512", [&] { copy_tma<2><<<296, 512, 98304>>>(s, d, rows, piece, 24576, 2); });
  timeit("tma 2 stages, 1 warp x 24 KB, 148 x 32 (48 KB)", [&] { copy_tma<2><<<148, 32, 49152>>>(s, d, rows, piece, 24576, 1); });
  timeit("tma 2 stages, 4 warps x 12 KB, 148 x 128", [&] { copy_tma<2><<<148, 128, 98304>>>(s, d, rows, piece, 12288, 4); });
  return 0;
}
