// Do the scattered-RED row pass of insert.cu and the TMA row copy slow each other down when they share the SMs?
// (Round 1 measured a pipelined commit -- copy of add k on a side stream next to the row phases of add k+1 -- at
// the SUM of the two and blamed the L2.) Two plain kernels on two streams: the classify-like pass of
// scatter_probe.cu (threshold gather + 2 red.max + 2 red.add per row, 2M rows into 250k cells) and a 2 x 200 MB
// TMA bulk copy; each alone, then together.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o overlap_probe overlap_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

struct __align__(32) Rec { unsigned long long a, b, c, d; };
__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void rows_kernel(const int32_t* __restrict__ idx, int64_t B, Rec* rec, const unsigned long long* __restrict__ tab) {
  constexpr int R = 4;
  const int64_t S = (int64_t)gridDim.x * blockDim.x;
  for (int64_t base = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; base < B; base += R * S) {
    int32_t c[R];
    unsigned long long t[R];
#pragma unroll
    for (int u = 0; u < R; ++u) c[u] = base + u * S < B ? idx[base + u * S] : -1;
#pragma unroll
    for (int u = 0; u < R; ++u)
      if (c[u] >= 0) t[u] = tab[c[u]];
#pragma unroll
    for (int u = 0; u < R; ++u)
      if (c[u] >= 0 && t[u] != 12345ull) {
        atomicMax(&rec[c[u]].a, (unsigned long long)(base + u));
        atomicMax(&rec[c[u]].b, (unsigned long long)(base + 2 * u));
        atomicAdd(&rec[c[u]].c, 3ull);
        atomicAdd(&rec[c[u]].d, 5ull);
      }
  }
}

// HINT: the copy's loads and stores carry an L2 evict_first policy -- 400 MB streaming through the L2 otherwise evict
// the 10 MB of per-cell records and thresholds the REDs of the row pass live on
template <int STAGES, bool HINT>
__global__ void copy_tma(const char* __restrict__ s, char* __restrict__ d, int64_t rows, int piece, int slot_bytes, int warps_used) {
  uint64_t pol = 0;
  if (HINT) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
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
      if (act) {
        if (HINT) asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(su32(base + st * slot_bytes + lane * piece)), "l"(s + item * piece), "r"(len), "r"(su32(&bar[warp * STAGES + st])), "l"(pol) : "memory");
        else asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(su32(base + st * slot_bytes + lane * piece)), "l"(s + item * piece), "r"(len), "r"(su32(&bar[warp * STAGES + st])) : "memory");
      }
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
    if (lane < L && item < rows) {
      if (HINT) asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(d + item * piece), "r"(su32(base + st * slot_bytes + lane * piece)), "r"(piece), "l"(pol) : "memory");
      else asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d + item * piece), "r"(su32(base + st * slot_bytes + lane * piece)), "r"(piece) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
    issue(st);
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void fill_idx(int32_t* idx, int64_t B, int64_t N) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long x = (unsigned long long)i * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32; x *= 0x94D049BB133111EBull; x ^= x >> 29;
    idx[i] = (int32_t)(x % (unsigned long long)N);
  }
}

int main() {
  const int64_t B = 2000000, N = 250000, rows = 250000; const int piece = 800; const int64_t bytes = rows * piece;
  int32_t* idx; Rec* rec; unsigned long long* tab; char *s, *d, *flush;
  CK(cudaMalloc(&idx, B * 4)); CK(cudaMalloc(&rec, N * sizeof(Rec))); CK(cudaMalloc(&tab, N * 8));
  CK(cudaMemset(rec, 0, N * sizeof(Rec))); CK(cudaMemset(tab, 0, N * 8));
  CK(cudaMalloc(&s, bytes)); CK(cudaMalloc(&d, bytes)); CK(cudaMemset(s, 1, bytes)); CK(cudaMalloc(&flush, 512 << 20));
  fill_idx<<<1024, 256>>>(idx, B, N);
  CK(cudaDeviceSynchronize());
  cudaStream_t sa, sb; CK(cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking));
  cudaEvent_t a0, a1, b0, b1; cudaEventCreate(&a0); cudaEventCreate(&a1); cudaEventCreate(&b0); cudaEventCreate(&b1);
  CK(cudaFuncSetAttribute(copy_tma<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304));
  CK(cudaFuncSetAttribute(copy_tma<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304));
  struct Cfg { int rg, rb, cg, cb, cw, slot, hint; };
  const Cfg cfgs[] = {{296, 512, 148, 128, 4, 12288, 0}, {296, 512, 148, 128, 4, 12288, 1}, {148, 512, 148, 128, 4, 12288, 1},
                      {296, 512, 148, 64, 2, 24576, 0},  {296, 512, 148, 64, 2, 24576, 1},  {296, 512, 296, 64, 2, 12288, 1}};
  for (const Cfg& c : cfgs) {
    const int csm = c.cw * 2 * c.slot;
    float ra = 1e9, ca = 1e9, rt = 1e9, ct = 1e9, both = 1e9;
    for (int r = 0; r < 5; ++r) {
      float ms;
      cudaMemsetAsync(flush, r, 512 << 20, sa); cudaStreamSynchronize(sa);
      cudaEventRecord(a0, sa); rows_kernel<<<c.rg, c.rb, 0, sa>>>(idx, B, rec, tab); cudaEventRecord(a1, sa); cudaEventSynchronize(a1);
      cudaEventElapsedTime(&ms, a0, a1); ra = ms < ra ? ms : ra;
      cudaMemsetAsync(flush, r, 512 << 20, sa); cudaStreamSynchronize(sa);
      cudaEventRecord(b0, sb); (c.hint ? copy_tma<2, true> : copy_tma<2, false>)<<<c.cg, c.cb, csm, sb>>>(s, d, rows, piece, c.slot, c.cw); cudaEventRecord(b1, sb); cudaEventSynchronize(b1);
      cudaEventElapsedTime(&ms, b0, b1); ca = ms < ca ? ms : ca;
      cudaMemsetAsync(flush, r, 512 << 20, sa); cudaStreamSynchronize(sa);
      cudaEventRecord(b0, sb); (c.hint ? copy_tma<2, true> : copy_tma<2, false>)<<<c.cg, c.cb, csm, sb>>>(s, d, rows, piece, c.slot, c.cw); cudaEventRecord(b1, sb);
      cudaEventRecord(a0, sa); rows_kernel<<<c.rg, c.rb, 0, sa>>>(idx, B, rec, tab); cudaEventRecord(a1, sa);
      cudaEventSynchronize(a1); cudaEventSynchronize(b1);
      float x, y, z;
      cudaEventElapsedTime(&x, a0, a1); cudaEventElapsedTime(&y, b0, b1); cudaEventElapsedTime(&z, b0, a1);
      float w; cudaEventElapsedTime(&w, b0, b1); z = z > w ? z : w;
      if (z < both) { both = z; rt = x; ct = y; }
    }
    cudaError_t e = cudaGetLastError();
    printf("rows %dx%d, copy %dx%d (%d warps x %d B x 2 stages, %s): rows alone %.1f us, copy alone %.1f us (%.0f GB/s) | together: rows %.1f us, copy %.1f us, both done after %.1f us %s\n",
           c.rg, c.rb, c.cg, c.cb, c.cw, c.slot, c.hint ? "L2 evict_first" : "no hint", ra * 1e3, ca * 1e3, 2.0 * bytes / ca / 1e6, rt * 1e3, ct * 1e3, both * 1e3, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
