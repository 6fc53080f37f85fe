// What does one scattered access per batch row cost on this part, by flavour? Decides the layout of the row
// passes of insert.cu (threshold gather + REDs on the per-cell record): B random cells out of N, the launch shape
// of insert_kernel (2 x 512 threads per SM, 4 rows in flight per thread), cycles per row per SM =
// t * f_sm * SMs / B.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scatter_probe scatter_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

struct __align__(32) Rec { unsigned long long a, b, c, d; };

enum { G_CA, G_CG, G_NC, G_CV, G_REC32, G_U32, R_MAX, R_ADD1, R_ADD2, R_ADD3, R_MAXADD2, A_ADD, S_8, S_16, S_32,
       C_CLASSIFY, C_CLASSIFY_CG, B_RED16, B_RED32, R_MAX32, R_ADD32, C_FOUR, B_BATCH16, C_MIX, C_MIX_NOSUM, NKIND };
const char* kNames[NKIND] = {"gather8 ld.ca", "gather8 ld.cg", "gather8 ld.nc", "gather8 ld.volatile", "gather 32B record (2x LDG.128 cg)",
                             "gather4 ld.cg", "red.max.u64", "red.add.u64 x1", "red.add.u64 x2 same sector", "red.add.u64 x3 same sector",
                             "red.max + 2 red.add same sector", "atom.add.u64 returning", "store 8B", "store 16B", "store 32B (2x16)",
                             "classify-like: gather8.ca + max + 2 add", "classify-like: gather8.cg + max + 2 add",
                             "cp.reduce.async.bulk add.u64 16B", "cp.reduce.async.bulk add.u64 32B", "red.max.u32", "red.add.u32",
                             "gather8 + 2 red.max + 2 red.add (one sector)", "bulk add.u64 16B, 4 ops per commit",
                             "gather8 + 2 red.max (LSU) + bulk add 16B (TMA)", "gather8 + 2 red.max"};

__device__ __forceinline__ unsigned long long ld_cg(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned long long ld_nc(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned long long ld_ca(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.global.ca.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned long long ld_cv(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int KIND>
__global__ void __launch_bounds__(512, 2) probe(const int32_t* __restrict__ idx, int64_t B, Rec* rec, unsigned long long* tab,
                                                unsigned long long* sink) {
  constexpr int R = 4;
  const int64_t S = (int64_t)gridDim.x * blockDim.x;
  unsigned long long acc = 0;
  __shared__ __align__(16) unsigned long long stage[512 * 4];
  for (int64_t base = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; base < B; base += R * S) {
    int32_t c[R];
#pragma unroll
    for (int u = 0; u < R; ++u) c[u] = base + u * S < B ? idx[base + u * S] : -1;
    if constexpr (KIND == C_CLASSIFY || KIND == C_CLASSIFY_CG) {
      unsigned long long t[R];
#pragma unroll
      for (int u = 0; u < R; ++u)
        if (c[u] >= 0) t[u] = KIND == C_CLASSIFY ? ld_ca(tab + c[u]) : ld_cg(tab + c[u]);
#pragma unroll
      for (int u = 0; u < R; ++u)
        if (c[u] >= 0 && t[u] != 12345ull) {
          atomicMax(&rec[c[u]].a, (unsigned long long)(base + u));
          atomicAdd(&rec[c[u]].b, 3ull);
          atomicAdd(&rec[c[u]].c, 5ull);
        }
      continue;
    }
#pragma unroll
    for (int u = 0; u < R; ++u) {
      if (c[u] < 0) continue;
      Rec* r = rec + c[u];
      if constexpr (KIND == G_CA) acc += ld_ca(tab + c[u]);
      if constexpr (KIND == G_CG) acc += ld_cg(tab + c[u]);
      if constexpr (KIND == G_NC) acc += ld_nc(tab + c[u]);
      if constexpr (KIND == G_CV) acc += ld_cv(tab + c[u]);
      if constexpr (KIND == G_U32) {
        uint32_t v;
        asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"((const uint32_t*)tab + c[u]));
        acc += v;
      }
      if constexpr (KIND == G_REC32) {
        const uint4 x = __ldcg((const uint4*)r), y = __ldcg((const uint4*)r + 1);
        acc += x.x + y.w;
      }
      if constexpr (KIND == R_MAX) atomicMax(&r->a, (unsigned long long)(base + u));
      if constexpr (KIND == R_MAX32) atomicMax((uint32_t*)&r->a, (uint32_t)(base + u));
      if constexpr (KIND == R_ADD32) atomicAdd((uint32_t*)&r->a, 3u);
      if constexpr (KIND == R_ADD1 || KIND == R_ADD2 || KIND == R_ADD3) atomicAdd(&r->a, 3ull);
      if constexpr (KIND == R_ADD2 || KIND == R_ADD3) atomicAdd(&r->b, 5ull);
      if constexpr (KIND == R_ADD3) atomicAdd(&r->c, 7ull);
      if constexpr (KIND == R_MAXADD2) {
        atomicMax(&r->a, (unsigned long long)(base + u));
        atomicAdd(&r->b, 3ull);
        atomicAdd(&r->c, 5ull);
      }
      if constexpr (KIND == A_ADD) acc += atomicAdd(&r->a, 1ull);
      if constexpr (KIND == S_8) r->a = base;
      if constexpr (KIND == S_16) *(uint4*)r = make_uint4(base, u, 3, 4);
      if constexpr (KIND == S_32) {
        *(uint4*)r = make_uint4(base, u, 3, 4);
        *((uint4*)r + 1) = make_uint4(base, u, 5, 6);
      }
    }
    if constexpr (KIND == C_FOUR || KIND == C_MIX || KIND == C_MIX_NOSUM || KIND == B_BATCH16) {
      unsigned long long t[R];
      unsigned long long* st = stage + threadIdx.x * 4;  // (two u64 per row would do; 4 rows x 2 = 8 words needs a second slot)
      __shared__ __align__(16) unsigned long long stage2[512 * 4];
      unsigned long long* st2 = stage2 + threadIdx.x * 4;
#pragma unroll
      for (int u = 0; u < R; ++u)
        if (c[u] >= 0 && KIND != B_BATCH16) t[u] = ld_ca(tab + c[u]);
      if constexpr (KIND == C_MIX || KIND == B_BATCH16) {
        st[0] = 3; st[1] = 5; st[2] = 7; st[3] = 9;
        st2[0] = 3; st2[1] = 5; st2[2] = 7; st2[3] = 9;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
#pragma unroll
      for (int u = 0; u < R; ++u)
        if (c[u] >= 0 && (KIND == B_BATCH16 || t[u] != 12345ull)) {
          if constexpr (KIND != B_BATCH16) {
            atomicMax(&rec[c[u]].a, (unsigned long long)(base + u));
            atomicMax(&rec[c[u]].b, (unsigned long long)(base + 2 * u));
          }
          if constexpr (KIND == C_FOUR) {
            atomicAdd(&rec[c[u]].c, 3ull);
            atomicAdd(&rec[c[u]].d, 5ull);
          }
          if constexpr (KIND == C_MIX || KIND == B_BATCH16) {
            const unsigned long long* src = (u < 2 ? st : st2) + 2 * (u & 1);
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u64 [%0], [%1], 16;" ::"l"(&rec[c[u]].c),
                         "r"(su32(src))
                         : "memory");
          }
        }
      if constexpr (KIND == C_MIX || KIND == B_BATCH16) {
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      continue;
    }
    if constexpr (KIND == B_RED16 || KIND == B_RED32) {
      constexpr int NB = KIND == B_RED16 ? 16 : 32;
#pragma unroll
      for (int u = 0; u < R; ++u) {
        if (c[u] < 0) continue;
        unsigned long long* s = stage + threadIdx.x * 4;
        s[0] = 3;
        s[1] = 5;
        s[2] = 7;
        s[3] = 9;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u64 [%0], [%1], %2;" ::"l"(rec + c[u]),
                     "r"(su32(s)), "n"(NB)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
    }
  }
  if constexpr (KIND == B_RED16 || KIND == B_RED32) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (acc == 0x123456789ull) *sink = acc;
}

__global__ void fill_idx(int32_t* idx, int64_t B, int64_t N, int clustered) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long x = (unsigned long long)i * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32; x *= 0x94D049BB133111EBull; x ^= x >> 29;
    idx[i] = clustered ? (int32_t)(((i >> 9) * 2477 + (x % 64)) % N) : (int32_t)(x % (unsigned long long)N);
  }
}

template <int KIND>
int run(const int32_t* idx, int64_t B, Rec* rec, unsigned long long* tab, unsigned long long* sink, double ghz, int sms) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int grid = 2 * sms;
  for (int k = 0; k < 3; ++k) probe<KIND><<<grid, 512>>>(idx, B, rec, tab, sink);
  CK(cudaDeviceSynchronize());
  const int it = 20;
  CK(cudaEventRecord(e0));
  for (int k = 0; k < it; ++k) probe<KIND><<<grid, 512>>>(idx, B, rec, tab, sink);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double us = ms * 1e3 / it;
  printf("  %-48s %8.2f us  %6.2f cyc/row/SM\n", kNames[KIND], us, us * 1e-6 * ghz * 1e9 * sms / B);
  return 0;
}

template <int K>
int run_all(const int32_t* idx, int64_t B, Rec* rec, unsigned long long* tab, unsigned long long* sink, double ghz, int sms) {
  if (run<K>(idx, B, rec, tab, sink, ghz, sms)) return 1;
  if constexpr (K + 1 < NKIND) return run_all<K + 1>(idx, B, rec, tab, sink, ghz, sms);
  return 0;
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  int khz = 0;
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  const double ghz = khz * 1e-6;
  const int sms = p.multiProcessorCount;
  printf("%s, %d SMs, %.3f GHz nominal\n", p.name, sms, ghz);
  const int64_t B = 2000000;
  int32_t* idx;
  unsigned long long* sink;
  CK(cudaMalloc(&idx, B * 4));
  CK(cudaMalloc(&sink, 8));
  for (int64_t N : {250000ll, 4000000ll}) {
    for (int clustered = 0; clustered < 2; ++clustered) {
      Rec* rec;
      unsigned long long* tab;
      CK(cudaMalloc(&rec, N * sizeof(Rec)));
      CK(cudaMalloc(&tab, N * 8));
      CK(cudaMemset(rec, 0, N * sizeof(Rec)));
      CK(cudaMemset(tab, 0, N * 8));
      fill_idx<<<1024, 256>>>(idx, B, N, clustered);
      CK(cudaDeviceSynchronize());
      printf("B = %lld rows, N = %lld cells, %s\n", (long long)B, (long long)N, clustered ? "clustered (512-row runs over 64 cells)" : "uniform");
      if (run_all<0>(idx, B, rec, tab, sink, ghz, sms)) return 1;
      CK(cudaFree(rec));
      CK(cudaFree(tab));
    }
  }
  return 0;
}
