// Micro-benchmarks behind DESIGN.md's fp64 notes: DFMA, DMMA (mma.sync m8n8k4 f64), IMAD.WIDE issue rates.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/fp64_peak.cu -o build/fp64_peak
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dmma_kernel(double* out, int iters, double a, double b) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) c0[i] = c1[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void imadwide_kernel(unsigned long long* out, int iters, uint32_t m) {
  unsigned long long x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = (unsigned long long)(uint32_t)x[i] * m + (x[i] >> 32);
  }
  unsigned long long s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Philox4x32-R cost: WIDE = one 32x32->64 multiply per word pair, else separate hi / lo multiplies
template <int ROUNDS, bool WIDE>
__global__ void philox_kernel(uint32_t* out, int iters, uint32_t k0, uint32_t k1) {
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    uint32_t c[4] = {(uint32_t)(blockIdx.x * blockDim.x + threadIdx.x), (uint32_t)it, 0u, 0x5A494731u};
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
      uint32_t hi0, lo0, hi1, lo1;
      if (WIDE) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c[0], p1 = (unsigned long long)0xCD9E8D57u * c[2];
        hi0 = (uint32_t)(p0 >> 32); lo0 = (uint32_t)p0; hi1 = (uint32_t)(p1 >> 32); lo1 = (uint32_t)p1;
      } else {
        asm("mul.hi.u32 %0, %1, %2;" : "=r"(hi0) : "r"(c[0]), "r"(0xD2511F53u));
        asm("mul.lo.u32 %0, %1, %2;" : "=r"(lo0) : "r"(c[0]), "r"(0xD2511F53u));
        asm("mul.hi.u32 %0, %1, %2;" : "=r"(hi1) : "r"(c[2]), "r"(0xCD9E8D57u));
        asm("mul.lo.u32 %0, %1, %2;" : "=r"(lo1) : "r"(c[2]), "r"(0xCD9E8D57u));
      }
      const uint32_t n0 = hi1 ^ c[1] ^ a, n2 = hi0 ^ c[3] ^ b;
      c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    acc ^= c[0] ^ c[1] ^ c[2] ^ c[3];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  const int iters = 4096, grid = sms * 4, block = 512;
  float ms = time_ms([&] { dfma_kernel<8><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
  double n = (double)grid * block * iters * 8;
  printf("{\"what\": \"DFMA\", \"fma_per_s\": %.4g, \"tflops\": %.2f, \"fma_per_clk_per_sm\": %.1f}\n", n / ms * 1e3,
         2 * n / ms * 1e3 / 1e12, n / ms * 1e3 / sms / (clk * 1e3));
  ms = time_ms([&] { dmma_kernel<8><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
  n = (double)grid * (block / 32) * iters * 8 * 256;
  printf("{\"what\": \"DMMA m8n8k4\", \"fma_per_s\": %.4g, \"tflops\": %.2f, \"fma_per_clk_per_sm\": %.1f}\n", n / ms * 1e3,
         2 * n / ms * 1e3 / 1e12, n / ms * 1e3 / sms / (clk * 1e3));
  ms = time_ms([&] { imadwide_kernel<8><<<grid, block>>>((unsigned long long*)out, iters, 0xD2511F53u); });
  n = (double)grid * block * iters * 8;
  printf("{\"what\": \"IMAD.WIDE.U32\", \"per_s\": %.4g, \"per_clk_per_sm\": %.1f, \"clock_khz\": %d}\n", n / ms * 1e3,
         n / ms * 1e3 / sms / (clk * 1e3), clk);
  const int pit = 1024;
  n = (double)grid * block * pit;
  ms = time_ms([&] { philox_kernel<10, true><<<grid, block>>>((uint32_t*)out, pit, 1u, 2u); });
  printf("{\"what\": \"Philox4x32-10 wide\", \"calls_per_s\": %.4g, \"calls_per_clk_per_sm\": %.2f}\n", n / ms * 1e3, n / ms * 1e3 / sms / (clk * 1e3));
  ms = time_ms([&] { philox_kernel<10, false><<<grid, block>>>((uint32_t*)out, pit, 1u, 2u); });
  printf("{\"what\": \"Philox4x32-10 hi/lo\", \"calls_per_s\": %.4g, \"calls_per_clk_per_sm\": %.2f}\n", n / ms * 1e3, n / ms * 1e3 / sms / (clk * 1e3));
  ms = time_ms([&] { philox_kernel<7, true><<<grid, block>>>((uint32_t*)out, pit, 1u, 2u); });
  printf("{\"what\": \"Philox4x32-7 wide\", \"calls_per_s\": %.4g, \"calls_per_clk_per_sm\": %.2f}\n", n / ms * 1e3, n / ms * 1e3 / sms / (clk * 1e3));
  ms = time_ms([&] { philox_kernel<7, false><<<grid, block>>>((uint32_t*)out, pit, 1u, 2u); });
  printf("{\"what\": \"Philox4x32-7 hi/lo\", \"calls_per_s\": %.4g, \"calls_per_clk_per_sm\": %.2f}\n", n / ms * 1e3, n / ms * 1e3 / sms / (clk * 1e3));
  return 0;
}
