// ABI bookkeeping: version, per-thread error text, launch counter.
#include <cstdarg>
#include <cstring>

#include "qd_common.cuh"

namespace qd {
static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace qd

extern "C" int qd_abi_version(void) { return QD_ABI_VERSION; }
extern "C" const char* qd_last_error(void) { return qd::g_err; }
extern "C" int64_t qd_launch_count(void) { return qd::g_launches.load(); }

extern "C" int qd_stream_synchronize(void* stream) {
  QD_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return QD_OK;
}

namespace qd {
__global__ void flags_or_kernel(const uint32_t* __restrict__ in, int n, uint32_t* out) {
  uint32_t v = 0;
  for (int i = threadIdx.x; i < n; i += 32) v |= in[i];
  v = __reduce_or_sync(0xffffffffu, v);
  if (threadIdx.x == 0) *out = v;
}
}  // namespace qd

extern "C" int qd_flags_or(const uint32_t* in, int n, uint32_t* out, void* stream) {
  QD_REQUIRE(in && out && n >= 1, "bad arguments");
  qd::flags_or_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(in, n, out);
  return qd::check_launch("flags_or_kernel");
}
