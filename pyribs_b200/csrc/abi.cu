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
