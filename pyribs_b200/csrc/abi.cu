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

// Thin runtime wrappers: the host side of add() issues its copies / event edges through these (one ctypes
// call each) instead of through framework objects whose per-call overhead exceeds the kernels' run time.
extern "C" int qd_memcpy_async(void* dst, const void* src, size_t nbytes, int kind, void* stream) {
  const cudaMemcpyKind k = kind == 1 ? cudaMemcpyHostToDevice
                           : kind == 2 ? cudaMemcpyDeviceToHost
                           : kind == 3 ? cudaMemcpyDeviceToDevice
                                       : cudaMemcpyDefault;
  QD_CHECK_CUDA(cudaMemcpyAsync(dst, src, nbytes, k, (cudaStream_t)stream));
  return QD_OK;
}
extern "C" int qd_event_record(void* event, void* stream) {
  QD_CHECK_CUDA(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream));
  return QD_OK;
}
extern "C" int qd_stream_wait_event(void* stream, void* event) {
  QD_CHECK_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0));
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

// ---- data-parallel insertion: push all-gather over NVLink peer memory -----------------------------
// Every rank writes its slice of every batch field straight into the global field arrays of every
// peer (P2P stores into symmetric memory): the exchange step of the data-parallel add() produces the
// contiguous (R x B_local)-row arrays the insert kernel reads, with no unpack pass and no library
// collective. One local read, R remote writes per 16-byte chunk.
namespace qd {
constexpr int kPushMaxPeers = 16;
constexpr int kPushMaxFields = QD_MAX_FIELDS + 4;
struct PushArgs {
  int n_fields, n_peers, rank;
  const char* src[kPushMaxFields];
  int64_t slice_bytes[kPushMaxFields];
  int64_t dst_off[kPushMaxFields];  // offset of the field's global array inside every peer's buffer
  char* peer[kPushMaxPeers];
};

__global__ void __launch_bounds__(256) dp_push_kernel(const __grid_constant__ PushArgs a) {
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int f = 0; f < a.n_fields; ++f) {
    const int64_t nb = a.slice_bytes[f];
    const char* __restrict__ s = a.src[f];
    const int64_t off = a.dst_off[f] + (int64_t)a.rank * nb;
    const bool vec = (((uintptr_t)s | (uintptr_t)nb | (uintptr_t)off) & 15) == 0;
    if (vec) {
      const int64_t n16 = nb >> 4;
      for (int64_t c = tid; c < n16; c += 2 * nthreads) {
        const int64_t c1 = c + nthreads;
        const uint4 v0 = __ldcs(reinterpret_cast<const uint4*>(s) + c);
        uint4 v1 = v0;
        if (c1 < n16) v1 = __ldcs(reinterpret_cast<const uint4*>(s) + c1);
        for (int p = 0; p < a.n_peers; ++p) {
          uint4* d = reinterpret_cast<uint4*>(a.peer[p] + off);
          d[c] = v0;
          if (c1 < n16) d[c1] = v1;
        }
      }
    } else {
      for (int64_t b = tid; b < nb; b += nthreads) {
        const char v = s[b];
        for (int p = 0; p < a.n_peers; ++p) a.peer[p][off + b] = v;
      }
    }
  }
}
}  // namespace qd

extern "C" int qd_dp_push(int n_fields, const void* const* src, const int64_t* slice_bytes, const int64_t* dst_off,
                          void* const* peer_bases, int n_peers, int rank, void* stream) {
  using namespace qd;
  QD_REQUIRE(n_fields >= 1 && n_fields <= kPushMaxFields, "too many fields");
  QD_REQUIRE(n_peers >= 1 && n_peers <= kPushMaxPeers && rank >= 0 && rank < n_peers, "bad peer count / rank");
  QD_REQUIRE(src && slice_bytes && dst_off && peer_bases, "null pointer");
  PushArgs a;
  a.n_fields = n_fields;
  a.n_peers = n_peers;
  a.rank = rank;
  int64_t total = 0;
  for (int f = 0; f < n_fields; ++f) {
    a.src[f] = (const char*)src[f];
    a.slice_bytes[f] = slice_bytes[f];
    a.dst_off[f] = dst_off[f];
    total += slice_bytes[f];
  }
  for (int p = 0; p < n_peers; ++p) a.peer[p] = (char*)peer_bases[p];
  dp_push_kernel<<<grid_for(total / 32 + 1, 256, 4), 256, 0, (cudaStream_t)stream>>>(a);
  return check_launch("dp_push_kernel");
}
