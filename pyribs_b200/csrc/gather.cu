// A6: ArrayStore.retrieve (reference ribs/archives/_array_store.py:330-485) -- row gather of
// every requested field, one warp per row, coalesced within the row.
#include "qd_common.cuh"

namespace qd {

struct GatherArgs {
  int n;
  const char* src[QD_MAX_FIELDS];
  char* dst[QD_MAX_FIELDS];
  int64_t row_bytes[QD_MAX_FIELDS];
  int vec[QD_MAX_FIELDS];
};

__global__ void __launch_bounds__(256) gather_kernel(int64_t M, const int32_t* __restrict__ idx,
                                                     const __grid_constant__ GatherArgs a,
                                                     const uint8_t* __restrict__ occupied,
                                                     uint8_t* __restrict__ occupied_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < M; r += nwarps) {
    const int64_t c = idx[r];
    if (lane == 0 && occupied_out) occupied_out[r] = occupied[c];
    for (int f = 0; f < a.n; ++f) {
      const char* s = a.src[f] + c * a.row_bytes[f];
      char* d = a.dst[f] + r * a.row_bytes[f];
      const int64_t nb = a.row_bytes[f];
      if (a.vec[f] == 16) {
        for (int64_t o = lane * 16; o < nb; o += 512) *(uint4*)(d + o) = *(const uint4*)(s + o);
      } else if (a.vec[f] == 8) {
        for (int64_t o = lane * 8; o < nb; o += 256) *(uint2*)(d + o) = *(const uint2*)(s + o);
      } else if (a.vec[f] == 4) {
        for (int64_t o = lane * 4; o < nb; o += 128) *(uint32_t*)(d + o) = *(const uint32_t*)(s + o);
      } else {
        for (int64_t o = lane; o < nb; o += 32) d[o] = s[o];
      }
    }
  }
}

}  // namespace qd

extern "C" int qd_store_gather(int64_t M, const int32_t* idx, int n_fields, const void* const* field_src,
                               void* const* field_out, const int64_t* field_row_bytes, const uint8_t* occupied,
                               uint8_t* occupied_out, void* stream) {
  using namespace qd;
  QD_REQUIRE(n_fields >= 0 && n_fields <= QD_MAX_FIELDS, "too many fields");
  QD_REQUIRE((occupied == nullptr) == (occupied_out == nullptr), "occupied in/out must both be given or both NULL");
  if (M == 0) return QD_OK;
  QD_REQUIRE(idx, "null idx");
  GatherArgs a;
  a.n = n_fields;
  for (int f = 0; f < n_fields; ++f) {
    a.src[f] = (const char*)field_src[f];
    a.dst[f] = (char*)field_out[f];
    a.row_bytes[f] = field_row_bytes[f];
    uintptr_t x = (uintptr_t)a.src[f] | (uintptr_t)a.dst[f] | (uintptr_t)a.row_bytes[f];
    a.vec[f] = (x & 15) == 0 ? 16 : (x & 7) == 0 ? 8 : (x & 3) == 0 ? 4 : 1;
  }
  int grid = grid_for(M * 32, 256, 8);
  gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(M, idx, a, occupied, occupied_out);
  return check_launch("gather_kernel");
}
