// Exact fp64 squared distance in NumPy's summation order, shared by the fp64 scan and the
// rescoring pass of the tensor-core search.
//
// Reference: ribs/archives/_cvt_archive.py:626-632 -- np.sum(np.square(x - c), axis=2).
// NumPy reduces the contiguous last axis with its pairwise routine: fewer than 8 terms are
// added left to right; 8..128 terms use eight strided accumulators combined as
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) followed by the remainder in order (verified against
// NumPy 2.3 in this image for every d <= 33). Square and add are rounded separately, so no
// FMA contraction is allowed here.
#pragma once
#include "qd_common.cuh"

namespace qd {

template <int D>
__device__ __forceinline__ double sqdist_np(const double* __restrict__ x, const double* __restrict__ c, int d_rt) {
  const int d = D > 0 ? D : d_rt;
  if (d < 8) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < (D > 0 ? D : 7); ++k) {
      if (k < d) {
        double t = __dsub_rn(x[k], c[k]);
        t = __dmul_rn(t, t);
        s = (k == 0) ? t : __dadd_rn(s, t);
      }
    }
    return s;
  }
  double r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    double t = __dsub_rn(x[j], c[j]);
    r[j] = __dmul_rn(t, t);
  }
  int i = 8;
  const int lim = d - (d % 8);
  for (; i < lim; i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      double t = __dsub_rn(x[i + j], c[i + j]);
      r[j] = __dadd_rn(r[j], __dmul_rn(t, t));
    }
  }
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (; i < d; ++i) {
    double t = __dsub_rn(x[i], c[i]);
    res = __dadd_rn(res, __dmul_rn(t, t));
  }
  return res;
}

}  // namespace qd
