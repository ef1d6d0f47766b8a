// Block-cooperative complex FFT in shared memory, float64.
//
// Stockham auto-sort, radix-4 passes with one radix-2 pass when log2(n) is odd.
// Data ping-pongs between two shared buffers (no bit reversal); twiddles come
// from a table W[m] = exp(-2 pi i m / tw_n) in global memory (L1-resident, one
// table per handle).  Every FFT of the analysis/synthesis path (sizes 512..8192)
// runs through this routine inside the fused per-frame kernels, so spectra never
// round-trip through HBM.
#pragma once
#include "wb_platform.h"

// dir = -1: forward (e^{-i...}), dir = +1: inverse WITHOUT the 1/n factor.
// Input in `a`; returns the buffer (a or b) that holds the result.  All threads
// of the block must call it; it ends with a barrier.
WB_DEV wb_cplx* wb_fft(wb_cplx* a, wb_cplx* b, int n, int dir, const wb_cplx* tw, int tw_n, int tid, int nthr) {
  const int tw_step = tw_n / n;
  wb_cplx* src = a;
  wb_cplx* dst = b;
  int ns = 1;
  while (ns < n) {
    const int rem = n / ns;
    if ((rem & 3) == 0) {
      const int q = n >> 2;
      const int tstep = tw_step * (n / (ns * 4));
      for (int j = tid; j < q; j += nthr) {
        const int k = j & (ns - 1);
        wb_cplx v0 = src[j], v1 = src[j + q], v2 = src[j + 2 * q], v3 = src[j + 3 * q];
        if (k) {
          wb_cplx w1 = wb_ldg_cplx(tw + k * tstep), w2 = wb_ldg_cplx(tw + 2 * k * tstep), w3 = wb_ldg_cplx(tw + 3 * k * tstep);
          if (dir > 0) {
            w1.y = -w1.y;
            w2.y = -w2.y;
            w3.y = -w3.y;
          }
          v1 = wb_cmul(v1, w1);
          v2 = wb_cmul(v2, w2);
          v3 = wb_cmul(v3, w3);
        }
        const wb_cplx t0 = wb_cadd(v0, v2), t1 = wb_csub(v0, v2), t2 = wb_cadd(v1, v3);
        const wb_cplx d = wb_csub(v1, v3);
        const wb_cplx t3 = dir < 0 ? wb_mk(d.y, -d.x) : wb_mk(-d.y, d.x);
        const int j0 = ((j - k) << 2) + k;
        dst[j0] = wb_cadd(t0, t2);
        dst[j0 + ns] = wb_cadd(t1, t3);
        dst[j0 + 2 * ns] = wb_csub(t0, t2);
        dst[j0 + 3 * ns] = wb_csub(t1, t3);
      }
      ns <<= 2;
    } else {
      const int h = n >> 1;
      const int tstep = tw_step * (n / (ns * 2));
      for (int j = tid; j < h; j += nthr) {
        const int k = j & (ns - 1);
        wb_cplx v0 = src[j], v1 = src[j + h];
        if (k) {
          wb_cplx w1 = wb_ldg_cplx(tw + k * tstep);
          if (dir > 0) w1.y = -w1.y;
          v1 = wb_cmul(v1, w1);
        }
        const int j0 = ((j - k) << 1) + k;
        dst[j0] = wb_cadd(v0, v1);
        dst[j0 + ns] = wb_csub(v0, v1);
      }
      ns <<= 1;
    }
    WB_SYNC();
    wb_cplx* t = src;
    src = dst;
    dst = t;
  }
  return src;
}
