// Block-cooperative complex FFT in shared memory, float64.
//
// Stockham auto-sort, radix-4 passes with one radix-2 pass when log2(n) is odd.  Data ping-pongs between
// two shared buffers (no bit reversal).  Twiddles come from a per-block shared-memory table of one eighth of
// the circle (see below), filled once per block from the handle's global table (the kernels run with the maximum
// shared-memory carve-out, so L1 is too small to keep a global twiddle table resident); w^2k and w^3k are
// formed from w^k by multiplication (table loads measured slower: shared memory is the scarce pipe).  Every FFT of the
// analysis/synthesis path (sizes 512..8192) runs through this routine inside the fused per-frame kernels,
// so spectra never round-trip through HBM.
#pragma once
#include "wb_platform.h"

// Only the first eighth of the circle is tabulated, T[m] = exp(-2 pi i m / (2 h)) for 0 <= m <= h/4 (h = half the
// largest transform size); the other octants follow from the symmetries of sine and cosine (exact: no arithmetic,
// only swaps and sign flips).  A 2048-point table is 4.7 KB instead of 16 KB, which is what lets the per-frame
// kernels keep one more block resident per SM.  The table is stored skewed -- entry m lives at
// m + m/8 + m/64 + m/512 -- so that the power-of-two strides the passes read it with spread over all banks
// (8 lanes x 16 bytes per shared-memory wavefront).  Tables need WB_FFT_TW_SLOTS(h) entries.
#define WB_FFT_TW_COUNT(h) (((h) >> 2) + 1)
#define WB_FFT_TW_SLOTS(h) (WB_FFT_TW_COUNT(h) + (WB_FFT_TW_COUNT(h) >> 3) + (WB_FFT_TW_COUNT(h) >> 6) + (WB_FFT_TW_COUNT(h) >> 9) + 1)
// Kernels whose occupancy is not limited by shared memory use the half-circle table instead (FULL = 1: h entries,
// a plain lookup plus one sign flip).
#define WB_FFT_TW_SLOTS_FULL(h) ((h) + ((h) >> 3) + ((h) >> 6) + ((h) >> 9) + 1)
WB_HD int wb_fft_tw_skew(int m) { return m + (m >> 3) + (m >> 6) + (m >> 9); }

// Fill the shared twiddle table for transforms up to size 2*h from the global table of tw_n entries.
template <int FULL = 0>
WB_DEV void wb_fft_load_twiddles(wb_cplx* T, int h, const wb_cplx* tw, int tw_n, int tid, int nthr) {
  const int step = tw_n / (2 * h);
  const int count = FULL ? h : (h >> 2) + 1;
  for (int m = tid; m < count; m += nthr) T[wb_fft_tw_skew(m)] = wb_ldg_cplx(tw + (size_t)m * step);
  WB_SYNC();
}

WB_HD int wb_fft_log2(int n) {  // n is a power of two
#if !defined(WB_HOST_EMU) && defined(__CUDA_ARCH__)
  return 31 - __clz(n);
#else
  int l = 0;
  while ((1 << l) < n) ++l;
  return l;
#endif
}
// exp(-2 pi i idx / (2 h)) for 0 <= idx < 2 h, idx = m << ts with ts = log2(2 h / n)
template <int FULL = 0>
WB_DEV wb_cplx wb_fft_tw_s(const wb_cplx* T, int h, int ts, int m) {
  int idx = m << ts;
  const bool neg = idx >= h;         // W(t + pi) = -W(t)
  if (neg) idx -= h;
  if (FULL) {
    const wb_cplx t = T[wb_fft_tw_skew(idx)];
    return neg ? wb_mk(-t.x, -t.y) : t;
  }
  const bool rot = idx > (h >> 1);   // t = pi/2 + t': W = (W'.y, -W'.x)
  if (rot) idx -= (h >> 1);
  const bool refl = idx > (h >> 2);  // t = pi/2 - p: W = (-T.y, -T.x)
  const wb_cplx t = T[wb_fft_tw_skew(refl ? (h >> 1) - idx : idx)];
  double wx = refl ? -t.y : t.x, wy = refl ? -t.x : t.y;
  if (rot) {
    const double q = wx;
    wx = wy;
    wy = -q;
  }
  return neg ? wb_mk(-wx, -wy) : wb_mk(wx, wy);
}
template <int FULL = 0>
WB_DEV wb_cplx wb_fft_tw(const wb_cplx* T, int h, int n, int m) {
  return wb_fft_tw_s<FULL>(T, h, wb_fft_log2(2 * h) - wb_fft_log2(n), m);
}

// dir = -1: forward (e^{-i...}), dir = +1: inverse WITHOUT the 1/n factor.
// Input in `a`; returns the buffer (a or b) that holds the result.  All threads of the block must call it;
// it ends with a barrier.  T/h: shared twiddle table as above.
template <int FULL = 0>
WB_DEV_NI wb_cplx* wb_fft(wb_cplx* a, wb_cplx* b, int n, int dir, const wb_cplx* T, int h, int tid, int nthr,
                         int nz = 0x7fffffff) {
  const int ln = wb_fft_log2(n);
  const int ts = wb_fft_log2(2 * h) - ln;
  wb_cplx* src = a;
  wb_cplx* dst = b;
  int ls = 0;  // log2(ns)
  while (ls < ln) {
    const int ns = 1 << ls;
    if (ln - ls >= 2) {
      const int q = n >> 2;
      const int shift = ln - ls - 2;  // twiddle of position k: exp(-2 pi i k / (4 ns)) = table index k << shift (of n)
      if (ls == 0 && nz <= 2 * q) {
        // Zero-padded input (only entries below nz are non-zero and need to have been written): the first pass
        // reads one or two of its four inputs.  nz <= n/4: every output is the input; nz <= n/2: a 2-point DFT
        // and its +-i rotations.
        const bool quarter = nz <= q;
        for (int j = tid; j < q; j += nthr) {
          const wb_cplx v0 = src[j];
          const int j0 = j << 2;
          if (quarter) {
            dst[j0] = v0;
            dst[j0 + 1] = v0;
            dst[j0 + 2] = v0;
            dst[j0 + 3] = v0;
          } else {
            const wb_cplx v1 = src[j + q];
            const wb_cplx r = dir < 0 ? wb_mk(v1.y, -v1.x) : wb_mk(-v1.y, v1.x);  // -+ i v1
            dst[j0] = wb_cadd(v0, v1);
            dst[j0 + 1] = wb_cadd(v0, r);
            dst[j0 + 2] = wb_csub(v0, v1);
            dst[j0 + 3] = wb_csub(v0, r);
          }
        }
        ls += 2;
        WB_SYNC();
        wb_cplx* t = src;
        src = dst;
        dst = t;
        continue;
      }
      for (int j = tid; j < q; j += nthr) {
        const int k = j & (ns - 1);
        wb_cplx v0 = src[j], v1 = src[j + q], v2 = src[j + 2 * q], v3 = src[j + 3 * q];
        if (k) {
          wb_cplx w1 = wb_fft_tw_s<FULL>(T, h, ts, k << shift);
          if (dir > 0) w1.y = -w1.y;
          const wb_cplx w2 = wb_cmul(w1, w1);
          const wb_cplx w3 = wb_cmul(w2, w1);
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
      ls += 2;
    } else {
      const int hh = n >> 1;
      const int shift = ln - ls - 1;
      for (int j = tid; j < hh; j += nthr) {
        const int k = j & (ns - 1);
        wb_cplx v0 = src[j], v1 = src[j + hh];
        if (k) {
          wb_cplx w1 = wb_fft_tw_s<FULL>(T, h, ts, k << shift);
          if (dir > 0) w1.y = -w1.y;
          v1 = wb_cmul(v1, w1);
        }
        const int j0 = ((j - k) << 1) + k;
        dst[j0] = wb_cadd(v0, v1);
        dst[j0 + ns] = wb_csub(v0, v1);
      }
      ls += 1;
    }
    WB_SYNC();
    wb_cplx* t = src;
    src = dst;
    dst = t;
  }
  return src;
}

// ---------------------------------------------------------------------------------------------------
// Real-input transforms through a half-size complex FFT.  A real sequence x[0..n) IS the complex array
// z[m] = (x[2m], x[2m+1]) of n/2 entries, so callers simply fill n doubles.  Buffers must hold n/2 + 1
// complex entries.  T/h: shared twiddle table with 2 h >= n.
//
// wb_rfft : in `a` (n doubles)            -> X[0..n/2]   (returns the buffer holding it)
// wb_irfft: in `a` (X[0..n/2], Hermitian) -> n doubles = n * irfft(X), i.e. sum_k X[k] e^{+2 pi i k m / n}
// ---------------------------------------------------------------------------------------------------
// A real input of n samples of which only the first len are non-zero has to be zero-filled up to here before
// wb_rfft(..., nz_real = len) (the pruned first pass never reads the rest): n/4, n/2 or n.
WB_HD int wb_rfft_fill(int n, int len) { return len <= (n >> 2) ? (n >> 2) : (len <= (n >> 1) ? (n >> 1) : n); }

template <int FULL = 0>
WB_DEV_NI wb_cplx* wb_rfft(wb_cplx* a, wb_cplx* b, int n, const wb_cplx* T, int h, int tid, int nthr,
                           int nz_real = 0x7ffffffe) {
  const int m = n >> 1;
  const int ts = wb_fft_log2(2 * h) - wb_fft_log2(n);
  wb_cplx* Z = wb_fft<FULL>(a, b, m, -1, T, h, tid, nthr, (nz_real + 1) >> 1);
  // X[k] = E + W^k O, X[m-k] = conj(E - W^k O), E = (Z[k] + conj(Z[m-k]))/2, O = -i (Z[k] - conj(Z[m-k]))/2
  for (int k = tid; k <= (m >> 1); k += nthr) {
    if (k == 0) {
      const wb_cplx z0 = Z[0];
      Z[0] = wb_mk(z0.x + z0.y, 0.0);
      Z[m] = wb_mk(z0.x - z0.y, 0.0);
    } else {
      const int kk = m - k;
      const wb_cplx zk = Z[k], zc = wb_conj(Z[kk]);
      const wb_cplx E = wb_mk(0.5 * (zk.x + zc.x), 0.5 * (zk.y + zc.y));
      const wb_cplx D = wb_mk(0.5 * (zk.x - zc.x), 0.5 * (zk.y - zc.y));
      const wb_cplx O = wb_mk(D.y, -D.x);  // -i D
      const wb_cplx WO = wb_cmul(wb_fft_tw_s<FULL>(T, h, ts, k), O);
      Z[k] = wb_cadd(E, WO);
      if (kk != k) Z[kk] = wb_conj(wb_csub(E, WO));
    }
  }
  WB_SYNC();
  return Z;
}

template <int FULL = 0>
WB_DEV_NI double* wb_irfft(wb_cplx* a, wb_cplx* b, int n, const wb_cplx* T, int h, int tid, int nthr) {
  const int m = n >> 1;
  const int ts = wb_fft_log2(2 * h) - wb_fft_log2(n);
  // Z[k] = A + i conj(W^k) Bd, Z[m-k] = conj(A) + i W^k conj(Bd), A = X[k] + conj(X[m-k]), Bd = X[k] - conj(X[m-k])
  for (int k = tid; k <= (m >> 1); k += nthr) {
    if (k == 0) {
      const double x0 = a[0].x, xm = a[m].x;
      a[0] = wb_mk(x0 + xm, x0 - xm);
    } else {
      const int kk = m - k;
      const wb_cplx xk = a[k], xc = wb_conj(a[kk]);
      const wb_cplx A = wb_cadd(xk, xc), Bd = wb_csub(xk, xc);
      const wb_cplx W = wb_fft_tw_s<FULL>(T, h, ts, k);
      const wb_cplx t1 = wb_cmul(wb_conj(W), Bd);  // conj(W^k) Bd
      a[k] = wb_mk(A.x - t1.y, A.y + t1.x);         // A + i t1
      if (kk != k) {
        const wb_cplx t2 = wb_cmul(W, wb_conj(Bd));
        a[kk] = wb_mk(A.x - t2.y, -A.y + t2.x);     // conj(A) + i t2
      }
    }
  }
  WB_SYNC();
  return (double*)wb_fft<FULL>(a, b, m, +1, T, h, tid, nthr);
}

// ---------------------------------------------------------------------------------------------------
// In-place forward complex FFT (radix-2 decimation in frequency): natural-order input, BIT-REVERSED
// output -- X[k] is found at index wb_bitrev(k, log2 n).  One buffer of n entries, for the one place
// (D4C's packed centroid transform) where two ping-pong buffers of n entries would not fit.
// ---------------------------------------------------------------------------------------------------
WB_HD int wb_bitrev(int k, int bits) {
#if !defined(WB_HOST_EMU) && defined(__CUDA_ARCH__)
  return (int)(__brev((unsigned)k) >> (32 - bits));
#endif
  int r = 0;
  for (int i = 0; i < bits; ++i) {
    r = (r << 1) | (k & 1);
    k >>= 1;
  }
  return r;
}

// one radix-4 DIF butterfly on (x0..x3) loaded from x[i0 + {0,1,2,3} q]; results go back in place
WB_DEV void wb_dif4_store(wb_cplx* x, int i0, int q, int pos, int shift, const wb_cplx* T, int h, int ts, wb_cplx x0,
                          wb_cplx x1, wb_cplx x2, wb_cplx x3) {
  const wb_cplx a0 = wb_cadd(x0, x2), a1 = wb_cadd(x1, x3);
  wb_cplx a2 = wb_csub(x0, x2);
  const wb_cplx d = wb_csub(x1, x3);
  wb_cplx a3 = wb_mk(d.y, -d.x);  // -i (x1 - x3)
  wb_cplx b1 = wb_csub(a0, a1);
  if (pos) {
    const wb_cplx w1 = wb_fft_tw_s(T, h, ts, pos << shift);
    const wb_cplx w2 = wb_cmul(w1, w1);
    a2 = wb_cmul(a2, w1);
    a3 = wb_cmul(a3, w1);
    b1 = wb_cmul(b1, w2);
    x[i0 + 3 * q] = wb_cmul(wb_csub(a2, a3), w2);
  } else {
    x[i0 + 3 * q] = wb_csub(a2, a3);
  }
  x[i0] = wb_cadd(a0, a1);
  x[i0 + q] = b1;
  x[i0 + 2 * q] = wb_cadd(a2, a3);
}

WB_DEV_NI void wb_fft_inplace_dif(wb_cplx* x, int n, const wb_cplx* T, int h, int tid, int nthr, int nz = 0x7fffffff) {
  const int ln = wb_fft_log2(n);
  const int ts = wb_fft_log2(2 * h) - ln;
  int lq = ln - 2;  // log2 of the quarter size of the current sub-transform
  if (lq >= 0 && nz <= (n >> 1)) {
    // Zero-padded input (entries at and above nz are zero and need not have been written): the first radix-4 step
    // with x2 = x3 = 0 (nz <= n/2) or x1 = x2 = x3 = 0 (nz <= n/4).
    const int q = n >> 2;
    const bool quarter = nz <= q;
    for (int t = tid; t < q; t += nthr) {
      const wb_cplx x0 = x[t];
      wb_cplx o0, o1, o2, o3;
      if (quarter) {
        o0 = o1 = o2 = o3 = x0;
      } else {
        const wb_cplx x1 = x[t + q];
        const wb_cplx r = wb_mk(x1.y, -x1.x);  // -i x1
        o0 = wb_cadd(x0, x1);
        o1 = wb_csub(x0, x1);
        o2 = wb_cadd(x0, r);
        o3 = wb_csub(x0, r);
      }
      if (t) {
        const wb_cplx w1 = wb_fft_tw_s(T, h, ts, t);
        const wb_cplx w2 = wb_cmul(w1, w1);
        o1 = wb_cmul(o1, w2);
        o2 = wb_cmul(o2, w1);
        o3 = wb_cmul(wb_cmul(o3, w1), w2);
      }
      x[t] = o0;
      x[t + q] = o1;
      x[t + 2 * q] = o2;
      x[t + 3 * q] = o3;
    }
    WB_SYNC();
    lq -= 2;
  }
  for (; lq >= 0; lq -= 2) {  // radix-4 step = two fused radix-2 DIF stages
    const int q = 1 << lq;
    const int shift = ln - lq - 2;  // W_{4q}^{pos} = table index pos << shift (of n)
    // two butterflies per trip, all eight loads issued before the first store (the butterflies of one pass touch
    // disjoint elements, which the compiler cannot know)
    for (int t = tid; t < (n >> 2); t += 2 * nthr) {
      const int pos = t & (q - 1);
      const int i0 = ((t - pos) << 2) + pos;
      const wb_cplx x0 = x[i0], x1 = x[i0 + q], x2 = x[i0 + 2 * q], x3 = x[i0 + 3 * q];
      const int t2 = t + nthr;
      if (t2 < (n >> 2)) {
        const int pos2 = t2 & (q - 1);
        const int j0 = ((t2 - pos2) << 2) + pos2;
        const wb_cplx y0 = x[j0], y1 = x[j0 + q], y2 = x[j0 + 2 * q], y3 = x[j0 + 3 * q];
        wb_dif4_store(x, i0, q, pos, shift, T, h, ts, x0, x1, x2, x3);
        wb_dif4_store(x, j0, q, pos2, shift, T, h, ts, y0, y1, y2, y3);
      } else {
        wb_dif4_store(x, i0, q, pos, shift, T, h, ts, x0, x1, x2, x3);
      }
    }
    WB_SYNC();
  }
  if (lq == -1) {  // one radix-2 stage left (odd log2 n): half size 1, no twiddle
    for (int t = tid; t < (n >> 1); t += nthr) {
      const wb_cplx a = x[2 * t], b = x[2 * t + 1];
      x[2 * t] = wb_cadd(a, b);
      x[2 * t + 1] = wb_csub(a, b);
    }
    WB_SYNC();
  }
}
