// Block-cooperative complex FFT in shared memory, float64.
//
// Every FFT of the analysis / synthesis path (256 .. 4096 complex points) runs through this header inside the fused
// per-frame kernels, so spectra never round-trip through HBM.  Two layers:
//   * wb_fft_fast<N, DIR> (second half of this file): the product path for N in 256 .. 2048 -- compile-time
//     Stockham passes, radix 8 (one radix-2 / radix-4 pass first), XOR-swizzled intermediates, twiddles w^k from
//     the first eighth of the circle with w^2k .. w^7k by multiplication; wb_fft_inplace_nat<N> runs the same
//     passes in one buffer with the registers as the staging area.
//   * wb_fft_generic: run-time sizes, radix 4 with one radix-2 pass when log2(n) is odd -- the fallback for the
//     shapes outside the fast path.
// Data ping-pongs between two shared buffers (no bit reversal).  Twiddles come from a per-block shared-memory
// table of one eighth of the circle (see below), filled once per block from the handle's global table; powers
// of w^k are formed by multiplication (table loads measured slower: a float64 multiply costs half a shared-memory
// wavefront).
#pragma once
#include "wb_platform.h"
#ifdef WB_HOST_EMU
#include <vector>
#endif

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
WB_DEV_NI wb_cplx* wb_fft_generic(wb_cplx* a, wb_cplx* b, int n, int dir, const wb_cplx* T, int h, int tid, int nthr,
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

// ===================================================================================================
// Fast path: compile-time size, radix-8 Stockham passes.
//
// N = R0 * 8^p with R0 in {1, 2, 4}: one twiddle-free pass of radix R0 (stride 1), then p radix-8 passes.
// Everything but the thread's butterfly index is a compile-time constant, so a pass is straight-line code:
// 8 loads of 16 bytes, one twiddle load (w^k; w^2k .. w^7k follow by multiplication -- with the radix-R0 pass
// first, every w^k lies in the first eighth of the circle, which is what the shared table holds, so there
// is no octant logic), 8 stores.  Data ping-pongs between the two buffers; intermediate results are stored
// XOR-swizzled (slot i ^ ((i >> 3) & 7)): with that, the stride-8 / stride-64 scatter of a radix-8 pass
// and the unit-stride gather of the next one both touch 8 distinct 16-byte bank groups per quarter-warp,
// without padding (checked for every pass shape by tools/fft_bank_check.py).  The first pass reads and the
// last pass writes natural order, so callers never see the swizzle.
// ===================================================================================================
WB_HD int wb_fft_swz(int i) { return i ^ ((i >> 3) & 7); }

template <int DIR>
WB_DEV wb_cplx wb_mul_i(wb_cplx v) {  // DIR < 0: -i v (forward), DIR > 0: +i v
  return DIR < 0 ? wb_mk(v.y, -v.x) : wb_mk(-v.y, v.x);
}

// DFT of R points in registers, natural order in and out (DIR < 0: e^{-2 pi i q m / R})
template <int R, int DIR>
WB_DEV void wb_dft_regs(wb_cplx (&v)[R]) {
  if (R == 2) {
    const wb_cplx a = v[0], b = v[1];
    v[0] = wb_cadd(a, b);
    v[1] = wb_csub(a, b);
  } else if (R == 4) {
    const wb_cplx t0 = wb_cadd(v[0], v[2]), t1 = wb_csub(v[0], v[2]), t2 = wb_cadd(v[1], v[3]);
    const wb_cplx t3 = wb_mul_i<DIR>(wb_csub(v[1], v[3]));
    v[0] = wb_cadd(t0, t2);
    v[1] = wb_cadd(t1, t3);
    v[2] = wb_csub(t0, t2);
    v[3] = wb_csub(t1, t3);
  } else {  // R == 8: one decimation-in-frequency step, then two 4-point transforms (even / odd outputs)
    const double r = 0.70710678118654752440;
    const wb_cplx s0 = wb_cadd(v[0], v[4 % R]), s1 = wb_cadd(v[1], v[5 % R]), s2 = wb_cadd(v[2], v[6 % R]),
                  s3 = wb_cadd(v[3], v[7 % R]);
    const wb_cplx d0 = wb_csub(v[0], v[4 % R]);
    wb_cplx d1 = wb_csub(v[1], v[5 % R]), d2 = wb_csub(v[2], v[6 % R]), d3 = wb_csub(v[3], v[7 % R]);
    // d_q *= W8^q, W8 = e^{-+ i pi / 4}
    d1 = DIR < 0 ? wb_mk((d1.x + d1.y) * r, (d1.y - d1.x) * r) : wb_mk((d1.x - d1.y) * r, (d1.x + d1.y) * r);
    d2 = wb_mul_i<DIR>(d2);
    d3 = DIR < 0 ? wb_mk((d3.y - d3.x) * r, -(d3.x + d3.y) * r) : wb_mk(-(d3.x + d3.y) * r, (d3.x - d3.y) * r);
    {
      const wb_cplx t0 = wb_cadd(s0, s2), t1 = wb_csub(s0, s2), t2 = wb_cadd(s1, s3);
      const wb_cplx t3 = wb_mul_i<DIR>(wb_csub(s1, s3));
      v[0] = wb_cadd(t0, t2);
      v[2] = wb_cadd(t1, t3);
      v[4 % R] = wb_csub(t0, t2);
      v[6 % R] = wb_csub(t1, t3);
    }
    {
      const wb_cplx t0 = wb_cadd(d0, d2), t1 = wb_csub(d0, d2), t2 = wb_cadd(d1, d3);
      const wb_cplx t3 = wb_mul_i<DIR>(wb_csub(d1, d3));
      v[1] = wb_cadd(t0, t2);
      v[3] = wb_cadd(t1, t3);
      v[5 % R] = wb_csub(t0, t2);
      v[7 % R] = wb_csub(t1, t3);
    }
  }
}

// One Stockham pass of radix R at stride NS over N points (N / R butterflies spread over the block).
// nzc: entries of src at and above it are zero and are not read (first pass only).
template <int N, int R, int NS, int DIR, bool SWZ_IN, bool SWZ_OUT>
WB_DEV void wb_fft_pass(const wb_cplx* src, wb_cplx* dst, const wb_cplx* T, int ts, int tid, int nthr, int nzc) {
  constexpr int TB = N / R;
  for (int j = tid; j < TB; j += nthr) {
    wb_cplx v[R];
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const int i = j + q * TB;
      if (NS == 1) v[q] = i < nzc ? src[SWZ_IN ? wb_fft_swz(i) : i] : wb_mk(0.0, 0.0);
      else v[q] = src[SWZ_IN ? wb_fft_swz(i) : i];
    }
    const int k = j & (NS - 1);
    if (NS > 1) {
      // w = e^{-+ 2 pi i k / (R NS)}: table index k * N / (R NS) of the N-point circle, below N / 8 for R = 8
      wb_cplx w1 = T[wb_fft_tw_skew((k * (N / (R * NS))) << ts)];
      if (DIR > 0) w1.y = -w1.y;
      v[1] = wb_cmul(v[1], w1);
      if (R >= 4) {
        const wb_cplx w2 = wb_mk(w1.x * w1.x - w1.y * w1.y, 2.0 * w1.x * w1.y);
        v[2 % R] = wb_cmul(v[2 % R], w2);
        const wb_cplx w3 = wb_cmul(w2, w1);
        v[3 % R] = wb_cmul(v[3 % R], w3);
        if (R >= 8) {
          const wb_cplx w4 = wb_mk(w2.x * w2.x - w2.y * w2.y, 2.0 * w2.x * w2.y);
          v[4 % R] = wb_cmul(v[4 % R], w4);
          v[5 % R] = wb_cmul(v[5 % R], wb_cmul(w4, w1));
          v[6 % R] = wb_cmul(v[6 % R], wb_mk(w3.x * w3.x - w3.y * w3.y, 2.0 * w3.x * w3.y));
          v[7 % R] = wb_cmul(v[7 % R], wb_cmul(w4, w3));
        }
      }
    }
    wb_dft_regs<R, DIR>(v);
    const int base = (j - k) * R + k;
#pragma unroll
    for (int m = 0; m < R; ++m) {
      const int o = base + m * NS;
      dst[SWZ_OUT ? wb_fft_swz(o) : o] = v[m];
    }
  }
  WB_SYNC();
}

// N = R0 * 8^P
template <int N>
struct wb_fft_plan {
  static constexpr int LN = N == 1 ? 0 : 1 + wb_fft_plan<(N > 1 ? N / 2 : 1)>::LN;
  static constexpr int R0 = 1 << (LN % 3);  // 1: no leading pass
  static constexpr int P = LN / 3;
};
template <>
struct wb_fft_plan<1> {
  static constexpr int LN = 0, R0 = 1, P = 0;
};

// the radix-8 passes I .. P-1 (stride R0 * 8^I), recursively so that every stride is a template constant
template <int N, int DIR, int I, int P, int NS, bool SWZ_LAST>
struct wb_fft_r8 {
  static WB_DEV wb_cplx* run(wb_cplx* src, wb_cplx* dst, const wb_cplx* T, int ts, int tid, int nthr, int nzc) {
    if (NS == 1) wb_fft_pass<N, 8, NS, DIR, false, (I + 1 < P) || SWZ_LAST>(src, dst, T, ts, tid, nthr, nzc);
    else wb_fft_pass<N, 8, NS, DIR, true, (I + 1 < P) || SWZ_LAST>(src, dst, T, ts, tid, nthr, nzc);
    return wb_fft_r8<N, DIR, I + 1, P, NS * 8, SWZ_LAST>::run(dst, src, T, ts, tid, nthr, nzc);
  }
};
template <int N, int DIR, int P, int NS, bool SWZ_LAST>
struct wb_fft_r8<N, DIR, P, P, NS, SWZ_LAST> {
  static WB_DEV wb_cplx* run(wb_cplx* src, wb_cplx*, const wb_cplx*, int, int, int, int) { return src; }
};

// Complex FFT of compile-time size N >= 8 (same contract as wb_fft_generic).  SWZ_LAST: leave the result
// swizzled too (entry i at slot wb_fft_swz(i)), for consumers whose threads each read a run of consecutive
// entries -- a lane stride of 8 entries that the swizzle spreads over all banks.
#if defined(WB_FFT_NOINLINE) && !defined(WB_HOST_EMU)
#define WB_FFT_FN __device__ __noinline__  // tuning variant: one copy of each transform per kernel
#else
#define WB_FFT_FN WB_DEV_NI
#endif
template <int N, int DIR, bool SWZ_LAST = false>
WB_FFT_FN wb_cplx* wb_fft_fast(wb_cplx* a, wb_cplx* b, const wb_cplx* T, int h, int tid, int nthr, int nzc) {
  typedef wb_fft_plan<N> PL;
  const int ts = wb_fft_log2(2 * h) - PL::LN;
  if (PL::R0 > 1) {
    wb_fft_pass<N, (PL::R0 > 1 ? PL::R0 : 2), 1, DIR, false, (PL::P > 0) || SWZ_LAST>(a, b, T, ts, tid, nthr, nzc);
    return wb_fft_r8<N, DIR, 0, PL::P, (PL::R0 > 1 ? PL::R0 : 8), SWZ_LAST>::run(b, a, T, ts, tid, nthr, nzc);
  }
  return wb_fft_r8<N, DIR, 0, PL::P, 1, SWZ_LAST>::run(a, b, T, ts, tid, nthr, nzc);
}
// element m of a real sequence held as n/2 swizzled complex entries
WB_DEV double wb_fft_swz_real(const double* x, int m) { return x[2 * wb_fft_swz(m >> 1) + (m & 1)]; }

// dir = -1: forward (e^{-i...}), dir = +1: inverse WITHOUT the 1/n factor.
// Input in `a`; returns the buffer (a or b) that holds the result.  All threads of the block must call it;
// it ends with a barrier.  T/h: shared twiddle table as above.  nz: entries of `a` at and above it are zero and need
// not have been written.  NC: the size when it is known at compile time (0: dispatch on n).
template <int FULL = 0, int NC = 0>
WB_DEV_NI wb_cplx* wb_fft(wb_cplx* a, wb_cplx* b, int n, int dir, const wb_cplx* T, int h, int tid, int nthr,
                         int nz = 0x7fffffff) {
#ifndef WB_FFT_GENERIC_ONLY
#define WB_FFT_CASE(SZ)                                                                                \
  if ((NC == SZ) || (NC == 0 && n == SZ))                                                              \
    return dir < 0 ? wb_fft_fast<SZ, -1>(a, b, T, h, tid, nthr, nz) : wb_fft_fast<SZ, +1>(a, b, T, h, tid, nthr, nz);
  WB_FFT_CASE(1024)
  WB_FFT_CASE(512)
  WB_FFT_CASE(2048)
  WB_FFT_CASE(256)
#undef WB_FFT_CASE
#endif
  return wb_fft_generic<FULL>(a, b, n, dir, T, h, tid, nthr, nz);
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

// The split / merge step pairs bin k with n/2 - k.  One table entry serves two such pairs: W^(n/4 - k) =
// -i conj(W^k), so a thread that takes k <= n/8 also takes n/4 - k, and every twiddle comes from the first
// eighth of the circle (no octant logic).
WB_DEV void wb_rfft_split(wb_cplx* Z, int k, int kk, wb_cplx W) {
  // X[k] = E + W^k O, X[m-k] = conj(E - W^k O), E = (Z[k] + conj(Z[m-k]))/2, O = -i (Z[k] - conj(Z[m-k]))/2
  const wb_cplx zk = Z[k], zc = wb_conj(Z[kk]);
  const wb_cplx E = wb_mk(0.5 * (zk.x + zc.x), 0.5 * (zk.y + zc.y));
  const wb_cplx D = wb_mk(0.5 * (zk.x - zc.x), 0.5 * (zk.y - zc.y));
  const wb_cplx WO = wb_cmul(W, wb_mk(D.y, -D.x));  // W (-i D)
  Z[k] = wb_cadd(E, WO);
  Z[kk] = wb_conj(wb_csub(E, WO));
}
WB_DEV void wb_irfft_merge(wb_cplx* a, int k, int kk, wb_cplx W) {
  // Z[k] = A + i conj(W^k) Bd, Z[m-k] = conj(A) + i W^k conj(Bd), A = X[k] + conj(X[m-k]), Bd = X[k] - conj(X[m-k])
  const wb_cplx xk = a[k], xc = wb_conj(a[kk]);
  const wb_cplx A = wb_cadd(xk, xc), Bd = wb_csub(xk, xc);
  const wb_cplx t1 = wb_cmul(wb_conj(W), Bd);
  const wb_cplx t2 = wb_cmul(W, wb_conj(Bd));
  a[k] = wb_mk(A.x - t1.y, A.y + t1.x);
  a[kk] = wb_mk(A.x - t2.y, -A.y + t2.x);
}

// NC: n when it is known at compile time (0: run-time dispatch)
template <int FULL = 0, int NC = 0>
WB_DEV_NI wb_cplx* wb_rfft(wb_cplx* a, wb_cplx* b, int n, const wb_cplx* T, int h, int tid, int nthr,
                           int nz_real = 0x7ffffffe) {
  if (NC) n = NC;
  const int m = n >> 1;
  const int ts = wb_fft_log2(2 * h) - wb_fft_log2(n);
  wb_cplx* Z = wb_fft<FULL, NC / 2>(a, b, m, -1, T, h, tid, nthr, (nz_real + 1) >> 1);
  const int q = m >> 2;  // n / 8
  for (int k = tid; k <= q; k += nthr) {
    if (k == 0) {
      const wb_cplx z0 = Z[0], zm = Z[m >> 1];
      Z[0] = wb_mk(z0.x + z0.y, 0.0);
      Z[m] = wb_mk(z0.x - z0.y, 0.0);
      Z[m >> 1] = wb_conj(zm);  // W^(n/4) = -i
    } else {
      const wb_cplx W = T[wb_fft_tw_skew(k << ts)];
      wb_rfft_split(Z, k, m - k, W);
      if (k != q) wb_rfft_split(Z, (m >> 1) - k, (m >> 1) + k, wb_mk(-W.y, -W.x));
    }
  }
  WB_SYNC();
  return Z;
}

template <int FULL = 0, int NC = 0>
WB_DEV_NI double* wb_irfft(wb_cplx* a, wb_cplx* b, int n, const wb_cplx* T, int h, int tid, int nthr) {
  if (NC) n = NC;
  const int m = n >> 1;
  const int ts = wb_fft_log2(2 * h) - wb_fft_log2(n);
  const int q = m >> 2;
  for (int k = tid; k <= q; k += nthr) {
    if (k == 0) {
      const double x0 = a[0].x, xm = a[m].x;
      const wb_cplx c = a[m >> 1];
      a[0] = wb_mk(x0 + xm, x0 - xm);
      a[m >> 1] = wb_mk(2.0 * c.x, -2.0 * c.y);
    } else {
      const wb_cplx W = T[wb_fft_tw_skew(k << ts)];
      wb_irfft_merge(a, k, m - k, W);
      if (k != q) wb_irfft_merge(a, (m >> 1) - k, (m >> 1) + k, wb_mk(-W.y, -W.x));
    }
  }
  WB_SYNC();
  return (double*)wb_fft<FULL, NC / 2>(a, b, m, +1, T, h, tid, nthr);
}

// ---------------------------------------------------------------------------------------------------
// In-place forward complex FFT with NATURAL-order output, for blocks of at least N/8 threads: the same
// passes as wb_fft_fast, but every thread keeps the 8 points of its butterflies in registers across a
// barrier between the gather and the scatter of a pass, so a single buffer of N entries suffices (D4C's
// packed centroid transform, where two buffers of N entries would not fit).  Two barriers per pass.
// ---------------------------------------------------------------------------------------------------
template <int N, int R, int NS, bool SWZ_IN>
WB_DEV void wb_fft_inplace_gather(const wb_cplx* x, wb_cplx (&v)[8 / R][R], int tid, int nzc) {
#pragma unroll
  for (int c = 0; c < 8 / R; ++c) {
    const int j = tid + c * (N / 8);
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const int i = j + q * (N / R);
      if (NS == 1) v[c][q] = i < nzc ? x[SWZ_IN ? wb_fft_swz(i) : i] : wb_mk(0.0, 0.0);
      else v[c][q] = x[SWZ_IN ? wb_fft_swz(i) : i];
    }
  }
}
template <int N, int R, int NS, bool SWZ_OUT>
WB_DEV void wb_fft_inplace_scatter(wb_cplx* x, wb_cplx (&v)[8 / R][R], const wb_cplx* T, int ts, int tid) {
#pragma unroll
  for (int c = 0; c < 8 / R; ++c) {
    const int j = tid + c * (N / 8);
    const int k = j & (NS - 1);
    if (NS > 1) {  // only radix-8 passes carry twiddles (the radix-2 / radix-4 pass comes first)
      const wb_cplx w1 = T[wb_fft_tw_skew((k * (N / (R * NS))) << ts)];
      v[c][1] = wb_cmul(v[c][1], w1);
      const wb_cplx w2 = wb_mk(w1.x * w1.x - w1.y * w1.y, 2.0 * w1.x * w1.y);
      v[c][2 % R] = wb_cmul(v[c][2 % R], w2);
      const wb_cplx w3 = wb_cmul(w2, w1);
      v[c][3 % R] = wb_cmul(v[c][3 % R], w3);
      const wb_cplx w4 = wb_mk(w2.x * w2.x - w2.y * w2.y, 2.0 * w2.x * w2.y);
      v[c][4 % R] = wb_cmul(v[c][4 % R], w4);
      v[c][5 % R] = wb_cmul(v[c][5 % R], wb_cmul(w4, w1));
      v[c][6 % R] = wb_cmul(v[c][6 % R], wb_mk(w3.x * w3.x - w3.y * w3.y, 2.0 * w3.x * w3.y));
      v[c][7 % R] = wb_cmul(v[c][7 % R], wb_cmul(w4, w3));
    }
    wb_dft_regs<R, -1>(v[c]);
    const int base = (j - k) * R + k;
#pragma unroll
    for (int m = 0; m < R; ++m) {
      const int o = base + m * NS;
      x[SWZ_OUT ? wb_fft_swz(o) : o] = v[c][m];
    }
  }
}
template <int N, int R, int NS, bool SWZ_IN, bool SWZ_OUT>
WB_DEV void wb_fft_pass_inplace(wb_cplx* x, const wb_cplx* T, int ts, int tid, int nzc) {
  constexpr int NT = N / 8;  // threads that take part
#ifdef WB_HOST_EMU
  // test-only emulation of the N/8 threads: all gathers, then all scatters
  static std::vector<wb_cplx> regs;
  regs.resize((size_t)NT * 8);
  typedef wb_cplx (*arr_t)[8 / R][R];
  for (int t = 0; t < NT; ++t) wb_fft_inplace_gather<N, R, NS, SWZ_IN>(x, *(arr_t)(regs.data() + (size_t)t * 8), t, nzc);
  for (int t = 0; t < NT; ++t) wb_fft_inplace_scatter<N, R, NS, SWZ_OUT>(x, *(arr_t)(regs.data() + (size_t)t * 8), T, ts, t);
  (void)tid;
#else
  wb_cplx v[8 / R][R];
  if (tid < NT) wb_fft_inplace_gather<N, R, NS, SWZ_IN>(x, v, tid, nzc);
  __syncthreads();
  if (tid < NT) wb_fft_inplace_scatter<N, R, NS, SWZ_OUT>(x, v, T, ts, tid);
  __syncthreads();
#endif
}
template <int N, int I, int P, int NS>
struct wb_fft_r8_inplace {
  static WB_DEV void run(wb_cplx* x, const wb_cplx* T, int ts, int tid, int nzc) {
    if (NS == 1) wb_fft_pass_inplace<N, 8, NS, false, (I + 1 < P)>(x, T, ts, tid, nzc);
    else wb_fft_pass_inplace<N, 8, NS, true, (I + 1 < P)>(x, T, ts, tid, nzc);
    wb_fft_r8_inplace<N, I + 1, P, NS * 8>::run(x, T, ts, tid, nzc);
  }
};
template <int N, int P, int NS>
struct wb_fft_r8_inplace<N, P, P, NS> {
  static WB_DEV void run(wb_cplx*, const wb_cplx*, int, int, int) {}
};

// All threads of the block call it (blockDim >= N / 8; the host emulation plays all of them).
template <int N>
WB_DEV_NI void wb_fft_inplace_nat(wb_cplx* x, const wb_cplx* T, int h, int tid, int nthr, int nzc) {
  typedef wb_fft_plan<N> PL;
  const int ts = wb_fft_log2(2 * h) - PL::LN;
  (void)nthr;
  if (PL::R0 > 1) {
    wb_fft_pass_inplace<N, (PL::R0 > 1 ? PL::R0 : 2), 1, false, (PL::P > 0)>(x, T, ts, tid, nzc);
    wb_fft_r8_inplace<N, 0, PL::P, (PL::R0 > 1 ? PL::R0 : 8)>::run(x, T, ts, tid, nzc);
  } else {
    wb_fft_r8_inplace<N, 0, PL::P, 1>::run(x, T, ts, tid, nzc);
  }
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
