// Per-frame spectral building blocks shared by CheapTrick and D4C.
//
// Everything operates on one frame held in shared memory by one thread block.
// Spectra of real segments are kept as their first half [0, n/2] plus, where a
// running integral is needed, a short mirrored margin above n/2.
#pragma once
#include "wb_fft.h"

enum { WB_WIN_HANN = 1, WB_WIN_BLACKMAN = 2 };

struct wb_window_sums {
  double sw;  // sum seg*win
  double w;   // sum win
  double ww;  // sum win^2
};

// Pitch-synchronous window around time `pos` (reference cheaptrick.py:79-99 and
// d4c.py:92-110): half = int(span*fs/f0 + 0.5) samples either side of the 1-based
// centre int(pos*fs + 0.501) + 1, indices clamped to [1, ns].  Writes
// dst_sw[i] = seg*win and dst_w[i] = win for i < min(len, cap) and returns the three
// sums over the whole window.  `subsample` adds (pos*fs - int(pos*fs + 0.5))/fs to the
// window's time axis (D4C only).  Returns the window length through *len_out.
WB_DEV_NI wb_window_sums wb_pitch_window(const double* x, int ns, int fs, double f0, double pos, double span, int kind,
                                      bool subsample, double* dst_sw, double* dst_w, int cap, int* len_out,
                                      double* scratch, int tid, int nthr) {
  const int half = (int)(span * fs / f0 + 0.5);
  const int len = 2 * half + 1;
  const int centre = (int)(pos * fs + 0.501) + 1;
  const double shift = subsample ? (pos * fs - (double)(int)(pos * fs + 0.5)) / fs : 0.0;
  double s_sw = 0.0, s_w = 0.0, s_ww = 0.0;
  // window argument pi * t * f0 with t = k/(fs*span) + shift is linear in the sample index: each thread
  // evaluates one sincos and advances by a fixed rotation (its samples are nthr apart)
  const double dtheta = WB_PI * f0 / ((double)fs * span);
  const double theta0 = WB_PI * f0 * ((double)(tid - half) / ((double)fs * span) + shift);
  double cr, ci, qr, qi;
  sincos(theta0, &ci, &cr);
  sincos(dtheta * (double)nthr, &qi, &qr);
  for (int i = tid; i < len; i += nthr) {
    const int k = i - half;
    int idx = centre + k;
    idx = idx < 1 ? 1 : (idx > ns ? ns : idx);
    const double seg = WB_LDG(x + idx - 1);
    const double c1 = cr;
    const double nr = cr * qr - ci * qi;
    ci = cr * qi + ci * qr;
    cr = nr;
    double win;
    if (kind == WB_WIN_HANN) {
      win = 0.5 * c1 + 0.5;
    } else {  // 0.08 cos(2a) + 0.5 cos(a) + 0.42 with cos(2a) = 2 cos(a)^2 - 1
      win = 0.08 * (2.0 * c1 * c1 - 1.0) + 0.5 * c1 + 0.42;
    }
    const double sw = seg * win;
    s_sw += sw;
    s_w += win;
    s_ww += win * win;
    if (i < cap) {
      dst_sw[i] = sw;
      dst_w[i] = win;
    }
  }
  wb_block_sum3(s_sw, s_w, s_ww, scratch, tid, nthr);
  *len_out = len;
  wb_window_sums r;
  r.sw = s_sw;
  r.w = s_w;
  r.ww = s_ww;
  return r;
}

// The same window with every thread keeping its samples in registers (position i = tid + c * nthr, c < MAXPT) instead
// of staging them in shared memory: callers combine them with the block sums and write the result straight into
// the transform's input.  Requires len <= MAXPT * nthr (the caller checks and falls back to wb_pitch_window).
template <int MAXPT>
WB_DEV wb_window_sums wb_pitch_window_regs(const double* x, int ns, int fs, double f0, double pos, double span, int kind,
                                           bool subsample, double (&sw)[MAXPT], double (&w)[MAXPT], double* scratch,
                                           int tid, int nthr) {
  const int half = (int)(span * fs / f0 + 0.5);
  const int len = 2 * half + 1;
  const int centre = (int)(pos * fs + 0.501) + 1;
  const double shift = subsample ? (pos * fs - (double)(int)(pos * fs + 0.5)) / fs : 0.0;
  double s_sw = 0.0, s_w = 0.0, s_ww = 0.0;
  const double dtheta = WB_PI * f0 / ((double)fs * span);
  const double theta0 = WB_PI * f0 * ((double)(tid - half) / ((double)fs * span) + shift);
  double cr = 1.0, ci = 0.0, qr = 1.0, qi = 0.0;
  if (tid < len) sincos(theta0, &ci, &cr);
  if (len > nthr) sincos(dtheta * (double)nthr, &qi, &qr);  // only threads with a second sample rotate
#pragma unroll
  for (int c = 0; c < MAXPT; ++c) {
    const int i = tid + c * nthr;
    double a = 0.0, b = 0.0;
    if (i < len) {
      int idx = centre + i - half;
      idx = idx < 1 ? 1 : (idx > ns ? ns : idx);
      const double seg = WB_LDG(x + idx - 1);
      const double c1 = cr;
      const double nr = cr * qr - ci * qi;
      ci = cr * qi + ci * qr;
      cr = nr;
      b = kind == WB_WIN_HANN ? 0.5 * c1 + 0.5 : 0.08 * (2.0 * c1 * c1 - 1.0) + 0.5 * c1 + 0.42;
      a = seg * b;
      s_sw += a;
      s_w += b;
      s_ww += b * b;
    }
    sw[c] = a;
    w[c] = b;
  }
  wb_block_sum3(s_sw, s_w, s_ww, scratch, tid, nthr);
  wb_window_sums r;
  r.sw = s_sw;
  r.w = s_w;
  r.ww = s_ww;
  return r;
}

// Bin frequency exactly as the reference forms it: arange(n)/n*fs.
// n is a power of two, so k * (1/n) is exactly k / n
WB_DEV double wb_bin_hz(int k, int n, int fs) { return (double)k * (1.0 / n) * fs; }

// Low-frequency replica (cheaptrick.py:66-74, d4c.py:213-222): the part of the
// half spectrum p[0..n/2] below f0 receives the spectrum mirrored about f0,
// linearly interpolated between the knots f0 - f_k (k < nk, nk = number of bins
// below `limit`), extrapolated from the outermost pair when a query falls
// outside.  `tmp` needs as many doubles as there are bins below f0.
WB_DEV_NI void wb_mirror_low_band(double* p, int n, int fs, double f0, double limit, double* tmp, int tid, int nthr) {
  const double df = (double)fs / n;
  int nk = (int)(limit / df) + 2;
  if (nk > n) nk = n;
  while (nk > 0 && !(wb_bin_hz(nk - 1, n, fs) < limit)) --nk;
  int nq = (int)(f0 / df) + 2;
  if (nq > n / 2 + 1) nq = n / 2 + 1;
  while (nq > 0 && !(wb_bin_hz(nq - 1, n, fs) < f0)) --nq;
  if (nq <= 0 || nk < 2) return;
  for (int j = tid; j < nq; j += nthr) {
    const double u = wb_bin_hz(j, n, fs);
    // k1 = largest k with knot(k) >= u, clipped so that (k1, k1+1) is a valid pair
    int k1 = (int)floor((f0 - u) / df);
    if (k1 < 0) k1 = 0;
    if (k1 > nk - 2) k1 = nk - 2;
    while (k1 + 1 <= nk - 2 && (f0 - wb_bin_hz(k1 + 1, n, fs)) >= u) ++k1;
    while (k1 > 0 && (f0 - wb_bin_hz(k1, n, fs)) < u) --k1;
    const double xh = f0 - wb_bin_hz(k1, n, fs), xl = f0 - wb_bin_hz(k1 + 1, n, fs);
    const double yh = p[k1], yl = p[k1 + 1];
    tmp[j] = (yh - yl) / (xh - xl) * (u - xl) + yl;
  }
  WB_SYNC();
  for (int j = tid; j < nq; j += nthr) p[j] += tmp[j];
  WB_SYNC();
}

// Rectangular smoothing through a running integral (cheaptrick.py:103-131,
// d4c.py:179-188).  p[0..n/2] is the first half of a symmetric spectrum.  The
// reference integrates the doubled spectrum from -fs; by symmetry every value it
// reads equals, up to the constant total that cancels in the difference, a
// prefix sum S over bins [0, n/2 + margin].  `S` (>= n doubles) receives that
// prefix sum; out[k] = I(f_k + hw) - I(f_k - hw) for k in [0, n/2].
// out may alias p only if the caller no longer needs p.
// Functor form: `load(j)` gives the spectrum at bin j in [0, n/2] and `store(k, v)` takes the smoothed value of bin
// k, so that callers fold the element-wise steps before and after a smoothing into it (D4C chains three).
// The fill of S is fused into the scan's first phase: every thread integrates one contiguous chunk.
template <class Load, class Store>
WB_DEV_NI void wb_box_integral_f(Load load, int n, int fs, double hw, double* S, double* carry, Store store, int tid,
                                 int nthr) {
  const int nh = n / 2;
  const double df = (double)fs / n;
  int margin = (int)(hw / df) + 3;
  if (margin > nh - 1) margin = nh - 1;
  const int m = nh + margin + 1;  // S covers bins [0, m)
  const double p0 = load(0) * df;
  {  // inclusive prefix sum of load(mirror(i)) * df over [0, m): wb_block_scan with the fill folded in
    const int chunk = (m + nthr - 1) / nthr;
    const int lo = wb_imin(m, tid * chunk), hi = wb_imin(m, lo + chunk);
    double run = 0.0;
    for (int i = lo; i < hi; ++i) {
      run += load(i <= nh ? i : n - i) * df;
      S[i] = run;
    }
#ifdef WB_HOST_EMU
    (void)carry;
    WB_SYNC();
#else
    const int lane = tid & 31, w = tid >> 5, nw = (nthr + 31) >> 5;
    double v = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) carry[w] = v;
    __syncthreads();
    if (w == 0) {
      double c = lane < nw ? carry[lane] : 0.0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, c, o);
        if (lane >= o) c += t;
      }
      if (lane < nw) carry[lane] = c;
    }
    __syncthreads();
    const double off = (v - run) + (w > 0 ? carry[w - 1] : 0.0);
    if (off != 0.0)
      for (int i = lo; i < hi; ++i) S[i] += off;
    __syncthreads();
#endif
  }
  const double x0 = (0.0 / n * fs - fs) + df / 2.0;
  const double x1 = (1.0 / n * fs - fs) + df / 2.0;
  const double dx = x1 - x0;
  const double xlast = ((double)(2 * n - 1) / n * fs - fs) + df / 2.0;
  const double inv_dx = 1.0 / dx;  // the integral is continuous, so a last-bit difference in pos is harmless
  for (int k = tid; k <= nh; k += nthr) {
    const double fc = wb_bin_hz(k, n, fs);
    double v[2];
    for (int s = 0; s < 2; ++s) {
      double xi = s == 0 ? fc + hw : fc - hw;
      xi = wb_dmax(x0, wb_dmin(xlast, xi));
      const double pos = (xi - x0) * inv_dx;
      const double fb = floor(pos);
      const double frac = pos - fb;
      const int b = (int)fb;
      double y0, y1;
      {
        int j = b;
        if (j >= n) {
          int q = j - n;
          y0 = q < m ? S[q] : S[m - 1];
        } else {
          int q = n - 1 - j;
          y0 = -((q < m ? S[q] : S[m - 1]) - p0);
        }
        j = b + 1;
        if (j >= 2 * n) {
          y1 = y0;
        } else if (j >= n) {
          int q = j - n;
          y1 = q < m ? S[q] : S[m - 1];
        } else {
          int q = n - 1 - j;
          y1 = -((q < m ? S[q] : S[m - 1]) - p0);
        }
      }
      v[s] = y0 + (y1 - y0) * frac;
    }
    store(k, v[0] - v[1]);
  }
  WB_SYNC();
}

WB_DEV_NI void wb_box_integral(const double* p, int n, int fs, double hw, double* S, double* carry, double* out, int tid,
                            int nthr) {
  const int nh = n / 2;
  const double df = (double)fs / n;
  int margin = (int)(hw / df) + 3;
  if (margin > nh - 1) margin = nh - 1;
  const int m = nh + margin + 1;  // S covers bins [0, m)
  for (int i = tid; i < m; i += nthr) S[i] = (i <= nh ? p[i] : p[n - i]) * df;
  WB_SYNC();
  wb_block_scan(S, m, carry, tid, nthr);
  const double p0 = p[0] * df;
  // integral relative to the total over one period:
  //   index j >= n : S[j-n];   index j < n : -(S[n-1-j] - p0)
  const double x0 = (0.0 / n * fs - fs) + df / 2.0;
  const double x1 = (1.0 / n * fs - fs) + df / 2.0;
  const double dx = x1 - x0;
  const double xlast = ((double)(2 * n - 1) / n * fs - fs) + df / 2.0;
  const double inv_dx = 1.0 / dx;  // the integral is continuous, so a last-bit difference in pos is harmless
  for (int k = tid; k <= nh; k += nthr) {
    const double fc = wb_bin_hz(k, n, fs);
    double v[2];
    for (int s = 0; s < 2; ++s) {
      double xi = s == 0 ? fc + hw : fc - hw;
      xi = wb_dmax(x0, wb_dmin(xlast, xi));
      const double pos = (xi - x0) * inv_dx;
      const double fb = floor(pos);
      const double frac = pos - fb;
      const int b = (int)fb;
      double y0, y1;
      {
        int j = b;
        if (j >= n) {
          int q = j - n;
          y0 = q < m ? S[q] : S[m - 1];
        } else {
          int q = n - 1 - j;
          y0 = -((q < m ? S[q] : S[m - 1]) - p0);
        }
        j = b + 1;
        if (j >= 2 * n) {
          y1 = y0;
        } else if (j >= n) {
          int q = j - n;
          y1 = q < m ? S[q] : S[m - 1];
        } else {
          int q = n - 1 - j;
          y1 = -((q < m ? S[q] : S[m - 1]) - p0);
        }
      }
      v[s] = y0 + (y1 - y0) * frac;
    }
    out[k] = v[0] - v[1];
  }
  WB_SYNC();
}
