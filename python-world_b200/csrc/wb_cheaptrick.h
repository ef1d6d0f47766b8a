// CheapTrick spectral envelope: one thread block per (utterance, frame).
//
// Replaces the per-frame Python loop of the reference (cheaptrick.py:9-39 driver,
// :43-60 estimate_one_slice).  The whole slice -- pitch-synchronous gather, window,
// FFT, power, low-band replica, box smoothing, log, cepstral lifter, inverse FFT,
// exp -- stays in shared memory; HBM sees the waveform once (L1/L2-cached gather)
// and one write of the envelope (plus the optional complex "ps spectrogram").
#pragma once
#include "wb_spectral.h"

struct wb_cheaptrick_params {
  // inputs
  const double* x;         // [B, x_stride]
  const int* n_samples;    // [B]
  const double* tpos;      // [B, f_stride] seconds
  const double* f0;        // [B, f_stride]
  const double* vuv;       // [B, f_stride]
  const int* n_frames;     // [B]
  const double* dither;    // [B, f_stride, n/2+1] or nullptr (hash dither)
  const wb_cplx* tw;
  int tw_n;
  int x_stride, f_stride, fs, n;
  double q1;
  unsigned long long seed;
  // outputs
  double* f0_used;  // [B, f_stride]  what CheapTrick leaves in source['f0'] (cheaptrick.py:27,33)
  double* spec;     // [B, f_stride, n/2+1]
  wb_cplx* ps;      // [B, f_stride, n] or nullptr

  // two complex buffers of n/2+1 (= n+2 doubles each), S (n doubles), W (n doubles: window values, later
  // power + temp), carry, scratch, twiddles (n/2 complex)
  static size_t smem_bytes(int n, int nthr) {
    return (size_t)(n / 2 + 1) * 2 * sizeof(wb_cplx) + ((size_t)2 * n + 2 + nthr + 2 + WB_REDUCE_SCRATCH + 64) * sizeof(double) +
           (size_t)WB_FFT_TW_SLOTS(n / 2) * sizeof(wb_cplx);
  }

};

// NC / NT: FFT size and block size when the launcher knows them at compile time (0: run-time values)
template <int NC = 0, int NT = 0>
struct wb_cheaptrick_body_t : wb_cheaptrick_params {
  WB_DEV void operator()(int block, int tid, int nthr_rt, double* smem) const {
    const int nthr = NT ? NT : nthr_rt;
    const int n = NC ? NC : this->n;
    const int u = block / f_stride, f = block - u * f_stride;
    if (f >= n_frames[u]) return;
    const int nh = n / 2;
    wb_cplx* A = (wb_cplx*)smem;          // nh + 1 complex
    wb_cplx* B = A + (nh + 1);
    double* S = (double*)(B + (nh + 1));  // n doubles
    double* Wv = S + n;                   // n + 2 doubles
    double* carry = Wv + n + 2;           // nthr + 2
    double* scratch = carry + nthr + 2;   // WB_REDUCE_SCRATCH
    wb_cplx* twS = (wb_cplx*)(scratch + WB_REDUCE_SCRATCH + ((nthr + WB_REDUCE_SCRATCH) & 1));
    const int twH = nh;
    wb_fft_load_twiddles(twS, twH, tw, tw_n, tid, nthr);
    const size_t fi = (size_t)u * f_stride + f;
    const double* xu = x + (size_t)u * x_stride;
    const int ns = n_samples[u];

    // cheaptrick.py:24-33
    const double limit = fs * 3.0 / (n - 3.0);
    double f0e = (vuv[fi] == 0.0) ? 500.0 : f0[fi];
    if (f0e < limit) f0e = 500.0;
    if (tid == 0) f0_used[fi] = f0e;

    // step 1 (cheaptrick.py:79-99): window, unit energy, weighted-mean removal
    double* Ad = (double*)A;
    int cap;
    const int wlen = 2 * (int)(1.5 * fs / f0e + 0.5) + 1;
    if (wlen <= 8 * nthr && wlen <= n) {
      // window samples in registers (<= 8 per thread), written straight into the transform input
      double sw[8], w[8];
      const wb_window_sums ws = wb_pitch_window_regs<8>(xu, ns, fs, f0e, tpos[fi], 1.5, WB_WIN_HANN, false, sw, w, scratch, tid, nthr);
      const double inv_norm = 1.0 / sqrt(ws.ww);
      const double ratio = ws.sw / ws.w;
      cap = wlen;
      // the window covers 3 pitch periods of an n-sample buffer: the transform's first pass skips the zero padding
      const int nfill = wb_rfft_fill(n, cap);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int i = tid + c * nthr;
        if (i < nfill) Ad[i] = (sw[c] - w[c] * ratio) * inv_norm;  // zero beyond the window
      }
      for (int i = 8 * nthr + tid; i < nfill; i += nthr) Ad[i] = 0.0;
    } else {
      int len;
      wb_window_sums ws = wb_pitch_window(xu, ns, fs, f0e, tpos[fi], 1.5, WB_WIN_HANN, false, S, Wv, n, &len, scratch, tid, nthr);
      const double inv_norm = 1.0 / sqrt(ws.ww);
      const double ratio = ws.sw / ws.w;
      cap = len < n ? len : n;
      const int nfill = wb_rfft_fill(n, cap);
      for (int i = tid; i < nfill; i += nthr) Ad[i] = i < cap ? (S[i] - Wv[i] * ratio) * inv_norm : 0.0;
    }
    WB_SYNC();
    wb_cplx* X = wb_rfft<0, NC>(A, B, n, twS, twH, tid, nthr, cap);
    wb_cplx* Y = (X == A) ? B : A;  // the free buffer
    if (ps) {  // 'ps spectrogram': the full-length spectrum of the real segment
      wb_cplx* o = ps + fi * (size_t)n;
      for (int k = tid; k < n; k += nthr) o[k] = k <= nh ? X[k] : wb_conj(X[n - k]);
    }
    // power of the first half (cheaptrick.py:66)
    double* P = Wv;                 // nh + 1
    double* T = (double*)Y;         // temp, nh + 1
    for (int k = tid; k <= nh; k += nthr) P[k] = X[k].x * X[k].x + X[k].y * X[k].y;
    WB_SYNC();
    wb_mirror_low_band(P, n, fs, f0e, f0e + (double)fs / n, T, tid, nthr);

    // step 2 (cheaptrick.py:103-118)
    const double* dz = dither ? dither + fi * (size_t)(nh + 1) : nullptr;
    double* Ld = (double*)X;  // the spectrum is no longer needed: log spectrum as a real even sequence
    const double scale = 1.5 / f0e;  // T * 1.5 / f0 with one division per frame (last-bit difference under the log)
    // the smoothing's store side takes the scaling, the eps-dither and the logarithm (cheaptrick.py:113-118)
    wb_box_integral_f([&](int j) { return P[j]; }, n, fs, f0e / 3.0, S, carry, [&](int k, double v) {
      double d;
      if (dz) {
        d = dz[k];
      } else {  // deterministic stand-in for |rand()|*eps (cheaptrick.py:117)
        unsigned long long h = seed ^ ((unsigned long long)fi * 0x9E3779B97F4A7C15ull + (unsigned long long)k);
        h ^= h >> 33;
        h *= 0xff51afd7ed558ccdull;
        h ^= h >> 33;
        h *= 0xc4ceb9fe1a85ec53ull;
        h ^= h >> 33;
        d = ((double)(h >> 11) + 1.0) * (1.0 / 9007199254740992.0) * WB_EPS;  // in (0, eps]
      }
      T[k] = log(v * scale + d);
    }, tid, nthr);
    for (int i = tid; i < n; i += nthr) Ld[i] = T[i <= nh ? i : n - i];
    WB_SYNC();

    // step 3 (cheaptrick.py:136-157): lifter in the quefrency domain
    wb_cplx* Cq = wb_rfft<0, NC>(X, Y, n, twS, twH, tid, nthr);
    wb_cplx* Cf = (Cq == X) ? Y : X;
    {
      // sinc(pi f0 q) ((1 - 2 q1) + 2 q1 cos(2 pi f0 q)) at q = k / fs; cos(2a) = 1 - 2 sin(a)^2, and a advances by a
      // fixed step per trip, so the thread evaluates two sincos and then rotates
      double sn, cs, dsn, dcs;
      sincos(WB_PI * f0e * ((double)tid / fs), &sn, &cs);
      sincos(WB_PI * f0e * ((double)nthr / fs), &dsn, &dcs);
      for (int k = tid; k <= nh; k += nthr) {
        const double q = (double)k / fs;
        double lift = 1.0;
        if (k > 0) lift = sn / (WB_PI * f0e * q);
        lift *= (1.0 - 2.0 * q1) + 2.0 * q1 * (1.0 - 2.0 * sn * sn);
        const double ns_ = sn * dcs + cs * dsn;
        cs = cs * dcs - sn * dsn;
        sn = ns_;
        const wb_cplx c = Cq[k];
        Cq[k] = wb_mk(c.x * lift, (k == 0 || k == nh) ? 0.0 : c.y * lift);
      }
    }
    WB_SYNC();
    const double* E = wb_irfft<0, NC>(Cq, Cf, n, twS, twH, tid, nthr);
    double* o = spec + fi * (size_t)(nh + 1);
    const double inv_n = 1.0 / n;
    for (int k = tid; k <= nh; k += nthr) o[k] = exp(E[k] * inv_n);
  }
};
typedef wb_cheaptrick_body_t<> wb_cheaptrick_body;
