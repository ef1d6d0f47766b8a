// Waveform synthesis -- kernel bodies.  Replaces world/synthesis.py:21-250 (pulse-by-pulse minimum-phase
// overlap-add) and world/synthesisRequiem.py:12-141 (excitation + per-frame minimum-phase filtering).
//
//   Y1 sy_timebase   per utterance: sample-rate F0/VUV, phase accumulation, pulse list       (synthesis.py:120-140)
//   Y2 sy_pulses     per pulse: spectral slices, two minimum-phase responses, noise convolution,
//                    overlap-add with the reference's clamped-index semantics                 (synthesis.py:61-116, 144-250)
//   R1 rq_excite     per utterance: band aperiodicity at sample rate, cyclic noise seeds, pulse seeds (synthesisRequiem.py:27-71,120-141)
//   R2 rq_frames     per frame: Hann-windowed excitation x minimum-phase envelope, overlap-add  (synthesisRequiem.py:74-101)
//   Y3 sy_normalise  per utterance: divide by the peak when it exceeds 1                       (main.py:209-212)
//
// "y[idx] += v" with clamped, hence duplicated, indices is last-write-wins in NumPy: at the low end only the
// element that lands exactly on sample 1 survives, at the high end only the final element.
#pragma once
#include "wb_fft.h"

// Decision-critical interpolation must not be contracted into FMAs (scipy rounds the product, then the sum).
#if defined(WB_HOST_EMU) || !defined(__CUDA_ARCH__)
#define WB_MUL(a, b) ((a) * (b))
#define WB_ADD(a, b) ((a) + (b))
#else
#define WB_MUL(a, b) __dmul_rn((a), (b))
#define WB_ADD(a, b) __dadd_rn((a), (b))
#endif

struct wb_sy_plan {
  int batch, fs, n, f_stride, y_stride, n_bins;  // n = fft size, n_bins = n/2+1
  // inputs [B, f_stride(, bins)]
  const double* tpos;
  const double* f0;
  const double* vuv;
  const double* spec;
  const double* ap;  // synthesis: [B, F, n_bins] linear; requiem: [B, F, n_ap] dB
  const int* n_frames;
  int n_ap;          // requiem: rows of the band aperiodicity (bands + 2)
  // workspace
  double* wrap;          // [B, y_stride] wrapped phase
  unsigned char* vuv_i;  // [B, y_stride]
  double* p_loc;         // [B, p_cap] pulse time
  int* p_idx;            // [B, p_cap] 1-based sample index
  double* p_shift;       // [B, p_cap] fractional delay (s)
  int* p_noise_off;      // [B, p_cap] offset of the pulse's noise run
  int p_cap;
  int* n_pulses;         // [B]
  int* noise_total;      // [B]
  int* out_len;          // [B]
  int* pulse_base;       // [B + 1] prefix of n_pulses over the batch
  double* ap_i;          // requiem: [B, n_ap, y_stride]
  double* exc;           // requiem: [B, y_stride]
  // outputs
  double* y;             // [B, y_stride]
};

WB_HD int wb_sy_length(double t0, double t_end, int fs) {
  const double step = 1.0 / fs;
  const double len = ceil(((t_end + step) - t0) / step);
  return len > 0.0 ? (int)len : 0;
}

// scipy interp1d(kind='linear', fill_value='extrapolate') on sorted knots; `hi` is the bracketing index
WB_DEV int wb_sy_bracket(const double* xk, int n, double x, int guess) {
  int hi = guess < 1 ? 1 : (guess > n - 1 ? n - 1 : guess);
  while (hi < n - 1 && xk[hi] < x) ++hi;      // searchsorted left: first knot >= x
  while (hi > 1 && xk[hi - 1] >= x) --hi;
  return hi;
}
WB_DEV double wb_sy_lerp(const double* xk, const double* yk, int hi, double x) {
  const double slope = (yk[hi] - yk[hi - 1]) / (xk[hi] - xk[hi - 1]);
  return WB_ADD(WB_MUL(slope, x - xk[hi - 1]), yk[hi - 1]);
}

// element k of numpy.arange(t0, stop, step): the first two are start and start + step, the rest
// start + k * (second - first)
WB_DEV double wb_sy_time(double t0, double step, double delta, int k) {
  if (k == 0) return t0;
  if (k == 1) return t0 + step;
  return WB_ADD(t0, WB_MUL((double)k, delta));
}

// exclusive scan of one value per thread (once per utterance: a serial pass by thread 0 is fine)
WB_DEV double wb_exscan_d(double v, double* scratch, int tid, int nthr, double* total) {
  scratch[tid] = v;
  WB_SYNC();
  if (tid == 0) {
    double a = 0.0;
    for (int t = 0; t < nthr; ++t) {
      const double q = scratch[t];
      scratch[t] = a;
      a += q;
    }
    scratch[nthr] = a;
  }
  WB_SYNC();
  const double r = scratch[tid];
  *total = scratch[nthr];
  WB_SYNC();
  return r;
}

// ------------------------------------------------------------------------------------ Y1
struct wb_sy_timebase {
  wb_sy_plan p;
  static size_t smem_bytes(int nthr) { return (size_t)(nthr + 8) * sizeof(double); }

  WB_DEV void operator()(int block, int tid, int nthr, double* smem) const {
    const int u = block;
    const int F = p.n_frames[u];
    const double* tp = p.tpos + (size_t)u * p.f_stride;
    const double* f0 = p.f0 + (size_t)u * p.f_stride;
    const double* vv = p.vuv + (size_t)u * p.f_stride;
    double* wrap = p.wrap + (size_t)u * p.y_stride;
    unsigned char* vi = p.vuv_i + (size_t)u * p.y_stride;
    if (F < 2) {
      if (tid == 0) {
        p.out_len[u] = 0;
        p.n_pulses[u] = 0;
        p.noise_total[u] = 0;
      }
      return;
    }
    const double t0 = tp[0];
    int L = wb_sy_length(t0, tp[F - 1], p.fs);
    if (L > p.y_stride) L = p.y_stride;
    const double step = 1.0 / p.fs;
    const double delta = (t0 + step) - t0;  // numpy.arange fills start + i * (second - first)
    const double two_pi = 2.0 * WB_PI;
    const int chunk = (L + nthr - 1) / nthr;
    const int lo = wb_imin(L, tid * chunk), hi = wb_imin(L, lo + chunk);
    // pass 1: phase increments, chunk-local running sums
    double run = 0.0;
    int br = 1;
    for (int k = lo; k < hi; ++k) {
      const double t = wb_sy_time(t0, step, delta, k);
      br = wb_sy_bracket(tp, F, t, br);
      const double fi = wb_sy_lerp(tp, f0, br, t);
      const bool voiced = wb_sy_lerp(tp, vv, br, t) > 0.5;
      double f = voiced ? fi : 0.0;  // f0_interpolated_raw * vuv_interpolated
      if (f == 0.0) f = f + 500.0;
      run += two_pi * f / p.fs;
      wrap[k] = run;
      vi[k] = voiced ? 1 : 0;
    }
    double total;
    const double off = wb_exscan_d(run, smem, tid, nthr, &total);
    for (int k = lo; k < hi; ++k) wrap[k] = fmod(wrap[k] + off, two_pi);
    WB_SYNC();
    // pass 2: pulses where the wrapped phase drops
    int cnt = 0;
    for (int k = lo; k < hi; ++k)
      if (k + 1 < L && fabs(wrap[k + 1] - wrap[k]) > WB_PI) ++cnt;
    double tot_d;
    int at = (int)wb_exscan_d((double)cnt, smem, tid, nthr, &tot_d);
    int n_p = (int)tot_d;
    if (n_p > p.p_cap) n_p = p.p_cap;
    double* loc = p.p_loc + (size_t)u * p.p_cap;
    int* idx = p.p_idx + (size_t)u * p.p_cap;
    double* shf = p.p_shift + (size_t)u * p.p_cap;
    for (int k = lo; k < hi; ++k) {
      if (k + 1 < L && fabs(wrap[k + 1] - wrap[k]) > WB_PI) {
        if (at < p.p_cap) {
          const double t = wb_sy_time(t0, step, delta, k);
          loc[at] = t;
          const double v = t * p.fs;  // Decimal(...).quantize(0, ROUND_HALF_UP) on the exact value, + 1
          const double fl = floor(v);
          int id = (int)fl + ((v - fl) >= 0.5 ? 1 : 0) + 1;
          if (id < 1) id = 1;
          if (id > L - 1) id = L - 1;
          idx[at] = id;
          const double y1 = wrap[id - 1] - two_pi, y2 = wrap[id];
          shf[at] = (-y1 / (y2 - y1)) / p.fs;
        }
        ++at;
      }
    }
    WB_SYNC();
    // noise run of each pulse: max(3, next index - this index) (synthesis.py:65, 93)
    const int pchunk = (n_p + nthr - 1) / nthr;
    const int plo = wb_imin(n_p, tid * pchunk), phi = wb_imin(n_p, plo + pchunk);
    int need = 0;
    for (int i = plo; i < phi; ++i) {
      const int ns = idx[wb_imin(n_p - 1, i + 1)] - idx[i];
      need += ns > 3 ? ns : 3;
    }
    double need_tot;
    int noff = (int)wb_exscan_d((double)need, smem, tid, nthr, &need_tot);
    int* pno = p.p_noise_off + (size_t)u * p.p_cap;
    for (int i = plo; i < phi; ++i) {
      pno[i] = noff;
      const int ns = idx[wb_imin(n_p - 1, i + 1)] - idx[i];
      noff += ns > 3 ? ns : 3;
    }
    if (tid == 0) {
      p.out_len[u] = L;
      p.n_pulses[u] = n_p;
      p.noise_total[u] = (int)need_tot;
    }
  }
};

// prefix of the pulse counts over the batch (one thread)
struct wb_sy_prefix {
  wb_sy_plan p;
  WB_DEV void operator()(long long) const {
    int a = 0;
    for (int u = 0; u < p.batch; ++u) {
      p.pulse_base[u] = a;
      a += p.n_pulses[u];
    }
    p.pulse_base[p.batch] = a;
  }
};

// Minimum-phase spectrum the way the reference builds it (synthesis.py:87-92, 104-111): cepstrum of
// log|S|/2 over the symmetric spectrum, kept at quefrency 0 and doubled on the upper half, back to the
// spectral domain, exp.  All sequences are real, so both transforms are half-size real FFTs.
// In: Ad[0..n/2] = log(|s|)/2 (doubles in buffer A).  Out: the half spectrum Z[0..n/2] (the other half is
// its conjugate mirror); returns the buffer holding it.
template <int NC = 0>  // NC: n when it is known at compile time
WB_DEV wb_cplx* wb_sy_minphase(wb_cplx* A, wb_cplx* B, int n, const wb_cplx* twS, int twH, int tid, int nthr) {
  const int nh = n / 2;
  double* Ad = (double*)A;
  for (int k = tid; k < nh - 1; k += nthr) Ad[n - 1 - k] = Ad[k + 1];  // symmetric extension
  WB_SYNC();
  wb_cplx* Cq = wb_rfft<0, NC>(A, B, n, twS, twH, tid, nthr);
  wb_cplx* Ot = (Cq == A) ? B : A;
  double* cc = (double*)Ot;  // folded cepstrum: c[0], zeros, 2 c[i] for i >= n/2 (c is even: c[i] = c[n-i])
  for (int i = tid; i < n; i += nthr) {
    double v = 0.0;
    if (i == 0) v = Cq[0].x;
    else if (i >= nh) v = Cq[n - i].x * 2.0;
    cc[i] = v;
  }
  WB_SYNC();
  wb_cplx* Z = wb_rfft<0, NC>(Ot, Cq, n, twS, twH, tid, nthr);
  const double inv_n = 1.0 / n;
  for (int k = tid; k <= nh; k += nthr) {  // exp(ifft(cc)[k]) with ifft(x)[k] = conj(fft(x)[k]) / n for real x
    const double re = Z[k].x * inv_n, im = -Z[k].y * inv_n;
    const double e = exp(re);
    double sn, cs;
    sincos(im, &sn, &cs);
    Z[k] = wb_mk(e * cs, e * sn);
  }
  WB_SYNC();
  return Z;
}

// overlap-add of v[0..n) at 1-based targets first, first+1, ... with NumPy's duplicate-index semantics
WB_DEV void wb_sy_scatter(double* y, int len, int first, const double* v, int n, double gain, int tid, int nthr) {
  const bool hi_clamp = first + n - 1 > len;
  for (int i = tid; i < n; i += nthr) {
    const int tgt = first + i;
    if (tgt < 1) continue;                     // only the element landing on sample 1 survives the low clamp
    if (hi_clamp && tgt >= len) {
      if (i == n - 1) wb_atomic_add(y + len - 1, v[i] * gain);  // only the final element survives the high clamp
      continue;
    }
    wb_atomic_add(y + tgt - 1, v[i] * gain);
  }
}

// ------------------------------------------------------------------------------------ Y2
// NC / NT: FFT size and block size when the launcher knows them at compile time (0: run-time values)
template <int NC = 0, int NT = 0>
struct wb_sy_pulses_t {
  wb_sy_plan p;
  const wb_cplx* tw;
  int tw_n;
  const double* dc_base;   // hanning(n+2)[1:-1] / sum  (synthesis.py:57-58)
  const double* noise;     // [B, noise_stride] standard normals in the reference's draw order, or nullptr
  int noise_stride;
  unsigned long long seed; // device generator when noise == nullptr
  int n_slots;
  int max_noise;           // capacity of the shared noise buffer

  static size_t smem_bytes(int n, int max_noise, int nthr) {
    return (size_t)(n / 2 + 1) * 2 * sizeof(wb_cplx) + ((size_t)n + max_noise + 3 * ((size_t)n / 2 + 1) + WB_REDUCE_SCRATCH + 16) *
                                                 sizeof(double) + (size_t)WB_FFT_TW_SLOTS(n / 2) * sizeof(wb_cplx) + 0 * nthr;
  }

  WB_DEV double normal(int u, long long k) const {  // counter-based N(0,1): two hashed uniforms, Box-Muller
    unsigned long long h = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(u + 1) + (unsigned long long)k * 0xBF58476D1CE4E5B9ull;
    unsigned long long a = h, b;
    a ^= a >> 30; a *= 0xBF58476D1CE4E5B9ull; a ^= a >> 27; a *= 0x94D049BB133111EBull; a ^= a >> 31;
    b = a + 0x9E3779B97F4A7C15ull;
    b ^= b >> 30; b *= 0xBF58476D1CE4E5B9ull; b ^= b >> 27; b *= 0x94D049BB133111EBull; b ^= b >> 31;
    const double u1 = ((double)(a >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    const double u2 = ((double)(b >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    return sqrt(-2.0 * log(u1)) * cos(2.0 * WB_PI * u2);
  }

  WB_DEV void operator()(int block, int tid, int nthr_rt, double* smem) const {
    const int nthr = NT ? NT : nthr_rt;
    const int n = NC ? NC : p.n, nh = n / 2, nb = nh + 1;
    wb_cplx* A = (wb_cplx*)smem;         // nh + 1 complex
    wb_cplx* B = A + (nh + 1);
    double* resp = (double*)(B + (nh + 1));  // n
    double* nz = resp + n;               // max_noise
    double* Ssl = nz + max_noise;        // nb: spectrum slice
    double* Psl = Ssl + nb;              // nb: periodic amplitude slice
    double* Asl = Psl + nb;              // nb: aperiodic amplitude slice
    double* scratch = Asl + nb;
    wb_cplx* twS = (wb_cplx*)(scratch + WB_REDUCE_SCRATCH + ((max_noise + 3 * nb + WB_REDUCE_SCRATCH) & 1));
    const int twH = nh;
    wb_fft_load_twiddles(twS, twH, tw, tw_n, tid, nthr);
    const int total = p.pulse_base[p.batch];
    for (int gp = block; gp < total; gp += n_slots) {
      int u = 0;
      {  // utterance of this pulse
        int lo = 0, hi = p.batch;
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (p.pulse_base[mid] <= gp) lo = mid; else hi = mid;
        }
        u = lo;
      }
      const int i = gp - p.pulse_base[u];
      const int F = p.n_frames[u], L = p.out_len[u], n_p = p.n_pulses[u];
      const double* tp = p.tpos + (size_t)u * p.f_stride;
      const int* idx = p.p_idx + (size_t)u * p.p_cap;
      const double loc = p.p_loc[(size_t)u * p.p_cap + i];
      const int id = idx[i];
      const int noise_size = idx[wb_imin(n_p - 1, i + 1)] - id;
      // frame position of the pulse (synthesis.py:50-52, 144-180)
      double a_w, b_w;
      int f_lo, f_hi;
      {
        // interp1d(tp, 1..F)(loc): value at knot j (0-based) is j + 1
        const int br = wb_sy_bracket(tp, F, loc, (int)(loc / (tp[1] - tp[0])) + 1);
        const double slope = ((double)(br + 1) - (double)br) / (tp[br] - tp[br - 1]);
        double pos = WB_ADD(WB_MUL(slope, loc - tp[br - 1]), (double)br);
        pos = wb_dmax(1.0, wb_dmin((double)F, pos));
        f_lo = (int)floor(pos) - 1;
        f_hi = (int)ceil(pos) - 1;
        const double t1 = tp[f_lo], t2 = tp[f_hi];
        const double xq = wb_dmax(t1, wb_dmin(t2, loc));
        if (t1 == t2) {
          a_w = 1.0;
          b_w = 0.0;
        } else {
          b_w = (xq - t1) / (t2 - t1);
          a_w = 1.0 - b_w;
        }
      }
      const double* S0 = p.spec + ((size_t)u * p.f_stride + f_lo) * nb;
      const double* S1 = p.spec + ((size_t)u * p.f_stride + f_hi) * nb;
      const double* Q0 = p.ap + ((size_t)u * p.f_stride + f_lo) * nb;
      const double* Q1 = p.ap + ((size_t)u * p.f_stride + f_hi) * nb;
      const bool same = (f_lo == f_hi) || (tp[f_lo] == tp[f_hi]);
      for (int k = tid; k < nb; k += nthr) {
        const double q0 = Q0[k] * Q0[k], q1 = Q1[k] * Q1[k];           // amplitude_aperiodic = ap ** 2
        const double r0 = wb_dmax(0.001, 1.0 - q0), r1 = wb_dmax(0.001, 1.0 - q1);
        if (same) {
          Ssl[k] = S0[k];
          Psl[k] = r0;
          Asl[k] = q0;
        } else {
          Ssl[k] = a_w * S0[k] + b_w * S1[k];
          Psl[k] = a_w * r0 + b_w * r1;
          Asl[k] = a_w * q0 + b_w * q1;
        }
      }
      WB_SYNC();
      const bool voiced = p.vuv_i[(size_t)u * p.y_stride + id - 1] && (Asl[0] <= 0.999);
      double* yu = p.y + (size_t)u * p.y_stride;
      const int first = id + (-nh + 1);  // target of element 0 (base_index starts at -n/2 + 1)

      if (voiced) {  // get_periodic_response (synthesis.py:100-116) + DC removal (:71-74)
        double* Ad = (double*)A;
        for (int k = tid; k <= nh; k += nthr) {
          double v = Ssl[k] * Psl[k];
          if (v == 0.0) v = WB_EPS;
          Ad[k] = log(fabs(v)) / 2.0;
        }
        WB_SYNC();
        wb_cplx* Z = wb_sy_minphase<NC>(A, B, n, twS, twH, tid, nthr);
        wb_cplx* O = (Z == A) ? B : A;
        const double coef = 2.0 * WB_PI * p.fs / n;
        const double sh = p.p_shift[(size_t)u * p.p_cap + i];
        for (int k = tid; k <= nh; k += nthr) {  // fractional delay; the spectrum is Hermitian by construction
          double sn, cs;
          sincos(-coef * sh * (double)k, &sn, &cs);
          const wb_cplx z = Z[k];
          Z[k] = wb_mk(z.x * cs - z.y * sn, (k == 0 || k == nh) ? 0.0 : z.x * sn + z.y * cs);
        }
        WB_SYNC();
        const double* R = wb_irfft<0, NC>(Z, O, n, twS, twH, tid, nthr);
        const double inv_n = 1.0 / n;
        double sum = 0.0;
        for (int k = tid; k < n; k += nthr) {  // fftshift
          const double v = R[(k + nh) & (n - 1)] * inv_n;
          resp[k] = v;
          sum += v;
        }
        sum = wb_block_sum(sum, scratch, tid, nthr);
        for (int k = tid; k < n; k += nthr) resp[k] += WB_LDG(dc_base + k) * -sum;
        WB_SYNC();
        wb_sy_scatter(yu, L, first, resp, n, sqrt((double)(noise_size > 1 ? noise_size : 1)), tid, nthr);
        WB_SYNC();
      }
      // get_aperiodic_response (synthesis.py:86-96)
      {
        double* Ad = (double*)A;
        for (int k = tid; k <= nh; k += nthr) {
          double v = voiced ? Ssl[k] * Asl[k] : Ssl[k];
          if (v == 0.0) v = WB_EPS;
          Ad[k] = log(fabs(v)) / 2.0;
        }
        WB_SYNC();
        wb_cplx* Z = wb_sy_minphase<NC>(A, B, n, twS, twH, tid, nthr);
        wb_cplx* O = (Z == A) ? B : A;
        if (tid == 0) {  // ifft(...).real of a Hermitian spectrum: the two self-conjugate bins contribute their real part
          Z[0].y = 0.0;
          Z[nh].y = 0.0;
        }
        WB_SYNC();
        const double* R = wb_irfft<0, NC>(Z, O, n, twS, twH, tid, nthr);
        const double inv_n = 1.0 / n;
        for (int k = tid; k < n; k += nthr) resp[k] = R[(k + nh) & (n - 1)] * inv_n;
      }
      // the mean is taken over ALL max(3, noise_size) draws of the pulse (synthesis.py:93-95); only the first
      // fft_size of them reach the output, because fftfilt truncates the convolution to len(response)
      const int nn_all = noise_size > 3 ? noise_size : 3;
      const int nn = nn_all < max_noise ? nn_all : max_noise;
      const int noff = p.p_noise_off[(size_t)u * p.p_cap + i];
      double msum = 0.0;
      for (int k = tid; k < nn_all; k += nthr) {
        const double v = noise ? WB_LDG(noise + (size_t)u * noise_stride + noff + k) : normal(u, (long long)noff + k);
        if (k < nn) nz[k] = v;
        msum += v;
      }
      msum = wb_block_sum(msum, scratch, tid, nthr);
      const double mean = msum / nn_all;
      WB_SYNC();
      // fftfilt(noise - mean, response) = linear convolution truncated to n samples (synthesis.py:95, 189-250)
      WB_SYNC();
      double* out = (double*)A;  // n doubles (the buffers hold n + 2), both FFT buffers are free now
      for (int m = tid; m < n; m += nthr) {
        double acc = 0.0;
        const int kmax = m < nn - 1 ? m : nn - 1;
        for (int k = 0; k <= kmax; ++k) acc += (nz[k] - mean) * resp[m - k];
        out[m] = acc;
      }
      WB_SYNC();
      wb_sy_scatter(yu, L, first, out, n, 1.0, tid, nthr);
      WB_SYNC();
    }
  }
};

typedef wb_sy_pulses_t<> wb_sy_pulses;

// ------------------------------------------------------------------------------------ R1
// One block per utterance: sample-rate band aperiodicity, aperiodic component; then periodic pulses.
struct wb_rq_excite {
  wb_sy_plan p;
  const double* pulse_seed;  // [seed_n, n_ap]
  const double* noise_seed;  // [noise_len, n_ap]
  int seed_n, noise_len;
  const double* cursor_in;   // [n_ap] generate_noise.current_index on entry (synthesisRequiem.py:131-141)
  double* cursor_out;        // [B, n_ap] value on exit for each utterance processed alone

  WB_DEV void operator()(int block, int tid, int nthr, double*) const {
    const int u = block;
    const int F = p.n_frames[u], L = p.out_len[u];
    if (F < 2 || L <= 0) return;
    const double* tp = p.tpos + (size_t)u * p.f_stride;
    const double t0 = tp[0];
    const double step = 1.0 / p.fs;
    const double delta = (t0 + step) - t0;
    double* exc = p.exc + (size_t)u * p.y_stride;
    double* api = p.ap_i + (size_t)u * p.n_ap * p.y_stride;
    const int chunk = (L + nthr - 1) / nthr;
    const int lo = wb_imin(L, tid * chunk), hi = wb_imin(L, lo + chunk);
    int br = 1;
    for (int k = lo; k < hi; ++k) {
      const double t = wb_sy_time(t0, step, delta, k);
      br = wb_sy_bracket(tp, F, t, br);
      double acc = 0.0;
      for (int b = 0; b < p.n_ap; ++b) {
        // interp1d(tp, 10 ** (band_ap / 10))(t)  (synthesisRequiem.py:120-128)
        const double y0 = pow(10.0, p.ap[((size_t)u * p.f_stride + br - 1) * p.n_ap + b] / 10.0);
        const double y1 = pow(10.0, p.ap[((size_t)u * p.f_stride + br) * p.n_ap + b] / 10.0);
        const double slope = (y1 - y0) / (tp[br] - tp[br - 1]);
        const double a = WB_ADD(WB_MUL(slope, t - tp[br - 1]), y0);
        api[(size_t)b * p.y_stride + k] = a;
        const long long pos = ((long long)cursor_in[b] + k) % noise_len;
        acc += noise_seed[(size_t)pos * p.n_ap + b] * a;
      }
      exc[k] = acc;
    }
    if (tid == 0)
      for (int b = 0; b < p.n_ap; ++b) cursor_out[(size_t)u * p.n_ap + b] = (double)(((long long)cursor_in[b] + L - 1) % noise_len);
  }
};

// One block per pulse (persistent): band-weighted pulse seed, overlap-add (synthesisRequiem.py:53-71).
struct wb_rq_pulses {
  wb_sy_plan p;
  const double* pulse_seed;
  int seed_n, n_slots;
  static size_t smem_bytes(int seed_n) { return (size_t)(seed_n + 32) * sizeof(double); }
  WB_DEV void operator()(int block, int tid, int nthr, double* smem) const {
    double* resp = smem;
    const int total = p.pulse_base[p.batch];
    for (int gp = block; gp < total; gp += n_slots) {
      int lo = 0, hi = p.batch;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (p.pulse_base[mid] <= gp) lo = mid; else hi = mid;
      }
      const int u = lo, i = gp - p.pulse_base[u];
      const int L = p.out_len[u], n_p = p.n_pulses[u];
      const int* idx = p.p_idx + (size_t)u * p.p_cap;
      const int id = idx[i];
      const double* api = p.ap_i + (size_t)u * p.n_ap * p.y_stride;
      const bool skip = !p.vuv_i[(size_t)u * p.y_stride + id - 1] || api[id - 1] > 0.999;
      if (!skip) {
        const int nsz = idx[wb_imin(n_p - 1, i + 1)] - id;
        const double gain = sqrt((double)(nsz > 1 ? nsz : 1));
        for (int k = tid; k < seed_n; k += nthr) {
          double acc = 0.0;
          for (int b = 0; b < p.n_ap; ++b) acc += pulse_seed[(size_t)k * p.n_ap + b] * (1.0 - api[(size_t)b * p.y_stride + id - 1]);
          resp[k] = acc * gain;
        }
        WB_SYNC();
        wb_sy_scatter(p.exc + (size_t)u * p.y_stride, L, id + (-seed_n / 2 + 1), resp, seed_n, 1.0, tid, nthr);
      }
      WB_SYNC();
    }
  }
};

// ------------------------------------------------------------------------------------ R2
template <int NC = 0, int NT = 0>
struct wb_rq_frames_t {
  wb_sy_plan p;
  const wb_cplx* tw;
  int tw_n;
  const double* win;  // hanning(2*hop+1)[1:-1] for the common hop, or nullptr to compute per frame
  static size_t smem_bytes(int n) { return ((size_t)(n / 2 + 1) * 3 + WB_FFT_TW_SLOTS(n / 2)) * sizeof(wb_cplx) + 64 * sizeof(double); }

  WB_DEV void operator()(int block, int tid, int nthr_rt, double* smem) const {
    const int nthr = NT ? NT : nthr_rt;
    const int u = block / p.f_stride, fr = block - u * p.f_stride;  // fr = i of the reference loop (2 .. F-2)
    const int F = p.n_frames[u];
    if (fr < 2 || fr > F - 2) return;
    const int n = NC ? NC : p.n, nh = n / 2, nb = nh + 1;
    const int L = p.out_len[u];
    const double* tp = p.tpos + (size_t)u * p.f_stride;
    const int hop = (int)((tp[1] - tp[0]) * p.fs);  // truncates: 110 for 110.25 (synthesisRequiem.py:78)
    const int wl = hop * 2 - 1;
    if (hop < 1 || wl > n) return;
    wb_cplx* A = (wb_cplx*)smem;  // nh + 1 complex each
    wb_cplx* B = A + (nh + 1);
    wb_cplx* T = B + (nh + 1);
    wb_cplx* twS = T + (nh + 1);
    const int twH = nh;
    wb_fft_load_twiddles(twS, twH, tw, tw_n, tid, nthr);
    const double* exc = p.exc + (size_t)u * p.y_stride;
    const int origin = (fr - 1) * hop - (hop - 1);  // 1-based
    // windowed excitation -> spectrum
    double* Ad = (double*)A;
    for (int m = tid; m < n; m += nthr) {
      double v = 0.0;
      if (m < wl) {
        int si = origin + m;
        if (si > L) si = L;
        const double w = 0.5 - 0.5 * cos(2.0 * WB_PI * (double)(m + 1) / (double)(wl + 1));  // hanning(wl+2)[1:-1]
        v = exc[si - 1] * w;
      }
      Ad[m] = v;
    }
    WB_SYNC();
    wb_cplx* X = wb_rfft<0, NC>(A, B, n, twS, twH, tid, nthr);
    for (int k = tid; k <= nh; k += nthr) T[k] = X[k];
    WB_SYNC();
    // minimum-phase spectrum of the envelope of frame fr - 1
    const double* S = p.spec + ((size_t)u * p.f_stride + fr - 1) * nb;
    for (int k = tid; k <= nh; k += nthr) Ad[k] = log(fabs(S[k])) / 2.0;
    WB_SYNC();
    wb_cplx* Z = wb_sy_minphase<NC>(A, B, n, twS, twH, tid, nthr);
    wb_cplx* O = (Z == A) ? B : A;
    for (int k = tid; k <= nh; k += nthr) {
      wb_cplx v = wb_cmul(Z[k], T[k]);
      if (k == 0 || k == nh) v.y = 0.0;  // real part of the inverse transform of the Hermitian product
      Z[k] = v;
    }
    WB_SYNC();
    const double* R = wb_irfft<0, NC>(Z, O, n, twS, twH, tid, nthr);
    double* out = (double*)T;
    const double inv_n = 1.0 / n;
    for (int k = tid; k < n; k += nthr) out[k] = R[k] * inv_n;
    WB_SYNC();
    wb_sy_scatter(p.y + (size_t)u * p.y_stride, L, origin, out, n, 1.0, tid, nthr);
  }
};

typedef wb_rq_frames_t<> wb_rq_frames;

// ------------------------------------------------------------------------------------ Y3
struct wb_sy_normalise {
  wb_sy_plan p;
  int requiem;
  WB_DEV void operator()(int block, int tid, int nthr, double* smem) const {
    const int u = block;
    const int L = p.out_len[u];
    double* y = p.y + (size_t)u * p.y_stride;
    double m = 0.0;
    for (int k = tid; k < L; k += nthr) m = wb_dmax(m, fabs(y[k]));
    m = wb_block_max(m, smem, tid, nthr);
    if (m > 1.0)
      for (int k = tid; k < L; k += nthr) y[k] = y[k] / m;
  }
};
