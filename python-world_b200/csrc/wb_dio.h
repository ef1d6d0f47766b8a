// DIO contour selection and StoneMask refinement -- kernel bodies.
// Replaces world/dio.py:113-124, 216-340 (candidate sort, fix_f0_contour) and world/stonemask.py:8-76.
// DIO's decimation and its band filtering / event streams reuse the Harvest kernels
// (wb_harvest.h: decimator kind 1, channel mode 1).
#pragma once
#include "wb_harvest.h"

#define WB_DIO_MAXB 32  // bands (7 with the defaults: ceil(log2(800/71) * 2))

// float("{0:.6f}".format(v)) for 0 <= v < 1e9 (dio.py:243): correctly rounded decimal, ties to even on
// the exact binary value, then the nearest double of that decimal.
WB_HD double wb_round6(double v) {
  const double p = v * 1e6;
  const double e = fma(v, 1e6, -p);  // exact residual of the product
  double n = floor(p);
  const double d = (p - n) - 0.5;    // exact
  bool up;
  if (d > 0.0) up = true;
  else if (d < 0.0) up = false;
  else if (e > 0.0) up = true;
  else if (e < 0.0) up = false;
  else up = (fmod(n, 2.0) != 0.0);   // exact tie: round half to even
  if (up) n += 1.0;
  return n / 1e6;
}

struct wb_dio_contour {
  wb_hv_plan p;          // raw / stab maps, sizes
  int n_bands;
  double allowed_range;
  double* cand;          // [B, f1_stride, n_bands] candidates sorted by stability (workspace)
  double* work;          // [B, 4, f1_stride] step buffers
  int* sect;             // [B, 4, f1_stride] voiced sections (start, end) and the boundary list
  double* out_cand;      // optional [B, f_stride, n_bands]: 'f0_candidates' of the reference's return dict

  // select_best_f0 (dio.py:310-323)
  WB_DEV double pick(double cur, double past, const double* c) const {
    const double ref = (cur * 3.0 - past) / 2.0;
    double best = c[0], err = fabs(ref - c[0]);
    for (int i = 1; i < n_bands; ++i) {
      const double e = fabs(ref - c[i]);
      if (e < err) {
        err = e;
        best = c[i];
      }
    }
    if (fabs(1.0 - best / (ref + WB_EPS)) > allowed_range) best = 0.0;
    return best;
  }

  WB_DEV void operator()(long long item) const {
    const int u = (int)item;
    const int ns = p.n_samples[u];
    const int F = wb_hv_frames(ns, p.fs, p.frame_period);
    p.out_n_frames[u] = F;
    double* of0 = p.out_f0 + (size_t)u * p.f_stride;
    double* ovuv = p.out_vuv + (size_t)u * p.f_stride;
    double* otp = p.out_tpos + (size_t)u * p.f_stride;
    double* C = cand + (size_t)u * p.f1_stride * n_bands;
    double* s1 = work + (size_t)u * 4 * p.f1_stride;
    double* s2 = s1 + p.f1_stride;
    double* s3 = s2 + p.f1_stride;
    double* s4 = s3 + p.f1_stride;
    int* sec_st = sect + (size_t)u * 4 * p.f1_stride;
    int* sec_ed = sec_st + p.f1_stride;
    // sort_candidates (dio.py:113-124): descending stability, stable
    for (int j = 0; j < F; ++j) {
      double key[WB_DIO_MAXB], val[WB_DIO_MAXB];
      for (int b = 0; b < n_bands; ++b) {
        key[b] = p.stab[((size_t)u * p.n_ch + b) * p.f1_stride + j];
        val[b] = p.raw[((size_t)u * p.n_ch + b) * p.f1_stride + j];
      }
      for (int a = 1; a < n_bands; ++a) {  // insertion sort
        const double k = key[a], v = val[a];
        int q = a - 1;
        while (q >= 0 && key[q] < k) {
          key[q + 1] = key[q];
          val[q + 1] = val[q];
          --q;
        }
        key[q + 1] = k;
        val[q + 1] = v;
      }
      for (int b = 0; b < n_bands; ++b) {
        C[(size_t)j * n_bands + b] = val[b];
        if (out_cand) out_cand[((size_t)u * p.f_stride + j) * n_bands + b] = val[b];
      }
      otp[j] = (double)j * p.frame_period / 1000.0;
    }
    // fix_f0_contour (dio.py:216-230)
    const int vrm = (int)(1.0 / (p.frame_period / 1000.0) / p.f0_floor + 0.5) * 2 + 1;
    if (F < 2 * vrm + 2) {
      for (int j = 0; j < F; ++j) {
        of0[j] = 0.0;
        ovuv[j] = 0.0;
      }
      return;
    }
    // step 1 (dio.py:234-247): row 0 loses its first/last vrm frames (in place, as steps 3-4 read it back)
    for (int j = 0; j < F; ++j)
      if (j < vrm || j >= F - vrm) C[(size_t)j * n_bands] = 0.0;
    for (int j = 0; j < F; ++j) {
      double v = C[(size_t)j * n_bands];
      if (j >= vrm - 1) {
        const double r1 = wb_round6(v), r0 = wb_round6(C[(size_t)(j - 1) * n_bands]);
        if (fabs((r1 - r0) / (0.000001 + r1)) > allowed_range) v = 0.0;
      }
      s1[j] = v;
    }
    // step 2 (dio.py:252-259): a frame survives only inside a fully voiced window of vrm frames
    const int hw = (vrm - 1) / 2;
    for (int j = 0; j < F; ++j) {
      double v = s1[j];
      if (j >= hw && j < F - hw) {
        for (int q = -hw; q <= hw; ++q)
          if (s1[j + q] == 0.0) {
            v = 0.0;
            break;
          }
      }
      s2[j] = v;
      s3[j] = v;
    }
    // count_voiced_sections (dio.py:327-340): boundaries = [0] + every index where vuv changes + [F-2]
    int n_sec = 0;
    {
      int* bl = sec_ed + p.f1_stride;  // [F + 1]
      int nb = 0;
      bl[nb++] = 0;
      for (int k = 0; k < F - 1; ++k)
        if ((s2[k + 1] != 0.0) != (s2[k] != 0.0)) bl[nb++] = k;
      bl[nb++] = F - 2;
      const int k1 = bl[1];
      const int d1 = (int)(s2[k1 + 1] != 0.0) - (int)(s2[k1] != 0.0);
      const int first = (int)ceil(-0.5 * (double)d1);  // 1 when the contour starts voiced
      n_sec = (int)floor((nb - (1 - first)) / 2.0);
      for (int i = 0; i < n_sec; ++i) {
        sec_st[i] = 1 + bl[2 * i + (1 - first)];
        sec_ed[i] = bl[2 * i + 1 + (1 - first)];
      }
    }
    // step 3 (dio.py:264-277): forward tracking from each section end
    for (int i = 0; i < n_sec; ++i) {
      const int limit = (i == n_sec - 1) ? F - 1 : sec_st[i + 1] + 1;
      for (int j = sec_ed[i]; j < limit; ++j) {
        s3[j + 1] = pick(s3[j], s3[j - 1], C + (size_t)(j + 1) * n_bands);
        if (s3[j + 1] == 0.0) break;
      }
    }
    for (int j = 0; j < F; ++j) s4[j] = s3[j];
    // step 4 (dio.py:281-293): backward tracking from each section start
    for (int i = n_sec - 1; i >= 0; --i) {
      const int limit = (i == 0) ? 1 : sec_ed[i - 1];
      for (int j = sec_st[i]; j >= limit; --j) {
        s4[j - 1] = pick(s4[j], s4[j + 1], C + (size_t)(j - 1) * n_bands);
        if (s4[j - 1] == 0.0) break;
      }
    }
    for (int j = 0; j < F; ++j) {
      of0[j] = s4[j];
      ovuv[j] = s4[j] != 0.0 ? 1.0 : 0.0;
    }
  }
};

// ------------------------------------------------------------------------------------ StoneMask
// One block per (utterance, frame); unvoiced frames return at once.
struct wb_stonemask_body {
  const double* x;
  const int* n_samples;
  const double* tpos;
  const double* f0;
  const int* n_frames;
  const double* time_lut;  // float("{:.4f}".format(k / fs)) for k = -lut_half .. lut_half (stonemask.py:38)
  const wb_cplx* tw;
  int tw_n, lut_half, x_stride, f_stride, fs;
  double* out;

  static size_t smem_bytes(int lut_half, int nthr) {
    return ((size_t)2 * (2 * lut_half + 4) + WB_REDUCE_SCRATCH + 64) * sizeof(double) + 0 * nthr;
  }

  // spectra of seg*main and seg*diff_window at `count` bins round(f * nfft / fs * h), h = 1..count;
  // returns sum(amp * inst_freq) / sum(amp * h)  (stonemask.py:58-64, 68-74)
  WB_DEV double harmonic_mean(double f, int count, int len, int nfft, const double* mainw, const double* segw,
                              double* scratch, int tid, int nthr) const {
    double num = 0.0, den = 0.0;
    for (int h = 1; h <= count; ++h) {
      const double v = f * nfft / fs * h;
      const int bin = (int)(v > 0.0 ? v + 0.5 : v - 0.5) + 1 - 1;  // trunc(round_matlab) + 1, used as index - 1
      const int stepw = tw_n / nfft;
      double sr = 0.0, si = 0.0, dr = 0.0;
      double di = 0.0;
      for (int i = tid; i < len; i += nthr) {
        const wb_cplx w = wb_ldg_cplx(tw + (size_t)(((long long)bin * i) & (nfft - 1)) * stepw);
        const double a = segw[i] * mainw[i + 1];
        const double b = segw[i] * (-(mainw[i + 2] - mainw[i]) / 2.0);
        sr += a * w.x;
        si += a * w.y;
        dr += b * w.x;
        di += b * w.y;
      }
      wb_block_sum3(sr, si, dr, scratch, tid, nthr);
      di = wb_block_sum(di, scratch, tid, nthr);
      double pw = sr * sr + si * si;
      if (pw == 0.0) pw = WB_EPS;
      const double inst = (double)(bin & (nfft - 1)) / nfft * fs + (sr * di - si * dr) / pw * fs / 2.0 / WB_PI;
      const double amp = sqrt(pw);
      num += amp * inst;
      den += amp * h;
    }
    return num / den;
  }

  WB_DEV void operator()(int block, int tid, int nthr, double* smem) const {
    const int u = block / f_stride, j = block - u * f_stride;
    if (j >= n_frames[u]) return;
    const size_t fi = (size_t)u * f_stride + j;
    const double f_in = f0[fi];
    if (f_in == 0.0) {
      if (tid == 0) out[fi] = 0.0;
      return;
    }
    const int half = (int)ceil(3.0 * fs / f_in / 2.0);
    if (half > lut_half || half < 1) {  // outside the planned F0 range: leave the value untouched
      if (tid == 0) out[fi] = f_in;
      return;
    }
    double* mainw = smem;                       // len + 2, zero at both ends
    double* segw = mainw + (2 * lut_half + 4);  // len
    double* scratch = segw + (2 * lut_half + 4);
    const int len = 2 * half + 1;
    const double span = (double)len / fs;
    int lg = 0;
    while ((1 << lg) < len) ++lg;
    const int nfft = 1 << (lg + 1);
    const double t = tpos[fi];
    const double* xu = x + (size_t)u * x_stride;
    const int ns = n_samples[u];
    for (int i = tid; i < len; i += nthr) {
      const double bt = WB_LDG(time_lut + lut_half + (i - half));
      const double v = (t + bt) * fs;
      const double r = v > 0.0 ? v + 0.5 : v - 0.5;  // round_matlab, un-truncated (stonemask.py:39)
      const double wt = (r - 1.0) / fs - t;
      mainw[i + 1] = 0.42 + 0.5 * cos(2.0 * WB_PI * wt / span) + 0.08 * cos(4.0 * WB_PI * wt / span);
      const double rc = r < 1.0 ? 1.0 : (r > (double)ns ? (double)ns : r);
      segw[i] = WB_LDG(xu + ((int)rc - 1));
    }
    if (tid == 0) {
      mainw[0] = 0.0;
      mainw[len + 1] = 0.0;
    }
    WB_SYNC();
    double res;
    const double f1 = harmonic_mean(f_in, 2, len, nfft, mainw, segw, scratch, tid, nthr);
    if (f1 < 0.0) {
      res = 0.0;
    } else {
      res = harmonic_mean(f1, 6, len, nfft, mainw, segw, scratch, tid, nthr);
    }
    if (fabs(res - f_in) / f_in > 0.2 || !(res == res)) res = f_in;  // stonemask.py:25-26
    if (tid == 0) out[fi] = res;
  }
};
