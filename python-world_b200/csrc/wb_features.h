// Spectral feature heads, spectrum / time-axis edits and the PCM edge -- kernel bodies.
//
//   F1 ft_lfbank       log mel-filterbank energies of a magnitude spectrogram      (main.py:305-322, 274-303)
//   F2 ft_mcep         mel-warped log spectrum -> first n0 real-cepstrum terms     (main.py:324-342)
//   F3 ft_mcep_decode  cepstrum -> mel-warped log spectrum -> linear axis -> exp   (main.py:344-358)
//   F4 ft_interp_rows  numpy.interp of every row at fixed query points            (warp_spectrum, main.py:189-194)
//   F5 ft_interp_knots numpy.interp of every element against a short knot list    (modify_duration, main.py:178-187)
//   F6 io_pcm16_in / io_pcm16_out  int16 <-> float64 at the wav edge               (example/prosody.py:12-13, 57)
//
// Rows are frames (bin fast): exactly the [B, F, bins] layout the analysis kernels write, so the heads run on
// the resident spectrogram and only n_filt / n0 values per frame travel to the host.  Everything that depends
// only on the sizes (mel bin tables, pre-emphasis response, interpolation brackets) is a small table prepared
// once by the host side with the reference's own expressions; the kernels do the per-frame arithmetic.
#pragma once
#include "wb_platform.h"

#if defined(WB_HOST_EMU) || !defined(__CUDA_ARCH__)
#define WB_FT_MUL(a, b) ((a) * (b))
#define WB_FT_ADD(a, b) ((a) + (b))
#else
#define WB_FT_MUL(a, b) __dmul_rn((a), (b))  // numpy.interp rounds the product, then the sum
#define WB_FT_ADD(a, b) __dadd_rn((a), (b))
#endif

// numpy.interp for one query x whose bracket j (largest index with xp[j] <= x, clipped to [0, len-1]) is known.
WB_DEV double wb_np_interp_at(const double* xp, const double* fp, int len, int j, double x) {
  if (j >= len - 1) return fp[len - 1];
  const double x0 = xp[j];
  if (x0 == x) return fp[j];
  const double slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - x0);
  return WB_FT_ADD(WB_FT_MUL(slope, x - x0), fp[j]);
}

// ------------------------------------------------------------------------------------ F1
struct wb_ft_lfbank {
  const double* spec;  // [rows, D] magnitude
  const double* habs;  // [D] |1 - prefac e^{-jw}| (freqz, main.py:313)
  const double* fb;    // [n_filt, D] triangular mel filters (get_filterbanks, main.py:274-303)
  int rows, D, n_filt;
  double inv_nfft;     // 1 / nfft, nfft = 2 (D - 1)
  double* out;         // [rows, n_filt]
  static size_t smem_bytes(int D) { return (size_t)(D + 2) * sizeof(double); }
  WB_DEV void operator()(int block, int tid, int nthr, double* smem) const {
    const double* s = spec + (size_t)block * D;
    double* P = smem;
    for (int k = tid; k < D; k += nthr) {
      const double v = s[k] * WB_LDG(habs + k);
      P[k] = inv_nfft * (v * v);
    }
    WB_SYNC();
    const int lane = tid % WB_LANES, wp = tid / WB_LANES, nwp = nthr / WB_LANES;
    for (int j = wp; j < n_filt; j += nwp) {
      const double* w = fb + (size_t)j * D;
      double acc = 0.0;
      for (int k = lane; k < D; k += WB_LANES) acc += P[k] * WB_LDG(w + k);
      acc = wb_lanes_sum(acc);
      if (lane == 0) out[(size_t)block * n_filt + j] = log(acc == 0.0 ? WB_EPS : acc);
    }
  }
};

// ------------------------------------------------------------------------------------ F2
struct wb_ft_mcep {
  const double* spec;  // [rows, D] magnitude
  const int* bin;      // [D] source bin of every mel point (floor(...), main.py:337); numpy.interp at an integer
                       // position is a gather, positions past the last bin repeat it
  const double* ctab;  // [n] cos(2 pi m / n), n = 2 (D - 1)
  int rows, D, n0;
  double* out;         // [rows, n0]
  static size_t smem_bytes(int D) { return (size_t)(D + 2) * sizeof(double); }
  WB_DEV void operator()(int block, int tid, int nthr, double* smem) const {
    const double* s = spec + (size_t)block * D;
    double* L = smem;
    const int n = 2 * (D - 1);
    for (int k = tid; k < D; k += nthr) {
      int b = bin[k];
      b = b < 0 ? 0 : (b > D - 1 ? D - 1 : b);
      L[k] = log(s[b]);
    }
    WB_SYNC();
    // irfft of a real half spectrum: c[m] = (X0 + (-1)^m X_{n/2} + 2 sum_{0<k<n/2} X_k cos(2 pi k m / n)) / n
    const int lane = tid % WB_LANES, wp = tid / WB_LANES, nwp = nthr / WB_LANES;
    for (int m = wp; m < n0; m += nwp) {
      double acc = 0.0;
      for (int k = lane; k < D; k += WB_LANES) {
        const double w = (k == 0 || k == D - 1) ? 1.0 : 2.0;
        acc += w * L[k] * WB_LDG(ctab + (int)(((long long)k * m) % n));
      }
      acc = wb_lanes_sum(acc);
      if (lane == 0) out[(size_t)block * n0 + m] = acc / n;
    }
  }
};

// ------------------------------------------------------------------------------------ F3
struct wb_ft_mcep_decode {
  const double* cep;   // [rows, n0]
  const double* ctab;  // [N] cos(2 pi m / N)
  const double* xp;    // [Dout] mel bin positions (non-decreasing, main.py:355)
  const int* jb;       // [Dout] numpy.interp bracket of query i
  const double* xq;    // [Dout] query (i, or xp[0] where i < xp[0])
  int rows, n0, N, Dout;
  double* out;         // [rows, Dout]
  static size_t smem_bytes(int n0, int Dout) { return (size_t)(n0 + Dout + 2) * sizeof(double); }
  WB_DEV void operator()(int block, int tid, int nthr, double* smem) const {
    double* c = smem;
    double* Y = smem + n0;
    for (int m = tid; m < n0; m += nthr) c[m] = cep[(size_t)block * n0 + m];
    WB_SYNC();
    // real part of the rfft of the symmetric extension c0, c1 .. c_{n0-1}, 0 .. 0, c_{n0-1} .. c1 (main.py:351-353)
    for (int k = tid; k < Dout; k += nthr) {
      double acc = c[0];
      for (int m = 1; m < n0; ++m) acc += 2.0 * c[m] * WB_LDG(ctab + (int)(((long long)k * m) % N));
      Y[k] = acc;
    }
    WB_SYNC();
    for (int i = tid; i < Dout; i += nthr)
      out[(size_t)block * Dout + i] = exp(wb_np_interp_at(xp, Y, Dout, jb[i], xq[i]));
  }
};

// ------------------------------------------------------------------------------------ F4
struct wb_ft_interp_rows {
  const double* in;   // [rows, Din]
  const double* xp;   // [Din] knots
  const int* jb;      // [Dout]
  const double* xq;   // [Dout]
  int rows, Din, Dout;
  double* out;        // [rows, Dout]; may alias `in` when Dout == Din
  static size_t smem_bytes(int Din) { return (size_t)(Din + 2) * sizeof(double); }
  WB_DEV void operator()(int block, int tid, int nthr, double* smem) const {
    double* Y = smem;
    for (int k = tid; k < Din; k += nthr) Y[k] = in[(size_t)block * Din + k];
    WB_SYNC();
    for (int i = tid; i < Dout; i += nthr) out[(size_t)block * Dout + i] = wb_np_interp_at(xp, Y, Din, jb[i], xq[i]);
  }
};

// ------------------------------------------------------------------------------------ F5
struct wb_ft_interp_knots {
  const double* x;
  const double* xp;  // [n_knots] strictly increasing
  const double* fp;
  int n_knots;
  double* out;       // may alias x
  WB_DEV void operator()(long long item) const {
    const double v = x[item];
    double r;
    if (v > xp[n_knots - 1]) r = fp[n_knots - 1];
    else if (v < xp[0]) r = fp[0];
    else {
      int lo = 0, hi = n_knots - 1;  // largest j with xp[j] <= v
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (xp[mid] <= v) lo = mid;
        else hi = mid - 1;
      }
      r = wb_np_interp_at(xp, fp, n_knots, lo, v);
    }
    out[item] = r;
  }
};

// ------------------------------------------------------------------------------------ F6
struct wb_io_pcm16_in {  // x = x_int16 / (2**15 - 1)  (example/prosody.py:13, test/speed.py:14)
  const short* in;
  int in_stride, out_stride;
  const int* n_samples;
  double divisor;
  double* out;
  WB_DEV void operator()(long long item) const {
    const int u = (int)(item / out_stride), i = (int)(item - (long long)u * out_stride);
    out[item] = i < n_samples[u] ? (double)in[(size_t)u * in_stride + i] / divisor : 0.0;
  }
};

struct wb_io_pcm16_out {  // (out * 2**15).astype(np.int16)  (example/prosody.py:57): truncation, 16-bit wrap-around
  const double* in;
  int in_stride, out_stride;
  const int* n_samples;
  double gain;
  short* out;
  WB_DEV void operator()(long long item) const {
    const int u = (int)(item / out_stride), i = (int)(item - (long long)u * out_stride);
    short r = 0;
    if (i < n_samples[u]) {
      const double v = in[(size_t)u * in_stride + i] * gain;
      const long long q = (v >= 2147483648.0 || v < -2147483648.0 || v != v) ? (long long)(-2147483647 - 1) : (long long)v;
      r = (short)(unsigned short)((unsigned long long)q & 0xffffull);
    }
    out[item] = r;
  }
};

// Optional float32 transport of per-frame matrices across PCIe (the batch API's opt-in `spectrogram_dtype`):
// round-to-nearest on the way out, exact widening on the way in.  Two elements per thread.
struct wb_io_f64_to_f32 {
  const double* in;
  float* out;
  long long n;
  WB_DEV void operator()(long long item) const {
    const long long i = item * 2;
    if (i + 1 < n) {
      out[i] = (float)in[i];
      out[i + 1] = (float)in[i + 1];
    } else if (i < n) {
      out[i] = (float)in[i];
    }
  }
};
struct wb_io_f32_to_f64 {
  const float* in;
  double* out;
  long long n;
  WB_DEV void operator()(long long item) const {
    const long long i = item * 2;
    if (i + 1 < n) {
      out[i] = (double)in[i];
      out[i + 1] = (double)in[i + 1];
    } else if (i < n) {
      out[i] = (double)in[i];
    }
  }
};
