// C-ABI: Harvest F0 estimation (replaces world/harvest.py:17 harvest()).
#include <algorithm>

#include "wb_filter_tables.h"
#include "wb_handle.h"
#include "wb_harvest.h"

// zero-input basis responses H[c][n] and the chunk transition matrix M[r][c] of the decimation filter in cb
void wb_hv_fill_zir(std::vector<double>& o, int kind) {
  for (int c = 0; c < 3; ++c) {
    double s0 = c == 0, s1 = c == 1, s2 = c == 2;
    for (int n = 0; n < WB_HV_CHUNK; ++n) {
      double out;
      if (kind == 0) {
        out = s0;
        s0 = -o[5] * out + s1;
        s1 = -o[6] * out + s2;
        s2 = -o[7] * out;
      } else {
        const double wt = 0.0 + o[5] * s0 + o[6] * s1 + o[7] * s2;
        out = o[0] * wt + o[1] * s0 + o[1] * s1 + o[0] * s2;
        s2 = s1;
        s1 = s0;
        s0 = wt;
      }
      o[11 + c * WB_HV_CHUNK + n] = out;
    }
    o[11 + 3 * WB_HV_CHUNK + 0 + c] = s0;
    o[11 + 3 * WB_HV_CHUNK + 3 + c] = s1;
    o[11 + 3 * WB_HV_CHUNK + 6 + c] = s2;
  }
}

#ifndef WB_HV_REFINE_MINB
#ifndef WB_HV_REFINE_MINB
#define WB_HV_REFINE_MINB 4  // blocks of 128 threads per SM of the refinement kernel (register cap 128)
#endif
#endif

namespace {

struct hv_sizes {
  int ratio, pad, n_ch, max_taps, max_win;
  double afs;
  int ext_stride, y_stride, f1_stride, edge_cap, n_slots, dec_chunks;
  int fft_nch, fft_blocks, fft_groups;  // overlap-save path (fft_nch = 0: unused); groups by filter length
  int fft_gc[WB_HV_FFT_GROUPS + 1], fft_gV[WB_HV_FFT_GROUPS], fft_gA[WB_HV_FFT_GROUPS], fft_gblocks[WB_HV_FFT_GROUPS];
  long long fft_goff[WB_HV_FFT_GROUPS];
  long long ctr_stride;
  size_t off[18];
  size_t total;
};

inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

// Filters with more taps than this run through the overlap-save kernel.  A 2048-point inverse real FFT per 1554
// output samples costs about as much as a ~100-tap direct filter, but the per-tile cost around the filter (event
// detection, barriers) is lower in the overlap-save kernel, so even Harvest's shortest filters (37 taps at
// 8 kHz) are at least as fast there: measured 17.5 ms with every channel on it against 18.3 ms with the
// threshold at 96 taps.  The direct kernel remains for DIO (circular FFT semantics of dio.py:74-88) and for
// configurations whose filters do not fit the block.  WB_HV_FFT_MIN_TAPS overrides the threshold (tuning).
int hv_fft_min_taps() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("WB_HV_FFT_MIN_TAPS");
    v = e ? std::atoi(e) : 24;
    if (v < 1) v = 1;
  }
  return v;
}

// Harvest filter half length of channel c (harvest.py:253, Decimal ROUND_HALF_UP)
int hv_half(double afs, double lo, int c) {
  const double edge = std::pow(2.0, (double)(c + 1) / 40) * lo;
  const double v = afs / edge * 2;
  const double fl = std::floor(v);
  return (int)fl + ((v - fl) >= 0.5 ? 1 : 0);
}

int hv_plan_sizes(int batch, int max_samples, int fs, double f0_floor, double f0_ceil, int n_slots, hv_sizes* z) {
  if (fs <= 0 || batch < 0 || max_samples < 0 || !(f0_floor > 0) || !(f0_ceil > f0_floor)) return WB_E_INVALID;
  z->ratio = (int)(fs / 8000.0 + 0.5);  // harvest.py:59
  if (fs <= 8000) z->ratio = 1;
  if (z->ratio > WB_CHEBY_MAX_RATIO) return WB_E_UNSUPPORTED;
  // the reference filters whenever fs > 8000, also when the ratio rounds to 1 (8 kHz < fs < 12 kHz): harvest.py:61-69
  z->pad = fs > 8000 ? (int)std::ceil(140.0 / z->ratio) * z->ratio : 0;  // harvest.py:65
  z->afs = fs > 8000 ? (double)fs / z->ratio : (double)fs;
  const double lo = f0_floor * 0.9, hi = f0_ceil * 1.1;
  z->n_ch = (int)std::ceil(std::log2(hi / lo) * 40);  // harvest.py:26
  if (z->n_ch < 3 || z->n_ch > 1024) return WB_E_UNSUPPORTED;
  const double e0 = lo * std::pow(2.0, 1.0 / 40);
  z->max_taps = 2 * ((int)(z->afs / e0 * 2 + 0.5) + 1) + 1;
  z->max_win = 2 * (int)std::ceil(3.0 * z->afs / f0_floor / 2.0) + 3;
  z->ext_stride = max_samples + 2 * z->pad + 18 + 2;
  z->dec_chunks = (z->ext_stride + WB_HV_CHUNK - 1) / WB_HV_CHUNK + 1;
  z->y_stride = (max_samples + 2 * z->pad) / z->ratio + 4;
  z->f1_stride = wb_hv_frames(max_samples, fs, 1.0) + 1;
  z->edge_cap = z->y_stride / 2 + 4;
  z->n_slots = n_slots;
  z->ctr_stride = wb_hv_contour::scratch_doubles(z->f1_stride);
  {  // channels are ordered by rising edge frequency, i.e. falling filter length: the first fft_nch are "long"
    const int h_max = hv_half(z->afs, lo, 0);
    z->fft_nch = 0;
    while (z->fft_nch < z->n_ch && 2 * hv_half(z->afs, lo, z->fft_nch) + 1 > hv_fft_min_taps()) ++z->fft_nch;
    if (WB_HV_FFT_N - 2 - 2 * h_max < WB_HV_FFT_N / 4) z->fft_nch = 0;  // filters too long for this transform size: all direct
    // group g holds the channels whose half length is at most h_max >> g (contiguous: the lengths fall with the
    // channel index); a block of group g yields 2046 - 2 h_top(g) output positions
    z->fft_groups = 0;
    z->fft_blocks = 0;
    long long off = 0;
    int c = 0;
    while (c < z->fft_nch && z->fft_groups < WB_HV_FFT_GROUPS) {
      const int g = z->fft_groups++;
      const int h_top = hv_half(z->afs, lo, c);
      z->fft_gc[g] = c;
      z->fft_gV[g] = WB_HV_FFT_N - 2 - 2 * h_top;
      z->fft_gA[g] = -h_top + 1;
      z->fft_gblocks[g] = z->y_stride / z->fft_gV[g] + 2;
      z->fft_goff[g] = off;
      off += (long long)batch * z->fft_gblocks[g] * (WB_HV_FFT_N / 2 + 1);
      z->fft_blocks += z->fft_gblocks[g];
      if (g + 1 == WB_HV_FFT_GROUPS) {
        c = z->fft_nch;
      } else {
        while (c < z->fft_nch && hv_half(z->afs, lo, c) > (h_top >> 1)) ++c;
      }
    }
    for (int g = z->fft_groups; g <= WB_HV_FFT_GROUPS; ++g) z->fft_gc[g] = z->fft_nch;
    for (int g = z->fft_groups; g < WB_HV_FFT_GROUPS; ++g) {
      z->fft_gV[g] = z->fft_gA[g] = z->fft_gblocks[g] = 0;
      z->fft_goff[g] = 0;
    }
  }
  const size_t B = (size_t)batch, F1 = (size_t)z->f1_stride;
  size_t o = 0;
  int i = 0;
  auto put = [&](size_t bytes) {
    z->off[i++] = o;
    o += align_up(bytes);
  };
  put(2 * B * z->ext_stride * sizeof(double) + 4 * B * (size_t)z->dec_chunks * 3 * sizeof(double));  // 0 fwd,bwd,states
  put(B * z->y_stride * sizeof(double));                    // 1 y
  put(B * sizeof(int));                                     // 2 y_len
  put(B * z->n_ch * F1 * sizeof(double));                   // 3 raw
  put((size_t)n_slots * 4 * z->edge_cap * sizeof(double));  // 4 edge_buf
  put(B * F1 * WB_HV_MAXC * sizeof(double));                // 5 base_c
  put(B * F1 * sizeof(int));                                // 6 base_n
  put(B * F1 * WB_HV_SLOTS * sizeof(double));               // 7 l_f0
  put(B * F1 * WB_HV_SLOTS * sizeof(double));               // 8 l_sc
  put(B * F1 * WB_HV_SLOTS);                                // 9 l_slot
  put(B * F1 * WB_HV_SLOTS);                                // 10 l_keep
  put(B * F1 * sizeof(int));                                // 11 l_n
  put(B * (size_t)z->ctr_stride * sizeof(double));          // 12 contour scratch
  put(256);                                                 // 13 status
  put((2 * WB_HV_NCLS + 8) * sizeof(int));                  // 14 refine class counters / cursors
  put(B * F1 * WB_HV_SLOTS * sizeof(unsigned long long));   // 15 refine work items (worst case)
  put(B * (size_t)z->fft_blocks * (WB_HV_FFT_N / 2 + 1) * sizeof(wb_cplx));  // 16 block spectra of y
  z->total = o;
  return WB_OK;
}

int hv_default_slots(wb_handle* h, int batch, int n_ch) {
#ifdef WB_HOST_EMU
  (void)h;
  long long items = (long long)batch * n_ch;
  return (int)(items < 2 ? (items < 1 ? 1 : items) : 2);
#else
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
  long long items = (long long)batch * n_ch;
  long long s = 4LL * sms;
  if (s > items) s = items;
  return (int)(s < 1 ? 1 : s);
#endif
}

struct hv_tables {
  const double* edges;
  const int* halfs;
  const int* ch_off;
  const int* tap_off;
  const double* taps;
  const double* cb;
  const wb_cplx* fft_H;
};

// in-place forward complex FFT of a power-of-two length (host, set-up only)
void hv_host_fft(std::vector<long double>& re, std::vector<long double>& im) {
  const int n = (int)re.size();
  for (int i = 1, j = 0; i < n; ++i) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) {
      std::swap(re[i], re[j]);
      std::swap(im[i], im[j]);
    }
  }
  const long double pi = 3.14159265358979323846264338327950288L;
  for (int len = 2; len <= n; len <<= 1) {
    for (int k = 0; k < len / 2; ++k) {
      const long double wr = cosl(2 * pi * k / len), wi = -sinl(2 * pi * k / len);
      for (int i = k; i < n; i += len) {
        const int j = i + len / 2;
        const long double xr = re[j] * wr - im[j] * wi, xi = re[j] * wi + im[j] * wr;
        re[j] = re[i] - xr;
        im[j] = im[i] - xi;
        re[i] += xr;
        im[i] += xi;
      }
    }
  }
}

int hv_get_tables(wb_handle* h, const hv_sizes& z, double f0_floor, double f0_ceil, hv_tables* t) {
  char key[160];
  snprintf(key, sizeof key, "hv:%.17g:%.17g:%.17g:%d", z.afs, f0_floor, f0_ceil, z.ratio);
  const std::string k(key);
  const int n_ch = z.n_ch;
  const double afs = z.afs, lo = f0_floor * 0.9;
  std::vector<double> edges(n_ch);
  std::vector<int> halfs(n_ch), offs(n_ch), lens(n_ch), first(n_ch);
  int total = 0;
  for (int c = 0; c < n_ch; ++c) {
    edges[c] = std::pow(2.0, (double)(c + 1) / 40) * lo;  // harvest.py:26-29
    const double v = afs / edges[c] * 2;                  // harvest.py:253 (Decimal ROUND_HALF_UP)
    const double fl = std::floor(v);
    halfs[c] = (int)fl + ((v - fl) >= 0.5 ? 1 : 0);
    offs[c] = total;
    total += 2 * halfs[c] + 1;
    lens[c] = 2 * halfs[c] + 1;   // filtered[n] = conv(y, taps)[n + half + 1] (harvest.py:258-262)
    first[c] = -halfs[c] + 1;
  }
  t->edges = wb_table<double>(h, k + ":edges", [&](std::vector<double>& o) { o = edges; });
  t->halfs = wb_table<int>(h, k + ":lens", [&](std::vector<int>& o) { o = lens; });
  t->ch_off = wb_table<int>(h, k + ":first", [&](std::vector<int>& o) { o = first; });
  t->tap_off = wb_table<int>(h, k + ":offs", [&](std::vector<int>& o) { o = offs; });
  t->taps = wb_table<double>(h, k + ":taps", [&](std::vector<double>& o) {
    o.resize(total);
    std::vector<double> win;
    for (int c = 0; c < n_ch; ++c) {  // nuttall * cosine carrier (harvest.py:254-256), stored reversed
      const int hh = halfs[c], L = 2 * hh + 1;
      wb_nuttall(L, win);
      for (int i = 0; i < L; ++i) {
        const double tap = win[i] * std::cos(2 * WB_PI * edges[c] * (double)(i - hh) / afs);
        o[offs[c] + (L - 1 - i)] = tap;
      }
    }
  });
  t->cb = wb_table<double>(h, k + ":cheby", [&](std::vector<double>& o) {
    o.assign(11 + 3 * WB_HV_CHUNK + 9, 0.0);
    for (int i = 0; i < 4; ++i) o[i] = wb_cheby_b[z.ratio][i];
    for (int i = 0; i < 4; ++i) o[4 + i] = wb_cheby_a[z.ratio][i];
    for (int i = 0; i < 3; ++i) o[8 + i] = wb_cheby_zi[z.ratio][i];
    wb_hv_fill_zir(o, 0);
  });
  t->fft_H = nullptr;
  if (z.fft_nch > 0) {
    char k2[64];
    snprintf(k2, sizeof k2, ":fftH:%d:%d:%d:%d:%d", z.fft_nch, WB_HV_FFT_N, z.fft_groups, z.fft_gc[1], z.fft_gc[2]);
    t->fft_H = wb_table<wb_cplx>(h, k + k2, [&](std::vector<wb_cplx>& o) {
      const int N = WB_HV_FFT_N, NH = N / 2;
      o.resize((size_t)z.fft_nch * (NH + 1));
      std::vector<double> win;
      std::vector<long double> re(N), im(N);
      for (int c = 0; c < z.fft_nch; ++c) {
        int g = 0;
        while (g + 1 < z.fft_groups && c >= z.fft_gc[g + 1]) ++g;
        const int hh = halfs[c], L = 2 * hh + 1, d = halfs[z.fft_gc[g]] - hh;  // d = off0_c - A(group)
        wb_nuttall(L, win);
        std::fill(re.begin(), re.end(), 0.0L);
        std::fill(im.begin(), im.end(), 0.0L);
        for (int i = 0; i < L; ++i)  // reversed tap L-1-i sits at offset d + (L-1-i)
          re[d + (L - 1 - i)] = win[i] * std::cos(2 * WB_PI * edges[c] * (double)(i - hh) / afs);
        hv_host_fft(re, im);
        for (int q = 0; q <= NH; ++q) o[(size_t)c * (NH + 1) + q] = wb_mk((double)(re[q] / N), (double)(-im[q] / N));
      }
    });
    if (!t->fft_H) return WB_E_NOMEM;
  }
  if (!t->edges || !t->halfs || !t->ch_off || !t->tap_off || !t->taps || !t->cb) return WB_E_NOMEM;
  return WB_OK;
}

}  // namespace

extern "C" {

int wb_harvest_workspace_bytes(wb_handle* h, int batch, int max_samples, int fs, double f0_floor, double f0_ceil,
                               size_t* bytes) {
  if (!h || !bytes) return WB_E_INVALID;
  hv_sizes z;
  const double lo = f0_floor * 0.9, hi = f0_ceil * 1.1;
  const int n_ch = (lo > 0 && hi > lo) ? (int)std::ceil(std::log2(hi / lo) * 40) : 0;
  int rc = hv_plan_sizes(batch, max_samples, fs, f0_floor, f0_ceil, hv_default_slots(h, batch, n_ch), &z);
  if (rc) return wb_fail(h, rc, "wb_harvest_workspace_bytes: unsupported configuration (fs=%d)", fs);
  *bytes = z.total;
  return WB_OK;
}

/* Diagnostic: byte offsets of the intermediates inside the workspace (tests pin every stage):
 * 1 y, 2 y_len, 3 raw, 5 base_c, 6 base_n, 7 l_f0, 8 l_sc, 9 l_slot, 10 l_keep, 11 l_n; strides in dims[]. */
int wb_harvest_workspace_layout(wb_handle* h, int batch, int max_samples, int fs, double f0_floor, double f0_ceil,
                                size_t* offsets16, int* dims8) {
  if (!h || !offsets16 || !dims8) return WB_E_INVALID;
  hv_sizes z;
  const double lo = f0_floor * 0.9, hi = f0_ceil * 1.1;
  const int n_ch = (lo > 0 && hi > lo) ? (int)std::ceil(std::log2(hi / lo) * 40) : 0;
  int rc = hv_plan_sizes(batch, max_samples, fs, f0_floor, f0_ceil, hv_default_slots(h, batch, n_ch), &z);
  if (rc) return wb_fail(h, rc, "wb_harvest_workspace_layout: unsupported configuration");
  for (int i = 0; i < 16; ++i) offsets16[i] = i < 14 ? z.off[i] : 0;
  dims8[0] = z.y_stride;
  dims8[1] = z.f1_stride;
  dims8[2] = z.n_ch;
  dims8[3] = WB_HV_MAXC;
  dims8[4] = WB_HV_SLOTS;
  dims8[5] = z.ratio;
  dims8[6] = z.max_taps;
  dims8[7] = z.n_slots;
  return WB_OK;
}

int wb_harvest_stages(wb_handle* h, void* stream, const double* d_x, int x_stride, const int* d_n_samples, int batch,
                      int max_samples, int fs, double f0_floor, double f0_ceil, double frame_period_ms,
                      void* d_workspace, size_t workspace_bytes, int f_stride, double* d_tpos, double* d_f0,
                      double* d_vuv, int* d_n_frames, int stage_first, int stage_last);

int wb_harvest(wb_handle* h, void* stream, const double* d_x, int x_stride, const int* d_n_samples, int batch,
               int max_samples, int fs, double f0_floor, double f0_ceil, double frame_period_ms, void* d_workspace,
               size_t workspace_bytes, int f_stride, double* d_tpos, double* d_f0, double* d_vuv, int* d_n_frames) {
  return wb_harvest_stages(h, stream, d_x, x_stride, d_n_samples, batch, max_samples, fs, f0_floor, f0_ceil,
                           frame_period_ms, d_workspace, workspace_bytes, f_stride, d_tpos, d_f0, d_vuv, d_n_frames,
                           0, 5);
}

/* Diagnostic variant: run only kernels stage_first..stage_last (0 decimate, 1 channels, 2 detect, 3 refine,
 * 4 prune, 5 contour) on a workspace that already holds the earlier stages' results; used by bench.py to
 * time each kernel with CUDA events. */
int wb_harvest_stages(wb_handle* h, void* stream, const double* d_x, int x_stride, const int* d_n_samples, int batch,
                      int max_samples, int fs, double f0_floor, double f0_ceil, double frame_period_ms,
                      void* d_workspace, size_t workspace_bytes, int f_stride, double* d_tpos, double* d_f0,
                      double* d_vuv, int* d_n_frames, int stage_first, int stage_last) {
  if (!h) return WB_E_INVALID;
  if (!d_x || !d_n_samples || !d_workspace || !d_tpos || !d_f0 || !d_vuv || !d_n_frames || batch < 0 ||
      max_samples > x_stride || !(frame_period_ms > 0))
    return wb_fail(h, WB_E_INVALID, "wb_harvest: null pointer or inconsistent sizes");
  if (batch == 0) return WB_OK;
  hv_sizes z;
  const double lo = f0_floor * 0.9, hi = f0_ceil * 1.1;
  const int n_ch0 = (lo > 0 && hi > lo) ? (int)std::ceil(std::log2(hi / lo) * 40) : 0;
  int rc = hv_plan_sizes(batch, max_samples, fs, f0_floor, f0_ceil, hv_default_slots(h, batch, n_ch0), &z);
  if (rc) return wb_fail(h, rc, "wb_harvest: unsupported configuration (fs=%d floor=%g ceil=%g)", fs, f0_floor, f0_ceil);
  if (workspace_bytes < z.total)
    return wb_fail(h, WB_E_INVALID, "wb_harvest: workspace %zu < %zu bytes", workspace_bytes, z.total);
  if (f_stride < wb_hv_frames(max_samples, fs, frame_period_ms))
    return wb_fail(h, WB_E_INVALID, "wb_harvest: f_stride %d too small", f_stride);
  WB_SET_DEVICE(h);
  hv_tables t;
  rc = hv_get_tables(h, z, f0_floor, f0_ceil, &t);
  if (rc) return wb_fail(h, rc, "wb_harvest: table allocation failed");
  char* ws = (char*)d_workspace;
  wb_stream_t st = (wb_stream_t)stream;
  wb_hv_plan p;
  memset(&p, 0, sizeof p);
  p.batch = batch;
  p.fs = fs;
  p.ratio = z.ratio;
  p.pad = z.pad;
  p.afs = z.afs;
  p.f0_floor = f0_floor;
  p.f0_ceil = f0_ceil;
  p.frame_period = frame_period_ms;
  p.n_ch = z.n_ch;
  p.max_taps = z.max_taps;
  p.edges = t.edges;
  p.halfs = t.halfs;
  p.ch_off = t.ch_off;
  p.wrap_n = 0;
  p.pow2_quirk = nullptr;
  p.dec_kind = 0;
  p.mode = 0;
  p.grid_ms = 1.0;
  p.stab = nullptr;
  p.four = nullptr;
  p.tap_off = t.tap_off;
  p.taps = t.taps;
  p.cb = t.cb;
  p.x = d_x;
  p.n_samples = d_n_samples;
  p.x_stride = x_stride;
  p.fwd = (double*)(ws + z.off[0]);
  p.bwd = p.fwd + (size_t)batch * z.ext_stride;
  p.ext_stride = z.ext_stride;
  p.dec_chunks = z.dec_chunks;
  p.dec_s1 = p.bwd + (size_t)batch * z.ext_stride;
  p.dec_init = p.dec_s1 + (size_t)batch * z.dec_chunks * 3;
  p.dec_s2 = p.dec_init + (size_t)batch * z.dec_chunks * 3;
  p.dec_initb = p.dec_s2 + (size_t)batch * z.dec_chunks * 3;
  p.y = (double*)(ws + z.off[1]);
  p.y_len = (int*)(ws + z.off[2]);
  p.y_stride = z.y_stride;
  p.f1_stride = z.f1_stride;
  p.raw = (double*)(ws + z.off[3]);
  p.edge_buf = (double*)(ws + z.off[4]);
  p.edge_cap = z.edge_cap;
  p.n_slots = z.n_slots;
  p.base_c = (double*)(ws + z.off[5]);
  p.base_n = (int*)(ws + z.off[6]);
  p.l_f0 = (double*)(ws + z.off[7]);
  p.l_sc = (double*)(ws + z.off[8]);
  p.l_slot = (unsigned char*)(ws + z.off[9]);
  p.l_keep = (unsigned char*)(ws + z.off[10]);
  p.l_n = (int*)(ws + z.off[11]);
  p.ctr = (double*)(ws + z.off[12]);
  p.ctr_stride = z.ctr_stride;
  p.status = (int*)(ws + z.off[13]);
  p.fft_nch = z.fft_nch;
  p.fft_blocks = z.fft_blocks;
  p.fft_groups = z.fft_groups;
  for (int g = 0; g <= WB_HV_FFT_GROUPS; ++g) p.fft_gc[g] = z.fft_gc[g];
  for (int g = 0; g < WB_HV_FFT_GROUPS; ++g) {
    p.fft_gV[g] = z.fft_gV[g];
    p.fft_gA[g] = z.fft_gA[g];
    p.fft_gblocks[g] = z.fft_gblocks[g];
    p.fft_goff[g] = z.fft_goff[g];
  }
  p.fft_H = t.fft_H;
  p.fft_Y = (wb_cplx*)(ws + z.off[16]);
  p.out_tpos = d_tpos;
  p.out_f0 = d_f0;
  p.out_vuv = d_vuv;
  p.out_n_frames = d_n_frames;
  p.f_stride = f_stride;

  if (stage_first <= 0 && wb_dev_memset(p.status, 0, 256, st)) return wb_fail(h, WB_E_CUDA, "wb_harvest: memset failed");
  if (stage_first <= 0 && 0 <= stage_last) {
    const long long chunks = (long long)batch * z.dec_chunks;
    wb_hv_dec_fwd k1;
    k1.p = p;
    WB_CHECK_LAUNCH(h, wb_launch_flat(k1, chunks, 64, st), "hv_dec_fwd");
    wb_hv_dec_scan k2;
    k2.p = p;
    k2.backward = 0;
    WB_CHECK_LAUNCH(h, wb_launch_flat(k2, batch, 32, st), "hv_dec_scan");
    wb_hv_dec_bwd k3;
    k3.p = p;
    WB_CHECK_LAUNCH(h, wb_launch_flat(k3, chunks, 64, st), "hv_dec_bwd");
    wb_hv_dec_scan k4;
    k4.p = p;
    k4.backward = 1;
    WB_CHECK_LAUNCH(h, wb_launch_flat(k4, batch, 32, st), "hv_dec_scan_b");
    wb_hv_dec_pick k5;
    k5.p = p;
    WB_CHECK_LAUNCH(h, wb_launch(k5, batch, 256, (WB_REDUCE_SCRATCH + 8) * sizeof(double), st), "hv_dec_pick");
  }
  if (stage_first <= 1 && 1 <= stage_last) {
    const int nthr = WB_HV_TILE / WB_HV_OPT;
    if (z.fft_nch > 0) {  // long filters: block spectra of the signal once, then one inverse FFT per (channel, block)
      wb_hv_fft_fwd kf;
      kf.p = p;
      kf.tw = h->tw;
      kf.tw_n = WB_TW_N;
      WB_CHECK_LAUNCH(h, wb_launch_spectral(kf, (long long)batch * z.fft_blocks, 256, wb_hv_fft_fwd::smem_bytes(), st),
                      "hv_fft_fwd");
      wb_hv_channels_fft kc;
      kc.p = p;
      kc.tw = h->tw;
      kc.tw_n = WB_TW_N;
      const long long items = (long long)batch * z.fft_nch;
      kc.p.n_slots = (int)(items < z.n_slots ? items : z.n_slots);
      WB_CHECK_LAUNCH(h, (wb_launch_b<wb_hv_channels_fft, WB_HV_TILE / WB_HV_OPT, 4>(kc, kc.p.n_slots, nthr, wb_hv_channels_fft::smem_bytes(nthr), st)),
                      "hv_channels_fft");
    }
    if (z.fft_nch < z.n_ch) {
      wb_hv_channels k;
      k.p = p;
      const long long items = (long long)batch * (z.n_ch - z.fft_nch);
      k.p.n_slots = (int)(items < z.n_slots ? items : z.n_slots);
      WB_CHECK_LAUNCH(h, (wb_launch_b<wb_hv_channels, WB_HV_TILE / WB_HV_OPT, 4>(k, k.p.n_slots, nthr, wb_hv_channels::smem_bytes(z.max_taps, nthr), st)),
                      "hv_channels");
    }
  }
  if (stage_first <= 2 && 2 <= stage_last) {
    wb_hv_detect k;
    k.p = p;
    WB_CHECK_LAUNCH(h, wb_launch_flat(k, (long long)batch * z.f1_stride, 128, st), "hv_detect");
  }
  if (stage_first <= 3 && 3 <= stage_last) {
    wb_hv_refine_prep k;
    k.p = p;
    k.tw = h->tw;
    k.tw_n = WB_TW_N;
    k.cls_count = (int*)(ws + z.off[14]);
    k.cls_cursor = k.cls_count + WB_HV_NCLS + 4;
    k.items = (unsigned long long*)(ws + z.off[15]);
    k.capacity = (long long)batch * z.f1_stride * WB_HV_SLOTS;
    k.frames_per_block = 256;
    if (wb_dev_memset(k.cls_count, 0, (2 * WB_HV_NCLS + 8) * sizeof(int), st)) return wb_fail(h, WB_E_CUDA, "memset");
    const long long frame_blocks = ((long long)batch * z.f1_stride + k.frames_per_block - 1) / k.frames_per_block;
    k.mode = 0;
    WB_CHECK_LAUNCH(h, (wb_launch_b<wb_hv_refine_prep, 256, 4>(k, frame_blocks, 256, wb_hv_refine_items::smem_bytes(), st)), "hv_refine_count");
    wb_hv_refine_scan ks;
    ks.cls_count = k.cls_count;
    ks.cls_cursor = k.cls_cursor;
    WB_CHECK_LAUNCH(h, wb_launch_flat(ks, 1, 32, st), "hv_refine_scan");
    k.mode = 1;
    WB_CHECK_LAUNCH(h, (wb_launch_b<wb_hv_refine_prep, 256, 4>(k, frame_blocks, 256, wb_hv_refine_items::smem_bytes(), st)), "hv_refine_scatter");
    k.mode = 2;
    // one persistent block per resident slot: WB_HV_REFINE_MINB blocks per SM (128 registers at 4)
    k.p.n_slots = wb_imax(1, hv_default_slots(h, batch, z.n_ch) / 4 * WB_HV_REFINE_MINB);
    WB_CHECK_LAUNCH(h, (wb_launch_b<wb_hv_refine_items, 128, WB_HV_REFINE_MINB>((const wb_hv_refine_items&)k, k.p.n_slots, 128, 0, st)),
                    "hv_refine");
  }
  if (stage_first <= 4 && 4 <= stage_last) {
    wb_hv_prune k;
    k.p = p;
    WB_CHECK_LAUNCH(h, wb_launch_flat(k, (long long)batch * z.f1_stride * WB_LANES, 128, st), "hv_prune");
  }
  if (stage_first <= 5 && 5 <= stage_last) {
    wb_hv_contour k;
    k.p = p;
    WB_CHECK_LAUNCH(h, wb_launch(k, batch, 8 * WB_LANES, (WB_REDUCE_SCRATCH + 8) * sizeof(double), st), "hv_contour");
  }
  return WB_OK;
}

}  // extern "C"
