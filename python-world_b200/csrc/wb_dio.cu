// C-ABI: DIO F0 estimation (replaces world/dio.py:10 dio()) and StoneMask (world/stonemask.py:8).
#include "wb_dio.h"
#include "wb_handle.h"

void wb_hv_fill_zir(std::vector<double>& o, int kind);  // wb_harvest.cu

namespace {

// dio.py:359-436: a0, a1, a2, b0, b1 per decimation ratio 2..12
const double kDioDec[13][5] = {
    {0, 0, 0, 0, 0},
    {0, 0, 0, 0, 0},
    {0.041156734567757189, -0.42599112459189636, 0.041037215479961225, 0.16797464681802227, 0.50392394045406674},
    {0.95039378983237421, -0.67429146741526791, 0.15412211621346475, 0.071221945171178636, 0.21366583551353591},
    {1.4499664446880227, -0.98943497080950582, 0.24578252340690215, 0.036710750339322612, 0.11013225101796784},
    {1.7610939654280557, -1.2554914843859768, 0.3237186507788215, 0.021334858522387423, 0.06400457556716227},
    {1.9715352749512141, -1.4686795689225347, 0.3893908434965701, 0.013469181309343825, 0.040407543928031475},
    {2.1225239019534703, -1.6395144861046302, 0.44469707800587366, 0.0090366882681608418, 0.027110064804482525},
    {2.2357462340187593, -1.7780899984041358, 0.49152555365968692, 0.0063522763407111993, 0.019056829022133598},
    {2.3236003491759578, -1.8921545617463598, 0.53148928133729068, 0.0046331164041389372, 0.013899349212416812},
    {2.3936475118069387, -1.9873904075111861, 0.5658879979027055, 0.0034818622251927556, 0.010445586675578267},
    {2.450743295230728, -2.06794904601978, 0.59574774438332101, 0.0026822508007163792, 0.0080467524021491377},
    {2.4981398605924205, -2.1368928194784025, 0.62187513816221485, 0.0021097275904709001, 0.0063291827714127002},
};

inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

struct dio_sizes {
  int ratio, n_bands, max_taps, wrap_add;
  int ext_stride, y_stride, f_stride_w, edge_cap, n_slots, dec_chunks;
  size_t off[12];
  size_t total;
};

int dio_plan_sizes(int batch, int max_samples, int fs, double f0_floor, double f0_ceil, int channels_in_octave,
                   int target_fs, double frame_period, int n_slots, dio_sizes* z) {
  if (fs <= 0 || batch < 0 || max_samples < 0 || !(f0_floor > 0) || !(f0_ceil > f0_floor) || channels_in_octave <= 0 ||
      target_fs <= 0 || !(frame_period > 0))
    return WB_E_INVALID;
  z->ratio = (int)(fs / (double)target_fs);  // dio.py:38
  if (z->ratio < 2 || z->ratio > 12) return WB_E_UNSUPPORTED;  // the reference's coefficient table (dio.py:365-436)
  z->n_bands = (int)std::ceil(std::log2(f0_ceil / f0_floor) * channels_in_octave);  // dio.py:32
  if (z->n_bands < 1 || z->n_bands > WB_DIO_MAXB) return WB_E_UNSUPPORTED;
  const double afs = target_fs;
  const int cc = (int)(afs / 50 + 0.5);
  const double e0 = f0_floor * std::pow(2.0, 1.0 / channels_in_octave);
  z->max_taps = 2 * cc + 1 + 4 * ((int)(afs / e0 / 2 + 0.5) + 1) + 8;
  z->wrap_add = (int)(afs / f0_floor / 2 + 0.5) * 4;  // dio.py:78
  z->ext_stride = max_samples + 18 + 2;
  z->dec_chunks = (z->ext_stride + WB_HV_CHUNK - 1) / WB_HV_CHUNK + 1;
  z->y_stride = (max_samples + 9) / z->ratio + 4;
  z->f_stride_w = wb_hv_frames(max_samples, fs, frame_period) + 1;
  z->edge_cap = z->y_stride / 2 + 4;
  z->n_slots = n_slots;
  const size_t B = (size_t)batch, F = (size_t)z->f_stride_w, NB = (size_t)z->n_bands;
  size_t o = 0;
  int i = 0;
  auto put = [&](size_t bytes) {
    z->off[i++] = o;
    o += align_up(bytes);
  };
  put(2 * B * z->ext_stride * sizeof(double) + 4 * B * (size_t)z->dec_chunks * 3 * sizeof(double));  // 0 decimator
  put(B * z->y_stride * sizeof(double));                    // 1 y
  put(B * sizeof(int));                                     // 2 y_len
  put(B * NB * F * sizeof(double));                         // 3 raw
  put(B * NB * F * sizeof(double));                         // 4 stab
  put(B * NB * 4 * F * sizeof(double));                     // 5 four
  put((size_t)n_slots * 4 * z->edge_cap * sizeof(double));  // 6 edge_buf
  put(B * F * NB * sizeof(double));                         // 7 sorted candidates
  put(B * 4 * F * sizeof(double));                          // 8 step buffers
  put(B * 4 * F * sizeof(int));                             // 9 sections
  put(256);                                                 // 10 status
  z->total = o;
  return WB_OK;
}

int dio_slots(wb_handle* h, int batch, int n_bands) {
#ifdef WB_HOST_EMU
  (void)h;
  long long items = (long long)batch * n_bands;
  return (int)(items < 2 ? (items < 1 ? 1 : items) : 2);
#else
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
  long long items = (long long)batch * n_bands;
  long long s = 4LL * sms;
  if (s > items) s = items;
  return (int)(s < 1 ? 1 : s);
#endif
}

}  // namespace

extern "C" {

int wb_dio_workspace_bytes(wb_handle* h, int batch, int max_samples, int fs, double f0_floor, double f0_ceil,
                           int channels_in_octave, int target_fs, double frame_period_ms, size_t* bytes) {
  if (!h || !bytes) return WB_E_INVALID;
  dio_sizes z;
  const int nb = (f0_floor > 0 && f0_ceil > f0_floor) ? (int)std::ceil(std::log2(f0_ceil / f0_floor) * channels_in_octave) : 1;
  int rc = dio_plan_sizes(batch, max_samples, fs, f0_floor, f0_ceil, channels_in_octave, target_fs, frame_period_ms,
                          dio_slots(h, batch, nb), &z);
  if (rc) return wb_fail(h, rc, "wb_dio_workspace_bytes: unsupported configuration (fs=%d target_fs=%d)", fs, target_fs);
  *bytes = z.total;
  return WB_OK;
}

int wb_dio(wb_handle* h, void* stream, const double* d_x, int x_stride, const int* d_n_samples, int batch,
           int max_samples, int fs, double f0_floor, double f0_ceil, int channels_in_octave, int target_fs,
           double frame_period_ms, double allowed_range, void* d_workspace, size_t workspace_bytes, int f_stride,
           double* d_tpos, double* d_f0, double* d_vuv, int* d_n_frames, double* d_f0_candidates,
           double* d_raw_f0_candidates) {
  if (!h) return WB_E_INVALID;
  if (!d_x || !d_n_samples || !d_workspace || !d_tpos || !d_f0 || !d_vuv || !d_n_frames || batch < 0 ||
      max_samples > x_stride)
    return wb_fail(h, WB_E_INVALID, "wb_dio: null pointer or inconsistent sizes");
  if (batch == 0) return WB_OK;
  dio_sizes z;
  const int nb0 = (f0_floor > 0 && f0_ceil > f0_floor) ? (int)std::ceil(std::log2(f0_ceil / f0_floor) * channels_in_octave) : 1;
  int rc = dio_plan_sizes(batch, max_samples, fs, f0_floor, f0_ceil, channels_in_octave, target_fs, frame_period_ms,
                          dio_slots(h, batch, nb0), &z);
  if (rc)
    return wb_fail(h, rc, "wb_dio: unsupported configuration (fs=%d target_fs=%d: decimation ratio must be 2..12)", fs,
                   target_fs);
  if (workspace_bytes < z.total) return wb_fail(h, WB_E_INVALID, "wb_dio: workspace %zu < %zu bytes", workspace_bytes, z.total);
  if (f_stride < wb_hv_frames(max_samples, fs, frame_period_ms))
    return wb_fail(h, WB_E_INVALID, "wb_dio: f_stride %d too small", f_stride);
  WB_SET_DEVICE(h);
  const double afs = target_fs;  // dio.py:39: taken as exact whatever fs / ratio really is
  const int nb = z.n_bands;
  char key[200];
  snprintf(key, sizeof key, "dio:%d:%.17g:%.17g:%d:%d", z.ratio, f0_floor, f0_ceil, channels_in_octave, target_fs);
  const std::string k(key);
  // ---- band tables: combined low-cut (dio.py:80-85) * Nuttall low-pass (dio.py:129-131) taps, reversed
  std::vector<double> edges(nb);
  std::vector<int> lens(nb), first(nb), offs(nb);
  std::vector<std::vector<double>> gtaps(nb);
  {
    const int cc = (int)(afs / 50 + 0.5);
    std::vector<double> hwin(2 * cc + 1);
    double hs = 0.0;
    for (int i = 0; i < 2 * cc + 1; ++i) {  // scipy hann(2c+3)[1:-1]
      hwin[i] = 0.5 - 0.5 * std::cos(2.0 * WB_PI * (double)(i + 1) / (double)(2 * cc + 2));
      hs += hwin[i];
    }
    for (int i = 0; i < 2 * cc + 1; ++i) hwin[i] = -hwin[i] / hs;
    hwin[cc] += 1.0;
    int total = 0;
    std::vector<double> lpf;
    for (int b = 0; b < nb; ++b) {
      edges[b] = f0_floor * std::pow(2.0, (double)(b + 1) / channels_in_octave);  // dio.py:32-34
      const int half = (int)(afs / edges[b] / 2 + 0.5);
      if (half < 1) return wb_fail(h, WB_E_UNSUPPORTED, "wb_dio: band %d has an empty filter", b);
      wb_nuttall(4 * half, lpf);
      int bias = 0;
      for (int i = 1; i < 4 * half; ++i)
        if (lpf[i] > lpf[bias]) bias = i;  // numpy argmax: first maximum (dio.py:131)
      // g[m], m = -cc .. cc + 4 half - 1 : linear convolution of the centred low-cut with the low-pass
      const int L = 2 * cc + 4 * half;
      std::vector<double> g(L, 0.0);
      for (int a = 0; a < 2 * cc + 1; ++a)
        for (int q = 0; q < 4 * half; ++q) g[a + q] += hwin[a] * lpf[q];
      // signal index for output n and tap m: n + bias + 1 - m ; reversed taps start at m_max = cc + 4 half - 1
      gtaps[b].resize(L);
      for (int i = 0; i < L; ++i) gtaps[b][i] = g[L - 1 - i];
      lens[b] = L;
      first[b] = bias + 1 - (cc + 4 * half - 1);
      offs[b] = total;
      total += L;
    }
  }

  wb_hv_plan p;
  memset(&p, 0, sizeof p);
  p.edges = wb_table<double>(h, k + ":edges", [&](std::vector<double>& o) { o = edges; });
  p.halfs = wb_table<int>(h, k + ":lens", [&](std::vector<int>& o) { o = lens; });
  p.ch_off = wb_table<int>(h, k + ":first", [&](std::vector<int>& o) { o = first; });
  p.tap_off = wb_table<int>(h, k + ":offs", [&](std::vector<int>& o) { o = offs; });
  p.taps = wb_table<double>(h, k + ":taps", [&](std::vector<double>& o) {
    for (int b = 0; b < nb; ++b) o.insert(o.end(), gtaps[b].begin(), gtaps[b].end());
  });
  p.cb = wb_table<double>(h, k + ":dec", [&](std::vector<double>& o) {
    o.assign(11 + 3 * WB_HV_CHUNK + 9, 0.0);
    o[0] = kDioDec[z.ratio][3];
    o[1] = kDioDec[z.ratio][4];
    o[5] = kDioDec[z.ratio][0];
    o[6] = kDioDec[z.ratio][1];
    o[7] = kDioDec[z.ratio][2];
    wb_hv_fill_zir(o, 1);
  });
  p.pow2_quirk = wb_table<int>(h, "pow2quirk", [](std::vector<int>& o) {
    o.resize(32);
    for (int q = 0; q < 32; ++q) o[q] = q == 0 ? 0 : (int)std::ceil(std::log(std::ldexp(1.0, q)) / std::log(2.0));
  });
  if (!p.edges || !p.halfs || !p.ch_off || !p.tap_off || !p.taps || !p.cb || !p.pow2_quirk)
    return wb_fail(h, WB_E_NOMEM, "wb_dio: table allocation failed");
  char* ws = (char*)d_workspace;
  wb_stream_t st = (wb_stream_t)stream;
  p.batch = batch;
  p.fs = fs;
  p.ratio = z.ratio;
  p.pad = 0;
  p.afs = afs;
  p.f0_floor = f0_floor;
  p.f0_ceil = f0_ceil;
  p.frame_period = frame_period_ms;
  p.n_ch = nb;
  p.max_taps = z.max_taps;
  p.wrap_n = z.wrap_add;
  p.dec_kind = 1;
  p.mode = 1;
  p.grid_ms = frame_period_ms;
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
  p.f1_stride = z.f_stride_w;
  p.raw = (double*)(ws + z.off[3]);
  p.stab = (double*)(ws + z.off[4]);
  p.four = (double*)(ws + z.off[5]);
  p.edge_buf = (double*)(ws + z.off[6]);
  p.edge_cap = z.edge_cap;
  p.n_slots = z.n_slots;
  p.status = (int*)(ws + z.off[10]);
  p.out_tpos = d_tpos;
  p.out_f0 = d_f0;
  p.out_vuv = d_vuv;
  p.out_n_frames = d_n_frames;
  p.f_stride = f_stride;

  if (wb_dev_memset(p.status, 0, 256, st)) return wb_fail(h, WB_E_CUDA, "wb_dio: memset failed");
  {
    const long long chunks = (long long)batch * z.dec_chunks;
    wb_hv_dec_fwd k1;
    k1.p = p;
    WB_CHECK_LAUNCH(h, wb_launch_flat(k1, chunks, 64, st), "dio_dec_fwd");
    wb_hv_dec_scan k2;
    k2.p = p;
    k2.backward = 0;
    WB_CHECK_LAUNCH(h, wb_launch_flat(k2, batch, 32, st), "dio_dec_scan");
    wb_hv_dec_bwd k3;
    k3.p = p;
    WB_CHECK_LAUNCH(h, wb_launch_flat(k3, chunks, 64, st), "dio_dec_bwd");
    wb_hv_dec_scan k4;
    k4.p = p;
    k4.backward = 1;
    WB_CHECK_LAUNCH(h, wb_launch_flat(k4, batch, 32, st), "dio_dec_scan_b");
    wb_hv_dec_pick k5;
    k5.p = p;
    WB_CHECK_LAUNCH(h, wb_launch(k5, batch, 256, (WB_REDUCE_SCRATCH + 8) * sizeof(double), st), "dio_dec_pick");
  }
  {
    wb_hv_channels kc;
    kc.p = p;
    const int nthr = WB_HV_TILE / WB_HV_OPT;
    WB_CHECK_LAUNCH(h, (wb_launch_b<wb_hv_channels, WB_HV_TILE / WB_HV_OPT, 4>(kc, z.n_slots, nthr, wb_hv_channels::smem_bytes(z.max_taps, nthr), st)),
                    "dio_channels");
  }
  {
    wb_dio_contour kf;
    kf.p = p;
    kf.n_bands = nb;
    kf.allowed_range = allowed_range;
    kf.cand = (double*)(ws + z.off[7]);
    kf.work = (double*)(ws + z.off[8]);
    kf.sect = (int*)(ws + z.off[9]);
    kf.out_cand = d_f0_candidates;
    WB_CHECK_LAUNCH(h, wb_launch_flat(kf, batch, 32, st), "dio_contour");
  }
  if (d_raw_f0_candidates) {  // [B, n_bands, f_stride]  <- raw [B, n_bands, f1_stride]
    for (int u = 0; u < batch; ++u)
      for (int b = 0; b < nb; ++b) {
        const double* src = p.raw + ((size_t)u * nb + b) * p.f1_stride;
        double* dst = d_raw_f0_candidates + ((size_t)u * nb + b) * f_stride;
        const size_t n = (size_t)(f_stride < p.f1_stride ? f_stride : p.f1_stride) * sizeof(double);
#ifdef WB_HOST_EMU
        std::memcpy(dst, src, n);
#else
        if (cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
          return wb_fail(h, WB_E_CUDA, "wb_dio: copy of raw candidates failed");
#endif
      }
  }
  return WB_OK;
}

int wb_dio_band_count(double f0_floor, double f0_ceil, int channels_in_octave) {
  if (!(f0_floor > 0) || !(f0_ceil > f0_floor) || channels_in_octave <= 0) return 0;
  return (int)std::ceil(std::log2(f0_ceil / f0_floor) * channels_in_octave);
}

int wb_stonemask(wb_handle* h, void* stream, const double* d_x, int x_stride, const int* d_n_samples, int batch, int fs,
                 const double* d_tpos, const double* d_f0, const int* d_n_frames, int f_stride, double* d_refined_f0) {
  if (!h) return WB_E_INVALID;
  if (!d_x || !d_n_samples || !d_tpos || !d_f0 || !d_n_frames || !d_refined_f0 || batch < 0 || f_stride < 0 || fs <= 0)
    return wb_fail(h, WB_E_INVALID, "wb_stonemask: null pointer or negative size");
  WB_SET_DEVICE(h);
  const int lut_half = (int)std::ceil(3.0 * fs / 40.0 / 2.0);  // F0 >= 40 Hz
  const double* lut = wb_table<double>(h, "sm_lut:" + std::to_string(fs), [&](std::vector<double>& o) {
    o.resize(2 * lut_half + 1);
    char buf[64];
    for (int kq = -lut_half; kq <= lut_half; ++kq) {  // float("{0:.4f}".format(k / fs)), stonemask.py:38
      snprintf(buf, sizeof buf, "%.4f", (double)kq / fs);
      o[kq + lut_half] = strtod(buf, nullptr);
    }
  });
  if (!lut) return wb_fail(h, WB_E_NOMEM, "wb_stonemask: table allocation failed");
  int lg = 0;
  while ((1 << lg) < 2 * lut_half + 1) ++lg;
  if ((1 << (lg + 1)) > WB_TW_N) return wb_fail(h, WB_E_UNSUPPORTED, "wb_stonemask: fs=%d too high", fs);
  wb_stonemask_body k;
  k.x = d_x;
  k.n_samples = d_n_samples;
  k.tpos = d_tpos;
  k.f0 = d_f0;
  k.n_frames = d_n_frames;
  k.time_lut = lut;
  k.tw = h->tw;
  k.tw_n = WB_TW_N;
  k.lut_half = lut_half;
  k.x_stride = x_stride;
  k.f_stride = f_stride;
  k.fs = fs;
  k.out = d_refined_f0;
  const int nthr = 128;
  WB_CHECK_LAUNCH(h, wb_launch(k, (long long)batch * f_stride, nthr, wb_stonemask_body::smem_bytes(lut_half, nthr),
                               (wb_stream_t)stream),
                  "wb_stonemask");
  return WB_OK;
}

/* Diagnostic: the Nuttall window exactly as the library tabulates it (tests compare it bit for bit). */
int wb_debug_nuttall(int n, double* host_out) {
  if (n < 2 || !host_out) return WB_E_INVALID;
  std::vector<double> w;
  wb_nuttall(n, w);
  for (int i = 0; i < n; ++i) host_out[i] = w[i];
  return WB_OK;
}

}  // extern "C"
