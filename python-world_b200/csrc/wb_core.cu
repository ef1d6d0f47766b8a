// C-ABI: handle management, size helpers, CheapTrick, D4C, D4C-Requiem.
#include "wb_cheaptrick.h"
#include "wb_d4c.h"
#include "wb_handle.h"

extern "C" {

int wb_is_cuda_build(void) {
#ifdef WB_HOST_EMU
  return 0;
#else
  return 1;
#endif
}

const char* wb_version(void) { return "world_b200 0.1 (sm_100a)"; }

int wb_create(wb_handle** out, int device) {
  if (!out) return WB_E_INVALID;
  wb_handle* h = new (std::nothrow) wb_handle();
  if (!h) return WB_E_NOMEM;
  h->device = device;
  h->tw = nullptr;
#ifndef WB_HOST_EMU
  wb_device_guard guard(device);  // the caller's current device is restored on return
  if (!guard.ok) {
    delete h;
    return WB_E_CUDA;
  }
#endif
  const wb_cplx* tw = wb_table<wb_cplx>(h, "twiddle", [](std::vector<wb_cplx>& t) {
    t.resize(WB_TW_N);
    for (int m = 0; m < WB_TW_N; ++m) {
      const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)m / (long double)WB_TW_N;
      t[m].x = (double)cosl(a);
      t[m].y = (double)-sinl(a);
    }
  });
  if (!tw) {
    delete h;
    return WB_E_NOMEM;
  }
  h->tw = const_cast<wb_cplx*>(tw);
  *out = h;
  return WB_OK;
}

int wb_destroy(wb_handle* h) {
  if (!h) return WB_E_INVALID;
  for (auto& kv : h->tables) wb_dev_free(kv.second);
  delete h;
  return WB_OK;
}

const char* wb_last_error(const wb_handle* h) { return h ? h->err.c_str() : "null handle"; }

int wb_frame_count(int n_samples, int fs, double frame_period_ms) {
  return (int)(1000.0 * n_samples / fs / frame_period_ms + 1);
}

int wb_cheaptrick_fft_size(int fs) { return wb_pow2_ceil_log2(3.0 * fs / 71 + 1); }

int wb_d4c_band_count(int fs, int requiem) {
  int interval = 3000;
  if (!requiem && fs < 16000) interval = 2000;
  const double a = fs / 2.0 - interval;
  return (int)std::floor((a < 15000.0 ? a : 15000.0) / interval);
}

int wb_synthesis_length(double t0, double t_end, int fs) {
  // numpy.arange(start, stop, step) has ceil((stop - start)/step) elements
  const double step = 1.0 / fs;
  const double stop = t_end + 1.0 / fs;
  const double len = std::ceil((stop - t0) / step);
  return len > 0 ? (int)len : 0;
}

int wb_cheaptrick(wb_handle* h, void* stream, const double* d_x, int x_stride, const int* d_n_samples, int batch,
                  int fs, const double* d_tpos, const double* d_f0, const double* d_vuv, const int* d_n_frames,
                  int f_stride, double q1, int fft_size, const double* d_dither, uint64_t seed, double* d_f0_used,
                  double* d_spec, void* d_ps) {
  if (!h) return WB_E_INVALID;
  if (!d_x || !d_n_samples || !d_tpos || !d_f0 || !d_vuv || !d_n_frames || !d_f0_used || !d_spec || batch < 0 ||
      f_stride < 0 || fs <= 0)
    return wb_fail(h, WB_E_INVALID, "wb_cheaptrick: null pointer or negative size");
  const int n = fft_size > 0 ? fft_size : wb_cheaptrick_fft_size(fs);
  if (!wb_is_pow2(n) || n < 16 || n > WB_TW_N) return wb_fail(h, WB_E_UNSUPPORTED, "wb_cheaptrick: fft_size %d", n);
  WB_SET_DEVICE(h);
  wb_cheaptrick_params k;
  k.x = d_x;
  k.n_samples = d_n_samples;
  k.tpos = d_tpos;
  k.f0 = d_f0;
  k.vuv = d_vuv;
  k.n_frames = d_n_frames;
  k.dither = d_dither;
  k.tw = h->tw;
  k.tw_n = WB_TW_N;
  k.x_stride = x_stride;
  k.f_stride = f_stride;
  k.fs = fs;
  k.n = n;
  k.q1 = q1;
  k.seed = seed;
  k.f0_used = d_f0_used;
  k.spec = d_spec;
  k.ps = (wb_cplx*)d_ps;
  int nthr = n / 8;  // the half-size complex FFT has n/8 radix-4 butterflies per pass
  if (const char* e = std::getenv("WB_CT_THREADS")) nthr = std::atoi(e);  // tuning knob
  nthr = nthr < 128 ? 128 : (nthr > 512 ? 512 : nthr);
  const long long grid = (long long)batch * f_stride;
  const size_t smem = wb_cheaptrick_params::smem_bytes(n, nthr);
#ifndef WB_HOST_EMU
  if (n == 1024 && nthr == 128) {  // 16 / 22.05 kHz
    wb_cheaptrick_body_t<1024, 128> b;
    static_cast<wb_cheaptrick_params&>(b) = k;
    WB_CHECK_LAUNCH(h, (wb_launch_b<wb_cheaptrick_body_t<1024, 128>, 128, 6>(b, grid, 128, smem, (wb_stream_t)stream)), "wb_cheaptrick");
    return WB_OK;
  }
  if (n == 2048 && nthr == 256) {  // 44.1 / 48 kHz
    wb_cheaptrick_body_t<2048, 256> b;
    static_cast<wb_cheaptrick_params&>(b) = k;
    WB_CHECK_LAUNCH(h, (wb_launch_b<wb_cheaptrick_body_t<2048, 256>, 256, 4>(b, grid, 256, smem, (wb_stream_t)stream)), "wb_cheaptrick");
    return WB_OK;
  }
#endif
  wb_cheaptrick_body b;
  static_cast<wb_cheaptrick_params&>(b) = k;
  WB_CHECK_LAUNCH(h, wb_launch_spectral(b, grid, nthr, smem, (wb_stream_t)stream), "wb_cheaptrick");
  return WB_OK;
}

static int wb_d4c_common(wb_handle* h, void* stream, const double* d_x, int x_stride, const int* d_n_samples,
                         int batch, int fs, const double* d_tpos, const double* d_f0, const double* d_vuv,
                         const int* d_n_frames, int f_stride, double threshold, int n, int n_spec, int interval,
                         int requiem, double* d_f0_out, double* d_ap, double* d_coarse) {
  if (!d_x || !d_n_samples || !d_tpos || !d_f0 || !d_vuv || !d_n_frames || !d_f0_out || (!d_ap && !d_coarse) ||
      batch < 0 || f_stride < 0 || fs <= 0)
    return wb_fail(h, WB_E_INVALID, "wb_d4c: null pointer or negative size");
  const int n_bands = (int)std::floor(std::fmin(15000.0, fs / 2.0 - interval) / interval);
  if (n_bands <= 0)  // the reference asserts (d4c.py:35, d4cRequiem.py:21)
    return wb_fail(h, WB_E_INVALID, "wb_d4c: number_of_aperiodicity <= 0 for fs=%d", fs);
  if (n_bands > 16) return wb_fail(h, WB_E_UNSUPPORTED, "wb_d4c: %d bands", n_bands);
  const int n_love = wb_pow2_ceil_log2(3.0 * fs / 40 + 1);
  if (!wb_is_pow2(n) || n < 64 || n > WB_TW_N || n_love > WB_TW_N)
    return wb_fail(h, WB_E_UNSUPPORTED, "wb_d4c: fft sizes %d / %d", n, n_love);
  const int wlen = (int)(std::floor(interval / ((double)fs / n)) * 2 + 1);
  if (wlen > n || wlen < 3) return wb_fail(h, WB_E_UNSUPPORTED, "wb_d4c: band window %d vs fft %d", wlen, n);
  WB_SET_DEVICE(h);
  const double* win = wb_table<double>(h, "nuttall" + std::to_string(wlen),
                                       [wlen](std::vector<double>& w) { wb_nuttall(wlen, w); });
  if (!win) return wb_fail(h, WB_E_NOMEM, "wb_d4c: window table");
  wb_d4c_params k;
  k.x = d_x;
  k.n_samples = d_n_samples;
  k.tpos = d_tpos;
  k.f0 = d_f0;
  k.vuv = d_vuv;
  k.n_frames = d_n_frames;
  k.band_win = win;
  k.tw = h->tw;
  k.tw_n = WB_TW_N;
  k.x_stride = x_stride;
  k.f_stride = f_stride;
  k.fs = fs;
  k.n = n;
  k.n_love = n_love;
  k.nm = wb_d4c_params::buffer_capacity(fs, n, n_love);
  k.n_spec = n_spec;
  k.interval = interval;
  k.n_bands = n_bands;
  k.band_wlen = wlen;
  k.requiem = requiem;
  k.threshold = threshold;
  k.f0_out = d_f0_out;
  k.ap = d_ap;
  k.coarse = d_coarse;
  const size_t smem = wb_d4c_params::smem_bytes_tw(k.nm, n, n_love);
  if (smem > 227 * 1024) return wb_fail(h, WB_E_UNSUPPORTED, "wb_d4c: %zu bytes of shared memory", smem);
  int nthr = (n > n_love ? n : n_love) / 8;
  if (const char* e = std::getenv("WB_D4C_THREADS")) nthr = std::atoi(e);  // tuning knob
  nthr = nthr < 128 ? 128 : (nthr > 512 ? 512 : nthr);
  const long long grid = (long long)batch * f_stride;
  const bool four = smem + 1024 <= (size_t)227 * 1024 / 4;
#ifndef WB_HOST_EMU
  // the shapes of the 16 / 22.05 kHz configurations run with compile-time sizes (four blocks of 256 threads per SM)
  if (four && nthr == 256 && n == 2048 && n_love == 2048) {
    wb_d4c_body_t<2048, 2048, 256> b;
    static_cast<wb_d4c_params&>(b) = k;
    WB_CHECK_LAUNCH(h, (wb_launch_b<wb_d4c_body_t<2048, 2048, 256>, 256, 4>(b, grid, 256, smem, (wb_stream_t)stream)), "wb_d4c");
    return WB_OK;
  }
  if (four && nthr == 256 && n == 1024 && n_love == 2048) {
    wb_d4c_body_t<1024, 2048, 256> b;
    static_cast<wb_d4c_params&>(b) = k;
    WB_CHECK_LAUNCH(h, (wb_launch_b<wb_d4c_body_t<1024, 2048, 256>, 256, 4>(b, grid, 256, smem, (wb_stream_t)stream)), "wb_d4c");
    return WB_OK;
  }
#endif
  wb_d4c_body b;
  static_cast<wb_d4c_params&>(b) = k;
  // four blocks of 256 threads per SM when the frame fits 56 KB (64 registers), three (80 registers) otherwise
  if (four)
    WB_CHECK_LAUNCH(h, wb_launch_spectral(b, grid, nthr, smem, (wb_stream_t)stream), "wb_d4c");
  else
    WB_CHECK_LAUNCH(h, wb_launch_spectral3(b, grid, nthr, smem, (wb_stream_t)stream), "wb_d4c");
  return WB_OK;
}

int wb_d4c(wb_handle* h, void* stream, const double* d_x, int x_stride, const int* d_n_samples, int batch, int fs,
           const double* d_tpos, const double* d_f0, const double* d_vuv, const int* d_n_frames, int f_stride,
           double threshold, int fft_size_for_spectrum, double* d_f0_out, double* d_ap, double* d_coarse_ap) {
  if (!h) return WB_E_INVALID;
  const int n = wb_pow2_ceil_log2(4.0 * fs / 47 + 1);                                        // d4c.py:19-20
  const int n_spec = fft_size_for_spectrum > 0 ? fft_size_for_spectrum : wb_cheaptrick_fft_size(fs);  // d4c.py:21-23
  const int interval = fs < 16000 ? 2000 : 3000;                                             // d4c.py:24-27
  return wb_d4c_common(h, stream, d_x, x_stride, d_n_samples, batch, fs, d_tpos, d_f0, d_vuv, d_n_frames, f_stride,
                       threshold, n, n_spec, interval, 0, d_f0_out, d_ap, d_coarse_ap);
}

int wb_d4c_requiem(wb_handle* h, void* stream, const double* d_x, int x_stride, const int* d_n_samples, int batch,
                   int fs, const double* d_tpos, const double* d_f0, const double* d_vuv, const int* d_n_frames,
                   int f_stride, double threshold, int fft_size, double* d_f0_out, double* d_band_ap) {
  if (!h) return WB_E_INVALID;
  const int n = fft_size > 0 ? fft_size : wb_pow2_ceil_log2(3.0 * fs / 47 + 1);  // d4cRequiem.py:10-12
  return wb_d4c_common(h, stream, d_x, x_stride, d_n_samples, batch, fs, d_tpos, d_f0, d_vuv, d_n_frames, f_stride,
                       threshold, n, 0, 3000, 1, d_f0_out, d_band_ap, nullptr);
}

}  // extern "C"
