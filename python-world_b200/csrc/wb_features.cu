// C-ABI: spectral feature heads, row / knot interpolation, PCM16 edge.
#include "wb_features.h"
#include "wb_handle.h"

static const double* wb_cos_table(wb_handle* h, int n) {
  return wb_table<double>(h, "cos" + std::to_string(n), [n](std::vector<double>& t) {
    t.resize(n);
    for (int m = 0; m < n; ++m)
      t[m] = (double)cosl(2.0L * 3.14159265358979323846264338327950288L * (long double)m / (long double)n);
  });
}

static int wb_ft_threads(int D) { return D >= 1024 ? 256 : 128; }

extern "C" {

int wb_lfbank(wb_handle* h, void* stream, const double* d_spec, int rows, int n_bins, const double* d_preemph_abs,
              const double* d_filterbank, int n_filt, double* d_out) {
  if (!h) return WB_E_INVALID;
  if (!d_spec || !d_preemph_abs || !d_filterbank || !d_out || rows < 0 || n_bins < 2 || n_filt < 1)
    return wb_fail(h, WB_E_INVALID, "wb_lfbank: null pointer or bad size");
  if (wb_ft_lfbank::smem_bytes(n_bins) > 200 * 1024) return wb_fail(h, WB_E_UNSUPPORTED, "wb_lfbank: %d bins", n_bins);
  WB_SET_DEVICE(h);
  wb_ft_lfbank k;
  k.spec = d_spec;
  k.habs = d_preemph_abs;
  k.fb = d_filterbank;
  k.rows = rows;
  k.D = n_bins;
  k.n_filt = n_filt;
  k.inv_nfft = 1.0 / (double)((n_bins - 1) * 2);
  k.out = d_out;
  WB_CHECK_LAUNCH(h, wb_launch(k, rows, wb_ft_threads(n_bins), wb_ft_lfbank::smem_bytes(n_bins), (wb_stream_t)stream),
                  "wb_lfbank");
  return WB_OK;
}

int wb_mcep(wb_handle* h, void* stream, const double* d_spec, int rows, int n_bins, const int* d_mel_bin, int n0,
            double* d_out) {
  if (!h) return WB_E_INVALID;
  if (!d_spec || !d_mel_bin || !d_out || rows < 0 || n_bins < 2 || n0 < 1)
    return wb_fail(h, WB_E_INVALID, "wb_mcep: null pointer or bad size");
  if (n0 > 2 * (n_bins - 1)) return wb_fail(h, WB_E_INVALID, "wb_mcep: n0 %d exceeds the transform length", n0);
  if (wb_ft_mcep::smem_bytes(n_bins) > 200 * 1024) return wb_fail(h, WB_E_UNSUPPORTED, "wb_mcep: %d bins", n_bins);
  WB_SET_DEVICE(h);
  const double* ct = wb_cos_table(h, 2 * (n_bins - 1));
  if (!ct) return wb_fail(h, WB_E_NOMEM, "wb_mcep: cosine table");
  wb_ft_mcep k;
  k.spec = d_spec;
  k.bin = d_mel_bin;
  k.ctab = ct;
  k.rows = rows;
  k.D = n_bins;
  k.n0 = n0;
  k.out = d_out;
  WB_CHECK_LAUNCH(h, wb_launch(k, rows, wb_ft_threads(n_bins), wb_ft_mcep::smem_bytes(n_bins), (wb_stream_t)stream),
                  "wb_mcep");
  return WB_OK;
}

int wb_mcep_decode(wb_handle* h, void* stream, const double* d_cepstrum, int rows, int n0, int fft_size,
                   const double* d_mel_pos, const int* d_bracket, const double* d_query, double* d_out) {
  if (!h) return WB_E_INVALID;
  if (!d_cepstrum || !d_mel_pos || !d_bracket || !d_query || !d_out || rows < 0 || n0 < 1 || fft_size < 4 ||
      (fft_size & 1))
    return wb_fail(h, WB_E_INVALID, "wb_mcep_decode: null pointer or bad size");
  if (2 * n0 > fft_size) return wb_fail(h, WB_E_INVALID, "wb_mcep_decode: n0 %d too large for fft_size %d", n0, fft_size);
  const int D = fft_size / 2 + 1;
  if (wb_ft_mcep_decode::smem_bytes(n0, D) > 200 * 1024)
    return wb_fail(h, WB_E_UNSUPPORTED, "wb_mcep_decode: fft_size %d", fft_size);
  WB_SET_DEVICE(h);
  const double* ct = wb_cos_table(h, fft_size);
  if (!ct) return wb_fail(h, WB_E_NOMEM, "wb_mcep_decode: cosine table");
  wb_ft_mcep_decode k;
  k.cep = d_cepstrum;
  k.ctab = ct;
  k.xp = d_mel_pos;
  k.jb = d_bracket;
  k.xq = d_query;
  k.rows = rows;
  k.n0 = n0;
  k.N = fft_size;
  k.Dout = D;
  k.out = d_out;
  WB_CHECK_LAUNCH(h, wb_launch(k, rows, wb_ft_threads(D), wb_ft_mcep_decode::smem_bytes(n0, D), (wb_stream_t)stream),
                  "wb_mcep_decode");
  return WB_OK;
}

int wb_interp_rows(wb_handle* h, void* stream, const double* d_in, int rows, int n_in, const double* d_knots,
                   const int* d_bracket, const double* d_query, int n_out, double* d_out) {
  if (!h) return WB_E_INVALID;
  if (!d_in || !d_knots || !d_bracket || !d_query || !d_out || rows < 0 || n_in < 1 || n_out < 1)
    return wb_fail(h, WB_E_INVALID, "wb_interp_rows: null pointer or bad size");
  if (d_in == d_out && n_in != n_out) return wb_fail(h, WB_E_INVALID, "wb_interp_rows: in-place needs n_in == n_out");
  if (wb_ft_interp_rows::smem_bytes(n_in) > 200 * 1024) return wb_fail(h, WB_E_UNSUPPORTED, "wb_interp_rows: %d columns", n_in);
  WB_SET_DEVICE(h);
  wb_ft_interp_rows k;
  k.in = d_in;
  k.xp = d_knots;
  k.jb = d_bracket;
  k.xq = d_query;
  k.rows = rows;
  k.Din = n_in;
  k.Dout = n_out;
  k.out = d_out;
  WB_CHECK_LAUNCH(h, wb_launch(k, rows, wb_ft_threads(n_in), wb_ft_interp_rows::smem_bytes(n_in), (wb_stream_t)stream),
                  "wb_interp_rows");
  return WB_OK;
}

int wb_interp_knots(wb_handle* h, void* stream, const double* d_x, long long n, const double* d_knot_x,
                    const double* d_knot_y, int n_knots, double* d_out) {
  if (!h) return WB_E_INVALID;
  if (!d_x || !d_knot_x || !d_knot_y || !d_out || n < 0 || n_knots < 1)
    return wb_fail(h, WB_E_INVALID, "wb_interp_knots: null pointer or bad size");
  WB_SET_DEVICE(h);
  wb_ft_interp_knots k;
  k.x = d_x;
  k.xp = d_knot_x;
  k.fp = d_knot_y;
  k.n_knots = n_knots;
  k.out = d_out;
  WB_CHECK_LAUNCH(h, wb_launch_flat(k, n, 256, (wb_stream_t)stream), "wb_interp_knots");
  return WB_OK;
}

int wb_pcm16_to_f64(wb_handle* h, void* stream, const int16_t* d_pcm, int pcm_stride, const int* d_n_samples, int batch,
                    double divisor, double* d_x, int x_stride) {
  if (!h) return WB_E_INVALID;
  if (!d_pcm || !d_n_samples || !d_x || batch < 0 || pcm_stride < 0 || x_stride < 0 || !(divisor != 0.0))
    return wb_fail(h, WB_E_INVALID, "wb_pcm16_to_f64: null pointer or bad size");
  WB_SET_DEVICE(h);
  wb_io_pcm16_in k;
  k.in = d_pcm;
  k.in_stride = pcm_stride;
  k.out_stride = x_stride;
  k.n_samples = d_n_samples;
  k.divisor = divisor;
  k.out = d_x;
  WB_CHECK_LAUNCH(h, wb_launch_flat(k, (long long)batch * x_stride, 256, (wb_stream_t)stream), "wb_pcm16_to_f64");
  return WB_OK;
}

int wb_f64_to_pcm16(wb_handle* h, void* stream, const double* d_y, int y_stride, const int* d_n_samples, int batch,
                    double gain, int16_t* d_pcm, int pcm_stride) {
  if (!h) return WB_E_INVALID;
  if (!d_y || !d_n_samples || !d_pcm || batch < 0 || pcm_stride < 0 || y_stride < 0)
    return wb_fail(h, WB_E_INVALID, "wb_f64_to_pcm16: null pointer or bad size");
  WB_SET_DEVICE(h);
  wb_io_pcm16_out k;
  k.in = d_y;
  k.in_stride = y_stride;
  k.out_stride = pcm_stride;
  k.n_samples = d_n_samples;
  k.gain = gain;
  k.out = d_pcm;
  WB_CHECK_LAUNCH(h, wb_launch_flat(k, (long long)batch * pcm_stride, 256, (wb_stream_t)stream), "wb_f64_to_pcm16");
  return WB_OK;
}

int wb_f64_to_f32(wb_handle* h, void* stream, const double* d_in, long long n, float* d_out) {
  if (!h) return WB_E_INVALID;
  if (!d_in || !d_out || n < 0) return wb_fail(h, WB_E_INVALID, "wb_f64_to_f32: null pointer or negative size");
  WB_SET_DEVICE(h);
  wb_io_f64_to_f32 k;
  k.in = d_in;
  k.out = d_out;
  k.n = n;
  WB_CHECK_LAUNCH(h, wb_launch_flat(k, (n + 1) / 2, 256, (wb_stream_t)stream), "wb_f64_to_f32");
  return WB_OK;
}

int wb_f32_to_f64(wb_handle* h, void* stream, const float* d_in, long long n, double* d_out) {
  if (!h) return WB_E_INVALID;
  if (!d_in || !d_out || n < 0) return wb_fail(h, WB_E_INVALID, "wb_f32_to_f64: null pointer or negative size");
  WB_SET_DEVICE(h);
  wb_io_f32_to_f64 k;
  k.in = d_in;
  k.out = d_out;
  k.n = n;
  WB_CHECK_LAUNCH(h, wb_launch_flat(k, (n + 1) / 2, 256, (wb_stream_t)stream), "wb_f32_to_f64");
  return WB_OK;
}

}  // extern "C"
