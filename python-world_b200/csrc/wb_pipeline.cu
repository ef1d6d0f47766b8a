// C-ABI: the fused entry points -- World.encode (main.py:106-152) and World.decode (main.py:198-214) for a
// batch, each as a workspace query plus one call that enqueues every stage kernel on the caller's stream.
// A binder that does not want to re-implement the stage sequencing (world_b200/engine.py) needs only these
// two; they compose the stage entry points of this library and add no arithmetic of their own.  Also here:
// the 'coarse_ap' -> aperiodicity expansion used by the compact transport format.
#include "wb_d4c.h"
#include "wb_handle.h"

namespace {
inline size_t align_up256(size_t v) { return (v + 255) & ~(size_t)255; }

struct enc_layout {
  size_t tracker, f0_raw, f0_ref, f0_used, total;
};

// the F0 floor follows the spectrum FFT size when the caller fixes it (main.py:123-124)
inline double enc_floor(const wb_encode_params* q) { return q->fft_size > 0 ? 3.0 * q->fs / q->fft_size : q->f0_floor; }

int enc_plan(wb_handle* h, const wb_encode_params* q, int batch, int max_samples, enc_layout* L, int* f_count) {
  size_t t = 0;
  int rc;
  if (q->f0_method == WB_F0_HARVEST)
    rc = wb_harvest_workspace_bytes(h, batch, max_samples, q->fs, enc_floor(q), q->f0_ceil, &t);
  else if (q->f0_method == WB_F0_DIO)
    rc = wb_dio_workspace_bytes(h, batch, max_samples, q->fs, enc_floor(q), q->f0_ceil, q->channels_in_octave,
                                q->target_fs, q->frame_period_ms, &t);
  else
    return wb_fail(h, WB_E_INVALID, "wb_encode: unknown f0_method %d", q->f0_method);  // main.py:136-137
  if (rc != WB_OK) return rc;
  const int F = wb_frame_count(max_samples, q->fs, q->frame_period_ms);
  const size_t vec = align_up256((size_t)batch * (size_t)(F > 0 ? F : 1) * sizeof(double));
  L->tracker = align_up256(t);
  L->f0_raw = L->tracker;
  L->f0_ref = L->f0_raw + vec;
  L->f0_used = L->f0_ref + vec;
  L->total = L->f0_used + vec;
  *f_count = F;
  return WB_OK;
}
}  // namespace

extern "C" {

int wb_encode_workspace_bytes(wb_handle* h, const wb_encode_params* q, int batch, int max_samples, size_t* bytes) {
  if (!h || !q || !bytes || batch < 0 || max_samples < 0) return WB_E_INVALID;
  enc_layout L;
  int F;
  const int rc = enc_plan(h, q, batch, max_samples, &L, &F);
  if (rc != WB_OK) return rc;
  *bytes = L.total;
  return WB_OK;
}

int wb_encode(wb_handle* h, void* stream, const wb_encode_params* q, const double* d_x, int x_stride,
              const int* d_n_samples, int batch, int max_samples, void* d_workspace, size_t workspace_bytes,
              int f_stride, const double* d_dither, double* d_tpos, double* d_f0, double* d_vuv, int* d_n_frames,
              double* d_spectrogram, double* d_aperiodicity, double* d_coarse_ap, void* d_ps) {
  if (!h) return WB_E_INVALID;
  if (!q || !d_x || !d_n_samples || !d_workspace || !d_tpos || !d_f0 || !d_vuv || !d_n_frames || !d_spectrogram)
    return wb_fail(h, WB_E_INVALID, "wb_encode: null pointer");
  enc_layout L;
  int F;
  int rc = enc_plan(h, q, batch, max_samples, &L, &F);
  if (rc != WB_OK) return rc;
  if (workspace_bytes < L.total) return wb_fail(h, WB_E_INVALID, "wb_encode: workspace %zu < %zu", workspace_bytes, L.total);
  if (f_stride != F) return wb_fail(h, WB_E_INVALID, "wb_encode: f_stride %d != wb_frame_count(max_samples) = %d", f_stride, F);
  if (batch == 0) return WB_OK;
  char* ws = (char*)d_workspace;
  double* f0_raw = (double*)(ws + L.f0_raw);   // the tracker's contour
  double* f0_ref = (double*)(ws + L.f0_ref);   // after StoneMask (dio only)
  double* f0_used = (double*)(ws + L.f0_used); // what CheapTrick leaves in source['f0'] (cheaptrick.py:27,33)
  const double floor = enc_floor(q);
  if (q->f0_method == WB_F0_HARVEST) {
    rc = wb_harvest(h, stream, d_x, x_stride, d_n_samples, batch, max_samples, q->fs, floor, q->f0_ceil,
                    q->frame_period_ms, ws, L.tracker, f_stride, d_tpos, f0_raw, d_vuv, d_n_frames);
    if (rc != WB_OK) return rc;
    f0_ref = f0_raw;
  } else {
    rc = wb_dio(h, stream, d_x, x_stride, d_n_samples, batch, max_samples, q->fs, floor, q->f0_ceil,
                q->channels_in_octave, q->target_fs, q->frame_period_ms, q->allowed_range, ws, L.tracker, f_stride,
                d_tpos, f0_raw, d_vuv, d_n_frames, nullptr, nullptr);
    if (rc != WB_OK) return rc;
    rc = wb_stonemask(h, stream, d_x, x_stride, d_n_samples, batch, q->fs, d_tpos, f0_raw, d_n_frames, f_stride, f0_ref);
    if (rc != WB_OK) return rc;
  }
  rc = wb_cheaptrick(h, stream, d_x, x_stride, d_n_samples, batch, q->fs, d_tpos, f0_ref, d_vuv, d_n_frames, f_stride,
                     q->q1, q->fft_size, d_dither, q->seed, f0_used, d_spectrogram, d_ps);
  if (rc != WB_OK) return rc;
  if (q->requiem == WB_AP_NONE) {  // World.get_spectrum (main.py:52-80): f0 stays as CheapTrick left it
    return wb_d2d(d_f0, f0_used, (size_t)batch * f_stride * sizeof(double), (wb_stream_t)stream) == 0
               ? WB_OK
               : wb_fail(h, WB_E_CUDA, "wb_encode: copy of the F0 contour failed");
  }
  if (q->requiem == WB_AP_REQUIEM) {
    if (!d_aperiodicity) return wb_fail(h, WB_E_INVALID, "wb_encode: requiem needs d_aperiodicity");
    return wb_d4c_requiem(h, stream, d_x, x_stride, d_n_samples, batch, q->fs, d_tpos, f0_used, d_vuv, d_n_frames,
                          f_stride, q->threshold, q->fft_size, d_f0, d_aperiodicity);
  }
  return wb_d4c(h, stream, d_x, x_stride, d_n_samples, batch, q->fs, d_tpos, f0_used, d_vuv, d_n_frames, f_stride,
                q->threshold, q->fft_size, d_f0, d_aperiodicity, d_coarse_ap);
}

int wb_d4c_expand(wb_handle* h, void* stream, const double* d_coarse_ap, long long rows, int fs,
                  int fft_size_for_spectrum, double* d_aperiodicity) {
  if (!h) return WB_E_INVALID;
  if (!d_coarse_ap || !d_aperiodicity || rows < 0 || fs <= 0)
    return wb_fail(h, WB_E_INVALID, "wb_d4c_expand: null pointer or negative size");
  const int interval = fs < 16000 ? 2000 : 3000;  // d4c.py:24-27
  const int n_bands = wb_d4c_band_count(fs, 0);
  if (n_bands <= 0 || n_bands > 16) return wb_fail(h, WB_E_INVALID, "wb_d4c_expand: %d bands for fs=%d", n_bands, fs);
  const int n_spec = fft_size_for_spectrum > 0 ? fft_size_for_spectrum : wb_cheaptrick_fft_size(fs);
  if (!wb_is_pow2(n_spec)) return wb_fail(h, WB_E_UNSUPPORTED, "wb_d4c_expand: fft_size %d", n_spec);
  WB_SET_DEVICE(h);
  wb_d4c_expand_body k;
  k.coarse = d_coarse_ap;
  k.ap = d_aperiodicity;
  k.rows = rows;
  k.fs = fs;
  k.n_spec = n_spec;
  k.interval = interval;
  k.n_bands = n_bands;
  WB_CHECK_LAUNCH(h, wb_launch(k, rows, 128, 16 * sizeof(double), (wb_stream_t)stream), "wb_d4c_expand");
  return WB_OK;
}

// Diagnostic for bench.py's FP64 roofline: every thread runs `iters` rounds of 8 independent fused multiply-adds.
struct wb_probe_dfma_body {
  double* out;
  int iters;
  WB_DEV void operator()(long long item) const {
    double a0 = 1.0 + (double)item * 1e-9, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0, a4 = a0 + 4.0, a5 = a0 + 5.0,
           a6 = a0 + 6.0, a7 = a0 + 7.0;
    const double m = 0.999999, c = 1e-6;
    for (int i = 0; i < iters; ++i) {
      a0 = fma(a0, m, c);
      a1 = fma(a1, m, c);
      a2 = fma(a2, m, c);
      a3 = fma(a3, m, c);
      a4 = fma(a4, m, c);
      a5 = fma(a5, m, c);
      a6 = fma(a6, m, c);
      a7 = fma(a7, m, c);
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 123.456) out[0] = r;  // keeps the chains alive; never true
  }
};

int wb_probe_dfma(wb_handle* h, void* stream, long long threads, int iters, double* d_out, double* flops) {
  if (!h || !d_out || threads <= 0 || iters <= 0) return WB_E_INVALID;
  WB_SET_DEVICE(h);
  wb_probe_dfma_body k;
  k.out = d_out;
  k.iters = iters;
  WB_CHECK_LAUNCH(h, wb_launch_flat(k, threads, 256, (wb_stream_t)stream), "wb_probe_dfma");
  if (flops) *flops = 2.0 * 8.0 * (double)iters * (double)threads;
  return WB_OK;
}

int wb_decode_workspace_bytes(wb_handle* h, int batch, int y_stride, int requiem_rows, size_t* bytes) {
  const int rc = wb_synthesis_workspace_bytes(h, batch, y_stride, requiem_rows, bytes);
  if (rc == WB_OK) *bytes += 2 * align_up256((size_t)(batch > 0 ? batch : 1) * sizeof(int));  // pulse / noise counts
  return rc;
}

int wb_decode(wb_handle* h, void* stream, int fs, int fft_size, const double* d_tpos, const double* d_f0,
              const double* d_vuv, const double* d_spectrogram, const double* d_aperiodicity, const int* d_n_frames,
              int batch, int f_stride, int requiem_rows, const double* d_pulse_seed, int seed_fft,
              const double* d_noise_seed, int noise_len, const double* d_cursor_in, double* d_cursor_out,
              const double* d_noise, int noise_stride, uint64_t seed, void* d_workspace, size_t workspace_bytes,
              double* d_y, int y_stride, int normalize, int* d_out_len) {
  if (!h) return WB_E_INVALID;
  if (!d_out_len) return wb_fail(h, WB_E_INVALID, "wb_decode: null d_out_len");
  // the two per-utterance counts the split API reports (pulses, normals consumed) go to the workspace tail
  size_t need = 0;
  int rc = wb_synthesis_workspace_bytes(h, batch, y_stride, requiem_rows, &need);
  if (rc != WB_OK) return rc;
  const size_t extra = align_up256((size_t)(batch > 0 ? batch : 1) * sizeof(int));
  if (workspace_bytes < need + 2 * extra)
    return wb_fail(h, WB_E_INVALID, "wb_decode: workspace %zu < %zu", workspace_bytes, need + 2 * extra);
  int* n_pulses = (int*)((char*)d_workspace + need);
  int* noise_total = (int*)((char*)d_workspace + need + extra);
  rc = wb_synthesis_timebase(h, stream, d_tpos, d_f0, d_vuv, d_n_frames, batch, f_stride, fs, y_stride, d_workspace, need,
                             requiem_rows, d_out_len, n_pulses, noise_total);
  if (rc != WB_OK) return rc;
  if (requiem_rows > 0)
    return wb_synthesis_requiem(h, stream, d_tpos, d_f0, d_vuv, d_spectrogram, d_aperiodicity, d_n_frames, batch,
                                f_stride, fs, fft_size, requiem_rows, d_pulse_seed, seed_fft, d_noise_seed, noise_len,
                                d_cursor_in, d_cursor_out, d_workspace, need, d_y, y_stride, normalize);
  return wb_synthesis(h, stream, d_tpos, d_f0, d_vuv, d_spectrogram, d_aperiodicity, d_n_frames, batch, f_stride, fs,
                      fft_size, d_workspace, need, d_noise, noise_stride, seed, d_y, y_stride, normalize);
}

}  // extern "C"
