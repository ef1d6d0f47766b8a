// C-ABI: synthesis (replaces world/synthesis.py:21 synthesis() and world/synthesisRequiem.py:12
// synthesisRequiem(), plus the peak normalisation of main.py:209-212).
#include "wb_handle.h"
#include "wb_synthesis.h"

namespace {

inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

struct sy_sizes {
  int p_cap;
  size_t off[16];
  size_t total;
};

void sy_plan_sizes(int batch, int y_stride, int n_ap, sy_sizes* z) {
  z->p_cap = y_stride / 2 + 8;
  const size_t B = (size_t)batch, Y = (size_t)y_stride, P = (size_t)z->p_cap;
  size_t o = 0;
  int i = 0;
  auto put = [&](size_t bytes) {
    z->off[i++] = o;
    o += align_up(bytes);
  };
  put(B * Y * sizeof(double));          // 0 wrap
  put(B * Y);                           // 1 vuv_i
  put(B * P * sizeof(double));          // 2 p_loc
  put(B * P * sizeof(int));             // 3 p_idx
  put(B * P * sizeof(double));          // 4 p_shift
  put(B * P * sizeof(int));             // 5 p_noise_off
  put(B * sizeof(int));                 // 6 n_pulses
  put(B * sizeof(int));                 // 7 noise_total
  put(B * sizeof(int));                 // 8 out_len
  put((B + 1) * sizeof(int));           // 9 pulse_base
  put(B * (size_t)(n_ap > 0 ? n_ap : 0) * Y * sizeof(double));  // 10 ap_i
  put(n_ap > 0 ? B * Y * sizeof(double) : 0);                   // 11 exc
  z->total = o;
}

int sy_slots(wb_handle* h) {
#ifdef WB_HOST_EMU
  (void)h;
  return 2;
#else
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
  return 4 * sms;
#endif
}

void sy_fill(wb_sy_plan* p, const sy_sizes& z, char* ws, int batch, int fs, int n, int f_stride, int y_stride,
             const double* tpos, const double* f0, const double* vuv, const double* spec, const double* ap,
             const int* n_frames, int n_ap, double* y) {
  memset(p, 0, sizeof *p);
  p->batch = batch;
  p->fs = fs;
  p->n = n;
  p->n_bins = n / 2 + 1;
  p->f_stride = f_stride;
  p->y_stride = y_stride;
  p->tpos = tpos;
  p->f0 = f0;
  p->vuv = vuv;
  p->spec = spec;
  p->ap = ap;
  p->n_frames = n_frames;
  p->n_ap = n_ap;
  p->wrap = (double*)(ws + z.off[0]);
  p->vuv_i = (unsigned char*)(ws + z.off[1]);
  p->p_loc = (double*)(ws + z.off[2]);
  p->p_idx = (int*)(ws + z.off[3]);
  p->p_shift = (double*)(ws + z.off[4]);
  p->p_noise_off = (int*)(ws + z.off[5]);
  p->p_cap = z.p_cap;
  p->n_pulses = (int*)(ws + z.off[6]);
  p->noise_total = (int*)(ws + z.off[7]);
  p->out_len = (int*)(ws + z.off[8]);
  p->pulse_base = (int*)(ws + z.off[9]);
  p->ap_i = (double*)(ws + z.off[10]);
  p->exc = (double*)(ws + z.off[11]);
  p->y = y;
}

}  // namespace

extern "C" {

int wb_synthesis_workspace_bytes(wb_handle* h, int batch, int y_stride, int requiem_rows, size_t* bytes) {
  if (!h || !bytes || batch < 0 || y_stride < 0) return WB_E_INVALID;
  sy_sizes z;
  sy_plan_sizes(batch, y_stride, requiem_rows, &z);
  *bytes = z.total;
  return WB_OK;
}

/* Stage 1 of either synthesiser: pulse train of every utterance.  d_out_len / d_n_pulses / d_noise_total
 * [batch] receive the output length, the pulse count and the number of normals synthesis() will consume. */
int wb_synthesis_timebase(wb_handle* h, void* stream, const double* d_tpos, const double* d_f0, const double* d_vuv,
                          const int* d_n_frames, int batch, int f_stride, int fs, int y_stride, void* d_workspace,
                          size_t workspace_bytes, int requiem_rows, int* d_out_len, int* d_n_pulses,
                          int* d_noise_total) {
  if (!h) return WB_E_INVALID;
  if (!d_tpos || !d_f0 || !d_vuv || !d_n_frames || !d_workspace || batch < 0 || fs <= 0)
    return wb_fail(h, WB_E_INVALID, "wb_synthesis_timebase: null pointer or negative size");
  sy_sizes z;
  sy_plan_sizes(batch, y_stride, requiem_rows, &z);
  if (workspace_bytes < z.total) return wb_fail(h, WB_E_INVALID, "wb_synthesis: workspace %zu < %zu", workspace_bytes, z.total);
  if (batch == 0) return WB_OK;
  WB_SET_DEVICE(h);
  wb_sy_plan p;
  sy_fill(&p, z, (char*)d_workspace, batch, fs, 0, f_stride, y_stride, d_tpos, d_f0, d_vuv, nullptr, nullptr, d_n_frames,
          requiem_rows, nullptr);
  wb_stream_t st = (wb_stream_t)stream;
  wb_sy_timebase k1;
  k1.p = p;
  const int nthr = 256;
  WB_CHECK_LAUNCH(h, wb_launch(k1, batch, nthr, wb_sy_timebase::smem_bytes(nthr), st), "sy_timebase");
  wb_sy_prefix k2;
  k2.p = p;
  WB_CHECK_LAUNCH(h, wb_launch_flat(k2, 1, 32, st), "sy_prefix");
  if (d_out_len && wb_d2d(d_out_len, p.out_len, (size_t)batch * sizeof(int), st)) return wb_fail(h, WB_E_CUDA, "copy");
  if (d_n_pulses && wb_d2d(d_n_pulses, p.n_pulses, (size_t)batch * sizeof(int), st)) return wb_fail(h, WB_E_CUDA, "copy");
  if (d_noise_total && wb_d2d(d_noise_total, p.noise_total, (size_t)batch * sizeof(int), st))
    return wb_fail(h, WB_E_CUDA, "copy");
  return WB_OK;
}

/* Stage 2, synthesis.py: d_noise [batch, noise_stride] holds each utterance's normals in draw order
 * (np.random.randn pulse by pulse), or NULL for the device generator seeded with `seed`.
 * d_y [batch, y_stride] is overwritten.  normalize != 0 applies main.py:209-212. */
int wb_synthesis(wb_handle* h, void* stream, const double* d_tpos, const double* d_f0, const double* d_vuv,
                 const double* d_spectrogram, const double* d_aperiodicity, const int* d_n_frames, int batch,
                 int f_stride, int fs, int fft_size, void* d_workspace, size_t workspace_bytes, const double* d_noise,
                 int noise_stride, uint64_t seed, double* d_y, int y_stride, int normalize) {
  if (!h) return WB_E_INVALID;
  if (!d_tpos || !d_f0 || !d_vuv || !d_spectrogram || !d_aperiodicity || !d_n_frames || !d_workspace || !d_y ||
      batch < 0 || fs <= 0)
    return wb_fail(h, WB_E_INVALID, "wb_synthesis: null pointer or negative size");
  if (!wb_is_pow2(fft_size) || fft_size < 16 || fft_size > WB_TW_N)
    return wb_fail(h, WB_E_UNSUPPORTED, "wb_synthesis: fft_size %d", fft_size);
  sy_sizes z;
  sy_plan_sizes(batch, y_stride, 0, &z);
  if (workspace_bytes < z.total) return wb_fail(h, WB_E_INVALID, "wb_synthesis: workspace %zu < %zu", workspace_bytes, z.total);
  if (batch == 0) return WB_OK;
  WB_SET_DEVICE(h);
  const int n = fft_size;
  const double* dc = wb_table<double>(h, "sy_dc:" + std::to_string(n), [n](std::vector<double>& o) {
    o.resize(n);
    double s = 0.0;
    for (int i = 0; i < n; ++i) {  // scipy hann(n + 2)[1:-1]
      o[i] = 0.5 - 0.5 * std::cos(2.0 * WB_PI * (double)(i + 1) / (double)(n + 1));
      s += o[i];
    }
    for (int i = 0; i < n; ++i) o[i] /= s;
  });
  if (!dc) return wb_fail(h, WB_E_NOMEM, "wb_synthesis: table allocation failed");
  wb_sy_plan p;
  sy_fill(&p, z, (char*)d_workspace, batch, fs, n, f_stride, y_stride, d_tpos, d_f0, d_vuv, d_spectrogram, d_aperiodicity,
          d_n_frames, 0, d_y);
  wb_stream_t st = (wb_stream_t)stream;
  if (wb_dev_memset(d_y, 0, (size_t)batch * y_stride * sizeof(double), st)) return wb_fail(h, WB_E_CUDA, "memset");
  wb_sy_pulses k;
  k.p = p;
  k.tw = h->tw;
  k.tw_n = WB_TW_N;
  k.dc_base = dc;
  k.noise = d_noise;
  k.noise_stride = noise_stride;
  k.seed = seed;
  k.n_slots = sy_slots(h);
  k.max_noise = n;
  int nthr = n / 4;
  if (const char* e = std::getenv("WB_SY_THREADS")) nthr = std::atoi(e);  // tuning knob
  nthr = nthr < 128 ? 128 : (nthr > 512 ? 512 : nthr);
  const size_t smem = wb_sy_pulses::smem_bytes(n, k.max_noise, nthr);
  if (smem > 227 * 1024) return wb_fail(h, WB_E_UNSUPPORTED, "wb_synthesis: %zu bytes of shared memory", smem);
#ifndef WB_HOST_EMU
  if (n == 1024 && nthr == 256) {  // 16 / 22.05 kHz: compile-time sizes
    wb_sy_pulses_t<1024, 256> kt;
    kt.p = k.p; kt.tw = k.tw; kt.tw_n = k.tw_n; kt.dc_base = k.dc_base; kt.noise = k.noise; kt.noise_stride = k.noise_stride;
    kt.seed = k.seed; kt.n_slots = k.n_slots; kt.max_noise = k.max_noise;
    WB_CHECK_LAUNCH(h, (wb_launch_b<wb_sy_pulses_t<1024, 256>, 256, 4>(kt, kt.n_slots, 256, smem, st)), "sy_pulses");
  } else
#endif
  WB_CHECK_LAUNCH(h, wb_launch_spectral(k, k.n_slots, nthr, smem, st), "sy_pulses");
  if (normalize) {
    wb_sy_normalise kn;
    kn.p = p;
    kn.requiem = 0;
    WB_CHECK_LAUNCH(h, wb_launch(kn, batch, 256, (WB_REDUCE_SCRATCH + 8) * sizeof(double), st), "sy_normalise");
  }
  return WB_OK;
}

/* Stage 2, synthesisRequiem.py: d_band_ap [batch, f_stride, rows] dB; seeds as get_seeds_signals() returns them
 * (pulse [seed_fft, rows], noise [noise_len, rows], row-major); d_cursor_in [rows] = generate_noise.current_index
 * on entry (zeros for a fresh process), d_cursor_out [batch, rows] its value after each utterance. */
int wb_synthesis_requiem(wb_handle* h, void* stream, const double* d_tpos, const double* d_f0, const double* d_vuv,
                         const double* d_spectrogram, const double* d_band_ap, const int* d_n_frames, int batch,
                         int f_stride, int fs, int fft_size, int rows, const double* d_pulse_seed, int seed_fft,
                         const double* d_noise_seed, int noise_len, const double* d_cursor_in, double* d_cursor_out,
                         void* d_workspace, size_t workspace_bytes, double* d_y, int y_stride, int normalize) {
  if (!h) return WB_E_INVALID;
  if (!d_tpos || !d_f0 || !d_vuv || !d_spectrogram || !d_band_ap || !d_n_frames || !d_workspace || !d_y ||
      !d_pulse_seed || !d_noise_seed || !d_cursor_in || !d_cursor_out || batch < 0 || fs <= 0 || rows < 1)
    return wb_fail(h, WB_E_INVALID, "wb_synthesis_requiem: null pointer or negative size");
  if (!wb_is_pow2(fft_size) || fft_size < 16 || fft_size > WB_TW_N)
    return wb_fail(h, WB_E_UNSUPPORTED, "wb_synthesis_requiem: fft_size %d", fft_size);
  sy_sizes z;
  sy_plan_sizes(batch, y_stride, rows, &z);
  if (workspace_bytes < z.total)
    return wb_fail(h, WB_E_INVALID, "wb_synthesis_requiem: workspace %zu < %zu", workspace_bytes, z.total);
  if (batch == 0) return WB_OK;
  WB_SET_DEVICE(h);
  wb_sy_plan p;
  sy_fill(&p, z, (char*)d_workspace, batch, fs, fft_size, f_stride, y_stride, d_tpos, d_f0, d_vuv, d_spectrogram,
          d_band_ap, d_n_frames, rows, d_y);
  wb_stream_t st = (wb_stream_t)stream;
  if (wb_dev_memset(d_y, 0, (size_t)batch * y_stride * sizeof(double), st)) return wb_fail(h, WB_E_CUDA, "memset");
  {
    wb_rq_excite k;
    k.p = p;
    k.pulse_seed = d_pulse_seed;
    k.noise_seed = d_noise_seed;
    k.seed_n = seed_fft;
    k.noise_len = noise_len;
    k.cursor_in = d_cursor_in;
    k.cursor_out = d_cursor_out;
    WB_CHECK_LAUNCH(h, wb_launch(k, batch, 256, 64, st), "rq_excite");
  }
  {
    wb_rq_pulses k;
    k.p = p;
    k.pulse_seed = d_pulse_seed;
    k.seed_n = seed_fft;
    k.n_slots = sy_slots(h);
    WB_CHECK_LAUNCH(h, wb_launch(k, k.n_slots, 128, wb_rq_pulses::smem_bytes(seed_fft), st), "rq_pulses");
  }
  {
    wb_rq_frames k;
    k.p = p;
    k.tw = h->tw;
    k.tw_n = WB_TW_N;
    k.win = nullptr;
    int nthr = fft_size / 4;
    if (const char* e = std::getenv("WB_SY_THREADS")) nthr = std::atoi(e);  // tuning knob
    nthr = nthr < 128 ? 128 : (nthr > 512 ? 512 : nthr);
    const size_t smem = wb_rq_frames::smem_bytes(fft_size);
    if (smem > 227 * 1024) return wb_fail(h, WB_E_UNSUPPORTED, "wb_synthesis_requiem: %zu bytes of shared memory", smem);
#ifndef WB_HOST_EMU
    if (fft_size == 1024 && nthr == 256) {
      wb_rq_frames_t<1024, 256> kt;
      kt.p = k.p; kt.tw = k.tw; kt.tw_n = k.tw_n; kt.win = k.win;
      WB_CHECK_LAUNCH(h, (wb_launch_b<wb_rq_frames_t<1024, 256>, 256, 4>(kt, (long long)batch * f_stride, 256, smem, st)), "rq_frames");
    } else
#endif
    WB_CHECK_LAUNCH(h, wb_launch_spectral(k, (long long)batch * f_stride, nthr, smem, st), "rq_frames");
  }
  if (normalize) {
    wb_sy_normalise kn;
    kn.p = p;
    kn.requiem = 1;
    WB_CHECK_LAUNCH(h, wb_launch(kn, batch, 256, (WB_REDUCE_SCRATCH + 8) * sizeof(double), st), "sy_normalise");
  }
  return WB_OK;
}

}  // extern "C"
