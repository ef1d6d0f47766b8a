/* world_b200 -- C-ABI of the B200-native WORLD analysis/synthesis engine.
 *
 * The reference (tuanad121/Python-WORLD) has no FFI: its boundary is the Python
 * class world.main.World (main.py:26) and the stage functions it calls.  This
 * header is the binding surface a maintainer of the reference would target
 * (INTEGRATION.md shows the ctypes stub); each entry point names the reference
 * function it replaces.
 *
 * Conventions
 *  - Every pointer marked d_ is a DEVICE pointer (HBM) owned by the caller.  The
 *    library never allocates or frees caller-visible memory and keeps no pointer
 *    after the call has been enqueued.  Work is asynchronous on `stream`
 *    (a cudaStream_t passed as void*; NULL = default stream).
 *  - All signal data is float64 (the reference's dtype); complex128 is a pair of
 *    doubles (re, im).
 *  - Batches: utterance u of `batch` starts at d_x + u*x_stride and has
 *    d_n_samples[u] samples; per-frame arrays are [batch, f_stride] (frame fast)
 *    and per-frame matrices [batch, f_stride, bins] (bin fast).  Frames
 *    >= d_n_frames[u] are not touched.
 *  - Return value: 0 on success, negative on error (WB_E_*); wb_last_error()
 *    gives the message.  No C++ exception crosses the boundary.
 */
#ifndef WORLD_B200_H
#define WORLD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wb_handle wb_handle;

#define WB_OK 0
#define WB_E_INVALID (-1)      /* bad argument (the facade raises Exception / AssertionError) */
#define WB_E_UNSUPPORTED (-2)  /* fs / size outside the supported range */
#define WB_E_NOMEM (-3)
#define WB_E_CUDA (-4)         /* CUDA runtime error, see wb_last_error */

/* 1 when this library was built for the GPU (sm_100a); the test-only host
 * emulation build returns 0.  The Python package refuses to load a library that
 * returns 0. */
int wb_is_cuda_build(void);
const char* wb_version(void);

int wb_create(wb_handle** out, int device);
int wb_destroy(wb_handle* h);
const char* wb_last_error(const wb_handle* h);

/* ---- size helpers (pure host arithmetic) --------------------------------------- */
/* int(1000*n_samples/fs/frame_period_ms + 1): dio.py:28, harvest.py:21,46 */
int wb_frame_count(int n_samples, int fs, double frame_period_ms);
/* 2^ceil(log2(3 fs/71 + 1)): cheaptrick.py:20-22 */
int wb_cheaptrick_fft_size(int fs);
/* number of aperiodicity bands: d4c.py:23-34 (requiem=0), d4cRequiem.py:13-20 (requiem=1) */
int wb_d4c_band_count(int fs, int requiem);
/* len(np.arange(t0, t_end + 1/fs, 1/fs)): synthesis.py:39, synthesisRequiem.py:39 */
int wb_synthesis_length(double t0, double t_end, int fs);

/* ---- CheapTrick: replaces world/cheaptrick.py:9 cheaptrick() --------------------
 * d_f0/d_vuv: the F0 tracker's output.  d_f0_used receives what the reference
 * leaves in source['f0'] (500 at unvoiced and below-limit frames, cheaptrick.py:27,33).
 * d_dither: optional [batch, f_stride, fft/2+1] values added to the smoothed
 * spectrum (the reference adds |rand|*eps, cheaptrick.py:117); NULL selects a
 * deterministic hash of (seed, frame, bin) scaled by eps.
 * d_ps: optional complex128 [batch, f_stride, fft] pitch-synchronous spectrum
 * ('ps spectrogram'), NULL to skip.  fft_size 0 = default. */
int wb_cheaptrick(wb_handle* h, void* stream, const double* d_x, int x_stride, const int* d_n_samples, int batch,
                  int fs, const double* d_temporal_positions, const double* d_f0, const double* d_vuv,
                  const int* d_n_frames, int f_stride, double q1, int fft_size, const double* d_dither,
                  uint64_t seed, double* d_f0_used, double* d_spectrogram, void* d_ps);

/* ---- D4C: replaces world/d4c.py:10 d4c() -----------------------------------------
 * d_f0: source['f0'] as CheapTrick left it.  d_f0_out: 0 at unvoiced frames
 * (d4c.py:32).  d_aperiodicity [batch, f_stride, fft_size_for_spectrum/2+1] linear
 * amplitude; d_coarse_ap optional [batch, f_stride, bands]; d_aperiodicity may be NULL when
 * d_coarse_ap is given (see wb_d4c_expand). */
int wb_d4c(wb_handle* h, void* stream, const double* d_x, int x_stride, const int* d_n_samples, int batch, int fs,
           const double* d_temporal_positions, const double* d_f0, const double* d_vuv, const int* d_n_frames,
           int f_stride, double threshold, int fft_size_for_spectrum, double* d_f0_out, double* d_aperiodicity,
           double* d_coarse_ap);

/* ---- D4C-Requiem: replaces world/d4cRequiem.py:9 d4cRequiem() --------------------
 * d_band_aperiodicity [batch, f_stride, bands+2] in dB.  fft_size 0 = 3 fs/47 rule. */
int wb_d4c_requiem(wb_handle* h, void* stream, const double* d_x, int x_stride, const int* d_n_samples, int batch,
                   int fs, const double* d_temporal_positions, const double* d_f0, const double* d_vuv,
                   const int* d_n_frames, int f_stride, double threshold, int fft_size, double* d_f0_out,
                   double* d_band_aperiodicity);

/* ---- Harvest: replaces world/harvest.py:17 harvest() ------------------------------
 * F0 tracking at frame_period_ms.  The caller provides one workspace of
 * wb_harvest_workspace_bytes() bytes (max_samples = largest d_n_samples[u], <= x_stride)
 * and output arrays [batch, f_stride] with f_stride >= wb_frame_count(max_samples, fs, period).
 * d_n_frames[u] receives the frame count of utterance u.  An all-zero utterance yields
 * all-unvoiced frames (the reference raises IndexError there). */
int wb_harvest_workspace_bytes(wb_handle* h, int batch, int max_samples, int fs, double f0_floor, double f0_ceil,
                               size_t* bytes);
int wb_harvest(wb_handle* h, void* stream, const double* d_x, int x_stride, const int* d_n_samples, int batch,
               int max_samples, int fs, double f0_floor, double f0_ceil, double frame_period_ms, void* d_workspace,
               size_t workspace_bytes, int f_stride, double* d_temporal_positions, double* d_f0, double* d_vuv,
               int* d_n_frames);
/* Diagnostics (tests / bench): workspace layout of the intermediates, and a variant that runs only kernels
 * stage_first..stage_last (0 decimate, 1 channels, 2 detect, 3 refine, 4 prune, 5 contour). */
int wb_harvest_workspace_layout(wb_handle* h, int batch, int max_samples, int fs, double f0_floor, double f0_ceil,
                                size_t* offsets16, int* dims8);
int wb_harvest_stages(wb_handle* h, void* stream, const double* d_x, int x_stride, const int* d_n_samples, int batch,
                      int max_samples, int fs, double f0_floor, double f0_ceil, double frame_period_ms,
                      void* d_workspace, size_t workspace_bytes, int f_stride, double* d_temporal_positions,
                      double* d_f0, double* d_vuv, int* d_n_frames, int stage_first, int stage_last);

/* ---- DIO: replaces world/dio.py:10 dio() -------------------------------------------
 * fs / target_fs must give a decimation ratio of 2..12 (the reference's coefficient table, dio.py:365-436;
 * outside it the reference silently filters with zeros).  Outputs [batch, f_stride]; optional
 * d_f0_candidates [batch, f_stride, bands] (sorted by stability) and d_raw_f0_candidates
 * [batch, bands, f_stride], the other two keys of the reference's return dict. */
int wb_dio_band_count(double f0_floor, double f0_ceil, int channels_in_octave);
int wb_dio_workspace_bytes(wb_handle* h, int batch, int max_samples, int fs, double f0_floor, double f0_ceil,
                           int channels_in_octave, int target_fs, double frame_period_ms, size_t* bytes);
int wb_dio(wb_handle* h, void* stream, const double* d_x, int x_stride, const int* d_n_samples, int batch,
           int max_samples, int fs, double f0_floor, double f0_ceil, int channels_in_octave, int target_fs,
           double frame_period_ms, double allowed_range, void* d_workspace, size_t workspace_bytes, int f_stride,
           double* d_temporal_positions, double* d_f0, double* d_vuv, int* d_n_frames, double* d_f0_candidates,
           double* d_raw_f0_candidates);

/* ---- StoneMask: replaces world/stonemask.py:8 stonemask() --------------------------
 * Refines every non-zero d_f0 (>= 40 Hz); zeros stay zero. */
int wb_stonemask(wb_handle* h, void* stream, const double* d_x, int x_stride, const int* d_n_samples, int batch, int fs,
                 const double* d_temporal_positions, const double* d_f0, const int* d_n_frames, int f_stride,
                 double* d_refined_f0);

/* ---- Synthesis: replaces world/synthesis.py:21 synthesis() and world/synthesisRequiem.py:12
 * synthesisRequiem(), plus the peak normalisation of World.decode (main.py:209-212) ----
 * Two calls per batch.  wb_synthesis_timebase builds every utterance's pulse train and reports, per
 * utterance, the output length (<= y_stride), the pulse count and the number of standard normals
 * synthesis() consumes (the reference draws np.random.randn(max(3, noise_size)) per pulse,
 * synthesis.py:93).  Then wb_synthesis with d_noise [batch, noise_stride] = those normals in draw order
 * (or NULL for the built-in counter-based generator), or wb_synthesis_requiem with the seeds of
 * get_seeds_signals() (pulse [seed_fft, rows], noise [noise_len, rows], row-major), d_cursor_in [rows]
 * = generate_noise.current_index on entry and d_cursor_out [batch, rows] its value after each
 * utterance.  The workspace (wb_synthesis_workspace_bytes; requiem_rows = 0 for synthesis.py) carries
 * the pulse trains from the first call to the second.  d_y [batch, y_stride] is overwritten. */
int wb_synthesis_workspace_bytes(wb_handle* h, int batch, int y_stride, int requiem_rows, size_t* bytes);
int wb_synthesis_timebase(wb_handle* h, void* stream, const double* d_temporal_positions, const double* d_f0,
                          const double* d_vuv, const int* d_n_frames, int batch, int f_stride, int fs, int y_stride,
                          void* d_workspace, size_t workspace_bytes, int requiem_rows, int* d_out_len, int* d_n_pulses,
                          int* d_noise_total);
int wb_synthesis(wb_handle* h, void* stream, const double* d_temporal_positions, const double* d_f0,
                 const double* d_vuv, const double* d_spectrogram, const double* d_aperiodicity, const int* d_n_frames,
                 int batch, int f_stride, int fs, int fft_size, void* d_workspace, size_t workspace_bytes,
                 const double* d_noise, int noise_stride, uint64_t seed, double* d_y, int y_stride, int normalize);
int wb_synthesis_requiem(wb_handle* h, void* stream, const double* d_temporal_positions, const double* d_f0,
                         const double* d_vuv, const double* d_spectrogram, const double* d_band_aperiodicity,
                         const int* d_n_frames, int batch, int f_stride, int fs, int fft_size, int rows,
                         const double* d_pulse_seed, int seed_fft, const double* d_noise_seed, int noise_len,
                         const double* d_cursor_in, double* d_cursor_out, void* d_workspace, size_t workspace_bytes,
                         double* d_y, int y_stride, int normalize);

/* ---- Fused analysis / synthesis: World.encode (main.py:106-152) and World.decode (main.py:198-214) ----
 * One workspace query and one call each; the call enqueues every stage kernel on `stream` in the
 * reference's order (tracker [-> StoneMask] -> CheapTrick -> D4C | D4C-Requiem; time base -> synthesis |
 * synthesisRequiem -> peak rescale).  They add no arithmetic to the stage entry points above, so a binder
 * can use either level.
 *
 * wb_encode: f_stride must equal wb_frame_count(max_samples, fs, frame_period_ms).  Outputs as in the
 * stage calls: d_f0 is the final source['f0'] (0 at unvoiced frames; 500 where a voiced frame lies below
 * CheapTrick's limit, cheaptrick.py:32-33); d_aperiodicity is [batch, f_stride, fft/2+1] linear
 * (requiem = 0) or [batch, f_stride, bands+2] dB (requiem = 1).  For requiem = 0 d_aperiodicity may be
 * NULL when d_coarse_ap [batch, f_stride, bands] is given: the band values are the compact transport
 * form of the aperiodicity (wb_d4c_expand below rebuilds it bit for bit).  fft_size 0 = defaults; when
 * set, the F0 floor becomes 3 fs / fft_size (main.py:123-124).  d_dither / d_ps as in wb_cheaptrick. */
#define WB_F0_HARVEST 0
#define WB_F0_DIO 1 /* dio + stonemask */
#define WB_AP_D4C 0
#define WB_AP_REQUIEM 1
#define WB_AP_NONE 2
typedef struct wb_encode_params {
  int fs;
  int f0_method;          /* WB_F0_HARVEST | WB_F0_DIO; anything else: WB_E_INVALID (main.py:136-137) */
  double f0_floor, f0_ceil;
  int channels_in_octave; /* dio */
  int target_fs;          /* dio */
  double frame_period_ms;
  double allowed_range;   /* dio */
  int fft_size;           /* 0 = cheaptrick.py:20-22 default */
  int requiem;            /* WB_AP_D4C = d4c, WB_AP_REQUIEM = d4cRequiem, WB_AP_NONE = no aperiodicity stage:
                             World.get_spectrum (main.py:52-80), d_f0 = the contour as CheapTrick leaves it */
  double q1;              /* cheaptrick.py:9, -0.15 */
  double threshold;       /* love-train threshold, 0.85 */
  uint64_t seed;          /* hash dither seed when d_dither is NULL */
} wb_encode_params;
int wb_encode_workspace_bytes(wb_handle* h, const wb_encode_params* params, int batch, int max_samples, size_t* bytes);
int wb_encode(wb_handle* h, void* stream, const wb_encode_params* params, const double* d_x, int x_stride,
              const int* d_n_samples, int batch, int max_samples, void* d_workspace, size_t workspace_bytes,
              int f_stride, const double* d_dither, double* d_temporal_positions, double* d_f0, double* d_vuv,
              int* d_n_frames, double* d_spectrogram, double* d_aperiodicity, double* d_coarse_ap, void* d_ps);

/* aperiodicity [rows, fft/2+1] from 'coarse_ap' [rows, bands] (d4c.py:56-59): rows whose first band value
 * has the sign bit clear (+0.0: unvoiced, or rejected by the love-train gate, d4c.py:49-51) get
 * 1 - 1e-12; wb_d4c writes every other row with the sign bit set (-0.0 when the value clamps to zero).
 * Same code as the tail of the D4C kernel: identical bits. */
int wb_d4c_expand(wb_handle* h, void* stream, const double* d_coarse_ap, long long rows, int fs,
                  int fft_size_for_spectrum, double* d_aperiodicity);

/* wb_decode: requiem_rows = 0 runs synthesis.py (d_aperiodicity [batch, f_stride, fft/2+1]; d_noise as in
 * wb_synthesis, NULL = counter-based generator with `seed`), requiem_rows = bands+2 runs
 * synthesisRequiem.py with the seeds / cursors of wb_synthesis_requiem.  d_out_len [batch] receives the
 * output lengths; normalize applies main.py:209-212. */
int wb_decode_workspace_bytes(wb_handle* h, int batch, int y_stride, int requiem_rows, size_t* bytes);
int wb_decode(wb_handle* h, void* stream, int fs, int fft_size, const double* d_temporal_positions, const double* d_f0,
              const double* d_vuv, const double* d_spectrogram, const double* d_aperiodicity, const int* d_n_frames,
              int batch, int f_stride, int requiem_rows, const double* d_pulse_seed, int seed_fft,
              const double* d_noise_seed, int noise_len, const double* d_cursor_in, double* d_cursor_out,
              const double* d_noise, int noise_stride, uint64_t seed, void* d_workspace, size_t workspace_bytes,
              double* d_y, int y_stride, int normalize, int* d_out_len);

/* ---- Spectral feature heads on the resident spectrogram --------------------
 * Rows are frames, bin fast: the [batch, f_stride, bins] arrays above taken as [rows, bins].
 * Tables that depend only on the sizes are prepared by the caller with the reference's own
 * expressions (world_b200/features.py does it for the facade).
 *
 * wb_lfbank: replaces World.encode_lfbank (main.py:305-322).  d_spec [rows, n_bins] MAGNITUDE
 *   spectrum; d_preemph_abs [n_bins] = |freqz([1, -prefac], 1, n_bins)|; d_filterbank
 *   [n_filt, n_bins] = get_filterbanks() (main.py:274-303).  d_out [rows, n_filt] = log energies
 *   (exact zeros replaced by eps first).
 * wb_mcep: replaces World.encode_mcep (main.py:324-342).  d_mel_bin [n_bins] = the integer source
 *   bin of each mel point (floor(...), main.py:337).  d_out [rows, n0].
 * wb_mcep_decode: replaces World.decode_mcep (main.py:344-358).  d_mel_pos [fft_size/2+1] = the mel
 *   bin positions (non-decreasing); d_bracket / d_query [fft_size/2+1]: for output bin i the
 *   numpy.interp bracket (largest j with mel_pos[j] <= i, 0 when none) and the query (i, or
 *   mel_pos[0] when i lies below it).  d_out [rows, fft_size/2+1] magnitude spectrum. */
int wb_lfbank(wb_handle* h, void* stream, const double* d_spec, int rows, int n_bins, const double* d_preemph_abs,
              const double* d_filterbank, int n_filt, double* d_out);
int wb_mcep(wb_handle* h, void* stream, const double* d_spec, int rows, int n_bins, const int* d_mel_bin, int n0,
            double* d_out);
int wb_mcep_decode(wb_handle* h, void* stream, const double* d_cepstrum, int rows, int n0, int fft_size,
                   const double* d_mel_pos, const int* d_bracket, const double* d_query, double* d_out);

/* ---- Prosody / spectrum edits on resident data ----------------------------
 * wb_interp_rows: numpy.interp(query, knots, row) for every row -- World.warp_spectrum
 *   (main.py:189-194) with knots = arange(n)/n and query = knots**factor.  d_bracket[i] = largest j
 *   with knots[j] <= query[i] (0 when none; then pass query[i] = knots[0]).  d_out may alias d_in
 *   when n_out == n_in.
 * wb_interp_knots: numpy.interp(x, knot_x, knot_y) element-wise -- World.modify_duration
 *   (main.py:178-187) applied to the temporal positions.  d_out may alias d_x. */
int wb_interp_rows(wb_handle* h, void* stream, const double* d_in, int rows, int n_in, const double* d_knots,
                   const int* d_bracket, const double* d_query, int n_out, double* d_out);
int wb_interp_knots(wb_handle* h, void* stream, const double* d_x, long long n, const double* d_knot_x,
                    const double* d_knot_y, int n_knots, double* d_out);

/* ---- PCM edge (example/prosody.py:12-13, 57) --------------------------------
 * wb_pcm16_to_f64: x = pcm / divisor (the reference divides by 2**15 - 1), samples past
 *   d_n_samples[u] are written as 0.  wb_f64_to_pcm16: (y * gain).astype(int16) -- truncation
 *   toward zero and 16-bit wrap-around as NumPy does on x86-64 (1.0 * 2**15 -> -32768). */
int wb_pcm16_to_f64(wb_handle* h, void* stream, const int16_t* d_pcm, int pcm_stride, const int* d_n_samples, int batch,
                    double divisor, double* d_x, int x_stride);
int wb_f64_to_pcm16(wb_handle* h, void* stream, const double* d_y, int y_stride, const int* d_n_samples, int batch,
                    double gain, int16_t* d_pcm, int pcm_stride);

/* Optional float32 transport of per-frame matrices (the batch API's opt-in spectrogram_dtype=float32: half the
 * PCIe bytes at a rounding of 6e-8 relative, three orders of magnitude inside the parity tolerance of the
 * spectrogram): round-to-nearest narrowing, exact widening. */
int wb_f64_to_f32(wb_handle* h, void* stream, const double* d_in, long long n, float* d_out);
int wb_f32_to_f64(wb_handle* h, void* stream, const float* d_in, long long n, double* d_out);

/* Diagnostic (bench.py's FP64 roofline denominator): `threads` threads each run `iters` rounds of 8 independent
 * float64 fused multiply-adds; *flops receives the number of floating-point operations of the launch. */
int wb_probe_dfma(wb_handle* h, void* stream, long long threads, int iters, double* d_out, double* flops);
/* Diagnostic: the Nuttall window exactly as the library tabulates it (host buffer of n doubles). */
int wb_debug_nuttall(int n, double* host_out);

#ifdef __cplusplus
}
#endif
#endif /* WORLD_B200_H */
