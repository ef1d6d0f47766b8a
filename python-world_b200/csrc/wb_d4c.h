// D4C and D4C-Requiem band aperiodicity: one thread block per (utterance, frame).
//
// Replaces d4c.py:10-64 / d4cRequiem.py:9-44 and their shared per-frame estimator
// (d4c.py:68-222).  Love-train gate, the two quarter-period centroid windows, the
// smoothed power spectrum, the group-delay shaping and the per-band
// sort-and-accumulate all run out of shared memory in one launch.  The two real
// FFTs of each centroid window (of x and of n*x) are packed into one complex FFT.
#pragma once
#include "wb_spectral.h"

// In-place ascending bitonic sort of v[0..m), m a power of two.
WB_DEV void wb_bitonic_sort(double* v, int m, int tid, int nthr) {
  for (int size = 2; size <= m; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (m >> 1); t += nthr) {
        const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const double a = v[lo], b = v[hi];
        if ((a > b) == up) {
          v[lo] = b;
          v[hi] = a;
        }
      }
      WB_SYNC();
    }
  }
}

#ifndef WB_HOST_EMU
// Sum of all but the K largest of the block's values without sorting them all (GPU only; the host
// emulation keeps the full sort).  Every thread holds VPL values (any assignment), `extra` is one more
// value known to all threads.  Each warp sorts its 32*VPL values in registers (bitonic network: strides
// below VPL inside the thread, the others through shuffles, no barrier) and publishes its KC >= K largest;
// the K largest of the block all lie among those lists, and a published value whose rank among the
// published values is below KC has the same rank among all values (a larger unpublished value would have
// KC still larger published ones in front of it).  The rank comes from one binary search per list, the
// value of rank K-1 is the threshold T, and the answer is sum(v < T) + (copies of T left over) * T.
// `cand`: (nw + 1) * KC doubles of shared memory; KC is a multiple of VPL with K <= KC <= 32 * VPL.
template <int VPL>
WB_DEV_COLD double wb_sum_without_top(double (&v)[VPL], double extra, int K, int KC, double* cand, double* scratch,
                                 int tid, int nthr) {
  const int lane = tid & 31, w = tid >> 5, nw = nthr >> 5;
  // the (size, stride) loops stay rolled (the network would otherwise unroll into thousands of instructions);
  // only the per-thread element index is compile-time, so the value array stays in registers
#pragma unroll 1
  for (int size = 2; size <= 32 * VPL; size <<= 1) {
#pragma unroll 1
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride >= VPL) {
        const int ls = stride / VPL;
        const bool is_lo = (lane & ls) == 0;
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
          const bool desc = (((lane * VPL + j) & size) == 0);
          const double o = __shfl_xor_sync(0xffffffffu, v[j], ls);
          v[j] = (is_lo == desc) ? fmax(v[j], o) : fmin(v[j], o);
        }
      } else {
#pragma unroll
        for (int st = 1; st < VPL; st <<= 1) {
          if (st == stride) {
#pragma unroll
            for (int j = 0; j < VPL; ++j) {
              if ((j & st) == 0) {
                const bool desc = (((lane * VPL + j) & size) == 0);
                const double a = v[j], b = v[j ^ st];
                const double hi = fmax(a, b), lo = fmin(a, b);
                v[j] = desc ? hi : lo;
                v[j ^ st] = desc ? lo : hi;
              }
            }
          }
        }
      }
    }
  }
  // element e = lane * VPL + j of the warp is its e-th largest
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int e = lane * VPL + j;
    if (e < KC) cand[w * KC + e] = v[j];
  }
  for (int i = tid; i < KC; i += nthr) cand[nw * KC + i] = i == 0 ? extra : -1.0;  // the values are powers, >= 0
  __syncthreads();
  for (int ci = tid; ci < (nw + 1) * KC; ci += nthr) {
    const int cw = ci / KC, i = ci - cw * KC;
    const double c = cand[ci];
    if (c < 0.0) continue;
    int rank = i;
    for (int l = 0; l <= nw; ++l) {
      if (l == cw) continue;
      const double* L = cand + l * KC;  // descending; equal values: lower list first
      int lo = 0, hi = KC;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const double x = L[mid];
        if (l < cw ? x >= c : x > c) lo = mid + 1;
        else hi = mid;
      }
      rank += lo;
    }
    if (rank == K - 1) scratch[WB_REDUCE_SCRATCH - 1] = c;
  }
  __syncthreads();
  const double T = scratch[WB_REDUCE_SCRATCH - 1];
  double low = 0.0, n_gt = 0.0, n_eq = 0.0;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    if (v[j] < T) low += v[j];
    else if (v[j] > T) n_gt += 1.0;
    else n_eq += 1.0;
  }
  if (tid == 0) {
    if (extra < T) low += extra;
    else if (extra > T) n_gt += 1.0;
    else n_eq += 1.0;
  }
  wb_block_sum3(low, n_gt, n_eq, scratch, tid, nthr);
  return low + (n_eq - ((double)K - n_gt)) * T;
}
// The same selection on 32-bit keys (GPU only).  The values are powers (>= 0), so the upper word of the float64
// bit pattern orders them up to ties in the lower 32 mantissa bits.  Each warp sorts the KEYS of its 32*VPL values
// (one shuffle and an integer min/max per compare-exchange instead of two shuffles and a float64 select), publishes
// its KC largest, and the key H of block rank K-1 is found as above.  Values with a larger key are among the K
// largest, values with a smaller key are not; if the number of values whose key EQUALS H is exactly what is left
// of K, they all belong to the top and the answer is the sum of the values below H -- exact.  Otherwise (two
// values within 2^-20 of each other straddling rank K: rare) *tie is set and the caller runs the float64 selection.
// `cand`: (nw + 1) * KC 32-bit words.
template <int VPL>
WB_DEV_COLD double wb_sum_without_top_keys(const double (&v)[VPL], double extra, int K, int KC, unsigned* cand, double* scratch,
                                      bool* tie, int tid, int nthr) {
  const int lane = tid & 31, w = tid >> 5, nw = nthr >> 5;
  unsigned key[VPL];
#pragma unroll
  for (int j = 0; j < VPL; ++j) key[j] = (unsigned)__double2hiint(v[j]);
#pragma unroll 1
  for (int size = 2; size <= 32 * VPL; size <<= 1) {
#pragma unroll 1
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride >= VPL) {
        const int ls = stride / VPL;
        const bool is_lo = (lane & ls) == 0;
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
          const bool desc = (((lane * VPL + j) & size) == 0);
          const unsigned o = __shfl_xor_sync(0xffffffffu, key[j], ls);
          key[j] = (is_lo == desc) ? max(key[j], o) : min(key[j], o);
        }
      } else {
#pragma unroll
        for (int st = 1; st < VPL; st <<= 1) {
          if (st == stride) {
#pragma unroll
            for (int j = 0; j < VPL; ++j) {
              if ((j & st) == 0) {
                const bool desc = (((lane * VPL + j) & size) == 0);
                const unsigned a = key[j], b = key[j ^ st];
                const unsigned hi = max(a, b), lo = min(a, b);
                key[j] = desc ? hi : lo;
                key[j ^ st] = desc ? lo : hi;
              }
            }
          }
        }
      }
    }
  }
  // element e = lane * VPL + j of the warp is its e-th largest key
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int e = lane * VPL + j;
    if (e < KC) cand[w * KC + e] = key[j];
  }
  const unsigned xkey = (unsigned)__double2hiint(extra);
  if (tid == 0) cand[nw * KC] = xkey;  // a list of one entry
  __syncthreads();
  for (int ci = tid; ci < nw * KC + 1; ci += nthr) {
    const int cw = ci / KC, i = ci - cw * KC;
    const unsigned c = cand[ci];
    int rank = i;
    for (int l = 0; l <= nw; ++l) {
      if (l == cw) continue;
      const unsigned* L = cand + l * KC;  // descending; equal keys: lower list first
      int lo = 0, hi = l == nw ? 1 : KC;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const unsigned x = L[mid];
        if (l < cw ? x >= c : x > c) lo = mid + 1;
        else hi = mid;
      }
      rank += lo;
    }
    if (rank == K - 1) ((unsigned*)scratch)[2 * (WB_REDUCE_SCRATCH - 1)] = c;
  }
  __syncthreads();
  const unsigned H = ((const unsigned*)scratch)[2 * (WB_REDUCE_SCRATCH - 1)];
  double low = 0.0, n_gt = 0.0, n_eq = 0.0;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const unsigned kj = (unsigned)__double2hiint(v[j]);
    if (kj < H) low += v[j];
    else if (kj > H) n_gt += 1.0;
    else n_eq += 1.0;
  }
  if (tid == 0) {
    if (xkey < H) low += extra;
    else if (xkey > H) n_gt += 1.0;
    else n_eq += 1.0;
  }
  wb_block_sum3(low, n_gt, n_eq, scratch, tid, nthr);
  *tie = n_eq != (double)K - n_gt;
#ifdef WB_D4C_FORCE_TIE  // test builds: always take the float64 selection after the key pass
  *tie = true;
#endif
  return low;
}
// The selection the kernel normally takes (GPU only): no sorting at all.  A threshold key T is searched such that
// between K and WB_D4C_SEL_CAP values have a key >= T -- starting ten octaves below the block's largest key,
// stepping down exponentially or bisecting upwards; every probe is one comparison per value, a warp reduction and one
// shared-memory atomic per warp, one barrier -- the few candidates are compacted into shared memory and ranked
// exactly as float64 (count of larger candidates, ties by position), and the answer is the sum of the values below
// T plus the candidates of rank >= K.  Power spectra of windowed group delays put the K ~ 22 largest of 1025 values
// well apart from the bulk, so one or two probes settle it.  *failed is set (nothing else is valid) when no key
// separates K..CAP values within WB_D4C_SEL_ROUNDS probes (e.g. more than CAP equal values on top); the caller then
// runs the sorting selection.  `buf`: 2 * WB_D4C_SEL_CAP + 16 doubles of shared memory.
#define WB_D4C_SEL_CAP 64
#define WB_D4C_SEL_ROUNDS 12
template <int VPL>
WB_DEV double wb_sum_without_top_search(const double (&v)[VPL], double extra, int K, double* buf, double* scratch,
                                        bool* failed, int tid, int nthr) {
  const int lane = tid & 31;
  double* cand = buf;
  unsigned* ctr = (unsigned*)(buf + WB_D4C_SEL_CAP);  // [0] max key, [1 .. ROUNDS] probe counts, [ROUNDS + 1] candidates
  unsigned key[VPL];
  unsigned kmax = 0;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    key[j] = (unsigned)__double2hiint(v[j]);
    kmax = max(kmax, key[j]);
  }
  const unsigned xkey = (unsigned)__double2hiint(extra);
  if (tid == 0) kmax = max(kmax, xkey);
  if (tid < WB_D4C_SEL_ROUNDS + 2) ctr[tid] = 0u;
  __syncthreads();
  kmax = __reduce_max_sync(0xffffffffu, kmax);
  if (lane == 0) atomicMax(ctr, kmax);
  __syncthreads();
  kmax = ctr[0];
  // bracket: count(key >= t_hi) < K (true at kmax + 1), count(key >= t_lo) > CAP (unknown yet)
  // first probe at max / 1024: the K-th largest power of these spectra sits 2^8.8 below the largest in the median,
  // the 64th 2^11.3 (measured on the bench workload: 1.7 probes on average, 5 at most)
  unsigned t_hi = kmax + 1u, t_lo = 0u, step = 2u << 20, T = kmax > (10u << 20) ? kmax - (10u << 20) : 0u;
  bool have_lo = false, found = false;
  unsigned C = 0;
#pragma unroll 1
  for (int r = 0; r < WB_D4C_SEL_ROUNDS; ++r) {
    unsigned c = 0;
#pragma unroll
    for (int j = 0; j < VPL; ++j) c += key[j] >= T ? 1u : 0u;
    if (tid == 0 && xkey >= T) ++c;
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0 && c) atomicAdd(ctr + 1 + r, c);
    __syncthreads();
    C = ctr[1 + r];
    if (C >= (unsigned)K && C <= (unsigned)WB_D4C_SEL_CAP) {
      found = true;
      break;
    }
    if (C < (unsigned)K) {
      t_hi = T;
      if (have_lo) {
        T = t_lo + ((t_hi - t_lo) >> 1);
      } else {
        T = T > step ? T - step : 0u;
        step <<= 1;
      }
    } else {
      t_lo = T;
      have_lo = true;
      T = t_lo + ((t_hi - t_lo) >> 1);
    }
    if (have_lo && t_hi - t_lo <= 1u) break;  // no key separates K .. CAP values
  }
#ifdef WB_D4C_FORCE_SEARCH_FAIL  // test builds: always continue with the sorting selection
  found = false;
#endif
  *failed = !found;
  if (!found) return 0.0;
  // compact the candidates (key >= T), one atomic per warp and value slot
  double low = 0.0;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const bool in = key[j] >= T;
    const unsigned mask = __ballot_sync(0xffffffffu, in);
    if (mask) {
      unsigned base = 0;
      if (lane == 0) base = atomicAdd(ctr + WB_D4C_SEL_ROUNDS + 1, (unsigned)__popc(mask));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (in) cand[base + __popc(mask & ((1u << lane) - 1u))] = v[j];
    }
    if (!in) low += v[j];
  }
  if (tid == 0) {
    if (xkey >= T) cand[atomicAdd(ctr + WB_D4C_SEL_ROUNDS + 1, 1u)] = extra;
    else low += extra;
  }
  __syncthreads();
  // exact float64 ranks among the C candidates; those of rank >= K stay in the sum.  The compaction order depends on
  // the arrival order of the warps' atomics, so the candidates are first put in rank order and then added by fixed
  // lanes: identical frames give identical bits wherever they sit in the batch.
  double* sorted = cand + WB_D4C_SEL_CAP + 16;  // past the candidates and the counters
  if (tid < (int)C) {
    const double c = cand[tid];
    int rank = 0;
    for (int j = 0; j < (int)C; ++j) {
      const double o = cand[j];
      rank += (o > c || (o == c && j < tid)) ? 1 : 0;
    }
    sorted[rank] = c;
  }
  __syncthreads();
  if (tid < 32) {
    double t = 0.0;
    for (int i = K + tid; i < (int)C; i += 32) t += sorted[i];
    t = wb_warp_sum(t);
    if (tid == 0) low += t;
  }
  return wb_block_sum(low, scratch, tid, nthr);
}
#endif

// One bin of d4c.py:56-59: the band values (dB, <= 0) at the knots interval, 2 interval, ..., between -60 dB at 0 Hz
// and -1e-12 dB at fs/2, linearly interpolated at bin k of the n_spec-point axis and turned into a linear amplitude.
// Shared by the D4C kernel and by wb_d4c_expand so that both produce the same bits.
WB_DEV double wb_d4c_expand_bin(int k, int fs, double inv_nspec, int interval, int n_bands, const double* coarse) {
  const int nk = n_bands + 2;
  const double fq = (double)k * fs * inv_nspec;
  // knots: 0, interval, ..., n_bands*interval, fs/2 ; searchsorted-left then clip to [1, nk-1]
  int hi = 1;
  while (hi < nk - 1 && !((hi <= n_bands ? (double)hi * interval : fs / 2.0) >= fq)) ++hi;
  const int lo = hi - 1;
  const double xl = (double)lo * interval;
  const double xh = hi <= n_bands ? (double)hi * interval : fs / 2.0;
  const double yl = lo == 0 ? -60.0 : coarse[lo - 1];
  const double yh = hi == nk - 1 ? -0.000000000001 : coarse[hi - 1];
  const double v = (yh - yl) / (xh - xl) * (fq - xl) + yl;
  return exp(v * (2.302585092994046 / 20.0));  // 10 ** (v / 20), v in [-60, 0]: exp is a third of pow's cost
}

// aperiodicity [rows, n_spec/2+1] from the 'coarse_ap' transport [rows, n_bands]: a frame whose first band value has
// the sign bit clear (+0.0: unvoiced, or rejected by the love-train gate) gets 1 - 1e-12 everywhere (d4c.py:49-51).
struct wb_d4c_expand_body {
  const double* coarse;
  double* ap;
  long long rows;
  int fs, n_spec, interval, n_bands;
  WB_DEV void operator()(int block, int tid, int nthr, double* smem) const {
    const int bins = n_spec / 2 + 1;
    const double* c = coarse + (size_t)block * n_bands;
    double* o = ap + (size_t)block * bins;
    for (int k = tid; k < n_bands; k += nthr) smem[k] = c[k];
    WB_SYNC();
#ifdef WB_HOST_EMU
    const bool pass = std::signbit(smem[0]);
#else
    const bool pass = (__double2hiint(smem[0]) < 0);
#endif
    const double inv_nspec = 1.0 / n_spec;
    for (int k = tid; k < bins; k += nthr)
      o[k] = pass ? wb_d4c_expand_bin(k, fs, inv_nspec, interval, n_bands, smem) : 1 - 0.000000000001;
  }
};

struct wb_d4c_params {
  // inputs
  const double* x;
  const int* n_samples;
  const double* tpos;
  const double* f0;   // as left by CheapTrick (500 at unvoiced / below-limit frames)
  const double* vuv;
  const int* n_frames;
  const double* band_win;  // nuttall, band_wlen samples (d4c.py:38-39)
  const wb_cplx* tw;
  int tw_n;
  int x_stride, f_stride, fs;
  int n;        // estimator FFT size (d4c.py:20 / d4cRequiem.py:12)
  int n_love;   // love-train FFT size (d4c.py:75)
  int nm;       // capacity (doubles) of each of the two FFT buffers
  int n_spec;   // CheapTrick FFT size, rows of the D4C output (d4c.py:41); unused for requiem
  int interval; // band spacing in Hz
  int n_bands;
  int band_wlen;
  int requiem;  // 0: d4c (linear amplitude over all bins), 1: requiem (dB per band)
  double threshold;
  // outputs
  double* f0_out;   // [B, f_stride]  0 at unvoiced frames (d4c.py:32)
  double* ap;       // d4c: [B, f_stride, n_spec/2+1]; requiem: [B, f_stride, n_bands+2]
  double* coarse;   // d4c only: [B, f_stride, n_bands] (the 'coarse_ap' debug output), may be nullptr

  // capacity (doubles) of each of the two FFT buffers: the largest real FFT (n/2+1 complex) and the longest
  // window the estimator can ask for, 4*T0 at 47 Hz (d4c.py:52,95)
  static int buffer_capacity(int fs, int n, int n_love) {
    int w = 2 * (int)(2.0 * fs / 47.0 + 0.5) + 1;
    int nm = (n > n_love ? n : n_love) + 2;
    nm = nm > w ? nm : w;
    return (nm + 1) & ~1;
  }
  static size_t smem_bytes_tw(int nm, int n, int n_love) {
    return ((size_t)2 * nm + 2 * ((size_t)n / 2 + 2) + WB_REDUCE_SCRATCH + 16 + 48 + 64) * sizeof(double) +
           (size_t)WB_FFT_TW_SLOTS((n > n_love ? n : n_love) / 2) * sizeof(wb_cplx);
  }
};

// NC / NLC / NT: estimator FFT size, love-train FFT size and block size when the launcher knows them at
// compile time (0: run-time values) -- with them every per-bin loop has a constant trip count and the
// transforms are selected without a run-time switch.
template <int NC = 0, int NLC = 0, int NT = 0>
struct wb_d4c_body_t : wb_d4c_params {
  WB_DEV void write_fail(size_t fi, int tid, int nthr) const {
    if (requiem) {
      double* o = ap + fi * (size_t)(n_bands + 2);
      for (int k = tid; k < n_bands + 2; k += nthr) o[k] = -0.000000000001;
    } else {
      const int rows = n_spec / 2 + 1;
      double* o = ap + fi * (size_t)rows;
      if (ap)
        for (int k = tid; k < rows; k += nthr) o[k] = 1 - 0.000000000001;
      if (coarse)
        for (int k = tid; k < n_bands; k += nthr) coarse[fi * (size_t)n_bands + k] = 0.0;
    }
  }

  // Windowed, mean-removed segment (d4c.py:92-110): Ad[i] for i < min(len, limit), zeros up to `fill`.
  // Returns the energy of the full-length segment when want_energy.  Bd is scratch.
  WB_DEV_COLD double segment(const double* xu, int ns, double f, double pos, double span, int kind, double* Ad, double* Bd,
                        int limit, int fill, bool want_energy, double* scratch, int* nz_out, int tid, int nthr) const {
    int len;
    wb_window_sums ws = wb_pitch_window(xu, ns, fs, f, pos, span, kind, true, Bd, Ad, nm, &len, scratch, tid, nthr);
    const double ratio = ws.sw / ws.w;
    const int capb = len < nm ? len : nm;
    double e = 0.0;
    for (int i = tid; i < capb; i += nthr) {
      const double v = Bd[i] - Ad[i] * ratio;
      Ad[i] = v;
      e += v * v;
    }
    if (want_energy) e = wb_block_sum(e, scratch, tid, nthr);
    WB_SYNC();
    // zero padding / truncation to the FFT length `fill`; the transforms' pruned first pass only reads the first
    // quarter or half of a mostly empty buffer, so the padding stops there
    const int cap = len < limit ? len : limit;
    const int stop = wb_rfft_fill(fill, cap);
    for (int i = cap + tid; i < stop; i += nthr) Ad[i] = 0.0;
    WB_SYNC();
    *nz_out = cap;
    return e;
  }

  // The same segment with the window samples held in registers (wb_pitch_window_regs): the mean-removed samples go
  // straight into the transform's input -- as n doubles (zero-filled up to what the pruned first pass reads), or,
  // when Zc is given, as the packed complex sequence (a_i, (i + 1) a_i) / sqrt(energy) of the centroid transform
  // (d4c.py:141-146).  3 barriers (5 with the energy) and no staging traffic; needs len <= WB_D4C_MAXPT * nthr.
  // Returns false (nothing written) when the window is too long for the registers.
#define WB_D4C_MAXPT 6
  WB_DEV bool segment_fast(const double* xu, int ns, double f, double pos, double span, int kind, double* Ad, wb_cplx* Zc,
                           int fill, double* scratch, int* nz_out, int tid, int nthr) const {
    const int len = 2 * (int)(span * fs / f + 0.5) + 1;
    if (len > WB_D4C_MAXPT * nthr || len > fill) return false;
    double sw[WB_D4C_MAXPT], w[WB_D4C_MAXPT];
    const wb_window_sums ws = wb_pitch_window_regs<WB_D4C_MAXPT>(xu, ns, fs, f, pos, span, kind, true, sw, w, scratch, tid, nthr);
    const double ratio = ws.sw / ws.w;
    double e = 0.0;
#pragma unroll
    for (int c = 0; c < WB_D4C_MAXPT; ++c) {
      sw[c] = sw[c] - w[c] * ratio;  // zero beyond the window (both factors are)
      e += sw[c] * sw[c];
    }
    const int stop = wb_rfft_fill(fill, len);
    if (Zc) {
      e = wb_block_sum(e, scratch, tid, nthr);
      const double inv = 1.0 / sqrt(e);
#pragma unroll
      for (int c = 0; c < WB_D4C_MAXPT; ++c) {
        const int i = tid + c * nthr;
        if (i < stop) {
          const double a = sw[c] * inv;
          Zc[i] = wb_mk(a, a * (double)(i + 1));
        }
      }
    } else {
#pragma unroll
      for (int c = 0; c < WB_D4C_MAXPT; ++c) {
        const int i = tid + c * nthr;
        if (i < stop) Ad[i] = sw[c];
      }
    }
    for (int i = WB_D4C_MAXPT * nthr + tid; i < stop; i += nthr) {  // zero padding past the register tile
      if (Zc) Zc[i] = wb_mk(0.0, 0.0);
      else Ad[i] = 0.0;
    }
    WB_SYNC();
    *nz_out = len;
    return true;
  }

#ifndef WB_HOST_EMU
  template <int VPL>
  WB_DEV double band_select(const wb_cplx* X, double extra, int K, int KC, double* cand, double* scratch, double& tot,
                            int tid, int nthr) const {
    double pv[VPL];
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
      const wb_cplx z = X[tid + q * nthr];
      pv[q] = z.x * z.x + z.y * z.y;
      t += pv[q];
    }
    tot = wb_block_sum(t, scratch, tid, nthr) + extra;
    bool failed;
    const double low_s = wb_sum_without_top_search<VPL>(pv, extra, K, cand, scratch, &failed, tid, nthr);
    if (!failed) return low_s;
    __syncthreads();
    bool tie;
    const double low = wb_sum_without_top_keys<VPL>(pv, extra, K, KC, (unsigned*)cand, scratch, &tie, tid, nthr);
    if (!tie) return low;
    __syncthreads();
    return wb_sum_without_top<VPL>(pv, extra, K, KC, cand, scratch, tid, nthr);
  }
#endif

  // band value through the full sort (shapes the selection does not cover, and the host emulation)
  WB_DEV_COLD void bandv_by_sort(const wb_cplx* X, double* V, int b, int m_low, int nh, double* bandv, double* scratch,
                                 int tid, int nthr) const {
    double tot = 0.0;
    for (int k = tid; k < nh; k += nthr) {
      const double pw = X[k].x * X[k].x + X[k].y * X[k].y;
      tot += pw;
      V[k] = pw;
    }
    const double extra = X[nh].x * X[nh].x + X[nh].y * X[nh].y;
    tot = wb_block_sum(tot, scratch, tid, nthr) + extra;
    WB_SYNC();
    wb_bitonic_sort(V, nh, tid, nthr);
    const bool extra_in = (m_low >= 1) && (extra < V[m_low - 1]);
    const int take = extra_in ? m_low - 1 : m_low;
    double low = 0.0;
    for (int k = tid; k < take; k += nthr) low += V[k];
    low = wb_block_sum(low, scratch, tid, nthr);
    if (extra_in) low += extra;
    if (tid == 0) bandv[b] = -10.0 * log10(low / tot);
    WB_SYNC();
  }

  WB_DEV void operator()(int block, int tid, int nthr_rt, double* smem) const {
    const int nthr = NT ? NT : nthr_rt;
    const int n = NC ? NC : this->n;
    const int n_love = NLC ? NLC : this->n_love;
    const int u = block / f_stride, f = block - u * f_stride;
    if (f >= n_frames[u]) return;
    const int nh = n / 2;
    double* Ad = smem;               // nm doubles (>= n + 2): FFT buffer / window / scratch
    double* Bd = Ad + nm;            // nm doubles; Ad..Bd contiguous = one buffer of >= n complex
    double* R1 = Bd + nm;            // nh + 1 (+1 pad)
    double* R2 = R1 + (nh + 2);      // nh + 1 (+1 pad)
    double* scratch = R2 + (nh + 2);              // WB_REDUCE_SCRATCH
    double* bandv = scratch + WB_REDUCE_SCRATCH;  // up to 16 band values
    double* carry = bandv + 16;                   // 48
    wb_cplx* twS = (wb_cplx*)(carry + 48);
    const int twH = (n > n_love ? n : n_love) / 2;
    wb_cplx* A = (wb_cplx*)Ad;
    wb_cplx* B = (wb_cplx*)Bd;
    const size_t fi = (size_t)u * f_stride + f;
    const double* xu = x + (size_t)u * x_stride;
    const int ns = n_samples[u];
    const double pos = tpos[fi];
    const double f0v = (vuv[fi] == 0.0) ? 0.0 : f0[fi];
    if (tid == 0) f0_out[fi] = f0v;
    if (f0v == 0.0) {
      write_fail(fi, tid, nthr);
      return;
    }
    wb_fft_load_twiddles(twS, twH, tw, tw_n, tid, nthr);

    // The frame goes through 4 + n_bands steps that share ONE call site each for the window, the real transform
    // and the smoothing (the steps differ in parameters, not in code): the kernel's hot instructions then fit the
    // instruction cache, where three inlined copies of each did not (10 % of the warp stalls were instruction
    // fetches).  Steps: 0 love train (d4c.py:68-88), 1 / 2 the two centroid windows (d4c.py:132-153), 3 smoothed
    // power spectrum (d4c.py:157-161); before step 4 the group-delay shaping (d4c.py:165-188); 4.. one band each
    // (d4c.py:192-209).
    const double cf = wb_dmax(47.0, f0v);
    const int ln = wb_fft_log2(n);
    const double inv_cf = 1.0 / cf;
    double* S = Ad;    // prefix sums of the smoothings (<= n doubles)
    double* R3 = Bd;   // nh + 1 doubles
    const int boundary = (int)((double)n / band_wlen * 8 + 0.5);
    const int hw = band_wlen / 2;
    // The nh+1 power values are sorted as nh (a power of two) plus one extra value x = P[nh]:
    // the sum of the m smallest of the union is  sum(sorted[0..m))      if x >= sorted[m-1]
    //                                            sum(sorted[0..m-1)) + x  otherwise.
    const int m_low = nh - boundary;  // cumsum index nh - boundary - 1 of d4c.py:207-208
#ifdef WB_HOST_EMU
    const int nthr_fft = 1 << 20;  // the emulation plays every thread of the in-place transform itself
#else
    const int nthr_fft = nthr;
#endif
    for (int step = 0; step < 4 + n_bands; ++step) {
      const bool centroid = step == 1 || step == 2;
      if (step == 4) {
        // ---- low-band replicas, smoothed power, group delay shaping -------------------------------------
        for (int r = 0; r < 2; ++r) wb_mirror_low_band(r == 0 ? R1 : R2, n, fs, cf, 1.2 * cf, r == 0 ? Ad : Bd, tid, nthr);
        // three running-integral smoothings; the element-wise steps between them ride on their load / store sides
        // (divisions by per-frame constants as multiplications by the reciprocal: last-bit differences)
        for (int it = 0; it < 3; ++it) {
          wb_box_integral_f([&](int j) { return it == 1 ? R1[j] * cf / R3[j] : R2[j]; }, n, fs, it == 1 ? cf / 4.0 : cf / 2.0, S,
                            carry,
                            [&](int k, double v) {
                              if (it == 0) R3[k] = v;                         // smoothed power (d4c.py:160-161)
                              else if (it == 1) R2[k] = v * (2.0 * inv_cf);   // d4c.py:168-170
                              else R2[k] = R2[k] - v * inv_cf;                // d4c.py:172-174
                            },
                            tid, nthr);
        }
      }
      // ---- input of the step's transform -----------------------------------------------------------------
      int nz = 0x7ffffffe;
      const int n_step = step == 0 ? n_love : n;
      if (step < 4) {
        const double f = step == 0 ? wb_dmax(f0v, 40.0) : cf;
        const double p2 = step == 1 ? pos + 1.0 / cf / 4.0 : (step == 2 ? pos - 1.0 / cf / 4.0 : pos);
        const double span = step == 0 ? 1.5 : 2.0;
        const int kind = step == 3 ? WB_WIN_HANN : WB_WIN_BLACKMAN;
        if (!segment_fast(xu, ns, f, p2, span, kind, Ad, centroid ? A : nullptr, n_step, scratch, &nz, tid, nthr)) {
          const double e = segment(xu, ns, f, p2, span, kind, Ad, Bd, n_step, n_step, centroid, scratch, &nz, tid, nthr);
          if (centroid) {
            const int nb = wb_rfft_fill(n, nz);  // entries the pruned transform reads
            const double inv = 1.0 / sqrt(e);
            // expand a_i -> (a_i, (i+1) a_i) in place, top tile first so that no source is overwritten early
            const int tile = nthr * 8;
            for (int t1 = ((nb + tile - 1) / tile) * tile; t1 > 0; t1 -= tile) {
              const int t0 = t1 - tile;
              double reg[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const int i = t0 + q * nthr + tid;
                reg[q] = i < nb ? Ad[i] * inv : 0.0;
              }
              WB_SYNC();
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const int i = t0 + q * nthr + tid;
                if (i < nb) A[i] = wb_mk(reg[q], reg[q] * (double)(i + 1));
              }
              WB_SYNC();
            }
          }
        }
      } else {  // nuttall-windowed slice of the group delay around the band centre (d4c.py:196-203)
        const int centre = (int)floor((double)interval * (step - 3) / ((double)fs / n));
        for (int i = tid; i < n; i += nthr) {
          double v = 0.0;
          if (i < band_wlen) {
            int j = centre - hw + i;
            j &= (n - 1);
            v = (j <= nh ? R2[j] : R2[n - j]) * WB_LDG(band_win + i);
          }
          Ad[i] = v;
        }
        WB_SYNC();
      }
      if (centroid) {
        // the spectra of x and of n*x come from ONE complex transform of x + i n x.  Natural-order output: ping-pong
        // over the two buffers when each holds n complex entries (n < n_love), else in place with the registers as
        // the staging area (a thread per radix-8 butterfly); the radix-4 decimation-in-frequency transform
        // (bit-reversed output) serves the remaining shapes
        bool nat = true;
        const wb_cplx* Zr = A;
        if (2 * n <= nm) Zr = wb_fft<0, NC>(A, B, n, -1, twS, twH, tid, nthr, nz);
        else if (n == 2048 && nthr_fft >= 256) wb_fft_inplace_nat<2048>(A, twS, twH, tid, nthr, nz);
        else if (n == 4096 && nthr_fft >= 512) wb_fft_inplace_nat<4096>(A, twS, twH, tid, nthr, nz);
        else {
          nat = false;
          wb_fft_inplace_dif(A, n, twS, twH, tid, nthr, nz);
        }
        for (int k = tid; k <= nh; k += nthr) {
          const int kn = (n - k) & (n - 1);
          const wb_cplx z = Zr[nat ? k : wb_bitrev(k, ln)], y = Zr[nat ? kn : wb_bitrev(kn, ln)];
          const double ar = 0.5 * (z.x + y.x), ai = 0.5 * (z.y - y.y);
          const double br = 0.5 * (z.y + y.y), bi = -0.5 * (z.x - y.x);
          const double c = br * ar + ai * bi;
          R1[k] = step == 1 ? c : R1[k] + c;
        }
        WB_SYNC();
        continue;
      }
      // ---- real transform of the step (one call site when the love-train size equals the estimator's) --------
      const wb_cplx* X;
      if (NC != NLC && step == 0) X = wb_rfft<0, NLC>(A, B, n_love, twS, twH, tid, nthr, nz);
      else X = wb_rfft<0, NC>(A, B, n_step, twS, twH, tid, nthr, nz);
      if (step == 0) {  // love train: power below 4 kHz against power below 7.9 kHz (d4c.py:76-88)
        const double dfl = (double)fs / n_love;
        const int b0 = (int)(ceil(100.0 / dfl) + 1);
        const int b1 = (int)(ceil(4000.0 / dfl) + 1);
        const int b2 = (int)(ceil(7900.0 / dfl) + 1);
        double s1 = 0.0, s2 = 0.0, s3 = 0.0;
        const int top = b2 < n_love ? b2 : n_love;
        const int hl = n_love / 2;
        for (int k = b0 + tid; k < top; k += nthr) {
          const wb_cplx z = X[k <= hl ? k : n_love - k];
          const double pw = z.x * z.x + z.y * z.y;
          s2 += pw;
          if (k < b1) s1 += pw;
        }
        wb_block_sum3(s1, s2, s3, scratch, tid, nthr);
        if (!((s1 / s2) > threshold)) {
          write_fail(fi, tid, nthr);
          return;
        }
        WB_SYNC();
        continue;
      }
      if (step == 3) {
        for (int k = tid; k <= nh; k += nthr) R2[k] = X[k].x * X[k].x + X[k].y * X[k].y;
        WB_SYNC();
        continue;
      }
      // ---- band aperiodicity: all but the boundary + 1 largest of the nh + 1 power values -------------------
      const int b = step - 4;
      double* V = (X == A) ? Bd : Ad;
#ifndef WB_HOST_EMU
      {  // by selection instead of a full sort
        const int K = boundary + 1, vpl = nh / nthr;
        const int KC = (K + 7) & ~7;
        if (vpl * nthr == nh && (nthr & 31) == 0 && K >= 1 && KC <= 32 * vpl && (vpl == 2 || vpl == 4 || vpl == 8)) {
          const double extra = X[nh].x * X[nh].x + X[nh].y * X[nh].y;
          double low, tot = 0.0;
          if (vpl == 2) low = band_select<2>(X, extra, K, KC, V, scratch, tot, tid, nthr);
          else if (vpl == 4) low = band_select<4>(X, extra, K, KC, V, scratch, tot, tid, nthr);
          else low = band_select<8>(X, extra, K, KC, V, scratch, tot, tid, nthr);
          if (tid == 0) bandv[b] = -10.0 * log10(low / tot);
          WB_SYNC();
          continue;
        }
      }
#endif
      bandv_by_sort(X, V, b, m_low, nh, bandv, scratch, tid, nthr);
    }

    // ---- outputs ----------------------------------------------------------------------
    if (requiem) {  // d4cRequiem.py:26-40
      double* o = ap + fi * (size_t)(n_bands + 2);
      for (int k = tid; k < n_bands + 2; k += nthr) {
        double v;
        if (k == 0) v = -60.0;
        else if (k == n_bands + 1) v = -0.000000000001;
        else v = -wb_dmax(0.0, bandv[k - 1] - (cf - 100.0) * 2.0 / 100.0);
        o[k] = v;
      }
    } else {  // d4c.py:56-59
      const int rows = n_spec / 2 + 1;
      double* o = ap + fi * (size_t)rows;
      const double adj = (cf - 100.0) * 2.0 / 100.0;
      WB_SYNC();
      for (int k = tid; k < n_bands; k += nthr) {  // the 'coarse_ap' values: always sign-bit set (-0.0 when clamped)
        const double c = -wb_dmax(0.0, bandv[k] - adj);
        bandv[k] = c;
        if (coarse) coarse[fi * (size_t)n_bands + k] = c;
      }
      WB_SYNC();
      const double inv_nspec = 1.0 / n_spec;  // n_spec is a power of two: exact
      if (ap)  // the caller may take only the band values and expand later (wb_d4c_expand)
        for (int k = tid; k < rows; k += nthr) o[k] = wb_d4c_expand_bin(k, fs, inv_nspec, interval, n_bands, bandv);
    }
  }
};
typedef wb_d4c_body_t<> wb_d4c_body;
