// Harvest F0 estimator -- kernel bodies.  Replaces world/harvest.py:17-54.
//
//   H1 hv_decimate   zero-phase Chebyshev decimation to ~8 kHz        (harvest.py:58-71, 584-609)
//   H2 hv_channels   152-channel band-pass as a direct FIR on shared-memory tiles, fused with the
//                    four zero-crossing event streams and their interpolation onto the 1 ms grid
//                                                                      (harvest.py:75-84, 252-297, 499-529)
//   H3 hv_detect     runs of >= 10 consecutive channels -> base candidates (harvest.py:88-110)
//   H4 hv_refine_*   instantaneous-frequency refinement of every candidate offered to a frame
//                    (own frame and frames +-3), direct DFT at <= 6 harmonic bins instead of two FFTs;
//                    work items counting-sorted by window length, one thread per candidate
//                                                                      (harvest.py:114-125, 131-150, 169-211)
//   H5 hv_prune      neighbour-consistency pruning                    (harvest.py:215-234)
//   H6 hv_contour    base contour, 4 fix steps, smoothing, 5 ms pick   (harvest.py:301-495, 533-559, 46-53)
//
// The reference materialises 152 filtered signals per utterance through 65 536-point FFTs
// (harvest.py:259-262).  Here the filtered samples never leave shared memory: HBM sees the
// decimated signal, the [channel, frame] candidate map and the per-frame candidate lists.
//
// Candidate lists: the reference keeps a dense [7*max_candidates, frames] matrix whose row order
// only matters for tie-breaking.  Here each frame holds a compact list of (f0, score, slot) with
// slot = shift_index*15 + candidate_index, which preserves the reference's row order, and the
// tie rules are applied on the slot (first maximum / last minimum).
#pragma once
#include "wb_fft.h"

#define WB_HV_MAXC 15    // int(152/10 + 0.5) rows of DetectCandidates (harvest.py:90)
#define WB_HV_SLOTS 105  // 7 shifts * 15
#define WB_HV_TILE 2048  // filtered samples per tile
#define WB_HV_FFT_GROUPS 3  // block-overlap classes of the overlap-save path
#define WB_HV_FPT 8      // consecutive 1 ms frames per thread when the event streams are interpolated
#define WB_HV_OPT 8      // outputs per thread in the FIR (TILE / OPT = 256 threads per block; 16 measured slower)

struct wb_hv_plan {
  int batch, fs, ratio, pad;
  double afs, f0_floor, f0_ceil, frame_period;
  int n_ch, max_taps;
  const double* edges;  // [n_ch] boundary F0 of each channel
  const int* halfs;     // [n_ch] filter length L of each channel (2*half+1 for Harvest)
  const int* ch_off;    // [n_ch] signal index of the first (reversed) tap for output sample 0
  int wrap_n;           // > 0: DIO's FFT filtering: the signal is zero-padded to 2^ceil(log2(len + wrap_n)) and circular
  const int* pow2_quirk;  // [32] ceil(log(2^k)/log(2)) as the host's libm evaluates it (math.log(x, 2), dio.py:78)
  int mode;             // 0 Harvest (mean of 4, +-10 % gate), 1 DIO (mean and std of 4, octave gate)
  double grid_ms;       // frame grid the event streams are interpolated onto
  double* stab;         // DIO: [B, n_ch, f1_stride] stability score
  double* four;         // DIO: [B, n_ch, 4, f1_stride] the four interpolated streams
  const int* tap_off;   // [n_ch] offset into taps
  const double* taps;   // reversed taps of every channel, concatenated
  const double* cb;     // decimation filter b[4], a[4], zi[3], then H[3][CHUNK] and M[3][3]
  // inputs
  const double* x;
  const int* n_samples;
  int x_stride;
  // workspace
  double* fwd;       // [B, ext_stride]  forward pass, zero-state per chunk
  double* bwd;       // [B, ext_stride]  backward pass, zero-state per chunk
  int ext_stride;
  int dec_kind;      // 0: scipy lfilter form, steady-state seed (Harvest); 1: DIO's form, zero seed, no padding
  int dec_chunks;    // chunks per utterance (stride of the state arrays)
  double* dec_s1;    // [B, dec_chunks, 3] zero-state end state of each forward chunk
  double* dec_init;  // [B, dec_chunks, 3] true state entering each forward chunk
  double* dec_s2;    // same for the backward pass
  double* dec_initb;
  double* y;         // [B, y_stride]
  int* y_len;        // [B]
  int y_stride;
  int f1_stride;     // stride of the 1 ms frame axis
  double* raw;       // [B, n_ch, f1_stride]
  double* edge_buf;  // [n_slots, 4, edge_cap]
  int edge_cap, n_slots;
  double* base_c;    // [B, f1_stride, WB_HV_MAXC]
  int* base_n;       // [B, f1_stride]
  double* l_f0;      // [B, f1_stride, WB_HV_SLOTS]
  double* l_sc;
  unsigned char* l_slot;
  unsigned char* l_keep;
  int* l_n;          // [B, f1_stride]
  double* ctr;       // contour scratch, [B, ctr_stride]
  long long ctr_stride;
  // overlap-save path of the long band-pass filters (channels [0, fft_nch); 0 = everything by direct FIR)
  int fft_nch;             // channels handled by wb_hv_channels_fft
  // The channels are split into up to WB_HV_FFT_GROUPS contiguous groups by filter length; a group's blocks
  // overlap by twice ITS longest half length only, so the shorter filters get more output samples per block.
  int fft_groups;
  int fft_gc[WB_HV_FFT_GROUPS + 1];   // first channel of each group; fft_gc[fft_groups] = fft_nch
  int fft_gV[WB_HV_FFT_GROUPS];       // output positions per block
  int fft_gA[WB_HV_FFT_GROUPS];       // signal index of block 0's first sample
  int fft_gblocks[WB_HV_FFT_GROUPS];  // signal blocks per utterance
  long long fft_goff[WB_HV_FFT_GROUPS];  // offset (complex entries) of the group's spectra in fft_Y
  int fft_blocks;          // sum of fft_gblocks
  const wb_cplx* fft_H;    // [fft_nch, N/2+1] conj(FFT(taps at their offset)) / N
  wb_cplx* fft_Y;          // per group [B, fft_gblocks[g], N/2+1] block spectra of the decimated signal
  int* status;       // [1] sticky error flags (bit0 edge overflow, bit1 track pool overflow)
  // outputs
  double* out_tpos;  // [B, f_stride]
  double* out_f0;
  double* out_vuv;
  int* out_n_frames;  // [B]
  int f_stride;
};

WB_HD int wb_hv_frames(int n_samples, int fs, double period_ms) {
  return (int)(1000.0 * n_samples / fs / period_ms + 1);
}

// ------------------------------------------------------------------------------------ H1
// Zero-phase decimation filter = scipy.signal.filtfilt(cheby1(3, .05, .8/r), padlen=9) as called at
// harvest.py:601: odd extension by 9 samples, forward pass seeded with zi*first, backward pass seeded
// with zi*last.  The recursion is linear, so each pass is split into chunks of WB_HV_CHUNK samples
// that run concurrently from a zero state (D1/D3); the true state at every chunk boundary follows
// from a short per-utterance recurrence st' = M st + s_chunk (D2/D4), and the zero-input response of
// that state (3 tabulated basis sequences H) is added when the samples are consumed (D3/D5).
#define WB_HV_CHUNK 256

struct wb_hv_dec_common {
  wb_hv_plan p;
  WB_DEV double padded(const double* xu, int ns, int k) const {  // edge-replicated input (harvest.py:66)
    const int i = k - p.pad;
    return xu[i < 0 ? 0 : (i >= ns ? ns - 1 : i)];
  }
  WB_DEV double extended(const double* xu, int ns, int nd, int i) const {
    if (i < 9) return 2.0 * padded(xu, ns, 0) - padded(xu, ns, 9 - i);
    if (i < 9 + nd) return padded(xu, ns, i - 9);
    const int j = i - (9 + nd);
    return 2.0 * padded(xu, ns, nd - 1) - padded(xu, ns, nd - 2 - j);
  }
  WB_DEV bool passthrough() const { return p.dec_kind == 0 && p.fs <= 8000; }  // harvest.py:61-63
  // The filter coefficients live in global memory and every pass stores between its loads, so the compiler may not
  // keep them in registers on its own: each pass copies them once.
  struct coefs {
    double c0, c1, c2, c3, c5, c6, c7;
  };
  WB_DEV coefs load_coefs() const {
    coefs c;
    c.c0 = WB_LDG(p.cb + 0);
    c.c1 = WB_LDG(p.cb + 1);
    c.c2 = WB_LDG(p.cb + 2);
    c.c3 = WB_LDG(p.cb + 3);
    c.c5 = WB_LDG(p.cb + 5);
    c.c6 = WB_LDG(p.cb + 6);
    c.c7 = WB_LDG(p.cb + 7);
    return c;
  }
  // one sample of the 3rd-order recursion; (s0, s1, s2) is the filter state
  WB_DEV double step(const coefs& c, double e, double& s0, double& s1, double& s2) const {
    if (p.dec_kind == 0) {  // direct form II transposed (scipy.signal.lfilter)
      const double o = c.c0 * e + s0;
      s0 = c.c1 * e - c.c5 * o + s1;
      s1 = c.c2 * e - c.c6 * o + s2;
      s2 = c.c3 * e - c.c7 * o;
      return o;
    }
    // FilterForDecimate (dio.py:438-446): cb[5..7] = a0..a2, cb[0] = b0, cb[1] = b1
    const double wt = e + c.c5 * s0 + c.c6 * s1 + c.c7 * s2;
    const double o = c.c0 * wt + c.c1 * s0 + c.c1 * s1 + c.c0 * s2;
    s2 = s1;
    s1 = s0;
    s0 = wt;
    return o;
  }
  // forward value with the zero-input response of the chunk's true initial state (st0, st1, st2) added
  WB_DEV double fwd_value_at(int u, int i, int n, double st0, double st1, double st2) const {
    const double* H = p.cb + 11;
    return p.fwd[(size_t)u * p.ext_stride + i] + st0 * WB_LDG(H + n) + st1 * WB_LDG(H + WB_HV_CHUNK + n) +
           st2 * WB_LDG(H + 2 * WB_HV_CHUNK + n);
  }
  WB_DEV double fwd_value(int u, int i) const {
    const int k = i / WB_HV_CHUNK, n = i - k * WB_HV_CHUNK;
    const double* st = p.dec_init + ((size_t)u * p.dec_chunks + k) * 3;
    return fwd_value_at(u, i, n, st[0], st[1], st[2]);
  }
};

struct wb_hv_dec_fwd : wb_hv_dec_common {  // D1: one thread per (utterance, chunk)
  WB_DEV void operator()(long long item) const {
    const int u = (int)(item / p.dec_chunks), k = (int)(item - (long long)u * p.dec_chunks);
    if (passthrough()) return;
    const int ns = p.n_samples[u], nd = ns + 2 * p.pad, ne = nd + 18;
    const int lo = k * WB_HV_CHUNK, hi = wb_imin(ne, lo + WB_HV_CHUNK);
    if (lo >= ne) return;
    const double* xu = p.x + (size_t)u * p.x_stride;
    double* f = p.fwd + (size_t)u * p.ext_stride;
    const coefs c = load_coefs();
    double z0 = 0.0, z1 = 0.0, z2 = 0.0;
    for (int i = lo; i < hi; ++i) f[i] = step(c, extended(xu, ns, nd, i), z0, z1, z2);
    double* s = p.dec_s1 + ((size_t)u * p.dec_chunks + k) * 3;
    s[0] = z0;
    s[1] = z1;
    s[2] = z2;
  }
};

struct wb_hv_dec_scan : wb_hv_dec_common {  // D2 (backward = 0) / D4 (backward = 1): one thread per utterance
  int backward;
  WB_DEV void operator()(long long item) const {
    const int u = (int)item;
    if (passthrough()) return;
    const int ns = p.n_samples[u], nd = ns + 2 * p.pad, ne = nd + 18;
    const int nck = (ne + WB_HV_CHUNK - 1) / WB_HV_CHUNK;
    // M[r*3+c]: state r after a full chunk from unit state c.  Matrix and zero-state chunk results are read through
    // the read-only path (written by earlier launches), four chunks ahead of the recursion that consumes them.
    const double* Mg = p.cb + 11 + 3 * WB_HV_CHUNK;
    double M[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) M[q] = WB_LDG(Mg + q);
    double st0, st1, st2;
    if (!backward) {
      const double e0 = extended(p.x + (size_t)u * p.x_stride, ns, nd, 0);
      st0 = p.cb[8] * e0;
      st1 = p.cb[9] * e0;
      st2 = p.cb[10] * e0;
      const double* sb = p.dec_s1 + (size_t)u * p.dec_chunks * 3;
      double* ob = p.dec_init + (size_t)u * p.dec_chunks * 3;
      for (int k0 = 0; k0 < nck; k0 += 4) {
        double s[12];
#pragma unroll
        for (int q = 0; q < 12; ++q) s[q] = k0 * 3 + q < nck * 3 ? WB_LDG(sb + k0 * 3 + q) : 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (k0 + j < nck) {
            double* o = ob + (size_t)(k0 + j) * 3;
            o[0] = st0;
            o[1] = st1;
            o[2] = st2;
            const double n0 = M[0] * st0 + M[1] * st1 + M[2] * st2 + s[3 * j + 0];
            const double n1 = M[3] * st0 + M[4] * st1 + M[5] * st2 + s[3 * j + 1];
            const double n2 = M[6] * st0 + M[7] * st1 + M[8] * st2 + s[3 * j + 2];
            st0 = n0;
            st1 = n1;
            st2 = n2;
          }
        }
      }
    } else {
      const double e0 = fwd_value(u, ne - 1);
      st0 = p.cb[8] * e0;
      st1 = p.cb[9] * e0;
      st2 = p.cb[10] * e0;
      const coefs c = load_coefs();
      const double* sb = p.dec_s2 + (size_t)u * p.dec_chunks * 3;
      double* ob = p.dec_initb + (size_t)u * p.dec_chunks * 3;
      for (int k0 = nck - 1; k0 >= 0; k0 -= 4) {
        double s[12];  // chunk k0 - j in s[3 j ..]
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int q = 0; q < 3; ++q) s[3 * j + q] = k0 - j >= 0 ? WB_LDG(sb + (size_t)(k0 - j) * 3 + q) : 0.0;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = k0 - j;
          if (k >= 0) {
            double* o = ob + (size_t)k * 3;
            o[0] = st0;
            o[1] = st1;
            o[2] = st2;
            const int len = wb_imin(ne, (k + 1) * WB_HV_CHUNK) - k * WB_HV_CHUNK;
            double n0, n1, n2;
            if (len == WB_HV_CHUNK) {
              n0 = M[0] * st0 + M[1] * st1 + M[2] * st2;
              n1 = M[3] * st0 + M[4] * st1 + M[5] * st2;
              n2 = M[6] * st0 + M[7] * st1 + M[8] * st2;
            } else {  // the short chunk at the far end: run the zero-input recursion
              n0 = st0;
              n1 = st1;
              n2 = st2;
              for (int i = 0; i < len; ++i) step(c, 0.0, n0, n1, n2);
            }
            st0 = n0 + s[3 * j + 0];
            st1 = n1 + s[3 * j + 1];
            st2 = n2 + s[3 * j + 2];
          }
        }
      }
    }
  }
};

struct wb_hv_dec_bwd : wb_hv_dec_common {  // D3: one thread per (utterance, chunk), time-reversed pass
  WB_DEV void operator()(long long item) const {
    const int u = (int)(item / p.dec_chunks), k = (int)(item - (long long)u * p.dec_chunks);
    if (passthrough()) return;
    const int ns = p.n_samples[u], nd = ns + 2 * p.pad, ne = nd + 18;
    const int lo = k * WB_HV_CHUNK, hi = wb_imin(ne, lo + WB_HV_CHUNK);
    if (lo >= ne) return;
    double* g = p.bwd + (size_t)u * p.ext_stride;
    const coefs c = load_coefs();
    const double* st = p.dec_init + ((size_t)u * p.dec_chunks + k) * 3;  // the whole range lies in chunk k
    const double t0 = st[0], t1 = st[1], t2 = st[2];
    double z0 = 0.0, z1 = 0.0, z2 = 0.0;
    for (int i = hi - 1; i >= lo; --i) g[i] = step(c, fwd_value_at(u, i, i - lo, t0, t1, t2), z0, z1, z2);
    double* s = p.dec_s2 + ((size_t)u * p.dec_chunks + k) * 3;
    s[0] = z0;
    s[1] = z1;
    s[2] = z2;
  }
};

// D5: one block per utterance: pick the decimated samples (decimate_matlab phase, harvest.py:605-609, and
// the trim of harvest.py:70), remove the mean (harvest.py:71).
struct wb_hv_dec_pick : wb_hv_dec_common {
  WB_DEV void operator()(int block, int tid, int nthr, double* smem) const {
    const int u = block;
    const double* xu = p.x + (size_t)u * p.x_stride;
    const int ns = p.n_samples[u];
    double* yu = p.y + (size_t)u * p.y_stride;
    int ylen;
    double total = 0.0;
    if (passthrough()) {
      ylen = ns;
      for (int i = tid; i < ns; i += nthr) {
        yu[i] = xu[i];
        total += xu[i];
      }
    } else {
      const int r = p.ratio;
      const int nd = ns + 2 * p.pad, ne = nd + 18;
      int i_first, trim = 0;
      if (p.dec_kind == 0) {
        const int n_out = (nd + r - 1) / r;
        const int first = r - (r * n_out - nd);  // 1-based
        const int m_count = (nd - first) / r + 1;
        trim = p.pad / r;
        ylen = m_count - 2 * trim;
        i_first = 9 + (first - 1);
      } else {  // dio.py:470-476: nout = ceil(len/r + 1), nbeg = r - r*nout + len, samples nbeg + 8 + m*r
        const int n_out = (int)ceil((double)ns / r + 1.0);
        const int nbeg = r - r * n_out + ns;
        ylen = (ns + 9 - nbeg + r - 1) / r;
        i_first = nbeg + 8;
      }
      const double* g = p.bwd + (size_t)u * p.ext_stride;
      const double* H = p.cb + 11;
      for (int m = tid; m < ylen; m += nthr) {
        int i = i_first + (m + trim) * r;
        if (i < 0) i += ne;  // a negative NumPy index counts from the end
        const int k = i / WB_HV_CHUNK;
        const int hi = wb_imin(ne, (k + 1) * WB_HV_CHUNK);
        const int n = hi - 1 - i;  // steps since this chunk's (backward) start
        const double* st = p.dec_initb + ((size_t)u * p.dec_chunks + k) * 3;
        const double v = g[i] + st[0] * H[n] + st[1] * H[WB_HV_CHUNK + n] + st[2] * H[2 * WB_HV_CHUNK + n];
        yu[m] = v;
        total += v;
      }
    }
    if (ylen < 0) ylen = 0;
    total = wb_block_sum(total, smem, tid, nthr);
    const double mean = (ylen > 0 && p.dec_kind == 0) ? total / ylen : 0.0;  // DIO keeps the mean (dio.py:38)
    WB_SYNC();
    for (int i = tid; i < ylen; i += nthr) yu[i] -= mean;
    if (tid == 0) {
      p.y_len[u] = ylen;
      p.out_n_frames[u] = wb_hv_frames(ns, p.fs, p.frame_period);
    }
  }
};

// ------------------------------------------------------------------------------------ H2
// Persistent blocks; work item = (channel, utterance).  Two kernels share the event detection and the
// interpolation onto the frame grid:
//   wb_hv_channels      direct FIR on shared-memory tiles (short filters; all of DIO's bands)
//   wb_hv_channels_fft  overlap-save through 2048-point real FFTs in shared memory (Harvest's long filters):
//                       the block spectra of the signal are computed once per utterance (wb_hv_fft_fwd) and
//                       shared by all channels; a channel multiplies by its tabulated response and inverts.
struct wb_hv_channels_common {
  wb_hv_plan p;

#ifndef WB_HOST_EMU
  // Events of one tile.  Every thread owns WB_HV_OPT consecutive filtered samples sv[0..OPT) starting at tile
  // position tid * OPT, plus the next two (sv[OPT], sv[OPT+1]).  Positions m in [0, tl) are examined; position m
  // is sample n = t0 + m.  Stream 0/1: falling/rising zero crossings of the filtered signal, stream 2/3: of its
  // first difference (ZeroCrossingEngine, harvest.py:283-297).  `sb` (tile samples in shared memory) is filled
  // from the registers unless it already holds them; plist: [4][WB_HV_TILE] ushort; wsum: nthr/32 + 1 words.
  // Events among a thread's WB_HV_OPT samples sv[0..OPT) (+ the next two); sample j is signal sample n0 + j, `left`
  // positions are left in the tile.  The sign tests run on all samples, the positions outside the tile or too
  // close to the end of the signal (n + 1 resp. n + 2 beyond the last sample) are masked out afterwards.  *bits_out:
  // bit j*4+s = event of stream s at sample j; returns the four event counts in 16-bit fields.
  WB_DEV unsigned long long classify(const double (&sv)[WB_HV_OPT + 2], int n0, int left, int ylen, unsigned* bits_out) const {
    unsigned bits = 0;
#pragma unroll
    for (int j = 0; j < WB_HV_OPT; ++j) {
      const double s0 = sv[j], s1 = sv[j + 1];
      const double d0 = s1 - s0, d1 = sv[j + 2] - s1;
      if (s1 * s0 < 0.0) bits |= (s1 < s0 ? 1u : 2u) << (j * 4);
      if (d1 * d0 < 0.0) bits |= (d1 < d0 ? 4u : 8u) << (j * 4);
    }
    const int room = ylen - 1 - n0;  // sample j qualifies for streams 0/1 if j + 1 <= room, for 2/3 if j + 2 <= room
    int ja = wb_imin(left, room), jb = wb_imin(left, room - 1);
    ja = ja < 0 ? 0 : ja;
    jb = jb < 0 ? 0 : jb;
    const unsigned ma = ja >= WB_HV_OPT ? 0x33333333u : (0x33333333u & ((1u << (4 * ja)) - 1u));
    const unsigned mb = jb >= WB_HV_OPT ? 0xccccccccu : (0xccccccccu & ((1u << (4 * jb)) - 1u));
    bits &= ma | mb;
    *bits_out = bits;
    return (unsigned long long)__popc(bits & 0x11111111u) | ((unsigned long long)__popc(bits & 0x22222222u) << 16) |
           ((unsigned long long)__popc(bits & 0x44444444u) << 32) | ((unsigned long long)__popc(bits & 0x88888888u) << 48);
  }
  WB_DEV void detect_regs(const double (&sv)[WB_HV_OPT + 2], int t0, int tl, int ylen, double* sb, bool sb_ready,
                          unsigned short* plist, unsigned long long* wsum, int* run, double* E, int tid,
                          int nthr) const {
    const int lane = tid & 31, wp = tid >> 5, nwp = nthr >> 5;
    const int m0 = tid * WB_HV_OPT;
    unsigned bits;  // bit j*4+s: event of stream s at this thread's j-th sample (8 samples x 4 streams)
    // exclusive scan of the packed counts over the block (time order = thread order); 16 bits per stream
    const unsigned long long pack = classify(sv, t0 + m0, tl - m0, ylen, &bits);
    unsigned long long inc = pack;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long v2 = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v2;
    }
    __syncthreads();
    if (lane == 31) wsum[wp] = inc;
    __syncthreads();
    unsigned long long woff = 0, total = 0;
    for (int q = 0; q < nwp; ++q) {
      const unsigned long long v2 = wsum[q];
      if (q < wp) woff += v2;
      total += v2;
    }
    unsigned long long excl = woff + inc - pack;
    if (tid == 0) {
#pragma unroll
      for (int s = 0; s < 4; ++s) run[4 + s] = (int)((total >> (16 * s)) & 0xffffull);
    }
    // positions into per-stream lists (shared; time order = thread order): one trip per event, lowest bit first,
    // i.e. in sample order within every stream ...
    while (bits) {
      const int bpos = __ffs((int)bits) - 1;
      bits &= bits - 1;
      const int st = bpos & 3, sh = 16 * st;
      const int at = (int)((excl >> sh) & 0xffffull);
      excl += 1ull << sh;
      plist[st * WB_HV_TILE + at] = (unsigned short)(m0 + (bpos >> 2));
    }
    // ... the three filtered samples each event needs come from shared memory ...
    __syncthreads();
    if (!sb_ready) {
#pragma unroll
      for (int j = 0; j < WB_HV_OPT; ++j) sb[m0 + j] = sv[j];
      __syncthreads();
    }
    // ... and a dense pass refines one event per thread (one division each, coalesced writes)
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int ne = (int)((total >> (16 * s)) & 0xffffull);
      for (int e = tid; e < ne; e += nthr) {
        const int at = run[s] + e;
        if (at < p.edge_cap) {
          const int m = plist[s * WB_HV_TILE + e];
          const double s0 = sb[m], s1 = sb[m + 1];
          double a2, b2;
          if (s < 2) {
            a2 = s0;
            b2 = s1;
          } else {
            a2 = s1 - s0;
            b2 = sb[m + 2] - s1;
          }
          // (-a)/((-b)-(-a)) == a/(b-a): the rising streams use the same expression
          E[(size_t)s * p.edge_cap + at] = (double)(t0 + m + 1) - a2 / (b2 - a2);
        }
      }
    }
  }
  // Variant for tiles that already sit in shared memory (`sb`): the running event counts of the four streams
  // live in registers (every thread derives the same totals from the warp sums), the warp-sum scratch is double
  // buffered by tile parity, and the caller provides the barrier that ends the tile: 2 barriers instead of 7.
  WB_DEV void detect_regs_fast(const double (&sv)[WB_HV_OPT + 2], int t0, int tl, int ylen, const double* sb,
                               unsigned short* plist, unsigned long long* wsum2, int parity, int (&runr)[4], double* E,
                               int tid, int nthr) const {
    const int lane = tid & 31, wp = tid >> 5, nwp = nthr >> 5;
    unsigned long long* wsum = wsum2 + parity * 16;
    const int m0 = tid * WB_HV_OPT;
    unsigned bits;
    const unsigned long long pack = classify(sv, t0 + m0, tl - m0, ylen, &bits);
    unsigned long long inc = pack;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long v2 = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v2;
    }
    if (lane == 31) wsum[wp] = inc;
    __syncthreads();
    unsigned long long woff = 0, total = 0;
    for (int q = 0; q < nwp; ++q) {
      const unsigned long long v2 = wsum[q];
      if (q < wp) woff += v2;
      total += v2;
    }
    unsigned long long excl = woff + inc - pack;
    while (bits) {
      const int bpos = __ffs((int)bits) - 1;
      bits &= bits - 1;
      const int st = bpos & 3, sh = 16 * st;
      const int at = (int)((excl >> sh) & 0xffffull);
      excl += 1ull << sh;
      plist[st * WB_HV_TILE + at] = (unsigned short)(m0 + (bpos >> 2));
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int ne = (int)((total >> (16 * s)) & 0xffffull);
      for (int e = tid; e < ne; e += nthr) {
        const int at = runr[s] + e;
        if (at < p.edge_cap) {
          const int m = plist[s * WB_HV_TILE + e];
          // the tile is the swizzled output of the inverse transform (wb_fft_fast<.., SWZ_LAST>)
          const double s0 = wb_fft_swz_real(sb, m), s1 = wb_fft_swz_real(sb, m + 1);
          double a2, b2;
          if (s < 2) {
            a2 = s0;
            b2 = s1;
          } else {
            a2 = s1 - s0;
            b2 = wb_fft_swz_real(sb, m + 2) - s1;
          }
          E[(size_t)s * p.edge_cap + at] = (double)(t0 + m + 1) - a2 / (b2 - a2);
        }
      }
      runr[s] += ne;
      if (runr[s] > p.edge_cap) {
        runr[s] = p.edge_cap;
        if (tid == 0) p.status[0] = 1;
      }
    }
  }
#else
  // Host emulation: the same events from the tile samples sb[0 .. tl + 2), two passes (count, then write).
  WB_DEV void detect_smem(const double* sb, int t0, int tl, int ylen, int* cnt, int* run, double* E, int tid,
                          int nthr) const {
    const int per_thread = (tl + nthr - 1) / nthr > 0 ? (tl + nthr - 1) / nthr : 1;
    const int mlo = wb_imin(tid * per_thread, tl), mhi = wb_imin(mlo + per_thread, tl);
    for (int pass = 0; pass < 2; ++pass) {
      int w[4] = {0, 0, 0, 0};
      int base4[4] = {0, 0, 0, 0};
      if (pass == 1) {
        for (int s = 0; s < 4; ++s) base4[s] = run[s] + cnt[s * nthr + tid];
      }
      for (int m = mlo; m < mhi; ++m) {
        const int n = t0 + m;
        const double s0 = sb[m], s1 = sb[m + 1];
        if (n + 1 <= ylen - 1 && s1 * s0 < 0.0) {
          const int s = (s1 < s0) ? 0 : 1;
          if (pass == 1) {
            const double a = s == 0 ? s0 : -s0, b = s == 0 ? s1 : -s1;
            const int at = base4[s] + w[s];
            if (at < p.edge_cap) E[(size_t)s * p.edge_cap + at] = (double)(n + 1) - a / (b - a);
          }
          ++w[s];
        }
        if (n + 2 <= ylen - 1) {
          const double d0 = s1 - s0, d1 = sb[m + 2] - s1;
          if (d1 * d0 < 0.0) {
            const int s = (d1 < d0) ? 2 : 3;
            if (pass == 1) {
              const double a = s == 2 ? d0 : -d0, b = s == 2 ? d1 : -d1;
              const int at = base4[s] + w[s];
              if (at < p.edge_cap) E[(size_t)s * p.edge_cap + at] = (double)(n + 1) - a / (b - a);
            }
            ++w[s];
          }
        }
      }
      if (pass == 0) {
        for (int s = 0; s < 4; ++s) cnt[s * nthr + tid] = w[s];
        WB_SYNC();
        for (int s = tid; s < 4; s += nthr) {
          int a = 0;
          for (int t = 0; t < nthr; ++t) {
            const int v2 = cnt[s * nthr + t];
            cnt[s * nthr + t] = a;
            a += v2;
          }
          run[4 + s] = a;
        }
        WB_SYNC();
      }
    }
  }
#endif

  // after a tile: fold its event counts (run[4..8)) into the running totals (run[0..4))
  WB_DEV void close_tile(int* run, int tid, int nthr) const {
    WB_SYNC();
    for (int s = tid; s < 4; s += nthr) {
      run[s] += run[4 + s];
      if (run[s] > p.edge_cap) {
        run[s] = p.edge_cap;
        p.status[0] = 1;
      }
    }
    WB_SYNC();
  }

  // ---- Harvest's case of finish_item below (mode 0), trimmed for the overlap-save kernel, which spends a quarter of
  // its instructions here: frame times advance as doubles (no int -> double conversion per frame), one running
  // pointer per group of frames, the four streams share one store path (prev + value, prev = 0 for the first
  // stream).  Streams with more intervals than the staging area holds are tabulated window by window: a window
  // serves the groups of frames whose brackets lie inside it, the next one starts at the bracket of the first
  // frame left over.  Same expressions, same values.  Returns false when the general routine has to run instead
  // (unusable item, or a window that does not get one group further).
  WB_DEV bool finish_item_hv(int c, int u, const int* run, const double* E, double* stage, int stage_cap, int tid,
                             int nthr) const {
    const int ne0 = run[0], ne1 = run[1], ne2 = run[2], ne3 = run[3];
    if (!(ne0 >= 4 && ne1 >= 4 && ne2 >= 4 && ne3 >= 4)) return false;
#ifdef WB_HOST_EMU
    stage_cap = wb_imin(stage_cap, 256);  // the CPU test tier walks through the window logic on its short fixtures
#endif
    const int cap = stage_cap / 2;  // intervals per window
    const double edge = p.edges[c];
    const double lim_hi = edge * 1.1, lim_lo = edge * 0.9;
    const int f1 = wb_hv_frames(p.n_samples[u], p.fs, p.grid_ms);
    double* R = p.raw + ((size_t)u * p.n_ch + c) * p.f1_stride;
    const int n_groups = (f1 + WB_HV_FPT - 1) / WB_HV_FPT;
    for (int s = 0; s < 4; ++s) {
      const double* Es = E + (size_t)s * p.edge_cap;
      const int ni = run[s] - 1;  // number of intervals
      int k0 = 0, g0 = 0;         // first interval of the window, first group still to do
      while (g0 < n_groups) {
        const int k1 = wb_imin(ni, k0 + cap), nk = k1 - k0;
        double* X = stage - k0;        // X[k], Yv[k] for k0 <= k < k1
        double* Yv = stage + nk - k0;
        for (int k = k0 + tid; k < k1; k += nthr) {
          const double e0 = Es[k], e1 = Es[k + 1];
          X[k] = (e0 + e1) / 2.0 / p.afs;
          Yv[k] = p.afs / (e1 - e0);
        }
        WB_SYNC();
        // groups [g0, g1) are served by this window: all that are left when it reaches the last interval, otherwise
        // those whose last frame lies at or before the window's last midpoint
        int g1 = n_groups, k0_next = k1;
        if (k1 < ni) {
          const double x_top = X[k1 - 1];
          int lo = g0, hi = n_groups;  // first group whose last frame lies beyond x_top
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const int jl = wb_imin(mid * WB_HV_FPT + WB_HV_FPT - 1, f1 - 1);
            if (wb_div1000((double)jl * p.grid_ms) <= x_top) lo = mid + 1;
            else hi = mid;
          }
          g1 = lo;
          // the next window starts one interval below the bracket of that group's first frame
          const double t_first = wb_div1000((double)(g1 * WB_HV_FPT) * p.grid_ms);
          int il = wb_imax(1, k0 + 1), ih = k1;  // smallest i in [il, k1) with x_i >= t_first, k1 if none
          while (il < ih) {
            const int mid = (il + ih) >> 1;
            if (X[mid] < t_first) il = mid + 1;
            else ih = mid;
          }
          k0_next = il - 1;
          if (g1 == g0 && k0_next <= k0) return false;  // every thread takes the same decision
        }
        for (int g = g0 + tid; g < g1; g += nthr) {
          const int j0 = g * WB_HV_FPT;
          const int nq = wb_imin(WB_HV_FPT, f1 - j0);
          double* Rg = R + j0;
          double prev[WB_HV_FPT];
#pragma unroll
          for (int q = 0; q < WB_HV_FPT; ++q) prev[q] = (s > 0 && q < nq) ? Rg[q] : 0.0;
          const double tj = (double)j0;
          double t = wb_div1000(tj * p.grid_ms);
          const int i_top = k1 - 1;  // == ni - 1 in the last window; never exceeded before (see g1)
          int lo = wb_imax(1, k0 + 1), hi = i_top;  // smallest i >= 1 with x_i >= t (the last one if none)
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (X[mid] < t) lo = mid + 1;
            else hi = mid;
          }
          int i = lo;
          double xl = X[i - 1], xh = X[i], yl = Yv[i - 1], yh = Yv[i];
          double slope = (yh - yl) / (xh - xl);
          auto frame = [&](int q) {
            t = wb_div1000((tj + (double)q) * p.grid_ms);
            while (i < i_top && xh < t) {
              ++i;
              xl = xh;
              yl = yh;
              xh = X[i];
              yh = Yv[i];
              slope = (yh - yl) / (xh - xl);
            }
            double v = prev[q] + (slope * (t - xl) + yl);
            if (s == 3) {
              v = v / 4.0;
              if (v > lim_hi) v = 0.0;
              if (v < lim_lo) v = 0.0;
              if (v > p.f0_ceil) v = 0.0;
              if (v < p.f0_floor) v = 0.0;
            }
            Rg[q] = v;
          };
#pragma unroll
          for (int q = 0; q < WB_HV_FPT; ++q) {
            if (q < nq) frame(q);
          }
        }
        WB_SYNC();
        g0 = g1;
        k0 = k0_next;
      }
    }
    return true;
  }

  // ---- interval F0 of each stream interpolated onto the frame grid (GetF0Candidates, harvest.py:499-529; DIO:
  // get_f0_candidates + get_raw_event, dio.py:128-185).  `stage` (stage_cap doubles of shared memory) holds the
  // events of one stream at a time.
  WB_DEV void finish_item(int c, int u, const int* run, const double* E, double* stage, int stage_cap, int tid,
                          int nthr) const {
    const double edge = p.edges[c];
    const int f1 = wb_hv_frames(p.n_samples[u], p.fs, p.grid_ms);
    double* R = p.raw + ((size_t)u * p.n_ch + c) * p.f1_stride;
    const int ne0 = run[0], ne1 = run[1], ne2 = run[2], ne3 = run[3];
    const bool usable = ne0 >= 4 && ne1 >= 4 && ne2 >= 4 && ne3 >= 4;  // >= 3 intervals each (harvest.py:504-507)
    if (!usable) {
      if (p.mode == 0) {
        for (int j = tid; j < f1; j += nthr) R[j] = 0.0;
      } else {  // dio.py:182-184, 144-150: no estimate, deviation 1000 -> overridden to 100000
        double* Sb = p.stab + ((size_t)u * p.n_ch + c) * p.f1_stride;
        for (int j = tid; j < f1; j += nthr) {
          R[j] = 0.0;
          Sb[j] = exp(-(100000.0 / 0.0000001));
        }
      }
      return;
    }
    double* V4 = p.mode ? p.four + ((size_t)u * p.n_ch + c) * 4 * p.f1_stride : nullptr;
    // Frame-major interpolation: a thread takes WB_HV_FPT consecutive frames, finds the pair of interval
    // midpoints around its first frame by bisection and then walks forward event by event.  Midpoint k
    // is x_k = (e_k + e_k+1)/2/afs with value afs/(e_k+1 - e_k); pair i (1 <= i <= ni-1) serves the frames
    // with x_i-1 < t <= x_i, the first and last pair extend to -inf / +inf (interp1d fill_value=
    // 'extrapolate').
    const int n_groups = (f1 + WB_HV_FPT - 1) / WB_HV_FPT;
    for (int s = 0; s < 4; ++s) {
      const double* Es = E + (size_t)s * p.edge_cap;
      const int ne = run[s], ni = ne - 1;  // number of intervals
      // Fast path: midpoints X[k] and interval values Yv[k] of the whole stream are tabulated once in shared
      // memory (one thread per interval), then every frame needs a bisection on X and one slope.  Streams too
      // long for the staging area are walked straight from the event list (same expressions, same values).
      const bool tab = 2 * ni <= stage_cap;
      const double* Ev = Es;
      double* X = stage;
      double* Yv = stage + ni;
      if (tab) {
        for (int k = tid; k < ni; k += nthr) {
          const double e0 = Es[k], e1 = Es[k + 1];
          X[k] = (e0 + e1) / 2.0 / p.afs;
          Yv[k] = p.afs / (e1 - e0);
        }
        WB_SYNC();
      } else if (ne <= stage_cap) {
        for (int i = tid; i < ne; i += nthr) stage[i] = Es[i];
        WB_SYNC();
        Ev = stage;
      }
      for (int g = tid; g < n_groups; g += nthr) {
        const int j0 = g * WB_HV_FPT, j1 = wb_imin(j0 + WB_HV_FPT, f1);
        double t = wb_div1000((double)j0 * p.grid_ms);
        // smallest i in [1, ni-1] with x_i >= t (ni-1 if none)
        int i;
        double eb = 0.0, ec = 0.0, xl, xh, yl, yh;
        if (tab) {
          int lo = 1, hi = ni - 1;
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (X[mid] < t) lo = mid + 1;
            else hi = mid;
          }
          i = lo;
          xl = X[i - 1];
          xh = X[i];
          yl = Yv[i - 1];
          yh = Yv[i];
        } else {  // bisection on the undivided sums, then the reference's own comparison settles the last step
          const double t2 = t * 2.0 * p.afs;
          int lo = 1, hi = ni - 1;
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (Ev[mid] + Ev[mid + 1] < t2) lo = mid + 1;
            else hi = mid;
          }
          while (lo > 1 && !((Ev[lo - 1] + Ev[lo]) / 2.0 / p.afs < t)) --lo;
          while (lo < ni - 1 && (Ev[lo] + Ev[lo + 1]) / 2.0 / p.afs < t) ++lo;
          i = lo;
          const double ea = Ev[i - 1];
          eb = Ev[i];
          ec = Ev[i + 1];
          xl = (ea + eb) / 2.0 / p.afs;
          xh = (eb + ec) / 2.0 / p.afs;
          yl = p.afs / (eb - ea);
          yh = p.afs / (ec - eb);
        }
        double slope = (yh - yl) / (xh - xl);
        double prev[WB_HV_FPT];
        if (!p.mode && s > 0) {
#pragma unroll
          for (int q = 0; q < WB_HV_FPT; ++q) prev[q] = j0 + q < j1 ? R[j0 + q] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < WB_HV_FPT; ++q) {
          const int j = j0 + q;
          if (j < j1) {
            t = wb_div1000((double)j * p.grid_ms);
            while (i < ni - 1 && xh < t) {
              ++i;
              xl = xh;
              yl = yh;
              if (tab) {
                xh = X[i];
                yh = Yv[i];
              } else {
                eb = ec;
                ec = Ev[i + 1];
                xh = (eb + ec) / 2.0 / p.afs;
                yh = p.afs / (ec - eb);
              }
              slope = (yh - yl) / (xh - xl);
            }
            const double val = slope * (t - xl) + yl;
            if (p.mode) {
              V4[(size_t)s * p.f1_stride + j] = val;
            } else if (s == 0) {
              R[j] = val;
            } else if (s < 3) {
              R[j] = prev[q] + val;
            } else {
              double est = (prev[q] + val) / 4.0;
              if (est > edge * 1.1) est = 0.0;
              if (est < edge * 0.9) est = 0.0;
              if (est > p.f0_ceil) est = 0.0;
              if (est < p.f0_floor) est = 0.0;
              R[j] = est;
            }
          }
        }
      }
      WB_SYNC();
    }
    if (p.mode) {  // get_f0_candidates + get_raw_event gates + stability (dio.py:176-181, 144-150, 106)
      double* Sb = p.stab + ((size_t)u * p.n_ch + c) * p.f1_stride;
      for (int j = tid; j < f1; j += nthr) {
        const double a = V4[j], b = V4[(size_t)p.f1_stride + j], c2 = V4[2 * (size_t)p.f1_stride + j],
                     d = V4[3 * (size_t)p.f1_stride + j];
        double est = (((a + b) + c2) + d) / 4.0;
        const double da = a - est, db = b - est, dc = c2 - est, dd = d - est;
        double dev = sqrt((((da * da + db * db) + dc * dc) + dd * dd) / 3.0);
        if (est > edge) est = 0.0;
        if (est < edge / 2.0) est = 0.0;
        if (est > p.f0_ceil) est = 0.0;
        if (est < p.f0_floor) est = 0.0;
        if (est == 0.0) dev = 100000.0;
        R[j] = est;
        Sb[j] = exp(-(dev / wb_dmax(est, 0.0000001)));
      }
    }
  }
};

struct wb_hv_channels : wb_hv_channels_common {
  // the signal tile is stored in groups of 16 samples padded to 18 doubles: consecutive threads (16 samples
  // apart) then hit distinct banks with 128-bit loads
  WB_HD static size_t ys_doubles(int max_taps) { return ((size_t)(WB_HV_TILE + max_taps + 40) / 16 + 2) * 18; }
  static size_t smem_bytes(int max_taps, int nthr) {
    return (ys_doubles(max_taps) + (size_t)((max_taps + 9) & ~1) + WB_HV_TILE + 8) * sizeof(double) +
           (size_t)(4 * nthr + 16) * sizeof(int);
  }
  WB_DEV static int skew(int i) { return (i >> 4) * 18 + (i & 15); }

  WB_DEV void operator()(int block, int tid, int nthr, double* smem) const {
    const int L_max = p.max_taps;
    double* ys = smem;
    double* rt = ys + ys_doubles(L_max);
    double* sb = rt + ((L_max + 9) & ~1);
    int* cnt = (int*)(sb + WB_HV_TILE + 8);  // [4][nthr] then 4 running totals + 4 tile totals
    int* run = cnt + 4 * nthr;
    double* E = p.edge_buf + (size_t)block * 4 * p.edge_cap;
    const int c_first = p.fft_nch;  // channels below run through wb_hv_channels_fft
    const long long n_items = (long long)(p.n_ch - c_first) * p.batch;

    for (long long item = block; item < n_items; item += p.n_slots) {
      const int c = c_first + (int)(item / p.batch), u = (int)(item % p.batch);
      const int L = p.halfs[c], off0 = p.ch_off[c];
      const double* yu = p.y + (size_t)u * p.y_stride;
      const int ylen = p.y_len[u];
      int wrap_n = 0;
      if (p.wrap_n > 0) {  // 2 ** ceil(log(ylen + wrap_add, 2)) (dio.py:78); pow2_quirk covers exact powers of two
        const int v = ylen + p.wrap_n;
        int k = 0;
        while ((1 << k) < v) ++k;
        if ((1 << k) == v) k = p.pow2_quirk[k];
        wrap_n = 1 << k;
      }
      for (int k = tid; k < L; k += nthr) rt[k] = WB_LDG(p.taps + p.tap_off[c] + k);
      for (int s = tid; s < 4; s += nthr) run[s] = 0;
      WB_SYNC();

      // ---- filter tile by tile and collect the four event streams -------------------
      for (int t0 = 0; t0 < ylen; t0 += WB_HV_TILE - 2) {
        const int need = WB_HV_TILE + L - 1;
        for (int i = tid; i < need; i += nthr) {
          int yi = t0 + off0 + i;
          if (wrap_n > 0) {
            yi %= wrap_n;
            if (yi < 0) yi += wrap_n;
          }
          ys[skew(i)] = (yi >= 0 && yi < ylen) ? WB_LDG(yu + yi) : 0.0;
        }
        WB_SYNC();
#ifndef WB_HOST_EMU
        double keep8[WB_HV_OPT];
#endif
        for (int m0 = tid * WB_HV_OPT; m0 < WB_HV_TILE; m0 += nthr * WB_HV_OPT) {
          double acc[WB_HV_OPT];
          double va[8], vb[8], cfs[8];
#pragma unroll
          for (int j = 0; j < WB_HV_OPT; ++j) acc[j] = 0.0;
          {  // the thread's first OPT samples (m0 is a multiple of 8: a group or half a group)
            const wb_cplx* g0 = (const wb_cplx*)(ys + (m0 >> 4) * 18 + (m0 & 15));
#pragma unroll
            for (int j = 0; j < WB_HV_OPT / 2; ++j) {
              const wb_cplx t2 = g0[j];
              va[2 * j] = t2.x;
              va[2 * j + 1] = t2.y;
            }
          }
          // eight taps against the 16-sample window (lo | hi); hi receives the next eight samples first
          auto fir8 = [&](int k8, double (&lo)[8], double (&hi)[8]) {
            const int nx = m0 + k8 + WB_HV_OPT;  // next 8 samples: a multiple of 8, i.e. half a group
            const wb_cplx* g = (const wb_cplx*)(ys + (nx >> 4) * 18 + (nx & 15));
            const wb_cplx* c2 = (const wb_cplx*)(rt + k8);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const wb_cplx t2 = g[j], t3 = c2[j];
              hi[2 * j] = t2.x;
              hi[2 * j + 1] = t2.y;
              cfs[2 * j] = t3.x;
              cfs[2 * j + 1] = t3.y;
            }
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
#pragma unroll
              for (int j = 0; j < WB_HV_OPT; ++j) acc[j] += cfs[kk] * (kk + j < 8 ? lo[kk + j] : hi[kk + j - 8]);
            }
          };
          int k = 0;
          for (; k + 16 <= L; k += 16) {  // the two sample buffers swap roles: no register shuffling
            fir8(k, va, vb);
            fir8(k + 8, vb, va);
          }
          if (k + 8 <= L) {
            fir8(k, va, vb);
            k += 8;
          }
          for (; k < L; ++k) {
            const double cf = rt[k];
#pragma unroll
            for (int j = 0; j < WB_HV_OPT; ++j) acc[j] += cf * ys[skew(m0 + k + j)];
          }
#ifdef WB_HOST_EMU
          for (int j = 0; j < WB_HV_OPT; ++j) sb[m0 + j] = acc[j];
#else
#pragma unroll
          for (int j = 0; j < WB_HV_OPT; ++j) keep8[j] = acc[j];  // one pass per tile on the GPU (OPT * nthr == TILE)
#endif
        }
        WB_SYNC();
#ifndef WB_HOST_EMU
        {
          // each thread owns WB_HV_OPT consecutive filtered samples (registers) and needs the next thread's first two
          sb[2 * tid] = keep8[0];
          sb[2 * tid + 1] = keep8[1];
          __syncthreads();
          double sv[WB_HV_OPT + 2];
#pragma unroll
          for (int j = 0; j < WB_HV_OPT; ++j) sv[j] = keep8[j];
          sv[WB_HV_OPT] = tid + 1 < nthr ? sb[2 * tid + 2] : 0.0;
          sv[WB_HV_OPT + 1] = tid + 1 < nthr ? sb[2 * tid + 3] : 0.0;
          detect_regs(sv, t0, WB_HV_TILE - 2, ylen, sb, false, (unsigned short*)ys,
                      (unsigned long long*)(sb + 2 * nthr), run, E, tid, nthr);
        }
#else
        detect_smem(sb, t0, WB_HV_TILE - 2, ylen, cnt, run, E, tid, nthr);
#endif
        close_tile(run, tid, nthr);
      }
      finish_item(c, u, run, E, ys, (int)(((double*)cnt) - ys), tid, nthr);
      WB_SYNC();
    }
  }
};

// ---- overlap-save path ------------------------------------------------------------------------------
// Block b of an utterance is the 2048 samples y[b * V + A + j], j < 2048 (zero outside the signal), with
// A = the most negative tap offset of any channel and V = WB_HV_FFT_V new output positions per block.  For
// channel c (first tap offset off0_c, reversed taps r_c) the filtered sample at position b V + m is the
// circular correlation  sum_i r_c[i] yb[m + (off0_c - A) + i], exact for m < 2048 - (off0_c - A + L_c - 1);
// with symmetric filters that bound is >= 2048 - 2 half_max.  In the frequency domain this is
// conj(FFT(g_c)) FFT(yb) / N with g_c the taps placed at offset off0_c - A; the table fft_H holds
// conj(FFT(g_c)) / N for k <= N/2.
#define WB_HV_FFT_N 2048

struct wb_hv_fft_fwd {  // one block per (utterance, signal block): spectrum of the block -> fft_Y
  wb_hv_plan p;
  const wb_cplx* tw;
  int tw_n;
  static size_t smem_bytes() {
    return (size_t)(2 * (WB_HV_FFT_N / 2 + 2)) * sizeof(wb_cplx) + (size_t)WB_FFT_TW_SLOTS(WB_HV_FFT_N / 2) * sizeof(wb_cplx);
  }
  WB_DEV void operator()(int block, int tid, int nthr, double* smem) const {
    const int u = block / p.fft_blocks;
    int b = block - u * p.fft_blocks, g = 0;
    while (g + 1 < p.fft_groups && b >= p.fft_gblocks[g]) b -= p.fft_gblocks[g++];
    const int ylen = p.y_len[u];
    if (b * p.fft_gV[g] >= ylen) return;
    wb_cplx* A = (wb_cplx*)smem;
    wb_cplx* B = A + (WB_HV_FFT_N / 2 + 2);
    wb_cplx* twS = B + (WB_HV_FFT_N / 2 + 2);
    wb_fft_load_twiddles(twS, WB_HV_FFT_N / 2, tw, tw_n, tid, nthr);
    const double* yu = p.y + (size_t)u * p.y_stride;
    double* Ad = (double*)A;
    const int base = b * p.fft_gV[g] + p.fft_gA[g];
    for (int j = tid; j < WB_HV_FFT_N; j += nthr) {
      const int yi = base + j;
      Ad[j] = (yi >= 0 && yi < ylen) ? yu[yi] : 0.0;
    }
    WB_SYNC();
    const wb_cplx* X = wb_rfft<0, WB_HV_FFT_N>(A, B, WB_HV_FFT_N, twS, WB_HV_FFT_N / 2, tid, nthr);
    wb_cplx* out = p.fft_Y + p.fft_goff[g] + ((size_t)u * p.fft_gblocks[g] + b) * (WB_HV_FFT_N / 2 + 1);
    for (int k = tid; k <= WB_HV_FFT_N / 2; k += nthr) out[k] = X[k];
  }
};

struct wb_hv_channels_fft : wb_hv_channels_common {
  const wb_cplx* tw;
  int tw_n;
  static size_t smem_bytes(int nthr) {
    // two transform buffers, the staging buffer the next block spectrum is bulk-copied into, twiddles, event scratch
    // (the per-thread event counters `cnt` are only used by the host emulation's two-pass detection)
#ifdef WB_HOST_EMU
    const size_t n_cnt = 4 * (size_t)nthr;
#else
    const size_t n_cnt = 0;
    (void)nthr;
#endif
    return (size_t)(3 * (WB_HV_FFT_N / 2 + 2)) * sizeof(wb_cplx) + (size_t)WB_FFT_TW_SLOTS(WB_HV_FFT_N / 2) * sizeof(wb_cplx) +
           (size_t)48 * sizeof(double) + (n_cnt + 16) * sizeof(int) + 16;
  }
  WB_DEV void operator()(int block, int tid, int nthr, double* smem) const {
    const int NH = WB_HV_FFT_N / 2;
    wb_cplx* A = (wb_cplx*)smem;
    wb_cplx* B = A + (NH + 2);
    wb_cplx* Ys = B + (NH + 2);   // block spectrum of the next transform, filled by cp.async.bulk
    wb_cplx* twS = Ys + (NH + 2);
    double* misc = (double*)(twS + WB_FFT_TW_SLOTS(NH));  // 48 doubles: warp sums of the event scan
    int* cnt = (int*)(misc + 48);
#ifdef WB_HOST_EMU
    int* run = cnt + 4 * nthr;
#else
    int* run = cnt;
#endif
#ifndef WB_HOST_EMU
    unsigned long long* ybar = (unsigned long long*)(((size_t)(run + 16) + 7) & ~(size_t)7);
    if (tid == 0) wb_mbar_init(ybar, 1);
    unsigned yphase = 0;
#endif
    double* E = p.edge_buf + (size_t)block * 4 * p.edge_cap;
    const long long n_items = (long long)p.fft_nch * p.batch;
    wb_fft_load_twiddles(twS, NH, tw, tw_n, tid, nthr);
    for (long long item = block; item < n_items; item += p.n_slots) {
      // utterance-major: the blocks in flight share a few utterances' spectra (L2-resident)
      const int u = (int)(item / p.fft_nch), c = (int)(item % p.fft_nch);
      const int ylen = p.y_len[u];
      const wb_cplx* H = p.fft_H + (size_t)c * (NH + 1);
      int g = 0;
      while (g + 1 < p.fft_groups && c >= p.fft_gc[g + 1]) ++g;
      const int V = p.fft_gV[g];
      const wb_cplx* Yu = p.fft_Y + p.fft_goff[g] + (size_t)u * p.fft_gblocks[g] * (NH + 1);
#ifndef WB_HOST_EMU
      int runr[4] = {0, 0, 0, 0};
      int b = 0;
      // The block spectra (16 KB each, L2-resident) reach shared memory through the TMA engine: the copy of
      // block b + 1 is issued as soon as the product of block b has consumed the staging buffer and lands while
      // the block runs its inverse transform and event detection.
      constexpr unsigned Y_BYTES = (unsigned)((NH + 1) * sizeof(wb_cplx));
      if (tid == 0 && ylen > 0) {
        wb_fence_proxy_async();  // the previous item's interpolation tables ran on into Ys
        wb_bulk_load(Ys, Yu, Y_BYTES, ybar);
      }
      for (int t0 = 0; t0 < ylen; t0 += V, ++b) {
        const wb_cplx* Y = Ys;
        wb_mbar_wait(ybar, yphase);
        yphase ^= 1u;
        // spectrum product fused with the first step of the inverse real transform (wb_irfft_merge): bins k and
        // NH - k give the entries k and NH - k of the half-size complex sequence; a thread that takes k <= NH/4
        // also takes the pair around NH/2 that shares its twiddle
        for (int k = tid; k <= (NH >> 2); k += nthr) {
          if (k == 0) {
            const double x0 = wb_cmul(wb_ldg_cplx(H), Y[0]).x, xm = wb_cmul(wb_ldg_cplx(H + NH), Y[NH]).x;
            const wb_cplx c = wb_cmul(wb_ldg_cplx(H + (NH >> 1)), Y[NH >> 1]);
            A[0] = wb_mk(x0 + xm, x0 - xm);
            A[NH >> 1] = wb_mk(2.0 * c.x, -2.0 * c.y);
          } else {
            const wb_cplx W = twS[wb_fft_tw_skew(k)];  // table of the 2 NH = WB_HV_FFT_N point circle
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              if (half == 1 && k == (NH >> 2)) break;
              const int k1 = half ? (NH >> 1) - k : k, k2 = NH - k1;
              const wb_cplx Wk = half ? wb_mk(-W.y, -W.x) : W;
              const wb_cplx xk = wb_cmul(wb_ldg_cplx(H + k1), Y[k1]);
              const wb_cplx xc = wb_conj(wb_cmul(wb_ldg_cplx(H + k2), Y[k2]));
              const wb_cplx S2 = wb_cadd(xk, xc), D2 = wb_csub(xk, xc);
              const wb_cplx t1 = wb_cmul(wb_conj(Wk), D2);
              const wb_cplx t2 = wb_cmul(Wk, wb_conj(D2));
              A[k1] = wb_mk(S2.x - t1.y, S2.y + t1.x);
              A[k2] = wb_mk(S2.x - t2.y, -S2.y + t2.x);
            }
          }
        }
        __syncthreads();
        if (tid == 0 && t0 + V < ylen) wb_bulk_load(Ys, Yu + (size_t)(b + 1) * (NH + 1), Y_BYTES, ybar);
        // out[m] = filtered sample t0 + m, left swizzled: every thread reads 10 consecutive samples
        double* out = (double*)wb_fft_fast<WB_HV_FFT_N / 2, +1, true>(A, B, twS, NH, tid, nthr, 0x7fffffff);
        double sv[WB_HV_OPT + 2];
        const int m0 = tid * WB_HV_OPT;
#pragma unroll
        for (int j = 0; j < WB_HV_OPT + 2; j += 2) {
          const int c = (m0 + j) >> 1;
          const wb_cplx z = c < NH ? ((const wb_cplx*)out)[wb_fft_swz(c)] : wb_mk(0.0, 0.0);
          sv[j] = z.x;
          sv[j + 1] = z.y;
        }
        unsigned short* plist = (unsigned short*)(out == (double*)A ? (double*)B : (double*)A);
        detect_regs_fast(sv, t0, V, ylen, out, plist, (unsigned long long*)misc, b & 1, runr, E, tid, nthr);
        __syncthreads();  // the next block's spectrum product overwrites the buffers the event pass read
      }
      if (tid == 0) {
#pragma unroll
        for (int s = 0; s < 4; ++s) run[s] = runr[s];
      }
      __syncthreads();
#else
      for (int s = tid; s < 4; s += nthr) run[s] = 0;
      WB_SYNC();
      int b = 0;
      for (int t0 = 0; t0 < ylen; t0 += V, ++b) {
        const wb_cplx* Y = Yu + (size_t)b * (NH + 1);
        for (int k = tid; k <= NH; k += nthr) A[k] = wb_cmul(wb_ldg_cplx(H + k), Y[k]);
        WB_SYNC();
        double* out = wb_irfft<0, WB_HV_FFT_N>(A, B, WB_HV_FFT_N, twS, NH, tid, nthr);  // out[m] = filtered sample t0 + m
        detect_smem(out, t0, V, ylen, cnt, run, E, tid, nthr);
        close_tile(run, tid, nthr);
      }
#endif
      // (the staging buffer of the block spectra is idle between items: the tables may run on into it)
      if (p.mode != 0 || !finish_item_hv(c, u, run, E, (double*)A, 3 * (NH + 2) * 2, tid, nthr))
        finish_item(c, u, run, E, (double*)A, 2 * (NH + 2) * 2, tid, nthr);
      WB_SYNC();
    }
  }
};

// ------------------------------------------------------------------------------------ H3
// One thread per (utterance, 1 ms frame).
struct wb_hv_detect {
  wb_hv_plan p;
  WB_DEV void operator()(long long item) const {
    const int u = (int)(item / p.f1_stride), j = (int)(item - (long long)u * p.f1_stride);
    const int f1 = wb_hv_frames(p.n_samples[u], p.fs, 1.0);
    if (j >= f1) return;
    const double* R = p.raw + (size_t)u * p.n_ch * p.f1_stride + j;
    double* bc = p.base_c + ((size_t)u * p.f1_stride + j) * WB_HV_MAXC;
    int count = 0, run_len = 0;
    double run_sum = 0.0;
    for (int c0 = 1; c0 <= p.n_ch - 1; c0 += 8) {  // eight channels' loads in flight (the map is read-only here)
      double v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = (c0 + q < p.n_ch - 1) ? WB_LDG(R + (size_t)(c0 + q) * p.f1_stride) : 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (c0 + q <= p.n_ch - 1) {  // the last channel only closes a run
          if (v[q] > 0.0) {
            run_sum += v[q];
            ++run_len;
          } else {
            if (run_len >= 10 && count < WB_HV_MAXC) bc[count++] = run_sum / run_len;
            run_len = 0;
            run_sum = 0.0;
          }
        }
      }
    }
    p.base_n[(size_t)u * p.f1_stride + j] = count;
  }
};

// ------------------------------------------------------------------------------------ H4 (one thread per candidate)
// Same computation as wb_hv_refine, turned sideways: every THREAD refines one candidate on its own (a serial
// walk over the window with a rotating window cosine and six rotating DFT phasors): no cross-lane reduction,
// no shared memory, set-up and tail once per thread.  To keep the lanes of a warp in step, the candidates
// offered to all frames are first counting-sorted by window half-length (hv_refine_count / _scan / _scatter),
// so that 32 consecutive work items have the same loop length.  Results go to fixed slots (frame, index);
// rejected candidates are written as zeros and dropped by hv_prune's keep flag.
#define WB_HV_NCLS 512  // window half-length classes
#define WB_HV_SRC_QUIRK 127  // source code of a work item whose candidate is row 6 of the frame itself

struct wb_hv_refine_items {
  wb_hv_plan p;
  const wb_cplx* tw;
  int tw_n;
  int* cls_count;                 // [WB_HV_NCLS] items per class; [WB_HV_NCLS] = total
  int* cls_cursor;                // [WB_HV_NCLS] next free position of each class
  unsigned long long* items;      // [capacity] work items, see refine_all
  long long capacity;
  int mode;                       // 0 count, 1 scatter (wb_hv_refine_prep); the refinement ignores it
  int frames_per_block;           // frames handled by one block in modes 0 / 1

  // candidates offered to frame j of utterance u, in the reference's row order (OverlapF0Candidates,
  // harvest.py:114-125): start[s] = first index of shift s, start[7] = total; quirk: row 0 keeps the 7th
  // candidate of the frame itself at frames 0..2
  WB_DEV int offered(int u, int j, int f1, int* start, int* quirk) const {
    const size_t fb = (size_t)u * p.f1_stride;
    *quirk = (j < 3 && p.base_n[fb + j] >= 7) ? 1 : 0;
    int n = *quirk;
    for (int s = 0; s < 7; ++s) {
      const int src = j - 3 + s;
      start[s] = n;
      n += (src >= 0 && src < f1) ? p.base_n[fb + src] : 0;
    }
    start[7] = n;
    return n > WB_HV_SLOTS ? WB_HV_SLOTS : n;
  }
  WB_DEV double candidate(int u, int j, int it, const int* start, int quirk, int* slot) const {
    const size_t fb = (size_t)u * p.f1_stride;
    if (it < quirk) {
      *slot = 0;
      return p.base_c[(fb + j) * WB_HV_MAXC + 6];
    }
    int s = 6;
    while (s > 0 && start[s] > it) --s;
    const int k = it - start[s];
    *slot = s * WB_HV_MAXC + k;
    return p.base_c[(fb + j - 3 + s) * WB_HV_MAXC + k];
  }
  WB_DEV int class_of(double c0) const {
    int half = (int)ceil(3.0 * p.afs / c0 / 2.0);
    return half < 0 ? 0 : (half > WB_HV_NCLS - 1 ? WB_HV_NCLS - 1 : half);
  }

  static size_t smem_bytes() { return (size_t)(2 * WB_HV_NCLS + 8) * sizeof(int); }

  WB_DEV void operator()(int block, int tid, int nthr, double* smem) const {
    (void)smem;
    refine_all(block, tid, nthr);
  }
  // ---- count / scatter: one thread per (utterance, frame), block-level histogram first (launched through
  // wb_hv_refine_prep: a kernel of its own, so that the refinement's registers do not limit its occupancy) ----
  WB_DEV void prepare(int block, int tid, int nthr, double* smem) const {
    int* hist = (int*)smem;            // [NCLS] this block's items per class
    int* base = hist + WB_HV_NCLS;     // [NCLS] start of this block's range inside each class
    for (int c = tid; c < WB_HV_NCLS; c += nthr) hist[c] = 0;
    WB_SYNC();
    const long long n_frames_all = (long long)p.batch * p.f1_stride;
    for (int rep = 0; rep < 2; ++rep) {  // rep 0: histogram; rep 1 (scatter mode only): placement
      if (rep == 1 && mode == 0) break;
      for (int q = tid; q < frames_per_block; q += nthr) {  // one frame per thread (per step)
        const long long fi = (long long)block * frames_per_block + q;
        if (fi < n_frames_all) {
          const int u = (int)(fi / p.f1_stride), j = (int)(fi - (long long)u * p.f1_stride);
          const int f1 = wb_hv_frames(p.n_samples[u], p.fs, 1.0);
          if (j < f1) {
            int start[8], quirk;
            const int n_items = offered(u, j, f1, start, &quirk);
            if (rep == 0 && mode == 0) p.l_n[fi] = n_items;
            for (int it = 0; it < n_items; ++it) {
              int slot;
              const double c0 = candidate(u, j, it, start, quirk, &slot);
              const int cls = class_of(c0);
              if (rep == 0) {
                wb_atomic_add_int(hist + cls, 1);
              } else {
                const int at = base[cls] + wb_atomic_add_int(hist + cls, 1);
                const int code = it < quirk ? WB_HV_SRC_QUIRK : slot;
                if ((long long)at < capacity)
                  items[at] = ((unsigned long long)fi << 14) | ((unsigned long long)it << 7) | (unsigned long long)code;
              }
            }
          }
        }
      }
      WB_SYNC();
      if (rep == 0) {
        for (int c = tid; c < WB_HV_NCLS; c += nthr) {
          const int v = hist[c];
          if (v) {
            if (mode == 0) wb_atomic_add_int(cls_count + c, v);
            else base[c] = wb_atomic_add_int(cls_cursor + c, v);
          }
          hist[c] = 0;
        }
        WB_SYNC();
      }
    }
  }

  // ---- refine: persistent blocks, one work item per thread, items in class order ----
  // GetRefinedF0 (harvest.py:169-211).  The two spectra are read at <= 6 bins only, so each bin is a Goertzel
  // recurrence over the window,
  //   s[n] = x[n] + 2 cos(w) s[n-1] - s[n-2],   s[N-1] - e^{-iw} s[N-2] = e^{iw(N-1)} sum_n x[n] e^{-iwn},
  // run for the windowed segment (ga) and the derivative-windowed segment (gb).  The common phase factor drops out
  // of everything read afterwards (|S|^2 and Im(conj(S) D)).
  struct refined {
    double f0, score;
  };
  struct geometry {  // of one candidate's window
    int half, len, nfft, n_harm, stepw;
    double bin_scale, inv_len;
  };
  WB_DEV geometry geometry_of(double c0) const {
    geometry q;
    q.half = (int)ceil(3.0 * p.afs / c0 / 2.0);
    q.len = 2 * q.half + 1;
    int lg = 0;
    while ((1 << lg) < q.len) ++lg;
    q.nfft = 1 << (lg + 1);
    q.inv_len = 1.0 / (double)q.len;
    q.n_harm = (int)(p.afs * 0.5 / c0);
    if (q.n_harm > 6) q.n_harm = 6;
    q.bin_scale = c0 * q.nfft / p.afs;
    q.stepw = tw_n / q.nfft;
    return q;
  }
  // instantaneous frequencies at the harmonic bins -> refined F0 and its score (harvest.py:193-210); newest
  // Goertzel states in ga2 / gb2, the ones before in ga1 / gb1
  WB_DEV refined score_of(const geometry& q, double c0, const double (&ga1)[6], const double (&ga2)[6],
                          const double (&gb1)[6], const double (&gb2)[6]) const {
    const double inv_c0 = 1.0 / c0, inv_nfft = 1.0 / (double)q.nfft;
    double num = 0.0, den = 0.0, var = 0.0;
#pragma unroll
    for (int hh = 0; hh < 6; ++hh) {
      if (hh < q.n_harm) {
        const int hnum = hh + 1;
        const int bin = (int)(q.bin_scale * hnum + 0.5);
        const wb_cplx e = wb_ldg_cplx(tw + (size_t)(bin & (q.nfft - 1)) * q.stepw);  // e^{-iw}
        const double sr = ga2[hh] - e.x * ga1[hh], si = -e.y * ga1[hh];
        const double dr = gb2[hh] - e.x * gb1[hh], di = -e.y * gb1[hh];
        const double pw = sr * sr + si * si;
        const double inst = ((double)bin * inv_nfft + (sr * di - si * dr) / pw * (0.5 / WB_PI)) * p.afs;
        const double amp = sqrt(pw);
        num += amp * inst;
        den += amp * hnum;
        var += fabs((inst / hnum - c0) * inv_c0);
      }
    }
    refined r;
    r.f0 = num / den;
    r.score = 1.0 / (0.000000000001 + var / q.n_harm);
    if (r.f0 < p.f0_floor || r.f0 > p.f0_ceil || r.score < 2.5 || !(r.f0 == r.f0) || !(r.score == r.score)) {
      r.f0 = 0.0;
      r.score = 0.0;
    }
    return r;
  }

  // Any window: the ones that start before t = 0 (round_matlab's -0.5 branch, harvest.py:158-160) or whose sample
  // index does not advance by exactly one per position.  Rare (the first `half` samples of an utterance): a real
  // function, so that its registers and code stay out of the hot loop.
  WB_DEV_COLD refined refine_general(const double* yu, int ylen, double t, double c0) const {
    const geometry q = geometry_of(c0);
    const int half = q.half, len = q.len;
    const double afs = p.afs, inv_afs = 1.0 / p.afs, inv_len = q.inv_len;
    double ga1[6], ga2[6], gb1[6], gb2[6], coef[6];
#pragma unroll
    for (int hh = 0; hh < 6; ++hh) {
      ga1[hh] = ga2[hh] = gb1[hh] = gb2[hh] = 0.0;
      const int bin = (int)(q.bin_scale * (hh + 1) + 0.5);
      coef[hh] = 2.0 * wb_ldg_cplx(tw + (size_t)(bin & (q.nfft - 1)) * q.stepw).x;
    }
    // window 0.42 + 0.5 cos(theta) + 0.08 cos(2 theta), theta_i/pi = 2 ((r_i - 1) - t afs)/len with the
    // un-truncated r_i = v_i +- 0.5 (harvest.py:178-181); theta advances by 2 pi/len per sample
    const bool fast = ((t + (double)(0 - half) * inv_afs) * afs + 0.001) > 0.0;  // no sample before t = 0
    double cr = 0.0, ci = 0.0, wr = 0.0, wi = 0.0;
    if (fast) {
      const double v0 = (t + (double)(0 - half) * inv_afs) * afs + 0.001;
      wb_sincospi(2.0 * ((v0 + 0.5 - 1.0) - t * afs) * inv_len, &ci, &cr);
      wb_sincospi(2.0 * inv_len, &wi, &wr);
    }
    auto produce = [&](int i, double& m_out, double& seg_out) {  // window value and signal sample of position i
      double c1;
      const double v = (t + (double)(i - half) * inv_afs) * afs + 0.001;
      const double r = v > 0.0 ? v + 0.5 : v - 0.5;
      if (fast) {
        c1 = cr;
        const double nr = cr * wr - ci * wi;
        ci = cr * wi + ci * wr;
        cr = nr;
      } else {
        double sn_;
        wb_sincospi(2.0 * ((r - 1.0) - t * afs) * inv_len, &sn_, &c1);
      }
      const double rc = r < 1.0 ? 1.0 : (r > (double)ylen ? (double)ylen : r);
      m_out = 0.42 + 0.5 * c1 + 0.08 * (2.0 * c1 * c1 - 1.0);
      seg_out = WB_LDG(yu + ((int)rc - 1));
    };
    // one Goertzel step of every bin for the sample `seg` with window values (before, at, after) it; the new
    // state overwrites the older one, so the two arrays swap roles at every step
    auto consume = [&](double seg, double m_before, double m_at, double m_after, double (&s1)[6], double (&s2)[6],
                       double (&d1)[6], double (&d2)[6]) {
      const double a = seg * m_at;
      const double b = seg * (-(m_after - m_before) / 2.0);
#pragma unroll
      for (int hh = 0; hh < 6; ++hh) {
        s2[hh] = fma(coef[hh], s1[hh], a - s2[hh]);
        d2[hh] = fma(coef[hh], d1[hh], b - d2[hh]);
      }
    };
    double m0 = 0.0, m1, sg1;
    produce(0, m1, sg1);
    for (int i = 1; i + 1 <= len; i += 2) {  // two samples per trip: no register shuffling between the state arrays
      double m2, sg2, m3 = 0.0, sg3 = 0.0;
      produce(i, m2, sg2);
      consume(sg1, m0, m1, m2, ga1, ga2, gb1, gb2);
      if (i + 1 < len) produce(i + 1, m3, sg3);
      consume(sg2, m1, m2, m3, ga2, ga1, gb2, gb1);
      m0 = m2;
      m1 = m3;
      sg1 = sg3;
    }
    // len is odd: one sample is left, and after it the newest state sits in ga2 / gb2
    consume(sg1, m0, m1, 0.0, ga1, ga2, gb1, gb2);
    return score_of(q, c0, ga1, ga2, gb1, gb2);
  }

  // work item: (frame index << 14) | (candidate index << 7) | source code; source code = shift * 15 + row of the
  // candidate in frame j - 3 + shift, or WB_HV_SRC_QUIRK (row 6 of frame j itself, stored in slot 0)
  WB_DEV const double* source_of(unsigned long long d) const {
    const long long fi = (long long)(d >> 14);
    const int code = (int)(d & 127ull);
    if (code == WB_HV_SRC_QUIRK) return p.base_c + (size_t)fi * WB_HV_MAXC + 6;
    const int s = code / WB_HV_MAXC, k = code - s * WB_HV_MAXC;
    return p.base_c + (size_t)(fi - 3 + s) * WB_HV_MAXC + k;
  }

  WB_DEV void refine_all(int block, int tid, int nthr) const {
    const long long total = cls_count[WB_HV_NCLS];
    const long long stride = (long long)p.n_slots * nthr;
    long long g = (long long)block * nthr + tid;
    if (g >= total) return;
    // the next item's descriptor, candidate and signal length are fetched while the current window is walked
    auto utterance_of = [&](unsigned long long dd) {
      const unsigned long long f = dd >> 14;
      return (f >> 32) ? (int)(f / (unsigned long long)p.f1_stride) : (int)((unsigned)f / (unsigned)p.f1_stride);
    };
    unsigned long long d = items[g];
    double c0 = WB_LDG(source_of(d));
    int u = utterance_of(d);
    int ylen = p.y_len[u];
    const double afs = p.afs, inv_afs = 1.0 / p.afs;
    for (; g < total; g += stride) {
      const bool more = g + stride < total;
      const unsigned long long d_next = more ? items[g + stride] : 0ull;
      const long long fi = (long long)(d >> 14);
      const int it = (int)((d >> 7) & 127ull), code = (int)(d & 127ull);
      const int slot = code == WB_HV_SRC_QUIRK ? 0 : code;
      const int j = (int)(fi - (long long)u * p.f1_stride);
      const double* yu = p.y + (size_t)u * p.y_stride;
      const double t = (double)j / 1000.0;
      const geometry q = geometry_of(c0);
      const int half = q.half, len = q.len;
      // Sample index of position i: trunc(r_i) - 1 with r_i = (t + (i - half)/afs) afs + 0.501.  r advances by one
      // per sample, so when the first and the last index are len - 1 apart every index in between is first + i;
      // otherwise (rounding put a step across an integer), and for windows that start before t = 0, each position
      // is evaluated on its own (refine_general).
      const double v0 = (t + (double)(0 - half) * inv_afs) * afs + 0.001;
      const double r_b = ((t + (double)(len - 1 - half) * inv_afs) * afs + 0.001) + 0.5;
      const int idx_first = (int)(v0 + 0.5) - 1;
      refined res;
      double c0_next = 0.0;
      int ylen_next = 0, u_next = 0;
      if (v0 > 0.0 && ((int)r_b - 1) - idx_first == len - 1) {
        // Hot path, 31 FP64 instructions per sample: the window cosine by the difference form of the Chebyshev
        // recurrence (c += d; d -= 4 sin^2(delta/2) c: errors grow linearly, not with 1/delta), the Blackman
        // polynomial in Horner form (0.42 + 0.5 c + 0.08 (2 c^2 - 1) = 0.34 + 0.5 c + 0.16 c^2), the derivative
        // window without its factor -1/2 (exact scaling, applied to the final state), 12 Goertzel steps written
        // s2 <- fma(coef, s1, x - s2): the subtraction does not wait for the newest state, so consecutive steps
        // are one FMA apart.  The inputs (a, b) of a step are prepared one step ahead, samples come through a
        // running pointer, four per trip, the next four already in flight.  Windows that leave the signal clamp
        // the index.
        double ga1[6], ga2[6], gb1[6], gb2[6], coef[6];
#pragma unroll
        for (int hh = 0; hh < 6; ++hh) {
          ga1[hh] = ga2[hh] = gb1[hh] = gb2[hh] = 0.0;
          const int bin = (int)(q.bin_scale * (hh + 1) + 0.5);
          coef[hh] = 2.0 * wb_ldg_cplx(tw + (size_t)(bin & (q.nfft - 1)) * q.stepw).x;
        }
        double sh_, ch_, sm_, s0_, cw;
        wb_sincospi(q.inv_len, &sh_, &ch_);
        const double a0 = 2.0 * ((v0 + 0.5 - 1.0) - t * afs) * q.inv_len;
        wb_sincospi(a0, &s0_, &cw);                // cos(theta_0)
        wb_sincospi(a0 + q.inv_len, &sm_, &ch_);   // sin(theta_0 + delta / 2)
        const double nkap = -4.0 * sh_ * sh_;
        double dw = -2.0 * sm_ * sh_;              // cos(theta_1) - cos(theta_0)
        if (more) {  // issued here, consumed after the walk
          c0_next = WB_LDG(source_of(d_next));
          u_next = utterance_of(d_next);
          ylen_next = p.y_len[u_next];
        }
        double mp, mc, A, Bv;  // m_i, m_{i+1}, inputs of step i
        auto advance = [&]() {
          cw += dw;
          dw = fma(nkap, cw, dw);
          return fma(fma(0.16, cw, 0.5), cw, 0.34);
        };
        auto update = [&](double (&s1)[6], double (&s2)[6], double (&d1)[6], double (&d2)[6]) {
#pragma unroll
          for (int hh = 0; hh < 6; ++hh) {
            s2[hh] = A - s2[hh];
            d2[hh] = Bv - d2[hh];
          }
#pragma unroll
          for (int hh = 0; hh < 6; ++hh) {
            s2[hh] = fma(coef[hh], s1[hh], s2[hh]);
            d2[hh] = fma(coef[hh], d1[hh], d2[hh]);
          }
        };
        // step i: prepare the inputs of step i + 1 from its sample, then update the states with the inputs of step i
        auto step = [&](double seg_next, double (&s1)[6], double (&s2)[6], double (&d1)[6], double (&d2)[6]) {
          const double mn = advance();  // m_{i+2}
          const double An = seg_next * mc, Bn = seg_next * (mn - mp);
          update(s1, s2, d1, d2);
          A = An;
          Bv = Bn;
          mp = mc;
          mc = mn;
        };
        const int n_gen = len - 2;  // steps followed by a full window value two positions on (odd)
        auto run = [&](auto ld) {
          {
            const double seg0 = ld(0);
            mp = fma(fma(0.16, cw, 0.5), cw, 0.34);  // m_0
            mc = advance();                          // m_1
            A = seg0 * mp;
            Bv = seg0 * (mc - 0.0);
          }
          int i = 0;
          double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
          if (4 <= n_gen) {
            x0 = ld(1);
            x1 = ld(2);
            x2 = ld(3);
            x3 = ld(4);
          }
          for (; i + 4 <= n_gen; i += 4) {
            double y0 = 0.0, y1 = 0.0, y2 = 0.0, y3 = 0.0;
            if (i + 8 <= n_gen) {
              y0 = ld(i + 5);
              y1 = ld(i + 6);
              y2 = ld(i + 7);
              y3 = ld(i + 8);
            }
            step(x0, ga1, ga2, gb1, gb2);
            step(x1, ga2, ga1, gb2, gb1);
            step(x2, ga1, ga2, gb1, gb2);
            step(x3, ga2, ga1, gb2, gb1);
            x0 = y0;
            x1 = y1;
            x2 = y2;
            x3 = y3;
          }
          if (n_gen - i == 3) {
            const double u0 = ld(i + 1), u1 = ld(i + 2);
            step(u0, ga1, ga2, gb1, gb2);
            step(u1, ga2, ga1, gb2, gb1);
            i += 2;
          }
          step(ld(i + 1), ga1, ga2, gb1, gb2);  // the last step with a full window value two positions on
          {                                     // step len - 2: the window value after the last sample counts as 0
            const double seg_last = ld(len - 1);
            const double An = seg_last * mc, Bn = seg_last * (0.0 - mp);
            update(ga2, ga1, gb2, gb1);
            A = An;
            Bv = Bn;
          }
          update(ga1, ga2, gb1, gb2);  // step len - 1
        };
        if (idx_first >= 0 && idx_first + len <= ylen) {
          const double* yp = yu + idx_first;
          run([&](int k) { return WB_LDG(yp + k); });
        } else {
          run([&](int k) {
            const int yi = idx_first + k;
            return WB_LDG(yu + (yi < 0 ? 0 : (yi > ylen - 1 ? ylen - 1 : yi)));
          });
        }
#pragma unroll
        for (int hh = 0; hh < 6; ++hh) {
          gb1[hh] *= -0.5;
          gb2[hh] *= -0.5;
        }
        res = score_of(q, c0, ga1, ga2, gb1, gb2);
      } else {
        if (more) {
          c0_next = WB_LDG(source_of(d_next));
          u_next = utterance_of(d_next);
          ylen_next = p.y_len[u_next];
        }
        res = refine_general(yu, ylen, t, c0);
      }
      const size_t o = (size_t)fi * WB_HV_SLOTS + it;
      p.l_f0[o] = res.f0;
      p.l_sc[o] = res.score;
      p.l_slot[o] = (unsigned char)slot;
      d = d_next;
      c0 = c0_next;
      u = u_next;
      ylen = ylen_next;
    }
  }
};

struct wb_hv_refine_prep : wb_hv_refine_items {  // modes 0 (count) and 1 (scatter)
  WB_DEV void operator()(int block, int tid, int nthr, double* smem) const { prepare(block, tid, nthr, smem); }
};

// class counts -> class offsets (exclusive prefix) and cursors; one thread
struct wb_hv_refine_scan {
  int* cls_count;
  int* cls_cursor;
  WB_DEV void operator()(long long) const {
    int a = 0;
    for (int c = 0; c < WB_HV_NCLS; ++c) {
      const int v = cls_count[c];
      cls_cursor[c] = a;
      a += v;
    }
    cls_count[WB_HV_NCLS] = a;
  }
};

// ------------------------------------------------------------------------------------ H5
// One thread per (utterance, frame): keep flag of every candidate (RemoveUnreliableCandidates).
struct wb_hv_prune {
  wb_hv_plan p;
  WB_DEV double nearest(double ref, size_t b) const {  // SelectBestF0(ref, column, 1)'s error
    // min_q fl(|ref - f_q| / ref) == fl(min_q |ref - f_q| / ref): rounding a quotient by the same positive divisor
    // is monotone, so one division serves the whole column
    const int n = p.l_n[b];
    if (n <= 0) return 1.0;
    double dmin = fabs(ref - p.l_f0[b * WB_HV_SLOTS]);
    for (int q = 1; q < n; ++q) dmin = wb_dmin(dmin, fabs(ref - p.l_f0[b * WB_HV_SLOTS + q]));
    const double e = dmin / ref;
    return e < 1.0 ? e : 1.0;
  }
  // one group of WB_LANES threads per frame: the lanes take the frame's candidates, every lane scans the two
  // neighbour columns (same addresses across the lanes: broadcast loads)
  WB_DEV void operator()(long long item) const {
    const long long frame = item / WB_LANES;
    const int lane = (int)(item - frame * WB_LANES);
    const int u = (int)(frame / p.f1_stride), j = (int)(frame - (long long)u * p.f1_stride);
    const int f1 = wb_hv_frames(p.n_samples[u], p.fs, 1.0);
    if (j >= f1) return;
    const size_t b = (size_t)u * p.f1_stride + j;
    const int n = p.l_n[b];
    for (int q = lane; q < n; q += WB_LANES) {
      unsigned char keep = 1;
      const double ref0 = p.l_f0[b * WB_HV_SLOTS + q];
      if (ref0 == 0.0) {  // rejected by the refinement
        p.l_keep[b * WB_HV_SLOTS + q] = 0;
        continue;
      }
      if (j >= 1 && j <= f1 - 2) {
        const double ref = ref0;
        const double e = wb_dmin(nearest(ref, b + 1), nearest(ref, b - 1));
        if (e > 0.05) keep = 0;
      }
      p.l_keep[b * WB_HV_SLOTS + q] = keep;
    }
  }
};

// ------------------------------------------------------------------------------------ H6
// One warp per utterance.  Control flow is uniform across lanes; lane 0 does the scalar writes,
// candidate scans and long sums are spread over the lanes.
struct wb_hv_contour {
  wb_hv_plan p;

  WB_HD static int max_runs(int f1) { return f1 / 2 + 2; }
  WB_HD static long long pool_len(int f1) { return (long long)f1 + 204LL * (f1 / 7 + 2); }
  WB_HD static long long scratch_doubles(int f1) {
    return 6LL * f1 + 2LL * (f1 + 600) + (long long)WB_LANES * (f1 + 600) + pool_len(f1) + 8LL * max_runs(f1) + 64;
  }

  // SelectBestF0 (harvest.py:238-248) over the kept candidates of one frame
  WB_DEV double select_best(double ref, size_t b, double tol, int lane, int lanes) const {
    double key = 1e300, payload = 0.0;
    int tag = -1;
    const int n = p.l_n[b];
    for (int q = lane; q < n; q += lanes) {
      if (!p.l_keep[b * WB_HV_SLOTS + q]) continue;
      const double f = p.l_f0[b * WB_HV_SLOTS + q];
      const double e = fabs(ref - f) / ref;
      const int slot = p.l_slot[b * WB_HV_SLOTS + q];
      if (e <= tol && (e < key || (e == key && slot > tag))) {
        key = e;
        tag = slot;
        payload = f;
      }
    }
    wb_lanes_argmin(key, tag, payload);
    return tag >= 0 ? payload : 0.0;
  }
  // SerachScore (harvest.py:488-495)
  WB_DEV double search_score(double value, size_t b) const {
    double s = 0.0;
    const int n = p.l_n[b];
    for (int q = 0; q < n; ++q)
      if (p.l_keep[b * WB_HV_SLOTS + q] && p.l_f0[b * WB_HV_SLOTS + q] == value && s < p.l_sc[b * WB_HV_SLOTS + q])
        s = p.l_sc[b * WB_HV_SLOTS + q];
    return s;
  }
  // GetBoundaryList (harvest.py:572-580): inclusive runs of non-zeros, ends treated as zero.  Lane 0 only.
  WB_DEV int list_runs(const double* f, int n, int* st, int* ed, int cap) const {
    int count = 0, start = -1;
    for (int i = 1; i < n; ++i) {
      const bool on = (i < n - 1) && (f[i] != 0.0);
      if (on && start < 0) start = i;
      if (!on && start >= 0) {
        if (count < cap) {
          st[count] = start;
          ed[count] = i - 1;
          ++count;
        }
        start = -1;
      }
    }
    return count;
  }

  // The same, run by all the lanes of ONE warp (every lane must call it; every lane gets the count): 32 elements per
  // trip through a ballot, run boundaries from the bit mask.  The serial walk above is one dependent global load
  // per element; this is one coalesced load per 32.
  WB_DEV int list_runs_lanes(const double* f, int n, int* st, int* ed, int cap, int lane) const {
#ifdef WB_HOST_EMU
    (void)lane;
    return list_runs(f, n, st, ed, cap);
#else
    int count = 0, start = -1;
    unsigned carry = 0;  // "on" state of the element before the chunk
    for (int base = 1; base < n; base += 32) {
      const int i = base + lane;
      const bool on = (i < n - 1) && (f[i] != 0.0);
      const unsigned m = __ballot_sync(0xffffffffu, on);
      const unsigned prev = (m << 1) | carry;
      unsigned ev = m ^ prev;  // positions where the state changes
      while (ev) {
        const int b = __ffs((int)ev) - 1;
        ev &= ev - 1;
        if ((m >> b) & 1u) {
          start = base + b;
        } else {
          if (count < cap && lane == 0) {
            st[count] = start;
            ed[count] = base + b - 1;
          }
          if (count < cap) ++count;
          start = -1;
        }
      }
      carry = m >> 31;
    }
    __syncwarp();
    return count;
#endif
  }

  WB_DEV void operator()(int block, int tid, int nthr, double* smem) const {
    const int u = block;
    // the block is a set of warps: block-wide loops use (tid, nthr), the per-run tracking of step 3 runs one
    // run per warp with (lane, lanes)
    const int lanes = WB_LANES < nthr ? WB_LANES : nthr;
    const int nw = nthr / lanes, w = tid / lanes, lane = tid - w * lanes;
    const int F = wb_hv_frames(p.n_samples[u], p.fs, 1.0);
    const int n5 = p.out_n_frames[u];
    double* outf = p.out_f0 + (size_t)u * p.f_stride;
    double* outv = p.out_vuv + (size_t)u * p.f_stride;
    double* outt = p.out_tpos + (size_t)u * p.f_stride;
    for (int k = tid; k < n5; k += nthr) outt[k] = (double)k * p.frame_period / 1000.0;
    if (F < 4) {
      for (int k = tid; k < n5; k += nthr) {
        outf[k] = 0.0;
        outv[k] = 0.0;
      }
      return;
    }
    double* W = p.ctr + (size_t)u * p.ctr_stride;
    double* base = W;
    double* s1 = base + F;
    double* s2 = s1 + F;
    double* s3 = s2 + F;
    double* s4 = s3 + F;
    double* spare = s4 + F;
    const int Lp = F + 600;
    double* P = spare + F;
    double* smo = P + Lp;
    double* lane_buf = smo + Lp;                       // WB_LANES rows of Lp
    double* pool = lane_buf + (size_t)WB_LANES * Lp;
    const long long pool_cap = pool_len(F);
    const int MR = max_runs(F);
    int* r_st = (int*)(pool + pool_cap);
    int* r_ed = r_st + MR;
    int* t_off = r_ed + MR;  // kept tracks: pool offset (int, pool_cap < 2^31 for F < ~70M)
    int* t_w0 = t_off + MR;  // first frame stored in the window
    int* t_w1 = t_w0 + MR;   // last frame stored
    int* t_lo = t_w1 + MR;   // span (range) of the track
    int* t_hi = t_lo + MR;
    int* order = t_hi + MR;
    int* scal = order + MR;  // [0] run count, [1] kept tracks
    // per-run scratch of step 3, overlaid on the smoothing arrays that are written only later
    int* q_off = (int*)P;
    int* q_lo = q_off + MR;
    int* q_hi = q_lo + MR;   // -1: track dropped
    const size_t fb = (size_t)u * p.f1_stride;

    // SearchF0Base (harvest.py:315-320): best-scored kept candidate, first maximum
    for (int j = tid; j < F; j += nthr) {
      const int n = p.l_n[fb + j];
      double best = 0.0, val = 0.0;
      int tag = 1 << 30;
      for (int q = 0; q < n; ++q) {
        if (!p.l_keep[(fb + j) * WB_HV_SLOTS + q]) continue;
        const double sc = p.l_sc[(fb + j) * WB_HV_SLOTS + q];
        const int slot = p.l_slot[(fb + j) * WB_HV_SLOTS + q];
        if (sc > best || (sc == best && sc > 0.0 && slot < tag)) {
          best = sc;
          tag = slot;
          val = p.l_f0[(fb + j) * WB_HV_SLOTS + q];
        }
      }
      base[j] = val;
    }
    WB_SYNC();
    // FixStep1 (harvest.py:324-338)
    for (int j = tid; j < F; j += nthr) {
      double v = base[j];
      if (j < 2) {
        v = 0.0;
      } else if (v != 0.0) {
        const double ref = base[j - 1] * 2.0 - base[j - 2];
        if (fabs((v - ref) / (ref + WB_EPS)) > 0.008 && fabs((v - base[j - 1]) / (base[j - 1] + WB_EPS)) > 0.008) v = 0.0;
      }
      s1[j] = v;
      s2[j] = v;
    }
    WB_SYNC();
    // FixStep2 (harvest.py:343-352), then the window every run gets in the track pool
    if (w == 0) {  // the first warp lists the runs together; lane 0 does the short serial parts
      const int nr = list_runs_lanes(s1, F, r_st, r_ed, MR, lane);
      if (lane == 0)
        for (int r = 0; r < nr; ++r)
          if (r_ed[r] - r_st[r] < 6)
            for (int i = r_st[r]; i <= r_ed[r]; ++i) s2[i] = 0.0;
      wb_lanes_sync();
      int nr2 = list_runs_lanes(s2, F, r_st, r_ed, MR, lane);
      wb_lanes_sync();
      if (lane == 0) {
        long long cur = 0;
        for (int r = 0; r < nr2; ++r) {
          const int w0 = wb_imax(0, r_st[r] - 101), w1 = wb_imin(F - 1, r_ed[r] + 101);
          if (cur + (w1 - w0 + 1) > pool_cap) {
            p.status[0] = 2;
            nr2 = r;
            break;
          }
          q_off[r] = (int)cur;
          cur += w1 - w0 + 1;
        }
        scal[0] = nr2;
        scal[1] = 0;
      }
    }
    WB_SYNC();
    // FixStep3 (harvest.py:357-384): extend each run along the candidates (one run per warp), keep the long ones
    const int n_runs = scal[0];
    for (int r = w; r < n_runs; r += nw) {
      const int st = r_st[r], ed = r_ed[r];
      const int w0 = wb_imax(0, st - 101), w1 = wb_imin(F - 1, ed + 101);
      double* seq = pool + q_off[r] - w0;  // seq[i] valid for w0 <= i <= w1
      for (int i = w0 + lane; i <= w1; i += lanes) seq[i] = (i >= st && i <= ed) ? s2[i] : 0.0;
      wb_lanes_sync();
      int hi = ed, lo = st;
      {
        double cur = seq[ed];
        int misses = 0;
        const int stop = wb_imin(F - 2, ed + 100);
        for (int i = ed; i <= stop; ++i) {
          const double v = select_best(cur, fb + i + 1, 0.18, lane, lanes);
          if (lane == 0) seq[i + 1] = v;
          if (v != 0.0) {
            cur = v;
            misses = 0;
            hi = i + 1;
          } else {
            ++misses;
          }
          if (misses == 4) break;
        }
      }
      {
        double cur = seq[st];
        int misses = 0;
        const int stop = wb_imax(1, st - 100);
        for (int i = st; i >= stop; --i) {
          const double v = select_best(cur, fb + i - 1, 0.18, lane, lanes);
          if (lane == 0) seq[i - 1] = v;
          if (v != 0.0) {
            cur = v;
            misses = 0;
            lo = i - 1;
          } else {
            ++misses;
          }
          if (misses == 4) break;
        }
      }
      wb_lanes_sync();
      double acc = 0.0;
      for (int i = lo + lane; i <= hi; i += lanes) acc += seq[i];
      acc = wb_lanes_sum(acc);
      const double mean = acc / (double)(hi - lo + 1);
      if (lane == 0) {
        q_lo[r] = lo;
        q_hi[r] = (2200.0 / mean < (double)(hi - lo)) ? hi : -1;
      }
    }
    WB_SYNC();
    if (tid == 0) {  // kept tracks, in run order
      int k = 0;
      for (int r = 0; r < n_runs; ++r) {
        if (q_hi[r] < 0) continue;
        t_off[k] = q_off[r];
        t_w0[k] = wb_imax(0, r_st[r] - 101);
        t_w1[k] = wb_imin(F - 1, r_ed[r] + 101);
        t_lo[k] = q_lo[r];
        t_hi[k] = q_hi[r];
        ++k;
      }
      scal[1] = k;
    }
    WB_SYNC();
    // MergeF0 (harvest.py:437-484)
    const int n_trk = scal[1];
    if (n_trk == 0) {
      for (int j = tid; j < F; j += nthr) s3[j] = s2[j];
      WB_SYNC();
    } else {
      if (tid == 0) {  // stable insertion sort by span start
        for (int k = 0; k < n_trk; ++k) {
          int pos = k;
          while (pos > 0 && t_lo[order[pos - 1]] > t_lo[k]) {
            order[pos] = order[pos - 1];
            --pos;
          }
          order[pos] = k;
        }
      }
      WB_SYNC();
      {
        const int k0 = order[0];
        const double* seq = pool + t_off[k0] - t_w0[k0];
        for (int j = tid; j < F; j += nthr) s3[j] = (j >= t_w0[k0] && j <= t_w1[k0]) ? seq[j] : 0.0;
      }
      WB_SYNC();
      int st1 = t_lo[order[0]], ed1 = t_hi[order[0]];
      for (int m = 1; m < n_trk; ++m) {
        const int k = order[m];
        const int st2 = t_lo[k], ed2 = t_hi[k];
        const double* seq = pool + t_off[k] - t_w0[k];
        const int w0 = t_w0[k], w1 = t_w1[k];
        if (st2 - ed1 > 0) {
          for (int j = st2 + tid; j <= ed2; j += nthr) s3[j] = (j >= w0 && j <= w1) ? seq[j] : 0.0;
          st1 = st2;
          ed1 = ed2;
        } else if (st1 <= st2 && ed1 >= ed2) {
          // completely covered: nothing to merge
        } else {
          double a = 0.0, b = 0.0;
          for (int i = st2 + tid; i <= ed1; i += nthr) {
            a += search_score(s3[i], fb + i);
            b += search_score((i >= w0 && i <= w1) ? seq[i] : 0.0, fb + i);
          }
          a = wb_block_sum(a, smem, tid, nthr);
          b = wb_block_sum(b, smem, tid, nthr);
          const int from = (a > b) ? ed1 : st2;
          WB_SYNC();
          for (int j = from + tid; j <= ed2; j += nthr) s3[j] = (j >= w0 && j <= w1) ? seq[j] : 0.0;
          ed1 = ed2;
        }
        WB_SYNC();
      }
    }
    // FixStep4 (harvest.py:389-405)
    for (int j = tid; j < F; j += nthr) s4[j] = s3[j];
    WB_SYNC();
    if (w == 0) {
      const int nr = list_runs_lanes(s3, F, r_st, r_ed, MR, lane);
      wb_lanes_sync();
      for (int r = 0; lane == 0 && r + 1 < nr; ++r) {
        const int e = r_ed[r], s = r_st[r + 1];
        const int gap = s - e - 1;
        if (gap >= 9) continue;
        const double lo = s3[e] + 1.0, hi = s3[s] - 1.0;
        const double slope = (hi - lo) / (double)(gap + 1);
        int c = 1;
        for (int j = e + 1; j < s; ++j, ++c) s4[j] = lo + slope * (double)c;
      }
    }
    WB_SYNC();
    // SmoothF0 (harvest.py:533-559)
    for (int i = tid; i < Lp; i += nthr) {
      const double v = (i >= 300 && i < 300 + F) ? s4[i - 300] : 0.0;
      P[i] = v;
      smo[i] = v;
    }
    WB_SYNC();
    if (w == 0) {
      const int nr = list_runs_lanes(P, Lp, r_st, r_ed, MR, lane);
      if (lane == 0) scal[0] = nr;
    }
    WB_SYNC();
    {
      const int nr = scal[0];
      const double b0 = 0.0078202080334971724, b1 = 0.015640416066994345, b2 = 0.0078202080334971724;
      const double a1 = -1.7347257688092754, a2 = 0.76600660094326412;
      double* fw = lane_buf + (size_t)lane * Lp;
      for (int r = (w == 0 ? lane : nr); r < nr; r += lanes) {  // one run per lane of the first warp
        const int st = r_st[r], ed = r_ed[r];
        // The reference filters the whole padded array; the 300-sample zero padding it adds is
        // its own bound on the transient (pole radius 0.875), so the passes start 300 samples out.
        const int a0 = wb_imax(0, st - 300), e1 = wb_imin(Lp - 1, ed + 600);
        const double cl = P[st], cr = P[ed];
        double z0 = 0.0, z1 = 0.0;
        for (int n = a0; n <= e1; ++n) {
          const double xin = n < st ? cl : (n > ed ? cr : P[n]);
          const double o = b0 * xin + z0;
          z0 = b1 * xin - a1 * o + z1;
          z1 = b2 * xin - a2 * o;
          fw[n - a0] = o;
        }
        z0 = 0.0;
        z1 = 0.0;
        for (int n = e1; n >= st; --n) {
          const double xin = fw[n - a0];
          const double o = b0 * xin + z0;
          z0 = b1 * xin - a1 * o + z1;
          z1 = b2 * xin - a2 * o;
          if (n <= ed) smo[n] = o;
        }
      }
    }
    WB_SYNC();
    // pick the frame_period grid (harvest.py:46-53)
    for (int k = tid; k < n5; k += nthr) {
      const double t = (double)k * p.frame_period / 1000.0;
      const double v = t * 1000.0;
      int idx = (int)(v > 0.0 ? v + 0.5 : v - 0.5);
      if (idx > F - 1) idx = F - 1;
      outf[k] = smo[300 + idx];
      outv[k] = s4[idx] != 0.0 ? 1.0 : 0.0;
    }
  }
};
