"""Batched, device-resident API: thin Python over the C-ABI (include/world_b200.h).

All tensors are torch CUDA tensors used as HBM containers; every method enqueues
work on the current CUDA stream and returns device tensors without synchronising.
Layouts: waveforms [B, S]; per-frame vectors [B, F]; per-frame matrices [B, F, bins].
"""
import ctypes

import torch

from . import _abi, _lib


class WorldB200Error(Exception):
    pass


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class Engine:
    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("world_b200: no CUDA device visible; this engine has no CPU path")
        self.L = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device) \
            if not isinstance(device, torch.device) else device
        h = ctypes.c_void_p()
        rc = self.L.wb_create(ctypes.byref(h), self.device.index or 0)
        if rc != 0:
            raise WorldB200Error("wb_create failed: %d" % rc)
        self.h = h
        self._ws = {}
        self._side = []

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.wb_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _check(self, rc):
        if rc != 0:
            msg = self.L.wb_last_error(self.h).decode()
            if rc == -1:
                raise AssertionError(msg)
            raise WorldB200Error("rc=%d: %s" % (rc, msg))

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def f64(self, a):
        return torch.as_tensor(a, dtype=torch.float64, device=self.device).contiguous()

    def i32(self, a):
        return torch.as_tensor(a, dtype=torch.int32, device=self.device).contiguous()

    def empty(self, *shape, dtype=torch.float64):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    ptr = staticmethod(_p)

    # ------------------------------------------------------------------ PCM edge (example/prosody.py:12-13, 57)
    def pcm16_to_f64(self, pcm, n_samples, divisor=2 ** 15 - 1, x_stride=None):
        """int16 [B, S] (device) -> float64 [B, x_stride] = pcm / divisor, zero past n_samples."""
        B, S = pcm.shape
        xs = int(S if x_stride is None else x_stride)
        x = self.empty(B, xs)
        self._check(self.L.wb_pcm16_to_f64(self.h, self._stream(), _p(pcm), S, _p(n_samples), B, float(divisor), _p(x), xs))
        return x

    def f64_to_pcm16(self, y, n_samples, gain=2 ** 15):
        """float64 [B, S] -> int16 [B, S] = (y * gain).astype(int16), zero past n_samples."""
        B, S = y.shape
        pcm = self.empty(B, S, dtype=torch.int16)
        self._check(self.L.wb_f64_to_pcm16(self.h, self._stream(), _p(y), S, _p(n_samples), B, float(gain), _p(pcm), S))
        return pcm

    def to_f32(self, t):
        """float64 tensor -> float32 (round to nearest) for the optional compact transport of the spectrogram."""
        t = t.contiguous()
        out = torch.empty(t.shape, dtype=torch.float32, device=self.device)
        self._check(self.L.wb_f64_to_f32(self.h, self._stream(), _p(t), t.numel(), _p(out)))
        return out

    def to_f64(self, t):
        """float32 tensor -> float64 (exact)."""
        t = t.contiguous()
        out = torch.empty(t.shape, dtype=torch.float64, device=self.device)
        self._check(self.L.wb_f32_to_f64(self.h, self._stream(), _p(t), t.numel(), _p(out)))
        return out

    # ------------------------------------------------------------------ stages
    def cheaptrick(self, x, n_samples, fs, tpos, f0, vuv, n_frames, q1=-0.15, fft_size=None,
                   dither=None, want_ps=False, seed=0):
        """world/cheaptrick.py:9.  Returns (f0_used [B,F], spectrogram [B,F,N/2+1], ps [B,F,N] | None)."""
        B, S = x.shape
        F = tpos.shape[1]
        n = int(fft_size) if fft_size else self.L.wb_cheaptrick_fft_size(int(fs))
        f0_used = self.empty(B, F)
        spec = self.empty(B, F, n // 2 + 1)
        ps = self.empty(B, F, n, dtype=torch.complex128) if want_ps else None
        self._check(self.L.wb_cheaptrick(self.h, self._stream(), _p(x), S, _p(n_samples), B, int(fs), _p(tpos),
                                         _p(f0), _p(vuv), _p(n_frames), F, float(q1), n, _p(dither), int(seed),
                                         _p(f0_used), _p(spec), _p(ps)))
        return f0_used, spec, ps

    def d4c(self, x, n_samples, fs, tpos, f0, vuv, n_frames, threshold=0.85, fft_size_for_spectrum=None,
            want_coarse=False):
        """world/d4c.py:10.  Returns (f0_out [B,F], aperiodicity [B,F,Ns/2+1], coarse_ap [B,F,bands] | None)."""
        B, S = x.shape
        F = tpos.shape[1]
        nsp = int(fft_size_for_spectrum) if fft_size_for_spectrum else self.L.wb_cheaptrick_fft_size(int(fs))
        nb = self.L.wb_d4c_band_count(int(fs), 0)
        f0_out = self.empty(B, F)
        ap = self.empty(B, F, nsp // 2 + 1)
        coarse = self.empty(B, F, max(nb, 1)) if want_coarse else None
        self._check(self.L.wb_d4c(self.h, self._stream(), _p(x), S, _p(n_samples), B, int(fs), _p(tpos), _p(f0),
                                  _p(vuv), _p(n_frames), F, float(threshold), nsp, _p(f0_out), _p(ap), _p(coarse)))
        return f0_out, ap, coarse

    def d4c_requiem(self, x, n_samples, fs, tpos, f0, vuv, n_frames, threshold=0.85, fft_size=None):
        """world/d4cRequiem.py:9.  Returns (f0_out [B,F], band_aperiodicity [B,F,bands+2] in dB)."""
        B, S = x.shape
        F = tpos.shape[1]
        nb = self.L.wb_d4c_band_count(int(fs), 1)
        f0_out = self.empty(B, F)
        ap = self.empty(B, F, max(nb, 0) + 2)
        self._check(self.L.wb_d4c_requiem(self.h, self._stream(), _p(x), S, _p(n_samples), B, int(fs), _p(tpos),
                                          _p(f0), _p(vuv), _p(n_frames), F, float(threshold),
                                          int(fft_size) if fft_size else 0, _p(f0_out), _p(ap)))
        return f0_out, ap



    # ------------------------------------------------------------------ F0
    def _workspace(self, key, nbytes):
        """Scratch buffer of a stage, one per CUDA stream: calls on different streams never share scratch."""
        key = (key, torch.cuda.current_stream(self.device).cuda_stream)
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    def harvest(self, x, n_samples, fs, f0_floor=71.0, f0_ceil=800.0, frame_period=5.0, max_samples=None):
        """world/harvest.py:17.  Returns (temporal_positions, f0, vuv [B,F] float64, n_frames [B] int32)."""
        B, S = x.shape
        smax = int(S if max_samples is None else max_samples)
        nbytes = ctypes.c_size_t()
        self._check(self.L.wb_harvest_workspace_bytes(self.h, B, smax, int(fs), float(f0_floor), float(f0_ceil),
                                                      ctypes.byref(nbytes)))
        ws = self._workspace("harvest", nbytes.value)
        F = self.L.wb_frame_count(smax, int(fs), float(frame_period))
        tpos, f0, vuv = self.empty(B, F), self.empty(B, F), self.empty(B, F)
        n_frames = self.empty(B, dtype=torch.int32)
        self._check(self.L.wb_harvest(self.h, self._stream(), _p(x), S, _p(n_samples), B, smax, int(fs),
                                      float(f0_floor), float(f0_ceil), float(frame_period), _p(ws), nbytes.value, F,
                                      _p(tpos), _p(f0), _p(vuv), _p(n_frames)))
        return tpos, f0, vuv, n_frames

    def dio(self, x, n_samples, fs, f0_floor=71.0, f0_ceil=800.0, channels_in_octave=2, target_fs=4000,
            frame_period=5.0, allowed_range=0.1, max_samples=None, want_candidates=False):
        """world/dio.py:10.  Returns (temporal_positions, f0, vuv [B,F], n_frames [B]) and, when
        want_candidates, (f0_candidates [B,F,bands], raw_f0_candidates [B,bands,F])."""
        B, S = x.shape
        smax = int(S if max_samples is None else max_samples)
        nbytes = ctypes.c_size_t()
        self._check(self.L.wb_dio_workspace_bytes(self.h, B, smax, int(fs), float(f0_floor), float(f0_ceil),
                                                  int(channels_in_octave), int(target_fs), float(frame_period),
                                                  ctypes.byref(nbytes)))
        ws = self._workspace("dio", nbytes.value)
        F = self.L.wb_frame_count(smax, int(fs), float(frame_period))
        nb = self.L.wb_dio_band_count(float(f0_floor), float(f0_ceil), int(channels_in_octave))
        tpos, f0, vuv = self.empty(B, F), self.empty(B, F), self.empty(B, F)
        n_frames = self.empty(B, dtype=torch.int32)
        cand = self.empty(B, F, nb) if want_candidates else None
        raw = self.empty(B, nb, F) if want_candidates else None
        self._check(self.L.wb_dio(self.h, self._stream(), _p(x), S, _p(n_samples), B, smax, int(fs), float(f0_floor),
                                  float(f0_ceil), int(channels_in_octave), int(target_fs), float(frame_period),
                                  float(allowed_range), _p(ws), nbytes.value, F, _p(tpos), _p(f0), _p(vuv),
                                  _p(n_frames), _p(cand), _p(raw)))
        if want_candidates:
            return tpos, f0, vuv, n_frames, cand, raw
        return tpos, f0, vuv, n_frames

    def stonemask(self, x, n_samples, fs, tpos, f0, n_frames):
        """world/stonemask.py:8.  Returns refined f0 [B,F]."""
        B, S = x.shape
        F = tpos.shape[1]
        out = self.empty(B, F)
        self._check(self.L.wb_stonemask(self.h, self._stream(), _p(x), S, _p(n_samples), B, int(fs), _p(tpos), _p(f0),
                                        _p(n_frames), F, _p(out)))
        return out

    # ------------------------------------------------------------------ fused analysis
    def encode(self, x, n_samples, fs, f0_method="harvest", f0_floor=71.0, f0_ceil=800.0, frame_period=5.0,
               fft_size=None, is_requiem=False, dither=None, want_ps=False, max_samples=None, seed=0,
               channels_in_octave=2, target_fs=4000, allowed_range=0.1, streams=1, aperiodicity="full", out=None,
               zero_fill=False):
        """Device-resident World.encode (main.py:106-152) for a batch: one wb_encode call (or one per stream).
        Returns a dict of device tensors: temporal_positions, f0, vuv [B,F]; n_frames [B]; spectrogram [B,F,N/2+1];
        aperiodicity ([B,F,N/2+1] linear, or [B,F,bands+2] dB for requiem); 'ps spectrogram' [B,F,N] when want_ps.
        aperiodicity="coarse" (d4c only) returns 'coarse_ap' [B,F,bands] instead of the expanded matrix -- the
        compact transport form, expand_aperiodicity() rebuilds the matrix bit for bit.  "both" returns both;
        "none" stops after CheapTrick (World.get_spectrum: 'f0' is then the contour as CheapTrick leaves it).
        streams > 1 splits the batch by utterance over that many CUDA streams so that the short, latency-bound
        kernels of one part (decimation scans, contour tracking) overlap the compute-bound kernels of another;
        every part writes its rows of the same output tensors.  `out`: preallocated output tensors to fill;
        zero_fill: allocate the others zeroed (frames past n_frames[u] of a ragged batch are never written)."""
        if f0_method not in _abi.F0_METHODS:
            raise Exception("world_b200: unknown f0_method %r" % (f0_method,))
        self._check_input(x, n_samples)
        B, S = x.shape
        x_stride = x.stride(0) if B > 1 else S  # a size-1 batch axis may carry any stride (NumPy's x[None]: 0)
        smax = int(S if max_samples is None else max_samples)
        F = self.L.wb_frame_count(smax, int(fs), float(frame_period))
        n = int(fft_size) if fft_size else self.L.wb_cheaptrick_fft_size(int(fs))
        q = _abi.EncodeParams(int(fs), _abi.F0_METHODS[f0_method], float(f0_floor), float(f0_ceil),
                              int(channels_in_octave), int(target_fs), float(frame_period), float(allowed_range),
                              int(fft_size) if fft_size else 0, int(bool(is_requiem)), -0.15, 0.85, int(seed))
        if aperiodicity == "none":  # World.get_spectrum (main.py:52-80): tracker + CheapTrick only
            q.requiem = 2
            ap_bins, want_full, want_coarse = 0, False, False
        elif is_requiem:
            ap_bins, want_full, want_coarse = max(self.L.wb_d4c_band_count(int(fs), 1), 0) + 2, True, False
        else:
            if aperiodicity not in ("full", "coarse", "both"):
                raise ValueError("aperiodicity must be 'full', 'coarse' or 'both'")
            ap_bins = n // 2 + 1
            want_full, want_coarse = aperiodicity != "coarse", aperiodicity != "full"
        nb = max(self.L.wb_d4c_band_count(int(fs), 0), 1)
        o = out or {}

        def get(key, shape, dtype=torch.float64):
            t = o.get(key)
            if t is None:
                t = (torch.zeros if zero_fill else torch.empty)(*shape, dtype=dtype, device=self.device)
            assert tuple(t.shape) == tuple(shape) and t.dtype == dtype and t.is_contiguous() and t.device == self.device, key
            return t
        d = {"temporal_positions": get("temporal_positions", (B, F)), "vuv": get("vuv", (B, F)), "fs": fs,
             "f0": get("f0", (B, F)),
             "aperiodicity": get("aperiodicity", (B, F, ap_bins)) if want_full else None,
             "ps spectrogram": get("ps spectrogram", (B, F, n), torch.complex128) if want_ps else None,
             "spectrogram": get("spectrogram", (B, F, n // 2 + 1)), "is_requiem": bool(is_requiem),
             "n_frames": get("n_frames", (B,), torch.int32)}
        if want_coarse:
            d["coarse_ap"] = get("coarse_ap", (B, F, nb))

        def run(lo, hi):
            nbytes = ctypes.c_size_t()
            self._check(self.L.wb_encode_workspace_bytes(self.h, ctypes.byref(q), hi - lo, smax, ctypes.byref(nbytes)))
            ws = self._workspace("encode", nbytes.value)
            sl = lambda t: None if t is None else _p(t[lo:hi])
            self._check(self.L.wb_encode(
                self.h, self._stream(), ctypes.byref(q), _p(x[lo:hi]), x_stride, _p(n_samples[lo:hi]), hi - lo, smax,
                _p(ws), nbytes.value, F, sl(dither), sl(d["temporal_positions"]), sl(d["f0"]), sl(d["vuv"]),
                sl(d["n_frames"]), sl(d["spectrogram"]), sl(d["aperiodicity"]), sl(d.get("coarse_ap")),
                sl(d["ps spectrogram"])))

        if streams > 1 and B >= 2 * streams:
            main = torch.cuda.current_stream(self.device)
            while len(self._side) < streams:
                self._side.append(torch.cuda.Stream(device=self.device))
            per = (B + streams - 1) // streams
            used = []
            for k in range(streams):
                lo, hi = k * per, min(B, (k + 1) * per)
                if lo >= hi:
                    break
                st = self._side[k]
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    run(lo, hi)
                used.append(st)
            for st in used:
                main.wait_stream(st)
            for v in list(d.values()) + [x, n_samples] + ([dither] if dither is not None else []):
                if isinstance(v, torch.Tensor):
                    for st in used:
                        v.record_stream(st)
        else:
            run(0, B)
        return d

    def expand_aperiodicity(self, coarse_ap, fs, fft_size=None, out=None):
        """d4c.py:56-59 on the device: 'coarse_ap' [.., bands] -> aperiodicity [.., fft/2+1] (same bits as the
        D4C kernel's own expansion)."""
        n = int(fft_size) if fft_size else self.L.wb_cheaptrick_fft_size(int(fs))
        c = coarse_ap.contiguous()
        rows = c.numel() // c.shape[-1]
        ap = out if out is not None else self.empty(*c.shape[:-1], n // 2 + 1)
        self._check(self.L.wb_d4c_expand(self.h, self._stream(), _p(c), rows, int(fs), n, _p(ap)))
        return ap

    def _check_input(self, x, n_samples):
        """The kernels take raw pointers: refuse anything but contiguous float64 rows / int32 lengths on this device."""
        if not (isinstance(x, torch.Tensor) and x.dtype == torch.float64 and x.dim() == 2 and x.stride(1) == 1
                and x.device == self.device):
            raise TypeError("world_b200: x must be a float64 [B, S] CUDA tensor with unit sample stride on %s" % (self.device,))
        if not (isinstance(n_samples, torch.Tensor) and n_samples.dtype == torch.int32 and n_samples.is_contiguous()
                and n_samples.device == self.device and n_samples.numel() == x.shape[0]):
            raise TypeError("world_b200: n_samples must be a contiguous int32 [B] CUDA tensor")

    # ------------------------------------------------------------------ synthesis
    def synthesis_length(self, t0, t_end, fs):
        return self.L.wb_synthesis_length(float(t0), float(t_end), int(fs))

    def _timebase(self, tpos, f0, vuv, n_frames, fs, y_stride, rows):
        B, F = tpos.shape
        nbytes = ctypes.c_size_t()
        self._check(self.L.wb_synthesis_workspace_bytes(self.h, B, int(y_stride), int(rows), ctypes.byref(nbytes)))
        ws = self._workspace("synthesis", nbytes.value)
        out_len = self.empty(B, dtype=torch.int32)
        n_pulses = self.empty(B, dtype=torch.int32)
        noise_total = self.empty(B, dtype=torch.int32)
        self._check(self.L.wb_synthesis_timebase(self.h, self._stream(), _p(tpos), _p(f0), _p(vuv), _p(n_frames), B, F,
                                                 int(fs), int(y_stride), _p(ws), nbytes.value, int(rows), _p(out_len),
                                                 _p(n_pulses), _p(noise_total)))
        return ws, nbytes.value, out_len, n_pulses, noise_total

    def synthesis(self, tpos, f0, vuv, spectrogram, aperiodicity, n_frames, fs, y_stride, noise="device", seed=0,
                  normalize=True):
        """world/synthesis.py:21 (+ main.py:209-212 when normalize).  spectrogram / aperiodicity [B, F, N/2+1].
        noise: "device" (counter-based generator), "legacy" (np.random.randn replayed in the reference's order --
        needs one host round trip for the draw sizes) or a [B, stride] tensor of normals.
        Returns (y [B, y_stride], out_len [B])."""
        B, F = tpos.shape
        n = (spectrogram.shape[2] - 1) * 2
        ws, wsb, out_len, n_pulses, noise_total = self._timebase(tpos, f0, vuv, n_frames, fs, y_stride, 0)
        nz, stride = None, 0
        if isinstance(noise, str) and noise == "legacy":
            import numpy as np
            tot = noise_total.cpu().numpy()
            stride = int(tot.max()) if B else 0
            host = np.zeros((B, max(stride, 1)))
            for i in range(B):
                host[i, :tot[i]] = np.random.randn(int(tot[i]))
            nz = self.f64(host)
            stride = host.shape[1]
        elif not isinstance(noise, str):
            nz = noise
            stride = noise.shape[1]
        y = self.empty(B, int(y_stride))
        self._check(self.L.wb_synthesis(self.h, self._stream(), _p(tpos), _p(f0), _p(vuv), _p(spectrogram),
                                        _p(aperiodicity), _p(n_frames), B, F, int(fs), n, _p(ws), wsb, _p(nz), stride,
                                        int(seed), _p(y), int(y_stride), int(bool(normalize))))
        return y, out_len

    def synthesis_requiem(self, tpos, f0, vuv, spectrogram, band_ap, n_frames, fs, y_stride, pulse_seed, noise_seed,
                          cursor=None, normalize=True):
        """world/synthesisRequiem.py:12.  band_ap [B, F, rows] dB; seeds as get_seeds_signals() returns them.
        Returns (y, out_len, cursor_out [B, rows])."""
        B, F = tpos.shape
        rows = band_ap.shape[2]
        n = (spectrogram.shape[2] - 1) * 2
        ws, wsb, out_len, n_pulses, noise_total = self._timebase(tpos, f0, vuv, n_frames, fs, y_stride, rows)
        cur_in = self.f64(torch.zeros(rows, dtype=torch.float64) if cursor is None else cursor)
        cur_out = self.empty(B, rows)
        y = self.empty(B, int(y_stride))
        self._check(self.L.wb_synthesis_requiem(self.h, self._stream(), _p(tpos), _p(f0), _p(vuv), _p(spectrogram),
                                                _p(band_ap), _p(n_frames), B, F, int(fs), n, rows, _p(pulse_seed),
                                                pulse_seed.shape[0], _p(noise_seed), noise_seed.shape[0], _p(cur_in),
                                                _p(cur_out), _p(ws), wsb, _p(y), int(y_stride), int(bool(normalize))))
        return y, out_len, cur_out

    def decode(self, tpos, f0, vuv, spectrogram, aperiodicity, n_frames, fs, y_stride, is_requiem=False, seeds=None,
               cursor=None, noise=None, seed=0, normalize=True):
        """Device-resident World.decode (main.py:198-214) for a batch through the fused wb_decode: time base, then
        synthesis.py (noise: [B, stride] normals in draw order, or None for the counter-based generator) or
        synthesisRequiem.py (seeds = (pulse, noise) of get_seeds_signals as device tensors).  Returns
        (y [B, y_stride], out_len [B], cursor_out [B, rows] | None)."""
        B, F = tpos.shape
        n = (spectrogram.shape[2] - 1) * 2
        rows = aperiodicity.shape[2] if is_requiem else 0
        nbytes = ctypes.c_size_t()
        self._check(self.L.wb_decode_workspace_bytes(self.h, B, int(y_stride), int(rows), ctypes.byref(nbytes)))
        ws = self._workspace("decode", nbytes.value)
        y = self.empty(B, int(y_stride))
        out_len = self.empty(B, dtype=torch.int32)
        ps = ns = cur_in = cur_out = None
        if is_requiem:
            ps, ns = seeds
            cur_in = self.f64(torch.zeros(rows, dtype=torch.float64) if cursor is None else cursor)
            cur_out = self.empty(B, rows)
        self._check(self.L.wb_decode(
            self.h, self._stream(), int(fs), n, _p(tpos), _p(f0), _p(vuv), _p(spectrogram), _p(aperiodicity), _p(n_frames),
            B, F, int(rows), _p(ps), ps.shape[0] if ps is not None else 0, _p(ns), ns.shape[0] if ns is not None else 0,
            _p(cur_in), _p(cur_out), _p(noise), noise.shape[1] if noise is not None else 0, int(seed), _p(ws), nbytes.value,
            _p(y), int(y_stride), int(bool(normalize)), _p(out_len)))
        return y, out_len, cur_out

    @staticmethod
    def launches_per_encode(f0_method, is_requiem):
        """Kernels of ours launched by one encode(): harvest 14 (5 decimation, block spectra + overlap-save channels,
        detect, 4 refine, prune, contour; the direct-FIR channel kernel only runs when some filter is shorter than
        the overlap-save threshold), dio 8; cheaptrick 1, d4c 1."""
        return {"harvest": 14, "dio": 8}[f0_method] + 2

    @staticmethod
    def launches_per_decode(is_requiem):
        """time base + prefix + (pulses | excitation, pulses, frames) + peak rescale."""
        return 6 if is_requiem else 4

    def profile_stages(self, x, n_samples, fs, f0_method="harvest", is_requiem=False, iters=3, f0_floor=71.0,
                       f0_ceil=800.0, frame_period=5.0, with_d4c=True):
        """Median CUDA-event time (ms) of every kernel of encode(), each launched alone on the current stream
        over the same inputs (the workspace keeps the earlier stages' results)."""
        B, S = x.shape
        names = ["hv_decimate", "hv_channels", "hv_detect", "hv_refine", "hv_prune", "hv_contour"]
        nbytes = ctypes.c_size_t()
        self._check(self.L.wb_harvest_workspace_bytes(self.h, B, S, int(fs), f0_floor, f0_ceil, ctypes.byref(nbytes)))
        ws = self._workspace("harvest", nbytes.value)
        F = self.L.wb_frame_count(S, int(fs), float(frame_period))
        tpos, f0, vuv = self.empty(B, F), self.empty(B, F), self.empty(B, F)
        nf = self.empty(B, dtype=torch.int32)

        def hv(a, b):
            self._check(self.L.wb_harvest_stages(self.h, self._stream(), _p(x), S, _p(n_samples), B, S, int(fs),
                                                 float(f0_floor), float(f0_ceil), float(frame_period), _p(ws),
                                                 nbytes.value, F, _p(tpos), _p(f0), _p(vuv), _p(nf), a, b))

        out = {}

        def timed(name, fn):
            ts = []
            for _ in range(iters):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            out[name] = sorted(ts)[len(ts) // 2]

        hv(0, 5)
        torch.cuda.synchronize()
        for i, name in enumerate(names):
            timed(name, lambda i=i: hv(i, i))
        timed("cheaptrick", lambda: self.cheaptrick(x, n_samples, fs, tpos, f0, vuv, nf))
        f0_used, _, _ = self.cheaptrick(x, n_samples, fs, tpos, f0, vuv, nf)
        if not with_d4c:
            return out
        if is_requiem:
            timed("d4c_requiem", lambda: self.d4c_requiem(x, n_samples, fs, tpos, f0_used, vuv, nf))
        else:
            timed("d4c", lambda: self.d4c(x, n_samples, fs, tpos, f0_used, vuv, nf))
        return out


    def profile_decode(self, feats, fs, y_stride, is_requiem=False, seeds=None, iters=3):
        """Median CUDA-event time (ms) of the two stages of decode(): the time base (pulse trains) and the
        synthesiser proper (pulse / frame responses, overlap-add, peak rescale)."""
        tp, f0, vuv, nf = feats["temporal_positions"], feats["f0"], feats["vuv"], feats["n_frames"]
        sp, ap = feats["spectrogram"], feats["aperiodicity"]
        rows = ap.shape[2] if is_requiem else 0
        out = {}

        def timed(name, fn):
            ts = []
            for _ in range(iters):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            out[name] = sorted(ts)[len(ts) // 2]

        timed("sy_timebase", lambda: self._timebase(tp, f0, vuv, nf, fs, y_stride, rows))
        if is_requiem:
            timed("rq_synthesis", lambda: self.synthesis_requiem(tp, f0, vuv, sp, ap, nf, fs, y_stride, seeds[0], seeds[1]))
            out["rq_synthesis"] -= out["sy_timebase"]
        else:
            timed("sy_synthesis", lambda: self.synthesis(tp, f0, vuv, sp, ap, nf, fs, y_stride, noise="device", seed=1))
            out["sy_synthesis"] -= out["sy_timebase"]
        return out


_default = {}


def default_engine(device=None):
    """One engine per device per process."""
    if not torch.cuda.is_available():
        raise RuntimeError("world_b200: no CUDA device visible; this engine has no CPU path")
    idx = torch.cuda.current_device() if device is None else (device.index if isinstance(device, torch.device) else int(device))
    if idx not in _default:
        _default[idx] = Engine(torch.device("cuda", idx))
    return _default[idx]
