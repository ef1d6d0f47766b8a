"""Batched, device-resident API: thin Python over the C-ABI (include/world_b200.h).

All tensors are torch CUDA tensors used as HBM containers; every method enqueues
work on the current CUDA stream and returns device tensors without synchronising.
Layouts: waveforms [B, S]; per-frame vectors [B, F]; per-frame matrices [B, F, bins].
"""
import ctypes

import torch

from . import _lib


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


_default = {}


def default_engine(device=None):
    """One engine per device per process."""
    if not torch.cuda.is_available():
        raise RuntimeError("world_b200: no CUDA device visible; this engine has no CPU path")
    idx = torch.cuda.current_device() if device is None else (device.index if isinstance(device, torch.device) else int(device))
    if idx not in _default:
        _default[idx] = Engine(torch.device("cuda", idx))
    return _default[idx]
