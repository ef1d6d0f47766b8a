"""Drop-in for the hot path of world/main.py: class World with encode() / decode().

Same method names, keyword arguments, defaults, dict keys, shapes and dtypes as the reference
(main.py:106-152, 198-214).  The single-utterance methods take and return NumPy float64 arrays
(they synchronise to hand the results back); encode_batch()/decode_batch() move a whole batch with
one H2D and one D2H copy per array.  Everything numeric runs in the CUDA library; there is no CPU
fallback (a missing library or GPU raises).
"""
import logging

import numpy as np
import torch

from . import engine as _engine

EPS = 2.220446049250313e-16


def _to_ref_layout(t):
    """[F, bins] device tensor -> NumPy [bins, F], C-contiguous like the reference's arrays."""
    return np.ascontiguousarray(t.cpu().numpy().T)


class World(object):
    def __init__(self, device=None):
        self._device = device
        self._pinned = {}

    def _host_buffer(self, key, like):
        """Pinned host staging buffer, reused across calls (allocation of pinned memory is slow)."""
        buf = self._pinned.get(key)
        if buf is None or buf.shape != like.shape or buf.dtype != like.dtype:
            buf = torch.empty(like.shape, dtype=like.dtype, pin_memory=True)
            self._pinned[key] = buf
        return buf

    @property
    def engine(self):
        return _engine.default_engine(self._device)

    # ------------------------------------------------------------------ main.py:106-152
    def encode(self, fs, x, f0_method='harvest', f0_floor=71, f0_ceil=800, channels_in_octave=2, target_fs=4000,
               frame_period=5, allowed_range=0.1, fft_size=None, is_requiem=False):
        E = self.engine
        if f0_method not in ('harvest',):
            if f0_method in ('dio', 'swipe'):
                raise NotImplementedError("world_b200: f0_method=%r is not built yet" % f0_method)
            raise Exception  # main.py:136-137
        x = np.ascontiguousarray(x, dtype=np.float64)
        n = E.L.wb_cheaptrick_fft_size(int(fs)) if fft_size is None else int(fft_size)
        floor = 3.0 * fs / fft_size if fft_size is not None else f0_floor
        F = E.L.wb_frame_count(len(x), int(fs), float(frame_period))
        # CheapTrick's eps-dither consumes np.random exactly as the reference does (cheaptrick.py:117)
        dither = np.abs(np.random.rand(F, n // 2 + 1)) * EPS
        X = E.f64(x[None])
        ns = E.i32([len(x)])
        d = E.encode(X, ns, int(fs), f0_method, float(floor), float(f0_ceil), float(frame_period), fft_size,
                     is_requiem, dither=E.f64(dither[None]), want_ps=True)
        torch.cuda.synchronize()
        return {'temporal_positions': d['temporal_positions'][0].cpu().numpy(),
                'vuv': d['vuv'][0].cpu().numpy(),
                'fs': fs,
                'f0': d['f0'][0].cpu().numpy(),
                'aperiodicity': _to_ref_layout(d['aperiodicity'][0]),
                'ps spectrogram': _to_ref_layout(d['ps spectrogram'][0]),
                'spectrogram': _to_ref_layout(d['spectrogram'][0]),
                'is_requiem': is_requiem}

    def encode_batch(self, fs, xs, n_samples=None, f0_method='harvest', f0_floor=71, f0_ceil=800, frame_period=5,
                     fft_size=None, is_requiem=False, want_ps=False):
        """Batched encode with HOST buffers: xs [B, S] float64 (NumPy or pinned torch tensor), optional
        n_samples [B].  Results are host tensors [B, F(, bins)] in pinned memory; the per-call copy volume is
        reported under '_h2d_bytes' / '_d2h_bytes'.  The returned host tensors are staging buffers owned by
        this World object and are overwritten by the next encode_batch() call."""
        E = self.engine
        xs_t = xs if isinstance(xs, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(xs, dtype=np.float64))
        B, S = xs_t.shape
        X = xs_t.to(E.device, non_blocking=True)
        ns_host = np.full(B, S, dtype=np.int32) if n_samples is None else np.asarray(n_samples, dtype=np.int32)
        ns = E.i32(ns_host)
        floor = 3.0 * fs / fft_size if fft_size is not None else f0_floor
        d = E.encode(X, ns, int(fs), f0_method, float(floor), float(f0_ceil), float(frame_period), fft_size,
                     is_requiem, want_ps=want_ps, max_samples=int(ns_host.max()))
        out = {'fs': fs, 'is_requiem': is_requiem}
        d2h = 0
        for k in ('temporal_positions', 'vuv', 'f0', 'aperiodicity', 'spectrogram', 'ps spectrogram', 'n_frames'):
            v = d[k]
            if v is None:
                continue
            hbuf = self._host_buffer(k, v)
            hbuf.copy_(v, non_blocking=True)
            out[k] = hbuf
            d2h += v.numel() * v.element_size()
        torch.cuda.synchronize()
        out['_h2d_bytes'] = xs_t.numel() * xs_t.element_size() + ns_host.nbytes
        out['_d2h_bytes'] = d2h
        return out

    # ------------------------------------------------------------------ main.py:198-214
    def decode(self, dat):
        raise NotImplementedError("world_b200: decode() is not built yet")
