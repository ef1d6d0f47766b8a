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


def expand_coarse_ap(coarse_ap, fs, fft_size=None):
    """d4c.py:56-59 on the host: aperiodicity [.., fft/2+1] from 'coarse_ap' [.., bands] with the reference's own
    expressions (scipy interp1d's slope form, 10 ** (v / 20)); frames whose band value has the sign bit clear
    (unvoiced, or rejected by the love-train gate: d4c.py:49-51) get 1 - 1e-12."""
    c = coarse_ap.numpy() if isinstance(coarse_ap, torch.Tensor) else np.asarray(coarse_ap)
    n = int(2 ** np.ceil(np.log2(3 * fs / 71 + 1))) if fft_size is None else int(fft_size)
    interval = 2000 if fs < 16000 else 3000
    nb = c.shape[-1]
    coarse_axis = np.r_[np.arange(nb + 1) * interval, fs / 2]
    freq = np.arange(n / 2 + 1) * fs / n
    hi = np.clip(np.searchsorted(coarse_axis, freq), 1, len(coarse_axis) - 1)
    lo = hi - 1
    y = np.concatenate([np.full(c.shape[:-1] + (1,), -60.0), c, np.full(c.shape[:-1] + (1,), -0.000000000001)], axis=-1)
    slope = (y[..., hi] - y[..., lo]) / (coarse_axis[hi] - coarse_axis[lo])
    ap = 10 ** ((slope * (freq - coarse_axis[lo]) + y[..., lo]) / 20)
    ap[~np.signbit(c[..., 0])] = 1 - 0.000000000001
    return ap


class EncodedBatch(dict):
    """What encode_batch() returns: the reference's keys as host tensors [B, F(, bins)].  With the compact
    transport (aperiodicity='coarse') the 'aperiodicity' matrix is rebuilt from 'coarse_ap' on first access."""
    _expand_args = None

    def __missing__(self, key):
        if key == 'aperiodicity' and dict.__contains__(self, 'coarse_ap'):
            ap = torch.from_numpy(expand_coarse_ap(dict.__getitem__(self, 'coarse_ap'), *self._expand_args))
            self[key] = ap
            return ap
        raise KeyError(key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or (key == 'aperiodicity' and dict.__contains__(self, 'coarse_ap'))

    def get(self, key, default=None):
        return self[key] if key in self else default


class World(object):
    def __init__(self, device=None):
        self._device = device
        self._pinned = {}

    def _host_buffer_shape(self, key, shape, dtype):
        buf = self._pinned.get(key)
        if buf is None or tuple(buf.shape) != tuple(shape) or buf.dtype != dtype:
            buf = torch.zeros(shape, dtype=dtype, pin_memory=True)
            self._pinned[key] = buf
        return buf

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
        if f0_method not in ('harvest', 'dio'):
            if f0_method == 'swipe':
                raise NotImplementedError("world_b200: f0_method='swipe' is outside the hot path (SURVEY.md section 2, row 11)")
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
                     is_requiem, dither=E.f64(dither[None]), want_ps=True, channels_in_octave=channels_in_octave,
                     target_fs=target_fs, allowed_range=allowed_range)
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
                     fft_size=None, is_requiem=False, want_ps=False, channels_in_octave=2, target_fs=4000,
                     allowed_range=0.1, device_resident=False, pipeline=16, aperiodicity='coarse', spectrogram_dtype=None):
        """Batched encode with HOST buffers: xs [B, S] float64 or int16 PCM (NumPy or torch, pinned for speed),
        optional n_samples [B].  Results are host tensors [B, F(, bins)] in pinned memory; the per-call copy volume
        is reported under '_h2d_bytes' / '_d2h_bytes'.  The returned host tensors are staging buffers owned by
        this World object and are overwritten by the next encode_batch() call.
        aperiodicity='coarse' (default; D4C only) brings back the band values 'coarse_ap' [B, F, bands] -- at
        16 kHz one float64 per frame instead of 513 -- and the returned dict rebuilds dat['aperiodicity'] from them
        on first access (d4c.py:56-59, the matrix is a deterministic function of the band values); 'full' copies
        the expanded matrix as the reference returns it.  spectrogram_dtype=torch.float32 (opt-in, lossy: 6e-8
        relative rounding) halves the bytes of the one large result that is left; decode_batch() widens it again.
        device_resident=True returns the CUDA tensors instead
        (no D2H): scale_pitch / scale_duration work on them in place and decode_batch() consumes them directly."""
        E = self.engine
        if isinstance(xs, torch.Tensor):
            xs_t = xs
            if xs_t.dtype != torch.int16:
                xs_t = xs_t.to(torch.float64)
        else:  # int16 PCM stays int16 on the wire (4x fewer H2D bytes) and is scaled on the device
            xs_t = torch.from_numpy(np.ascontiguousarray(xs, dtype=np.int16 if np.asarray(xs).dtype == np.int16 else np.float64))
        if xs_t.dim() != 2:
            raise ValueError("encode_batch: xs must be [B, S]")
        if xs_t.stride(1) != 1:
            xs_t = xs_t.contiguous()
        is_pcm = xs_t.dtype == torch.int16  # x = x_int16 / (2**15 - 1), example/prosody.py:13
        B, S = xs_t.shape
        ns_host = np.full(B, S, dtype=np.int32) if n_samples is None else np.asarray(
            n_samples.cpu() if isinstance(n_samples, torch.Tensor) else n_samples).astype(np.int32)
        if ns_host.shape != (B,) or (B and (ns_host.min() < 0 or ns_host.max() > S)):
            raise ValueError("encode_batch: n_samples must be [B] with 0 <= n_samples <= xs.shape[1]")
        smax = int(ns_host.max()) if B else 0
        ragged = bool(B) and int(ns_host.min()) != smax
        mode = 'full' if is_requiem else aperiodicity
        kw = dict(f0_method=f0_method, f0_floor=float(f0_floor), f0_ceil=float(f0_ceil), frame_period=float(frame_period),
                  fft_size=fft_size, is_requiem=is_requiem, want_ps=want_ps, channels_in_octave=channels_in_octave,
                  target_fs=target_fs, allowed_range=allowed_range, max_samples=smax, zero_fill=ragged)
        h2d = xs_t.numel() * xs_t.element_size() + ns_host.nbytes
        if device_resident:  # encode -> edit -> decode_batch without leaving HBM
            X = xs_t.to(E.device, non_blocking=True)
            ns_dev = E.i32(ns_host)
            if is_pcm:
                X = E.pcm16_to_f64(X, ns_dev)
            d = E.encode(X, ns_dev, int(fs), streams=2, aperiodicity='full', **kw)
            d['_h2d_bytes'] = h2d
            d['_d2h_bytes'] = 0
            return d
        # host buffers in, host buffers out: the batch goes through in `pipeline` parts, each on its own CUDA
        # stream, so the H2D copy / kernels / D2H copy of different parts overlap
        main = torch.cuda.current_stream(E.device)
        parts = max(1, min(int(pipeline), B))
        per = (B + parts - 1) // parts
        out = EncodedBatch(fs=fs, is_requiem=is_requiem)
        out._expand_args = (int(fs), fft_size)
        d2h = 0
        while len(E._side) < parts:
            E._side.append(torch.cuda.Stream(device=E.device))
        for k in range(parts):
            lo, hi = k * per, min(B, (k + 1) * per)
            if lo >= hi:
                break
            st = E._side[k]
            st.wait_stream(main)
            with torch.cuda.stream(st):
                X = xs_t[lo:hi].to(E.device, non_blocking=True)
                ns_dev = E.i32(ns_host[lo:hi])
                if is_pcm:
                    X = E.pcm16_to_f64(X, ns_dev)
                d = E.encode(X, ns_dev, int(fs), streams=1, aperiodicity=mode, **kw)
                for key in ('temporal_positions', 'vuv', 'f0', 'aperiodicity', 'coarse_ap', 'spectrogram',
                            'ps spectrogram', 'n_frames'):
                    v = d.get(key)
                    if v is None:
                        continue
                    if key == 'spectrogram' and spectrogram_dtype == torch.float32:
                        v = E.to_f32(v)
                    hbuf = self._host_buffer_shape(key, (B,) + tuple(v.shape[1:]), v.dtype)
                    hbuf[lo:hi].copy_(v, non_blocking=True)
                    out[key] = hbuf
                    d2h += v.numel() * v.element_size()
        for k in range(parts):
            main.wait_stream(E._side[k])
        torch.cuda.synchronize()
        out['_h2d_bytes'] = h2d
        out['_d2h_bytes'] = d2h
        return out

    # ------------------------------------------------------------------ partial pipelines (main.py:27-104)
    def _track(self, fs, x, f0_method, f0_floor, f0_ceil, channels_in_octave, target_fs, frame_period):
        from .dio import dio
        from .harvest import harvest
        from .stonemask import stonemask
        if f0_method == 'dio':
            source = dio(x, fs, f0_floor, f0_ceil, channels_in_octave, target_fs, frame_period)
            source['f0'] = stonemask(x, fs, source['temporal_positions'], source['f0'])
        elif f0_method == 'harvest':
            source = harvest(x, fs, f0_floor, f0_ceil, frame_period)
        else:
            raise Exception
        return source

    def get_f0(self, fs, x, f0_method='harvest', f0_floor=71, f0_ceil=800, channels_in_octave=2, target_fs=4000,
               frame_period=5):
        s = self._track(fs, x, f0_method, f0_floor, f0_ceil, channels_in_octave, target_fs, frame_period)
        return s['temporal_positions'], s['f0'], s['vuv']

    def get_spectrum(self, fs, x, f0_method='harvest', f0_floor=71, f0_ceil=800, channels_in_octave=2, target_fs=4000,
                     frame_period=5, fft_size=None):
        from .cheaptrick import cheaptrick
        s = self._track(fs, x, f0_method, f0_floor, f0_ceil, channels_in_octave, target_fs, frame_period)
        flt = cheaptrick(x, fs, s, fft_size=fft_size)
        return {'f0': s['f0'], 'temporal_positions': s['temporal_positions'], 'fs': fs,
                'ps spectrogram': flt['ps spectrogram'], 'spectrogram': flt['spectrogram']}

    def encode_w_gvn_f0(self, fs, x, source, fft_size=None, is_requiem=False):
        from .cheaptrick import cheaptrick
        from .d4c import d4c
        from .d4cRequiem import d4cRequiem
        assert np.all(source['f0'] >= 3 * fs / fft_size)
        flt = cheaptrick(x, fs, source, fft_size=fft_size)
        if is_requiem:
            source = d4cRequiem(x, fs, source, fft_size=fft_size)
        else:
            source = d4c(x, fs, source, fft_size_for_spectrum=fft_size)
        return {'temporal_positions': source['temporal_positions'], 'vuv': source['vuv'], 'f0': source['f0'],
                'fs': fs, 'spectrogram': flt['spectrogram'], 'aperiodicity': source['aperiodicity'],
                'coarse_ap': source['coarse_ap'],  # KeyError for requiem, as in the reference (main.py:102)
                'is_requiem': is_requiem}

    # ------------------------------------------------------------------ prosody edits on the dict (main.py:154-196)
    def scale_pitch(self, dat, factor):
        dat['f0'] *= factor
        return dat

    def set_pitch(self, dat, time, value):
        raise NotImplementedError  # as in the reference (main.py:165)

    def scale_duration(self, dat, factor):
        dat['temporal_positions'] *= factor
        return dat

    def modify_duration(self, dat, from_time, to_time):
        """main.py:178-187.  NumPy dicts take the reference's host path; a device-resident batch (encode_batch(...,
        device_resident=True)) is edited in HBM, every utterance against its own last frame time."""
        tp = dat['temporal_positions']
        if isinstance(tp, torch.Tensor):
            from . import features
            E = self.engine
            nf = dat['n_frames'].cpu().numpy()
            ends = tp.gather(1, (dat['n_frames'].long() - 1).clamp(min=0)[:, None])[:, 0].cpu().numpy()
            assert np.all(np.diff(from_time)) > 0 and np.all(np.diff(to_time)) > 0 and from_time[0] > 0
            for u in range(tp.shape[0]):
                end = float(ends[u])
                assert from_time[-1] < end
                ys = np.array(to_time, dtype=np.float64)
                if ys[-1] == -1:
                    ys[-1] = end
                row = tp[u, :int(nf[u])]
                features.interp_knots(E, row, np.r_[0, from_time, end], ys, out=row)
            return
        end = tp[-1]
        assert np.all(np.diff(from_time)) > 0
        assert np.all(np.diff(to_time)) > 0
        assert from_time[0] > 0
        assert from_time[-1] < end
        from_time = np.r_[0, from_time, end]
        if to_time[-1] == -1:
            to_time[-1] = end
        dat['temporal_positions'] = self._np_or_dev(features_fn='interp_knots', x=tp, args=(from_time, to_time))

    def _np_or_dev(self, features_fn, x, args):
        """Run one of the world_b200.features launchers on a NumPy array (upload, kernel, download)."""
        from . import features
        E = self.engine
        out = getattr(features, features_fn)(E, E.f64(x), *args)
        torch.cuda.synchronize()
        return out.cpu().numpy()

    def warp_spectrum(self, dat, factor):
        """main.py:189-194: every frame's spectrum resampled at (k/D)**factor, in place."""
        from . import features
        sp = dat['spectrogram']
        if isinstance(sp, torch.Tensor):  # device-resident [B, F, bins]
            features.warp_rows(self.engine, sp, factor, out=sp)
            return dat
        rows = np.ascontiguousarray(sp.T)  # reference layout is [bins, F]
        sp[:] = self._np_or_dev('warp_rows', rows, (factor,)).T
        return dat

    # ------------------------------------------------------------------ spectral feature heads (main.py:258-358)
    def hz2mel(self, hz):
        from . import features
        return features.hz2mel(hz)

    def mel2hz(self, mel):
        from . import features
        return features.mel2hz(mel)

    def get_filterbanks(self, nfilt=20, nfft=512, samplerate=16000, lowfreq=0, highfreq=None):
        from . import features
        return features.mel_filterbank(nfilt, nfft, samplerate, lowfreq, highfreq)

    def _head(self, fn, arr, *args):
        from . import features
        if isinstance(arr, torch.Tensor):  # resident [.., bins] tensor in, resident tensor out (no copies)
            return getattr(features, fn)(self.engine, arr.contiguous(), *args)
        return self._np_or_dev(fn, np.ascontiguousarray(arr, dtype=np.float64), args)

    def encode_lfbank(self, spec, prefac=0.97, fs=16000, nfilt=32, lowfreq=0, highfreq=None):
        """Log mel-filterbank energies of an [N, D] magnitude spectrum (main.py:305-322)."""
        return self._head('lfbank', spec, prefac, fs, nfilt, lowfreq, highfreq)

    def encode_mcep(self, spec, n0=12, fs=16000, lowhz=0, highhz=8000):
        """First n0 cepstral coefficients of the mel-warped log spectrum (main.py:324-342)."""
        return self._head('mcep', spec, n0, fs, lowhz, highhz)

    def decode_mcep(self, cepstrum, fft_size):
        """Magnitude spectrum [N, fft_size/2+1] from mel-cepstra (main.py:344-358)."""
        return self._head('mcep_decode', cepstrum, fft_size)

    # ------------------------------------------------------------------ main.py:198-214
    def decode(self, dat):
        """Combine F0, spectrogram and aperiodicity into a waveform; returns `dat` with dat['out'] added."""
        from .get_seeds_signals import get_seeds_signals
        from .synthesis import synthesis
        from .synthesisRequiem import synthesisRequiem
        if dat['is_requiem']:
            seeds_signals = get_seeds_signals(dat['fs'])
            y = synthesisRequiem(dat, dat, seeds_signals)
        else:
            y = synthesis(dat, dat)
        m = np.max(np.abs(y))
        if m > 1.0:
            logging.info('rescaling waveform')
            y /= m
        dat['out'] = y
        return dat

    def decode_batch(self, dat, noise="device", seed=0, pcm16=False):
        """Batched decode from the dict encode_batch() returns (host or device tensors, [B, F(, bins)] layout).
        Noise comes from the device generator (noise="device") -- the legacy np.random replay is a
        single-utterance feature.  Returns dict(out [B, S] pinned host tensor, out_len [B]); pcm16=True returns
        int16 samples encoded on the device."""
        E = self.engine
        fs = int(dat['fs'])
        dev = lambda v: v.to(E.device, non_blocking=True) if isinstance(v, torch.Tensor) else E.f64(v)
        tp, f0, vuv = dev(dat['temporal_positions']), dev(dat['f0']), dev(dat['vuv'])
        spec = dev(dat['spectrogram'])
        h2d_spec = spec
        if spec.dtype == torch.float32:  # the compact transport of encode_batch(spectrogram_dtype=torch.float32)
            spec = E.to_f64(spec)
        h2d_ap = None
        if not dat['is_requiem'] and dict.__contains__(dat, 'coarse_ap') and not dict.__contains__(dat, 'aperiodicity'):
            # compact transport: upload the band values and expand on the device (same bits as the D4C kernel)
            h2d_ap = dev(dat['coarse_ap'])
            ap = E.expand_aperiodicity(h2d_ap, fs, (spec.shape[2] - 1) * 2)
        else:
            ap = h2d_ap = dev(dat['aperiodicity'])
        nf = dat['n_frames'].to(E.device) if isinstance(dat['n_frames'], torch.Tensor) else E.i32(dat['n_frames'])
        tp_h = dat['temporal_positions'].cpu() if isinstance(dat['temporal_positions'], torch.Tensor) else torch.as_tensor(dat['temporal_positions'])
        nf_h = dat['n_frames'].cpu() if isinstance(dat['n_frames'], torch.Tensor) else torch.as_tensor(dat['n_frames'])
        last = tp_h.gather(1, (nf_h.long() - 1).clamp(min=0)[:, None])[:, 0]
        ylen = max(E.synthesis_length(float(tp_h[i, 0]), float(last[i]), fs) for i in range(tp_h.shape[0]))
        if dat['is_requiem']:
            from .get_seeds_signals import get_seeds_signals
            sd = get_seeds_signals(fs)
            y, out_len, _ = E.decode(tp, f0, vuv, spec, ap, nf, fs, ylen, is_requiem=True,
                                     seeds=(E.f64(sd['pulse']), E.f64(sd['noise'])))
        else:
            if not (isinstance(noise, str) and noise == "device"):
                raise ValueError('decode_batch draws its noise on the device (noise="device"); the legacy '
                                 'np.random replay is a single-utterance feature (World.decode)')
            y, out_len, _ = E.decode(tp, f0, vuv, spec, ap, nf, fs, ylen, seed=seed)
        if pcm16:  # (out * 2**15).astype(np.int16) on the device (example/prosody.py:57): 4x fewer D2H bytes
            y = E.f64_to_pcm16(y, out_len)
        hy = self._host_buffer('out', y)
        hy.copy_(y, non_blocking=True)
        hl = out_len.cpu()
        torch.cuda.synchronize()
        return {'out': hy, 'out_len': hl, '_h2d_bytes': sum(int(v.numel() * v.element_size()) for v in (tp, f0, vuv, h2d_spec, h2d_ap)),
                '_d2h_bytes': int(y.numel() * y.element_size())}
