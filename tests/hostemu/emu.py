"""TEST-ONLY: drive the kernel bodies compiled for the host (-DWB_HOST_EMU, one
emulated thread per block) through the same C-ABI, with NumPy arrays standing in
for device memory.  Lets the CPU test tier check kernel LOGIC against the oracle
without a GPU.  The product package never loads this library."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
PKG = os.path.join(ROOT, "python-world_b200")
sys.path.insert(0, PKG)
from world_b200 import _abi  # noqa: E402

SO = os.path.join(HERE, "libworld_b200_hostemu.so")


def build(force=False):
    srcs = sorted(os.path.join(PKG, "csrc", f) for f in os.listdir(os.path.join(PKG, "csrc")) if f.endswith(".cu"))
    deps = srcs + [os.path.join(PKG, "csrc", f) for f in os.listdir(os.path.join(PKG, "csrc")) if f.endswith(".h")]
    deps.append(os.path.join(ROOT, "include", "world_b200.h"))
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return SO
    cmd = ["g++", "-x", "c++", "-std=c++17", "-DWB_HOST_EMU", "-O2", "-fPIC", "-shared", "-o", SO] + srcs
    subprocess.check_call(cmd)
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _abi.declare(C.CDLL(build()))
        assert _lib.wb_is_cuda_build() == 0
    return _lib


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Emu:
    def __init__(self):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.wb_create(C.byref(h), 0)
        assert rc == 0
        self.h = h

    def close(self):
        if self.h:
            self.L.wb_destroy(self.h)
            self.h = None

    def check(self, rc):
        if rc != 0:
            raise RuntimeError("rc=%d: %s" % (rc, self.L.wb_last_error(self.h).decode()))

    # the small "ops" surface world_b200/features.py launches through (NumPy arrays stand in for device memory)
    ptr = staticmethod(ptr)
    _check = check

    @staticmethod
    def f64(a):
        return np.ascontiguousarray(a, dtype=np.float64)

    @staticmethod
    def i32(a):
        return np.ascontiguousarray(a, dtype=np.int32)

    @staticmethod
    def empty(*shape, dtype=np.float64):
        return np.zeros(shape, dtype=dtype)

    @staticmethod
    def _stream():
        return None

    def pcm16_to_f64(self, pcm, n_samples, divisor=2 ** 15 - 1):
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        B, S = pcm.shape
        x = np.zeros((B, S))
        self.check(self.L.wb_pcm16_to_f64(self.h, None, ptr(pcm), S, ptr(self.i32(n_samples)), B, float(divisor), ptr(x), S))
        return x

    def f64_to_pcm16(self, y, n_samples, gain=2 ** 15):
        y = self.f64(y)
        B, S = y.shape
        pcm = np.zeros((B, S), dtype=np.int16)
        self.check(self.L.wb_f64_to_pcm16(self.h, None, ptr(y), S, ptr(self.i32(n_samples)), B, float(gain), ptr(pcm), S))
        return pcm

    # batch helpers: x [B, S] float64, per-frame arrays [B, F]
    @staticmethod
    def _prep(x, tpos, f0, vuv):
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
        tpos = np.ascontiguousarray(np.atleast_2d(tpos), dtype=np.float64)
        f0 = np.ascontiguousarray(np.atleast_2d(f0), dtype=np.float64)
        vuv = np.ascontiguousarray(np.atleast_2d(vuv), dtype=np.float64)
        B, S = x.shape
        F = tpos.shape[1]
        ns = np.full(B, S, dtype=np.int32)
        nf = np.full(B, F, dtype=np.int32)
        return x, tpos, f0, vuv, B, S, F, ns, nf

    def cheaptrick(self, x, fs, tpos, f0, vuv, q1=-0.15, fft_size=0, dither=None, want_ps=True, seed=0):
        x, tpos, f0, vuv, B, S, F, ns, nf = self._prep(x, tpos, f0, vuv)
        n = fft_size or self.L.wb_cheaptrick_fft_size(fs)
        f0u = np.zeros((B, F))
        spec = np.zeros((B, F, n // 2 + 1))
        ps = np.zeros((B, F, n), dtype=np.complex128) if want_ps else None
        if dither is not None:
            dither = np.ascontiguousarray(dither, dtype=np.float64)
        self.check(self.L.wb_cheaptrick(self.h, None, ptr(x), S, ptr(ns), B, fs, ptr(tpos), ptr(f0), ptr(vuv),
                                         ptr(nf), F, q1, n, ptr(dither), seed, ptr(f0u), ptr(spec), ptr(ps)))
        return f0u, spec, ps

    def d4c(self, x, fs, tpos, f0, vuv, threshold=0.85, fft_size_for_spectrum=0):
        x, tpos, f0, vuv, B, S, F, ns, nf = self._prep(x, tpos, f0, vuv)
        nsp = fft_size_for_spectrum or self.L.wb_cheaptrick_fft_size(fs)
        nb = self.L.wb_d4c_band_count(fs, 0)
        f0o = np.zeros((B, F))
        ap = np.zeros((B, F, nsp // 2 + 1))
        co = np.zeros((B, F, nb))
        self.check(self.L.wb_d4c(self.h, None, ptr(x), S, ptr(ns), B, fs, ptr(tpos), ptr(f0), ptr(vuv), ptr(nf), F,
                                  threshold, nsp, ptr(f0o), ptr(ap), ptr(co)))
        return f0o, ap, co

    def d4c_requiem(self, x, fs, tpos, f0, vuv, threshold=0.85, fft_size=0):
        x, tpos, f0, vuv, B, S, F, ns, nf = self._prep(x, tpos, f0, vuv)
        nb = self.L.wb_d4c_band_count(fs, 1)
        f0o = np.zeros((B, F))
        ap = np.zeros((B, F, nb + 2))
        self.check(self.L.wb_d4c_requiem(self.h, None, ptr(x), S, ptr(ns), B, fs, ptr(tpos), ptr(f0), ptr(vuv),
                                          ptr(nf), F, threshold, fft_size, ptr(f0o), ptr(ap)))
        return f0o, ap

    def harvest(self, x, fs, f0_floor=71.0, f0_ceil=800.0, frame_period=5.0, n_samples=None, debug=False):
        """x [B, S] (or [S]).  Returns dict(temporal_positions, f0, vuv, n_frames) with [B, F] arrays;
        debug=True adds the workspace intermediates."""
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
        B, S = x.shape
        ns = np.full(B, S, dtype=np.int32) if n_samples is None else np.asarray(n_samples, dtype=np.int32)
        smax = int(ns.max())
        nbytes = C.c_size_t()
        self.check(self.L.wb_harvest_workspace_bytes(self.h, B, smax, fs, f0_floor, f0_ceil, C.byref(nbytes)))
        ws = np.zeros(nbytes.value // 8 + 1, dtype=np.float64)
        F = self.L.wb_frame_count(smax, fs, frame_period)
        tp, f0, vuv = np.zeros((B, F)), np.zeros((B, F)), np.zeros((B, F))
        nf = np.zeros(B, dtype=np.int32)
        self.check(self.L.wb_harvest(self.h, None, ptr(x), S, ptr(ns), B, smax, fs, f0_floor, f0_ceil, frame_period,
                                      ptr(ws), nbytes.value, F, ptr(tp), ptr(f0), ptr(vuv), ptr(nf)))
        out = {"temporal_positions": tp, "f0": f0, "vuv": vuv, "n_frames": nf}
        if debug:
            offs = (C.c_size_t * 16)()
            dims = (C.c_int * 8)()
            self.check(self.L.wb_harvest_workspace_layout(self.h, B, smax, fs, f0_floor, f0_ceil, offs, dims))
            ys, f1s, nch, maxc, slots = dims[0], dims[1], dims[2], dims[3], dims[4]
            raw8 = ws.view(np.uint8)

            def arr(i, dtype, shape):
                n = int(np.prod(shape)) * np.dtype(dtype).itemsize
                return raw8[offs[i]:offs[i] + n].view(dtype).reshape(shape).copy()

            out.update(y=arr(1, np.float64, (B, ys)), y_len=arr(2, np.int32, (B,)),
                       raw=arr(3, np.float64, (B, nch, f1s)), base_c=arr(5, np.float64, (B, f1s, maxc)),
                       base_n=arr(6, np.int32, (B, f1s)), l_f0=arr(7, np.float64, (B, f1s, slots)),
                       l_sc=arr(8, np.float64, (B, f1s, slots)), l_slot=arr(9, np.uint8, (B, f1s, slots)),
                       l_keep=arr(10, np.uint8, (B, f1s, slots)), l_n=arr(11, np.int32, (B, f1s)),
                       status=arr(13, np.int32, (1,)))
        return out

    def dio(self, x, fs, f0_floor=71.0, f0_ceil=800.0, channels_in_octave=2, target_fs=4000, frame_period=5.0,
            allowed_range=0.1, n_samples=None):
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
        B, S = x.shape
        ns = np.full(B, S, dtype=np.int32) if n_samples is None else np.asarray(n_samples, dtype=np.int32)
        smax = int(ns.max())
        nbytes = C.c_size_t()
        self.check(self.L.wb_dio_workspace_bytes(self.h, B, smax, fs, f0_floor, f0_ceil, channels_in_octave, target_fs,
                                                  frame_period, C.byref(nbytes)))
        ws = np.zeros(nbytes.value // 8 + 1, dtype=np.float64)
        F = self.L.wb_frame_count(smax, fs, frame_period)
        nb = self.L.wb_dio_band_count(f0_floor, f0_ceil, channels_in_octave)
        tp, f0, vuv = np.zeros((B, F)), np.zeros((B, F)), np.zeros((B, F))
        nf = np.zeros(B, dtype=np.int32)
        cand = np.zeros((B, F, nb))
        raw = np.zeros((B, nb, F))
        self.check(self.L.wb_dio(self.h, None, ptr(x), S, ptr(ns), B, smax, fs, f0_floor, f0_ceil, channels_in_octave,
                                  target_fs, frame_period, allowed_range, ptr(ws), nbytes.value, F, ptr(tp), ptr(f0),
                                  ptr(vuv), ptr(nf), ptr(cand), ptr(raw)))
        return {"temporal_positions": tp, "f0": f0, "vuv": vuv, "n_frames": nf, "f0_candidates": cand,
                "raw_f0_candidates": raw}

    def encode(self, x, fs, f0_method="harvest", is_requiem=False, n_samples=None, aperiodicity="both", dither=None,
               fft_size=0):
        """The fused wb_encode (csrc/wb_pipeline.cu) on the host emulation: x [B, S] -> dict of [B, F(, bins)] arrays."""
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
        B, S = x.shape
        ns = np.full(B, S, dtype=np.int32) if n_samples is None else np.asarray(n_samples, dtype=np.int32)
        smax = int(ns.max())
        mode = 2 if aperiodicity == "none" else int(bool(is_requiem))
        q = _abi.EncodeParams(int(fs), _abi.F0_METHODS[f0_method], 71.0, 800.0, 2, 4000, 5.0, 0.1, int(fft_size), mode,
                              -0.15, 0.85, 0)
        nbytes = C.c_size_t()
        self.check(self.L.wb_encode_workspace_bytes(self.h, C.byref(q), B, smax, C.byref(nbytes)))
        ws = np.zeros(nbytes.value // 8 + 1, dtype=np.float64)
        F = self.L.wb_frame_count(smax, fs, 5.0)
        n = fft_size or self.L.wb_cheaptrick_fft_size(fs)
        tp, f0, vuv = np.zeros((B, F)), np.zeros((B, F)), np.zeros((B, F))
        nf = np.zeros(B, dtype=np.int32)
        spec = np.zeros((B, F, n // 2 + 1))
        nb = self.L.wb_d4c_band_count(fs, 1 if is_requiem else 0)
        ap = co = None
        if mode == 1:
            ap = np.zeros((B, F, nb + 2))
        elif mode == 0:
            ap = np.zeros((B, F, n // 2 + 1)) if aperiodicity in ("full", "both") else None
            co = np.zeros((B, F, nb)) if aperiodicity in ("coarse", "both") else None
        if dither is not None:
            dither = np.ascontiguousarray(dither, dtype=np.float64)
        self.check(self.L.wb_encode(self.h, None, C.byref(q), ptr(x), S, ptr(ns), B, smax, ptr(ws), nbytes.value, F,
                                     ptr(dither), ptr(tp), ptr(f0), ptr(vuv), ptr(nf), ptr(spec), ptr(ap), ptr(co), None))
        return {"temporal_positions": tp, "f0": f0, "vuv": vuv, "n_frames": nf, "spectrogram": spec, "aperiodicity": ap,
                "coarse_ap": co}

    def expand_aperiodicity(self, coarse, fs, fft_size=0):
        c = np.ascontiguousarray(coarse, dtype=np.float64)
        n = fft_size or self.L.wb_cheaptrick_fft_size(fs)
        rows = c.size // c.shape[-1]
        out = np.zeros(c.shape[:-1] + (n // 2 + 1,))
        self.check(self.L.wb_d4c_expand(self.h, None, ptr(c), rows, fs, n, ptr(out)))
        return out

    def stonemask(self, x, fs, tpos, f0):
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
        tpos = np.ascontiguousarray(np.atleast_2d(tpos), dtype=np.float64)
        f0 = np.ascontiguousarray(np.atleast_2d(f0), dtype=np.float64)
        B, S = x.shape
        F = tpos.shape[1]
        ns = np.full(B, S, dtype=np.int32)
        nf = np.full(B, F, dtype=np.int32)
        out = np.zeros((B, F))
        self.check(self.L.wb_stonemask(self.h, None, ptr(x), S, ptr(ns), B, fs, ptr(tpos), ptr(f0), ptr(nf), F, ptr(out)))
        return out

    def nuttall(self, n):
        out = np.zeros(n)
        self.check(self.L.wb_debug_nuttall(n, ptr(out)))
        return out

    # ---------------------------------------------------------------- synthesis
    def _sy_common(self, dat_list, rows):
        B = len(dat_list)
        F = max(len(d["f0"]) for d in dat_list)
        fs = int(dat_list[0]["fs"])
        tp = np.zeros((B, F)); f0 = np.zeros((B, F)); vuv = np.zeros((B, F))
        nf = np.zeros(B, dtype=np.int32)
        for i, d in enumerate(dat_list):
            n = len(d["f0"])
            tp[i, :n] = d["temporal_positions"]; f0[i, :n] = d["f0"]; vuv[i, :n] = d["vuv"]; nf[i] = n
        ylen = max(self.L.wb_synthesis_length(float(d["temporal_positions"][0]), float(d["temporal_positions"][-1]), fs)
                   for d in dat_list)
        nbytes = C.c_size_t()
        self.check(self.L.wb_synthesis_workspace_bytes(self.h, B, ylen, rows, C.byref(nbytes)))
        ws = np.zeros(nbytes.value // 8 + 1)
        out_len = np.zeros(B, dtype=np.int32); n_p = np.zeros(B, dtype=np.int32); n_tot = np.zeros(B, dtype=np.int32)
        self.check(self.L.wb_synthesis_timebase(self.h, None, ptr(tp), ptr(f0), ptr(vuv), ptr(nf), B, F, fs, ylen,
                                                 ptr(ws), nbytes.value, rows, ptr(out_len), ptr(n_p), ptr(n_tot)))
        return B, F, fs, tp, f0, vuv, nf, ylen, ws, nbytes.value, out_len, n_p, n_tot

    def synthesis(self, dat_list, noise="legacy", normalize=True, seed=0):
        """dat_list: dicts in the reference layout (spectrogram / aperiodicity [bins, F])."""
        B, F, fs, tp, f0, vuv, nf, ylen, ws, wsb, out_len, n_p, n_tot = self._sy_common(dat_list, 0)
        nb = dat_list[0]["spectrogram"].shape[0]
        n = (nb - 1) * 2
        spec = np.ones((B, F, nb)); ap = np.zeros((B, F, nb))
        for i, d in enumerate(dat_list):
            k = len(d["f0"])
            spec[i, :k] = d["spectrogram"].T; ap[i, :k] = d["aperiodicity"].T
        nz = None
        stride = 0
        if noise == "legacy":
            stride = int(n_tot.max())
            nz = np.zeros((B, stride))
            for i in range(B):
                nz[i, :n_tot[i]] = np.random.randn(int(n_tot[i]))
        y = np.zeros((B, ylen))
        self.check(self.L.wb_synthesis(self.h, None, ptr(tp), ptr(f0), ptr(vuv), ptr(spec), ptr(ap), ptr(nf), B, F, fs, n,
                                        ptr(ws), wsb, ptr(nz), stride, seed, ptr(y), ylen, int(normalize)))
        return [y[i, :out_len[i]].copy() for i in range(B)], n_p

    def synthesis_requiem(self, dat_list, seeds, cursor=None, normalize=True):
        rows = dat_list[0]["aperiodicity"].shape[0]
        B, F, fs, tp, f0, vuv, nf, ylen, ws, wsb, out_len, n_p, n_tot = self._sy_common(dat_list, rows)
        nb = dat_list[0]["spectrogram"].shape[0]
        n = (nb - 1) * 2
        spec = np.ones((B, F, nb)); ap = np.zeros((B, F, rows))
        for i, d in enumerate(dat_list):
            k = len(d["f0"])
            spec[i, :k] = d["spectrogram"].T; ap[i, :k] = d["aperiodicity"].T
        pulse = np.ascontiguousarray(seeds["pulse"], dtype=np.float64)
        noise = np.ascontiguousarray(seeds["noise"], dtype=np.float64)
        cur_in = np.zeros(rows) if cursor is None else np.asarray(cursor, dtype=np.float64)
        cur_out = np.zeros((B, rows))
        y = np.zeros((B, ylen))
        self.check(self.L.wb_synthesis_requiem(self.h, None, ptr(tp), ptr(f0), ptr(vuv), ptr(spec), ptr(ap), ptr(nf), B, F,
                                                fs, n, rows, ptr(pulse), pulse.shape[0], ptr(noise), noise.shape[0],
                                                ptr(cur_in), ptr(cur_out), ptr(ws), wsb, ptr(y), ylen, int(normalize)))
        return [y[i, :out_len[i]].copy() for i in range(B)], cur_out
