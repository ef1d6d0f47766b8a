"""Harvest F0 estimator -- oracle restatement of world/harvest.py.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Stage functions mirror the
pipeline of harvest.py:17-54 and return their intermediates so that every stage
boundary can be pinned against tests/golden/syn16k_1s.npz.
"""
import numpy as np
from scipy import signal

from . import common as C

TARGET_FS = 8000
CHANNELS_PER_OCTAVE = 40


# --------------------------------------------------------------------------- stage 1
def zero_phase_iir(b, a, x, padlen):
    """scipy.signal.filtfilt(b, a, x, padlen=padlen) spelled out (harvest.py:599-603):
    odd extension by padlen samples, steady-state initial conditions scaled by the
    first sample of each pass, forward then time-reversed pass."""
    ext = np.concatenate([2 * x[0] - x[padlen:0:-1], x, 2 * x[-1] - x[-2:-padlen - 2:-1]])
    zi = signal.lfilter_zi(b, a)
    fwd, _ = signal.lfilter(b, a, ext, zi=zi * ext[0])
    bwd, _ = signal.lfilter(b, a, fwd[::-1], zi=zi * fwd[-1])
    return bwd[::-1][padlen:-padlen]


def downsample(x, fs):
    """CalculateDownsampledSignal + decimate_matlab (harvest.py:58-71, 584-609)."""
    x = np.asarray(x, dtype=np.float64)
    ratio = int(fs / TARGET_FS + 0.5)
    if fs <= TARGET_FS:
        y = x.copy()
        actual = fs
    else:
        pad = int(np.ceil(140 / ratio) * ratio)
        xe = np.concatenate([np.full(pad, x[0]), x, np.full(pad, x[-1])])
        b, a = signal.cheby1(3, 0.05, 0.8 / ratio)
        z = zero_phase_iir(b, a, xe, 3 * (max(len(a), len(b)) - 1))
        n_out = np.ceil(len(z) / ratio)
        first = int(ratio - (ratio * n_out - len(z)))
        z = z[first - 1::ratio]
        actual = fs / ratio
        y = z[int(pad / ratio):int(-pad / ratio)]
    return y - np.mean(y), actual


def band_edges(f0_floor, f0_ceil):
    """harvest.py:23-29."""
    lo, hi = f0_floor * 0.9, f0_ceil * 1.1
    k = np.arange(np.ceil(np.log2(hi / lo) * CHANNELS_PER_OCTAVE)) + 1
    return lo * 2.0 ** (k / CHANNELS_PER_OCTAVE)


# --------------------------------------------------------------------------- stage 2
def crossing_intervals(sig, fs):
    """ZeroCrossingEngine (harvest.py:283-297): positive-to-negative crossings,
    1-based positions refined by linear interpolation; returns interval mid-times
    and the interval-based F0."""
    nxt = np.concatenate([sig[1:], sig[-1:]])
    at = np.nonzero((nxt * sig < 0) & (nxt < sig))[0] + 1        # 1-based
    fine = at - sig[at - 1] / (sig[at] - sig[at - 1])
    return (fine[:-1] + fine[1:]) / 2 / fs, fs / np.diff(fine)


def channel_candidates(edge, afs, y_spec, y_len, times, f0_floor, f0_ceil):
    """CalculateRawEvent + GetF0Candidates (harvest.py:252-278, 499-529)."""
    half = int(C.round_half_up(afs / edge * 2))
    k = np.arange(-half, half + 1)
    taps = C.nuttall(2 * half + 1) * np.cos(2 * np.pi * edge * k / afs)
    filt = np.real(np.fft.ifft(np.fft.fft(taps, len(y_spec)) * y_spec))
    s = filt[half + 1:half + 1 + y_len]
    d = np.diff(s)
    streams = [crossing_intervals(v, afs) for v in (s, -s, d, -d)]
    usable = 1
    for loc, _ in streams:
        usable *= max(0, len(loc) - 2)
    if usable <= 0:
        return np.zeros(len(times))
    est = np.mean([C.lerp_extrap(loc, f, times) for loc, f in streams], axis=0)
    est[est > edge * 1.1] = 0
    est[est < edge * 0.9] = 0
    est[est > f0_ceil] = 0
    est[est < f0_floor] = 0
    return est


def raw_candidates(y, afs, fs, times, f0_floor, f0_ceil):
    """CalculateCandidates (harvest.py:75-84); FFT length from harvest.py:33 (uses fs,
    not the decimated rate)."""
    edges = band_edges(f0_floor, f0_ceil)
    n = int(2 ** np.ceil(np.log2(len(y) + int(fs / (f0_floor * 0.9) * 4 + 0.5) + 1)))
    spec = np.fft.fft(y, n)
    return np.stack([channel_candidates(e, afs, spec, len(y), times, f0_floor, f0_ceil) for e in edges])


# --------------------------------------------------------------------------- stage 3
def detect(raw, min_run=10):
    """DetectCandidates (harvest.py:88-110): in each frame, every run of more than
    min_run consecutive non-zero channels (first and last channel ignored) yields
    one candidate, the mean over the run."""
    n_ch, n_fr = raw.shape
    out = np.zeros((int(n_ch / 10 + 0.5), n_fr))
    most = 0
    on = (raw > 0).astype(np.int8)
    on[0] = 0
    on[-1] = 0
    step = np.diff(on, axis=0)
    for i in range(n_fr):
        starts = np.nonzero(step[:, i] == 1)[0]
        ends = np.nonzero(step[:, i] == -1)[0]
        c = 0
        for s, e in zip(starts, ends):
            if e - s >= min_run:
                out[c, i] = np.mean(raw[s + 1:e + 1, i])
                c += 1
        most = max(most, c)
    return out, most


def overlap(cand, most, reach=3):
    """OverlapF0Candidates (harvest.py:114-125): candidates of frames -reach..+reach
    are offered to each frame.  Row 0, columns 0..reach-1 keep whatever row 2*reach
    of the input held there (a leftover of the reference's first assignment)."""
    n_fr = cand.shape[1]
    width = 2 * reach + 1
    out = np.zeros((width * most, n_fr))
    if most == 0:
        return out
    out[0, :] = cand[width - 1, :]
    for s in range(width):
        shift = reach - s                       # frame f takes candidates of frame f - shift
        rows = slice(s * most, (s + 1) * most)
        if shift >= 0:
            out[rows, shift:] = cand[:most, :n_fr - shift]
        else:
            out[rows, :n_fr + shift] = cand[:most, -shift:]
    return out


# --------------------------------------------------------------------------- stage 4
def refine(y, afs, times, cand, f0_floor, f0_ceil):
    """RefineCandidates / GetRefinedF0 (harvest.py:131-150, 169-211) for every
    non-zero candidate; candidates with equal window length are processed together."""
    new = np.zeros_like(cand)
    score = np.zeros_like(cand)
    rows, cols = np.nonzero(cand)
    if len(rows) == 0:
        return new, score
    c0 = cand[rows, cols]
    t0 = times[cols]
    halves = np.ceil(3 * afs / c0 / 2)
    for h in np.unique(halves):
        sel = np.nonzero(halves == h)[0]
        f = c0[sel][:, None]
        t = t0[sel][:, None]
        h = int(h)
        length = 2 * h + 1
        span = length / afs
        base = np.arange(-h, h + 1)[None, :] / afs
        n_fft = int(2 ** (np.ceil(np.log2(length)) + 1))
        raw_idx = C.half_away((t + base) * afs + 0.001)
        ph = np.pi * ((raw_idx - 1) / afs - t) / span
        main = 0.42 + 0.5 * np.cos(2 * ph) + 0.08 * np.cos(4 * ph)
        dwin = np.empty_like(main)
        dwin[:, 0] = -main[:, 1] / 2
        dwin[:, -1] = main[:, -2] / 2
        dwin[:, 1:-1] = -(main[:, 2:] - main[:, :-2]) / 2
        seg = y[(np.clip(raw_idx, 1, len(y)) - 1).astype(np.int64)]
        S = np.fft.fft(seg * main, n_fft, axis=1)
        D = np.fft.fft(seg * dwin, n_fft, axis=1)
        power = np.abs(S) ** 2
        inst = (np.arange(n_fft)[None, :] / n_fft + (S.real * D.imag - S.imag * D.real) / power / 2 / np.pi) * afs
        n_harm = np.minimum(np.floor(afs / 2 / f[:, 0]), 6).astype(np.int64)
        harm = np.arange(1, 7)[None, :]
        live = harm <= n_harm[:, None]
        bins = np.where(live, C.half_away(f * n_fft / afs * harm).astype(np.int64), 0)
        r = np.arange(len(sel))[:, None]
        inst_h = inst[r, bins]
        amp_h = np.where(live, np.sqrt(power[r, bins]), 0.0)
        ref = np.sum(amp_h * inst_h, axis=1) / np.sum(amp_h * harm, axis=1)
        var = np.where(live, np.abs((inst_h / harm - f) / f), 0.0)
        sc = 1 / (0.000000000001 + var.sum(axis=1) / n_harm)
        bad = (ref < f0_floor) | (ref > f0_ceil) | (sc < 2.5)
        new[rows[sel], cols[sel]] = np.where(bad, 0.0, ref)
        score[rows[sel], cols[sel]] = np.where(bad, 0.0, sc)
    return new, score


def prune(cand, score, tol=0.05):
    """RemoveUnreliableCandidates (harvest.py:215-234): a candidate survives only if a
    candidate of the previous or the next frame lies within tol of it."""
    out_c, out_s = cand.copy(), score.copy()
    n_fr = cand.shape[1]
    ref = cand[:, None, 1:n_fr - 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        e_next = np.abs(ref - cand[None, :, 2:]) / ref
        e_prev = np.abs(ref - cand[None, :, :n_fr - 2]) / ref
    best = np.minimum(1.0, np.minimum(e_next.min(axis=1), e_prev.min(axis=1)))
    kill = (cand[:, 1:n_fr - 1] != 0) & (best > tol)
    out_c[:, 1:n_fr - 1][kill] = 0
    out_s[:, 1:n_fr - 1][kill] = 0
    return out_c, out_s


# --------------------------------------------------------------------------- stage 5
def voiced_runs(f0):
    """GetBoundaryList (harvest.py:572-580): inclusive (start, end) pairs of runs of
    non-zero values, with the first and last element treated as zero."""
    v = (np.asarray(f0) != 0).astype(np.int8)
    v[0] = 0
    v[-1] = 0
    d = np.diff(v)
    return list(zip(np.nonzero(d == 1)[0] + 1, np.nonzero(d == -1)[0]))


def nearest_candidate(ref, column, tol):
    """SelectBestF0 (harvest.py:238-248): the last candidate whose relative distance to
    ref does not exceed the running best (initially tol); 0 if none."""
    best, err = 0.0, tol
    for c in column:
        e = abs(ref - c) / ref
        if e > err:
            continue
        best, err = c, e
    return best


def _track(seq, origin, stop, step, cand, tol):
    """ExtendF0 (harvest.py:398-421)."""
    seq = seq.copy()
    cur = seq[origin]
    reached = origin
    misses = 0
    for i in range(origin, stop + step, step):
        seq[i + step] = nearest_candidate(cur, cand[:, i + step], tol)
        if seq[i + step] != 0:
            cur = seq[i + step]
            misses = 0
            reached = i + step
        else:
            misses += 1
        if misses == 4:
            break
    return seq, reached


def _score_of(value, column, scores):
    """SerachScore (harvest.py:488-495)."""
    s = 0.0
    for c, w in zip(column, scores):
        if value == c and s < w:
            s = w
    return s


def _merge(tracks, spans, cand, score):
    """MergeF0 / MergeF0Sub (harvest.py:437-484)."""
    order = np.argsort(spans[:, 0], kind="stable")
    f0 = tracks[order[0]].copy()
    st1, ed1 = int(spans[order[0], 0]), int(spans[order[0], 1])
    for j in order[1:]:
        st2, ed2 = int(spans[j, 0]), int(spans[j, 1])
        other = tracks[j]
        if st2 - ed1 > 0:
            f0[st2:ed2 + 1] = other[st2:ed2 + 1]
            st1, ed1 = st2, ed2
        elif st1 <= st2 and ed1 >= ed2:
            pass
        else:
            s1 = sum(_score_of(f0[i], cand[:, i], score[:, i]) for i in range(st2, ed1 + 1))
            s2 = sum(_score_of(other[i], cand[:, i], score[:, i]) for i in range(st2, ed1 + 1))
            if s1 > s2:
                f0[ed1:ed2 + 1] = other[ed1:ed2 + 1]
            else:
                f0[st2:ed2 + 1] = other[st2:ed2 + 1]
            ed1 = ed2
    return f0


def connect(cand, score):
    """FixF0Contour (harvest.py:301-311) = SearchF0Base + FixStep1..4."""
    n = cand.shape[1]
    base = cand[np.argmax(score, axis=0), np.arange(n)]
    # step 1 (harvest.py:324-338): drop frames that jump relative to both predictors
    s1 = base.copy()
    s1[:2] = 0
    pred = base[1:-1] * 2 - base[:-2]
    with np.errstate(divide="ignore", invalid="ignore"):
        jump = (np.abs((base[2:] - pred) / (pred + C.EPS)) > 0.008) & \
               (np.abs((base[2:] - base[1:-1]) / (base[1:-1] + C.EPS)) > 0.008)
    s1[2:][jump & (base[2:] != 0)] = 0
    # step 2 (harvest.py:343-352): drop short voiced runs
    s2 = s1.copy()
    for st, ed in voiced_runs(s1):
        if ed - st < 6:
            s2[st:ed + 1] = 0
    # step 3 (harvest.py:357-384): extend every run both ways along the candidates
    runs = voiced_runs(s2)
    tracks, spans = [], []
    for st, ed in runs:
        seq = np.zeros(n)
        seq[st:ed + 1] = s2[st:ed + 1]
        seq, hi = _track(seq, ed, min(n - 2, ed + 100), 1, cand, 0.18)
        seq, lo = _track(seq, st, max(1, st - 100), -1, cand, 0.18)
        if 2200 / np.mean(seq[lo:hi + 1]) < hi - lo:
            tracks.append(seq)
            spans.append((lo, hi))
    s3 = _merge(tracks, np.array(spans), cand, score) if tracks else s2.copy()
    # step 4 (harvest.py:389-405): bridge short unvoiced gaps linearly
    s4 = s3.copy()
    runs = voiced_runs(s3)
    for (_, ed), (st, _) in zip(runs[:-1], runs[1:]):
        gap = st - ed - 1
        if gap >= 9:
            continue
        lo, hi = s3[ed] + 1, s3[st] - 1
        slope = (hi - lo) / (gap + 1)
        for c, j in enumerate(range(ed + 1, st), start=1):
            s4[j] = lo + slope * c
    return s4, (s4 != 0).astype(np.float64)


def smooth(f0):
    """SmoothF0 / FilterF0 (harvest.py:533-559): every voiced run is held constant
    outside itself and filtered forward and backward with a 2nd-order low-pass."""
    b = np.array([0.0078202080334971724, 0.015640416066994345, 0.0078202080334971724])
    a = np.array([1.0, -1.7347257688092754, 0.76600660094326412])
    padded = np.concatenate([np.zeros(300), f0, np.zeros(300)])
    out = padded.copy()
    for st, ed in voiced_runs(padded):
        held = np.zeros_like(padded)
        held[st:ed + 1] = padded[st:ed + 1]
        held[:st] = held[st]
        held[ed + 1:] = held[ed]
        fwd = signal.lfilter(b, a, held)
        bwd = signal.lfilter(b, a, fwd[::-1])[::-1]
        out[st:ed + 1] = bwd[st:ed + 1]
    return out[300:len(out) - 300]


# --------------------------------------------------------------------------- driver
def harvest(x, fs, f0_floor=71, f0_ceil=800, frame_period=5, stages=None):
    """harvest.py:17-54.  Returns dict(temporal_positions, f0, vuv); when `stages` is a
    dict it receives every intermediate."""
    x = np.asarray(x, dtype=np.float64)
    n1 = C.frame_count(len(x), fs, 1)
    t1 = np.arange(0, n1) * 1 / 1000
    y, afs = downsample(x, fs)
    raw = raw_candidates(y, afs, fs, t1, f0_floor, f0_ceil)
    det, most = detect(raw)
    ov = overlap(det, most)
    rf, rs = refine(y, afs, t1, ov, f0_floor, f0_ceil)
    pf, psc = prune(rf, rs)
    conn, vuv1 = connect(pf, psc)
    sm = smooth(conn)
    n = C.frame_count(len(x), fs, frame_period)
    tp = np.arange(0, n) * frame_period / 1000
    pick = np.minimum(len(sm) - 1, C.half_away(tp * 1000)).astype(np.int64)
    if stages is not None:
        stages.update(y=y, actual_fs=afs, raw=raw, detect=det, ncand=most, refined_f0=rf, refined_score=rs,
                      reliable_f0=pf, reliable_score=psc, connected=conn, smoothed=sm)
    return {"temporal_positions": tp, "f0": sm[pick], "vuv": vuv1[pick]}
