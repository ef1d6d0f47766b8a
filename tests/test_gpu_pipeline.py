"""GPU tier: the fused C entry points (wb_encode / wb_decode), the compact aperiodicity transport, the batch API's
input checks and pipelining, the partial pipelines of the facade (main.py:27-104), and oracle parity on utterances
drawn from the workloads bench.py times (BASELINE configs 2, 3 and 5 at their full 4 s length)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, spec_close

pytestmark = pytest.mark.gpu

F0_RTOL = 1e-6
AP_ATOL = 1e-5   # linear aperiodicity
AP_DB_ATOL = 1e-3  # requiem band aperiodicity, dB


@pytest.fixture(scope="module")
def partial():
    return dict(np.load(os.path.join(GOLDEN, "partial.npz")))


def _voiced_rel(f0, g):
    m = g > 0
    return float(np.max(np.abs(f0[m] - g[m]) / g[m])) if m.any() else 0.0


# ------------------------------------------------------------------ oracle parity on the benchmarked workloads
@pytest.mark.parametrize("fs,config,req,idx", [
    (16000, 2, False, (0, 31, 64, 100, 127, 200, 230, 255)),   # config 2: Harvest + CheapTrick + D4C
    (16000, 3, True, (0, 85, 170, 255)),                        # config 3: full encode, D4C-Requiem
    (48000, 5, False, (0, 127)),                                # config 5: 48 kHz, FFT 2048
])
def test_bench_workload_vs_oracle(engine, fs, config, req, idx):
    import oracle_pool
    from world_b200 import synth_input
    xs = np.stack([synth_input.utterance(fs, 4.0, config, i) for i in idx])
    want = oracle_pool.encode_many([(fs, 4.0, config, i, req) for i in idx])
    X, ns = engine.f64(xs), engine.i32([xs.shape[1]] * len(idx))
    d = engine.encode(X, ns, fs, f0_method="harvest", is_requiem=req, streams=2)
    f0, vuv = d["f0"].cpu().numpy(), d["vuv"].cpu().numpy()
    sp, ap = d["spectrogram"].cpu().numpy(), d["aperiodicity"].cpu().numpy()
    assert list(d["n_frames"].cpu().numpy()) == [801] * len(idx)
    for u, w in enumerate(want):
        assert np.array_equal(d["temporal_positions"][u].cpu().numpy(), w["temporal_positions"])
        assert np.array_equal(vuv[u], w["vuv"]), "voiced/unvoiced decisions differ (utterance %d)" % idx[u]
        assert np.array_equal(f0[u] > 0, w["f0"] > 0)
        assert _voiced_rel(f0[u], w["f0"]) <= F0_RTOL
        p99, mx = spec_close(sp[u].T, w["spectrogram"])
        assert p99 <= 1e-4 and mx <= 1e-3, (p99, mx)
        if config != 5:  # config 5 is Harvest + CheapTrick only (BASELINE.json)
            assert np.max(np.abs(ap[u].T - w["aperiodicity"])) <= (AP_DB_ATOL if req else AP_ATOL)


# ------------------------------------------------------------------ fused entry points vs the stage calls
@pytest.mark.parametrize("method,req", [("harvest", False), ("dio", True)])
def test_fused_encode_equals_stages(engine, syn16k, method, req):
    import torch
    x = syn16k["x"]
    X = engine.f64(np.stack([x, np.r_[x[:12000], np.zeros(4000)]]))
    ns = engine.i32([16000, 12000])
    d = engine.encode(X, ns, 16000, f0_method=method, is_requiem=req, aperiodicity="full" if req else "both")
    if method == "harvest":
        tp, f0, vuv, nf = engine.harvest(X, ns, 16000)
    else:
        tp, f0, vuv, nf = engine.dio(X, ns, 16000)
        f0 = engine.stonemask(X, ns, 16000, tp, f0, nf)
    f0u, spec, _ = engine.cheaptrick(X, ns, 16000, tp, f0, vuv, nf)
    if req:
        f0o, ap = engine.d4c_requiem(X, ns, 16000, tp, f0u, vuv, nf)
    else:
        f0o, ap, coarse = engine.d4c(X, ns, 16000, tp, f0u, vuv, nf, want_coarse=True)
    for u, k in enumerate(nf.cpu().numpy()):
        assert torch.equal(d["temporal_positions"][u, :k], tp[u, :k]) and torch.equal(d["vuv"][u, :k], vuv[u, :k])
        assert torch.equal(d["f0"][u, :k], f0o[u, :k]) and torch.equal(d["spectrogram"][u, :k], spec[u, :k])
        assert torch.equal(d["aperiodicity"][u, :k], ap[u, :k])
        if not req:
            assert torch.equal(d["coarse_ap"][u, :k], coarse[u, :k])
    assert torch.equal(d["n_frames"], nf)


def test_fused_decode_equals_split_calls(engine, syn16k):
    import torch
    g = syn16k
    F = len(g["harvest_d4c_f0"])
    sp = engine.f64(np.ascontiguousarray(g["harvest_d4c_spectrogram"].T)[None])
    ap = engine.f64(np.ascontiguousarray(g["harvest_d4c_aperiodicity"].T)[None])
    tp, f0, vuv = (engine.f64(g["harvest_d4c_" + k][None]) for k in ("temporal_positions", "f0", "vuv"))
    nf = engine.i32([F])
    ylen = engine.synthesis_length(0.0, float(g["harvest_d4c_temporal_positions"][-1]), 16000)
    y1, l1 = engine.synthesis(tp, f0, vuv, sp, ap, nf, 16000, ylen, noise="device", seed=11)
    y2, l2, _ = engine.decode(tp, f0, vuv, sp, ap, nf, 16000, ylen, seed=11)
    # (the overlap-add accumulates with float64 atomics: the last bits depend on the arrival order)
    assert torch.equal(l1, l2) and float((y1 - y2).abs().max()) < 1e-12
    # requiem flavour
    bap = engine.f64(np.ascontiguousarray(g["harvest_req_aperiodicity"].T)[::1][None]) if g["harvest_req_aperiodicity"].shape[1] == F else None
    if bap is not None:
        ps, nz = engine.f64(g["harvest_req_seed_pulse"]), engine.f64(g["harvest_req_seed_noise"])
        sp2 = engine.f64(np.ascontiguousarray(g["harvest_req_spectrogram"].T)[None]) if g["harvest_req_spectrogram"].shape[1] == F else sp
        y3, l3, c3 = engine.synthesis_requiem(tp, f0, vuv, sp2, bap, nf, 16000, ylen, ps, nz)
        y4, l4, c4 = engine.decode(tp, f0, vuv, sp2, bap, nf, 16000, ylen, is_requiem=True, seeds=(ps, nz))
        assert float((y3 - y4).abs().max()) < 1e-12 and torch.equal(c3, c4)


# ------------------------------------------------------------------ compact aperiodicity transport
def test_coarse_transport(engine, syn16k, mwm):
    import torch
    from world_b200 import main
    W = main.World()
    for x, fs in ((syn16k["x"], 16000), (mwm["x"][:40000], int(mwm["fs"]))):
        xs = np.stack([x, x[::-1].copy()])
        full = W.encode_batch(fs, xs, aperiodicity="full")
        ap_full = full["aperiodicity"].clone()
        d2h_full = full["_d2h_bytes"]
        c = W.encode_batch(fs, xs, aperiodicity="coarse")
        assert "coarse_ap" in c and "aperiodicity" in c and not dict.__contains__(c, "aperiodicity")
        assert c["_d2h_bytes"] < 0.52 * d2h_full
        # device expansion: the same bits as the D4C kernel's own
        dev = engine.expand_aperiodicity(c["coarse_ap"].to(engine.device), fs)
        assert torch.equal(dev.cpu(), ap_full)
        # lazy host expansion (reference expressions): a few ulp from the device's exp()
        host = c["aperiodicity"]
        assert float((host - ap_full).abs().max()) < 1e-14
        # decoding either form gives the same samples
        y1 = W.decode_batch(dict(full), seed=5)["out"].clone()
        cc = {k: v for k, v in c.items() if k != "aperiodicity"}
        y2 = W.decode_batch(cc, seed=5)["out"]
        assert float((y1 - y2).abs().max()) < 1e-12  # atomics: last-bit order dependence only


def test_f32_spectrogram_transport(engine, syn16k):
    """Opt-in lossy transport: the spectrogram crosses PCIe as float32 = the float64 result rounded to nearest
    (6e-8 relative, vs. the 1e-4 p99 parity tolerance in log10); decode_batch widens it again."""
    import torch
    from world_b200 import main
    W = main.World()
    xs = np.stack([syn16k["x"], syn16k["x"][::-1].copy()])
    a = W.encode_batch(16000, xs, aperiodicity="full")
    sp64 = a["spectrogram"].clone()
    d2h = a["_d2h_bytes"]
    ya = W.decode_batch(dict(a), seed=2)["out"].clone()
    b = W.encode_batch(16000, xs, aperiodicity="full", spectrogram_dtype=torch.float32)
    assert b["spectrogram"].dtype == torch.float32 and torch.equal(b["spectrogram"], sp64.to(torch.float32))
    assert b["_d2h_bytes"] < d2h - sp64.numel() * 4 + 64
    yb = W.decode_batch(dict(b), seed=2)["out"]
    assert float(((ya - yb) ** 2).mean().sqrt()) < 1e-6


# ------------------------------------------------------------------ batch API: checks, pipelining, ragged batches
def test_encode_batch_inputs_and_pipelining(engine, syn16k):
    import torch
    from world_b200 import main
    W = main.World()
    x = syn16k["x"]
    lens = [16000, 9000, 12345, 16000, 4000]
    xs = np.zeros((5, 16000))
    for u, n in enumerate(lens):
        xs[u, :n] = x[:n]
    a = W.encode_batch(16000, xs, n_samples=lens, pipeline=3, aperiodicity="full")
    keep = {k: a[k].clone() for k in ("f0", "vuv", "spectrogram", "aperiodicity", "n_frames", "temporal_positions")}
    # float32 / non-contiguous torch inputs are converted, not reinterpreted (values equal after the cast)
    x32 = torch.from_numpy(xs.astype(np.float32))
    b32 = W.encode_batch(16000, x32, n_samples=torch.tensor(lens), pipeline=1, aperiodicity="full")
    b64 = W.encode_batch(16000, xs.astype(np.float32).astype(np.float64), n_samples=lens, pipeline=2, aperiodicity="full")
    assert torch.equal(b32["f0"], b64["f0"]) and torch.equal(b32["spectrogram"], b64["spectrogram"])
    wide = torch.from_numpy(np.ascontiguousarray(np.repeat(xs, 2, axis=1)))[:, ::2]
    c = W.encode_batch(16000, wide, n_samples=lens, pipeline=2, aperiodicity="full")
    assert torch.equal(c["f0"], keep["f0"]) and torch.equal(c["aperiodicity"], keep["aperiodicity"])
    with pytest.raises(ValueError):
        W.encode_batch(16000, xs, n_samples=[16001, 1, 1, 1, 1])
    with pytest.raises(ValueError):
        W.encode_batch(16000, xs[0])
    with pytest.raises(TypeError):
        engine.encode(torch.zeros(2, 100, dtype=torch.float32, device=engine.device), engine.i32([100, 100]), 16000)
    # every row equals the single-utterance call; frames past n_frames[u] are zero, not stale
    again = W.encode_batch(16000, xs[:, ::-1].copy(), n_samples=[16000] * 5, pipeline=3, aperiodicity="full")  # dirties the buffers
    a = W.encode_batch(16000, xs, n_samples=lens, pipeline=3, aperiodicity="full")
    for u, n in enumerate(lens):
        one = W.encode_batch(16000, xs[u:u + 1, :n], aperiodicity="full")
        k = int(one["n_frames"][0])
        assert int(a["n_frames"][u]) == k
        assert torch.equal(a["f0"][u, :k], one["f0"][0]) and torch.equal(a["vuv"][u, :k], one["vuv"][0])
        assert torch.equal(a["aperiodicity"][u, :k], one["aperiodicity"][0])
        # the hash dither of CheapTrick is keyed by the frame's position in the batch: eps-level differences
        p99, mx = spec_close(a["spectrogram"][u, :k].numpy(), one["spectrogram"][0].numpy())
        assert mx <= 1e-3 and p99 <= 1e-4
        for key in ("f0", "vuv", "spectrogram", "aperiodicity", "temporal_positions"):
            assert float(a[key][u, k:].abs().max()) == 0.0 if k < a[key].shape[1] else True


# ------------------------------------------------------------------ partial pipelines (main.py:27-104)
def test_partial_pipelines_vs_reference(engine, syn16k, partial):
    from world_b200 import main
    W = main.World()
    x, g = syn16k["x"], partial
    for m in ("harvest", "dio"):
        tp, f0, vuv = W.get_f0(16000, x.copy(), f0_method=m)
        assert np.array_equal(tp, g["get_f0_%s_tp" % m]) and np.array_equal(vuv, g["get_f0_%s_vuv" % m])
        assert _voiced_rel(f0, g["get_f0_%s_f0" % m]) <= 1e-9 and np.array_equal(f0 > 0, g["get_f0_%s_f0" % m] > 0)
    with pytest.raises(Exception):
        W.get_f0(16000, x.copy(), f0_method="nope")
    np.random.seed(0)
    sp = W.get_spectrum(16000, x.copy(), f0_method="dio")
    assert set(sp) == {"f0", "temporal_positions", "fs", "ps spectrogram", "spectrogram"}
    assert _voiced_rel(sp["f0"], g["get_spectrum_f0"]) <= 1e-9  # as CheapTrick leaves it: 500 at unvoiced frames
    p99, mx = spec_close(sp["spectrogram"][:, ::4], g["get_spectrum_spectrogram"])
    assert p99 <= 1e-6 and mx <= 1e-4
    F = len(g["gvn_d4c_f0_in"])
    src = {"temporal_positions": np.arange(F) * 0.005, "f0": g["gvn_d4c_f0_in"].copy(), "vuv": np.ones(F)}
    np.random.seed(0)
    d = W.encode_w_gvn_f0(16000, x.copy(), src, fft_size=1024)
    assert set(d) == {"temporal_positions", "vuv", "f0", "fs", "spectrogram", "aperiodicity", "coarse_ap", "is_requiem"}
    assert np.max(np.abs(d["f0"] - g["gvn_d4c_f0"])) < 1e-12
    p99, mx = spec_close(d["spectrogram"][:, ::4], g["gvn_d4c_spectrogram"])
    assert p99 <= 1e-6 and mx <= 1e-4
    assert np.max(np.abs(d["aperiodicity"][:, ::4] - g["gvn_d4c_aperiodicity"])) < 1e-8
    assert np.max(np.abs(d["coarse_ap"] - g["gvn_d4c_coarse_ap"])) < 1e-6
    with pytest.raises(AssertionError):  # main.py:86: every frame must lie above 3 fs / fft_size
        W.encode_w_gvn_f0(16000, x.copy(), {"temporal_positions": src["temporal_positions"], "f0": np.zeros(F), "vuv": np.ones(F)}, fft_size=1024)
    with pytest.raises(KeyError):  # main.py:102: the reference reads source['coarse_ap'], which d4cRequiem never sets
        W.encode_w_gvn_f0(16000, x.copy(), {"temporal_positions": src["temporal_positions"], "f0": src["f0"].copy(), "vuv": np.ones(F)},
                          fft_size=1024, is_requiem=True)


# ------------------------------------------------------------------ decode after a non-uniform duration edit
def test_decode_after_nonuniform_duration_edit(engine, partial, syn16k):
    import random
    from world_b200 import main, synthesisRequiem
    W = main.World()
    g = partial
    for tag, req in (("harvest_d4c_", False), ("harvest_req_", True)):
        dat = {"temporal_positions": syn16k["harvest_d4c_temporal_positions"].copy(), "f0": g["dur_" + tag + "f0"].copy(),
               "vuv": g["dur_" + tag + "vuv"].copy(), "fs": 16000, "is_requiem": req,
               "spectrogram": g["dur_" + tag + "spectrogram"].copy(), "aperiodicity": g["dur_" + tag + "aperiodicity"].copy()}
        W.modify_duration(dat, [0.3, 0.6], [0.0, 0.2, 0.8, -1])
        assert np.array_equal(dat["temporal_positions"], g["dur_" + tag + "tp"])
        np.random.seed(0)
        random.seed(0)
        synthesisRequiem.generate_noise.current_index = None
        W.decode(dat)
        want = g["dur_" + tag + "out"]
        assert dat["out"].shape == want.shape
        rms = float(np.sqrt(np.mean((dat["out"] - want) ** 2)))
        assert rms <= 1e-4, (tag, rms)  # BASELINE.json: within 1e-4 RMS with the replayed noise stream


# ------------------------------------------------------------------ DIO candidates, device noise statistics
def test_dio_sorted_candidates_vs_reference(engine, mwm, syn16k):
    for g in (mwm, syn16k):
        x, fs = g["x"], int(g["fs"])
        X, ns = engine.f64(x[None]), engine.i32([len(x)])
        tp, f0, vuv, nf, cand, raw = engine.dio(X, ns, fs, want_candidates=True)
        got = cand.cpu().numpy()[0].T      # [bands, F] like the reference's f0_candidates
        want = g["dio_d4c_dio_f0_candidates"]
        assert got.shape == want.shape
        assert np.max(np.abs(got - want)) < 1e-6


def test_device_noise_is_standard_normal(engine):
    """The counter-based generator behind decode_batch.  An all-unvoiced utterance with a flat spectrum turns the
    normals each pulse draws (synthesis.py:93-95: zero-mean randn(32) through the aperiodic response) into the output,
    so the output's mean, variance, lag-1 autocorrelation and kurtosis are those of the same configuration run by the
    oracle with np.random.randn -- a wrong variance or a correlated / non-Gaussian generator shows up here."""
    from oracle import synthesis as o_syn
    B, F, n, fs = 4, 401, 1024, 16000
    tp1 = np.arange(F) * 0.005
    dat = {"temporal_positions": tp1, "f0": np.zeros(F), "vuv": np.zeros(F), "fs": fs, "is_requiem": False,
           "spectrogram": np.ones((n // 2 + 1, F)), "aperiodicity": np.full((n // 2 + 1, F), 1 - 1e-12)}
    np.random.seed(0)
    ref = o_syn.synthesis(dat)[2000:-2000]

    def stats(v):
        m = float(np.mean(v))
        c = v - m
        var = float(np.mean(c * c))
        return m, var, float(np.mean(c[1:] * c[:-1]) / var), float(np.mean(c ** 4) / var ** 2)

    m0, v0, r0, k0 = stats(ref)
    tp = engine.f64(np.tile(tp1, (B, 1)))
    f0, vuv = engine.f64(np.zeros((B, F))), engine.f64(np.zeros((B, F)))
    sp = engine.f64(np.ones((B, F, n // 2 + 1)))
    ap = engine.f64(np.full((B, F, n // 2 + 1), 1 - 1e-12))
    nf = engine.i32([F] * B)
    ylen = engine.synthesis_length(0.0, float(tp1[-1]), fs)
    ys = []
    for seed in (1, 2):
        y, ln, _ = engine.decode(tp, f0, vuv, sp, ap, nf, fs, ylen, seed=seed, normalize=False)
        assert int(ln[0]) == len(ref) + 4000
        ys.append(y.cpu().numpy()[:, 2000:int(ln[0]) - 2000])
    assert not np.array_equal(ys[0], ys[1])        # the seed matters
    assert not np.array_equal(ys[0][0], ys[0][1])  # and so does the utterance
    sd = np.sqrt(v0 / len(ref))
    for u in range(B):
        m, v, r1, k = stats(ys[0][u])
        assert abs(m - m0) < 6 * sd + 1e-3, (m, m0)
        assert abs(v / v0 - 1) < 0.08, (v, v0)
        assert abs(r1 - r0) < 0.05, (r1, r0)
        assert abs(k - k0) < 0.25, (k, k0)
