"""GPU tier: BASELINE config 2 at full size (batch 256 x 16 kHz x 4 s, Harvest + CheapTrick + D4C) through
size-independent properties: batch-position independence (identical inputs at different batch positions give
bit-identical results), agreement with the single-utterance call, run-to-run determinism, value ranges."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_config2_full_size_properties(engine):
    import torch
    from world_b200 import synth_input
    fs, uniq, batch = 16000, 32, 256
    xs = synth_input.batch(fs, 4.0, 2, uniq)
    x = np.concatenate([xs] * (batch // uniq), axis=0)
    X = engine.f64(x)
    ns = engine.i32([x.shape[1]] * batch)
    d = engine.encode(X, ns, fs, f0_method="harvest", is_requiem=False, streams=2)
    torch.cuda.synchronize()
    F = 801
    assert d["f0"].shape == (batch, F) and d["spectrogram"].shape == (batch, F, 513)
    assert int(d["n_frames"].min()) == F and int(d["n_frames"].max()) == F
    f0, vuv = d["f0"].cpu().numpy(), d["vuv"].cpu().numpy()
    spec, ap = d["spectrogram"], d["aperiodicity"]
    # ranges
    assert set(np.unique(vuv)) <= {0.0, 1.0}
    v = vuv > 0
    assert np.all(f0[~v] == 0) and np.all((f0[v] >= 71 * 0.5) & (f0[v] <= 800 * 1.2))
    assert 0.4 < v.mean() < 0.9
    assert bool(torch.isfinite(spec).all()) and bool((spec > 0).all())
    assert bool(torch.isfinite(ap).all()) and float(ap.min()) >= 0.0 and float(ap.max()) <= 1.0 + 1e-12  # 10**(v/20) with the interpolated v within an ulp of -1e-12 at fs/2
    tp = d["temporal_positions"].cpu().numpy()
    assert np.allclose(tp[0], np.arange(F) * 0.005) and np.all(tp == tp[0])
    # batch-position independence: utterance k and k + 32 m are the same signal
    for m in range(1, batch // uniq):
        sl = slice(m * uniq, (m + 1) * uniq)
        assert np.array_equal(f0[sl], f0[:uniq]) and np.array_equal(vuv[sl], vuv[:uniq])
        # CheapTrick's eps-dither (cheaptrick.py:117: an ABSOLUTE 2e-16 x rand, i.e. 4e-8 relative on the smallest
        # bins, 2e-8) is keyed on the frame's position in the batch; D4C has no dither and must be bit-identical
        assert torch.allclose(spec[sl], spec[:uniq], rtol=1e-6, atol=0) and torch.equal(ap[sl], ap[:uniq])
    # the batch agrees with a single-utterance call
    d1 = engine.encode(X[5:6].contiguous(), ns[5:6].contiguous(), fs, f0_method="harvest", is_requiem=False)
    assert np.array_equal(d1["vuv"].cpu().numpy()[0], vuv[5])
    assert torch.allclose(d1["f0"][0], d["f0"][5], rtol=1e-12, atol=0)
    assert torch.allclose(d1["spectrogram"][0].log(), spec[5].log(), rtol=0, atol=1e-6)  # position-keyed dither, see above
    # run-to-run determinism of the whole batch
    d2 = engine.encode(X, ns, fs, f0_method="harvest", is_requiem=False, streams=2)
    assert torch.equal(d2["f0"], d["f0"]) and torch.equal(d2["spectrogram"], spec) and torch.equal(d2["aperiodicity"], ap)


def _tile(t, rep):
    return t.repeat((rep,) + (1,) * (t.dim() - 1)).contiguous()


def test_config3_shard_full_size_properties(engine):
    """BASELINE config 3, one GPU's shard (256 of the 2048 utterances; the shards are independent): full encode
    with is_requiem=True.  Batch-position independence, agreement with a small batch, value ranges."""
    import torch
    from world_b200 import synth_input
    fs, uniq, batch = 16000, 32, 256
    xs = synth_input.batch(fs, 4.0, 3, uniq)
    x = np.concatenate([xs] * (batch // uniq), axis=0)
    X, ns = engine.f64(x), engine.i32([x.shape[1]] * batch)
    d = engine.encode(X, ns, fs, f0_method="harvest", is_requiem=True, streams=2)
    torch.cuda.synchronize()
    ap, f0 = d["aperiodicity"], d["f0"]
    assert tuple(ap.shape) == (batch, 801, 3) and tuple(d["spectrogram"].shape) == (batch, 801, 513)
    assert bool(torch.isfinite(ap).all()) and float(ap.max()) <= 0.0 and float(ap.min()) >= -60.0
    assert bool((ap[..., 0] == -60.0).logical_or(ap[..., 0] == -1e-12).all()) and bool((ap[..., 2] == -1e-12).all())
    for m in range(1, batch // uniq):
        sl = slice(m * uniq, (m + 1) * uniq)
        assert torch.equal(f0[sl], f0[:uniq]) and torch.equal(ap[sl], ap[:uniq])
    d1 = engine.encode(X[:4].contiguous(), ns[:4].contiguous(), fs, f0_method="harvest", is_requiem=True)
    assert torch.equal(d1["f0"], f0[:4]) and torch.equal(d1["aperiodicity"], ap[:4])


def test_config4_full_size_decode_properties(engine):
    """BASELINE config 4: synthesis only from precomputed features, batch 4096 on one GPU, both flavours.  The batch
    is 128 copies of 32 utterances: every copy must give the same waveform (requiem: same seeds signals; synthesis:
    same noise stream when the copies are decoded as their own batch with the same seed), lengths
    len(arange(0, t_end + 1/fs, 1/fs)) = 64001, peak <= 1 after World.decode's rescale."""
    import torch
    from world_b200 import get_seeds_signals, synth_input
    fs, uniq, batch = 16000, 32, 4096
    xs = synth_input.batch(fs, 4.0, 4, uniq)
    X, ns = engine.f64(xs), engine.i32([xs.shape[1]] * uniq)
    for req in (True, False):
        d = engine.encode(X, ns, fs, f0_method="harvest", is_requiem=req)
        keys = ("temporal_positions", "f0", "vuv", "spectrogram", "aperiodicity", "n_frames")
        small = [d[k] for k in keys]
        big = [_tile(t, batch // uniq) for t in small]
        ylen = engine.synthesis_length(0.0, float(small[0][0, -1]), fs)
        assert ylen == 64001
        if req:
            sd = get_seeds_signals.get_seeds_signals(fs)
            P, N = engine.f64(sd["pulse"]), engine.f64(sd["noise"])
            run = lambda a: engine.synthesis_requiem(a[0], a[1], a[2], a[3], a[4], a[5], fs, ylen, P, N)[:2]
        else:
            run = lambda a: engine.synthesis(a[0], a[1], a[2], a[3], a[4], a[5], fs, ylen, noise="device", seed=11)
        y, out_len = run(big)
        torch.cuda.synchronize()
        assert tuple(y.shape) == (batch, ylen) and bool((out_len == ylen).all())
        assert bool(torch.isfinite(y).all()) and float(y.abs().max()) <= 1.0 + 1e-12
        rms = y.pow(2).mean(dim=1).sqrt()
        assert float(rms.min()) > 1e-3
        y0, _ = run(small)
        if req:  # deterministic excitation: every copy equals the small batch (overlap-add order aside)
            for m in range(0, batch // uniq, 16):
                assert torch.allclose(y[m * uniq:(m + 1) * uniq], y0, rtol=0, atol=1e-11)
        else:    # the device noise generator is keyed on (seed, utterance index): the first copy replays the small batch
            assert torch.allclose(y[:uniq], y0, rtol=0, atol=1e-11)
            r = rms.view(batch // uniq, uniq)
            assert float(((r - r[0]).abs() / r[0]).max()) < 0.05  # other copies: same signal, different noise draw
        del y, big
        torch.cuda.empty_cache()


def test_config5_shard_full_size_properties(engine):
    """BASELINE config 5, one GPU's shard (128 of the 512 utterances): fs = 48 kHz, Harvest + CheapTrick (FFT 2048)."""
    import torch
    from world_b200 import synth_input
    fs, uniq, batch = 48000, 16, 128
    xs = synth_input.batch(fs, 4.0, 5, uniq)
    x = np.concatenate([xs] * (batch // uniq), axis=0)
    X, ns = engine.f64(x), engine.i32([x.shape[1]] * batch)
    tp, f0, vuv, nf = engine.harvest(X, ns, fs)
    f0u, spec, _ = engine.cheaptrick(X, ns, fs, tp, f0, vuv, nf)
    torch.cuda.synchronize()
    assert tuple(f0.shape) == (batch, 801) and tuple(spec.shape) == (batch, 801, 1025)
    assert bool(torch.isfinite(spec).all()) and bool((spec > 0).all())
    v = vuv > 0
    assert 0.4 < float(v.double().mean()) < 0.9 and bool((f0[~v] == 0).all())
    assert bool(((f0u == 500.0) | (f0u >= 3.0 * fs / (2048 - 3.0))).all())  # cheaptrick.py:24-33
    for m in range(1, batch // uniq):
        sl = slice(m * uniq, (m + 1) * uniq)
        assert torch.equal(f0[sl], f0[:uniq]) and torch.allclose(spec[sl], spec[:uniq], rtol=1e-6, atol=0)
    t1, g1, v1, n1 = engine.harvest(X[3:5].contiguous(), ns[3:5].contiguous(), fs)
    assert torch.equal(g1, f0[3:5]) and torch.equal(v1, vuv[3:5])
