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
