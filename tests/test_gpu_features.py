"""GPU tier: feature heads, spectrum / time-axis edits and the PCM edge through the C-ABI against the reference
goldens (tests/golden/features.npz) -- SURVEY 8f rows 1, 3, 4."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import features as o_ft

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def feat():
    return dict(np.load(os.path.join(GOLDEN, "features.npz")))


def test_feature_heads_facade_numpy(engine, feat):
    from world_b200 import main
    W = main.World()
    for tag, fs in (("k16_", 16000), ("k48_", 48000)):
        mag = feat[tag + "spec"]
        assert np.max(np.abs(W.encode_lfbank(mag, fs=fs) - feat[tag + "lfbank"])) < TOL
        got = W.encode_lfbank(mag, prefac=0.9, fs=fs, nfilt=20, lowfreq=100, highfreq=fs / 2 - 500)
        assert np.max(np.abs(got - feat[tag + "lfbank_20_hi"])) < TOL
        m40 = W.encode_mcep(mag, n0=40, fs=fs, highhz=min(8000, fs // 2))
        assert m40.shape == feat[tag + "mcep40"].shape and np.max(np.abs(m40 - feat[tag + "mcep40"])) < TOL
        assert np.max(np.abs(W.encode_mcep(mag, fs=fs) - feat[tag + "mcep12"])) < TOL
        dec = W.decode_mcep(feat[tag + "mcep40"], fft_size=(mag.shape[1] - 1) * 2)
        assert np.max(np.abs(np.log(dec) - np.log(feat[tag + "decoded"]))) < TOL
        for f in (0.8, 1.25):
            d = {"spectrogram": np.ascontiguousarray(mag.T)}
            W.warp_spectrum(d, f)
            assert np.array_equal(d["spectrogram"], feat[tag + "warp_%g" % f])
    d = {"temporal_positions": feat["dur_tp"].copy()}
    W.modify_duration(d, [1, 1.5], [0, 1, 3, -1])
    assert np.array_equal(d["temporal_positions"], feat["dur_out"])
    assert np.array_equal(W.get_filterbanks(26, 1024, 16000), o_ft.filterbank(26, 1024, 16000))


def test_heads_on_resident_batch(engine, syn16k):
    """encode_batch(device_resident=True) -> heads / edits on the CUDA tensors, checked against the oracle run on
    the downloaded spectrogram; nothing but the features leaves HBM."""
    import torch
    from world_b200 import main
    W = main.World()
    x = syn16k["x"]
    xs = np.stack([x, x[::-1].copy(), np.r_[x[:12000], np.zeros(4000)]])
    d = W.encode_batch(16000, xs, n_samples=[16000, 16000, 12000], f0_method="harvest", device_resident=True)
    sp = d["spectrogram"]
    assert sp.is_cuda and sp.dim() == 3
    mag = sp.sqrt()
    lf = W.encode_lfbank(mag, fs=16000)
    mc = W.encode_mcep(mag, n0=24, fs=16000)
    assert lf.is_cuda and tuple(lf.shape) == (3, sp.shape[1], 32) and tuple(mc.shape) == (3, sp.shape[1], 24)
    mag_h = mag.cpu().numpy()
    nf = d["n_frames"].cpu().numpy()
    assert list(nf) == [201, 201, 151]
    for u in range(3):  # frames past n_frames[u] are padding (never written by the analysis kernels)
        k = nf[u]
        assert np.max(np.abs(lf[u, :k].cpu().numpy() - o_ft.encode_lfbank(mag_h[u, :k], fs=16000))) < TOL
        assert np.max(np.abs(mc[u, :k].cpu().numpy() - o_ft.encode_mcep(mag_h[u, :k], n0=24, fs=16000))) < TOL
    dec = W.decode_mcep(mc, fft_size=1024)
    assert np.max(np.abs(np.log(dec[1].cpu().numpy()) - np.log(o_ft.decode_mcep(mc[1].cpu().numpy(), 1024)))) < TOL
    before = sp.cpu().numpy().copy()
    W.warp_spectrum(d, 1.1)
    for u in range(3):
        k = nf[u]
        assert np.array_equal(d["spectrogram"][u, :k].cpu().numpy().T, o_ft.warp_spectrum(before[u, :k].T, 1.1))
    tp0 = d["temporal_positions"].cpu().numpy().copy()
    W.modify_duration(d, [0.2, 0.4], [0, 0.3, 0.5, -1])
    for u in range(3):
        want = o_ft.modify_duration(tp0[u, :nf[u]], [0.2, 0.4], [0, 0.3, 0.5, -1])
        assert np.array_equal(d["temporal_positions"][u, :nf[u]].cpu().numpy(), want)
    out = W.decode_batch(d)  # the edited batch still decodes
    for u in range(3):
        assert bool(torch.isfinite(out["out"][u, :int(out["out_len"][u])]).all())


def test_pcm_edge(engine, mwm):
    """int16 in: x = pcm / 32767 on the device gives the same analysis as the float64 input; int16 out matches
    (out * 2**15).astype(int16)."""
    import torch
    from world_b200 import main
    W = main.World()
    fs = int(mwm["fs"])
    pcm = mwm["x_int16"][None, :40000]
    a = W.encode_batch(fs, pcm, f0_method="dio")
    fa, sa = a["f0"].clone(), a["spectrogram"].clone()
    b = W.encode_batch(fs, pcm / 32767.0, f0_method="dio")
    assert a["_h2d_bytes"] * 4 - 12 == b["_h2d_bytes"] and torch.equal(fa, b["f0"]) and torch.equal(sa, b["spectrogram"])
    y = W.decode_batch(dict(b), seed=3)
    yf = y["out"].clone()
    n = int(y["out_len"][0])
    yi = W.decode_batch(dict(b), seed=3, pcm16=True)
    assert yi["out"].dtype == torch.int16 and yi["_d2h_bytes"] * 4 == y["_d2h_bytes"]
    assert np.array_equal(yi["out"][0, :n].numpy(), o_ft.float_to_pcm16(yf[0, :n].numpy()))
    edge = engine.f64([[1.0, -1.0, 0.99997, 1.5, -0.25, 3.0e-5]])
    assert np.array_equal(engine.f64_to_pcm16(edge, engine.i32([6])).cpu().numpy()[0],
                          o_ft.float_to_pcm16(edge.cpu().numpy()[0]))
