"""GPU tier: the World facade end to end (encode -> decode) against the reference goldens, with the
reference's noise stream replayed."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _reseed():
    np.random.seed(0)
    random.seed(0)
    from world_b200 import synthesisRequiem
    synthesisRequiem.generate_noise.current_index = None


def rms(a, b):
    return float(np.sqrt(np.mean((a - b) ** 2)))


def test_config1_mwm_dio_d4c_encode_decode(engine, mwm):
    """BASELINE config 1 / north_star: encode(dio, d4c) -> decode on test-mwm.wav within 1e-4 RMS."""
    from world_b200 import main
    W = main.World()
    g = mwm
    _reseed()
    dat = W.encode(int(g["fs"]), g["x"], f0_method="dio", is_requiem=False)
    assert set(dat) == {"temporal_positions", "vuv", "fs", "f0", "aperiodicity", "ps spectrogram", "spectrogram",
                        "is_requiem"}
    assert dat["spectrogram"].shape == (513, 929) and dat["ps spectrogram"].shape == (1024, 929)
    assert dat["spectrogram"].flags["C_CONTIGUOUS"] and dat["ps spectrogram"].dtype == np.complex128
    assert np.array_equal(dat["vuv"], g["dio_d4c_vuv"])
    v = g["dio_d4c_vuv"] > 0
    assert np.max(np.abs(dat["f0"][v] - g["dio_d4c_f0"][v]) / g["dio_d4c_f0"][v]) < 1e-6
    st = int(g["dio_d4c_frame_stride"])
    S, Sg = dat["spectrogram"][:, ::st], g["dio_d4c_spectrogram"]
    m = Sg > 1e-10
    d = np.abs(np.log10(S[m]) - np.log10(Sg[m]))
    assert np.percentile(d, 99) <= 1e-4 and d.max() <= 1e-3
    assert np.max(np.abs(dat["aperiodicity"][:, ::st] - g["dio_d4c_aperiodicity"])) <= 1e-5
    _reseed()
    out = W.decode(dat)
    assert out is dat and len(dat["out"]) == len(g["dio_d4c_out"])
    assert rms(dat["out"], g["dio_d4c_out"]) <= 1e-4
    print("config1 waveform RMS diff %.3e (signal RMS %.3e)" % (rms(dat["out"], g["dio_d4c_out"]),
                                                                 float(np.sqrt(np.mean(g["dio_d4c_out"] ** 2)))))


def test_prosody_path_mwm_harvest_requiem(engine, mwm):
    """example/prosody.py path: encode(harvest, requiem) -> decode on test-mwm.wav."""
    from world_b200 import main
    W = main.World()
    g = mwm
    _reseed()
    dat = W.encode(int(g["fs"]), g["x"], f0_method="harvest", is_requiem=True)
    assert dat["aperiodicity"].shape == (4, 929)
    assert np.array_equal(dat["vuv"], g["harvest_req_vuv"])
    st = int(g["harvest_req_frame_stride"])
    assert np.max(np.abs(dat["aperiodicity"][:, ::st] - g["harvest_req_aperiodicity"])) <= 1e-3
    _reseed()
    W.decode(dat)
    assert rms(dat["out"], g["harvest_req_out"]) <= 1e-4


def test_syn16k_all_flavours(engine, syn16k):
    from world_b200 import main
    W = main.World()
    g = syn16k
    for tag, method, req in (("harvest_d4c_", "harvest", False), ("harvest_req_", "harvest", True),
                             ("dio_d4c_", "dio", False)):
        _reseed()
        dat = W.encode(16000, g["x"], f0_method=method, is_requiem=req)
        assert np.array_equal(dat["vuv"], g[tag + "vuv"])
        _reseed()
        W.decode(dat)
        assert rms(dat["out"], g[tag + "out"]) <= 1e-4, tag


def test_stage_modules_and_errors(engine, syn16k):
    """Drop-in stage functions keep the reference's in-place behaviour; unknown methods raise."""
    from world_b200 import main
    from world_b200.cheaptrick import cheaptrick
    from world_b200.d4c import d4c
    from world_b200.harvest import harvest
    g = syn16k
    x = g["x"]
    src = harvest(x, 16000)
    assert np.array_equal(src["vuv"], g["harvest_d4c_vuv"])
    np.random.seed(0)
    flt = cheaptrick(x, 16000, src)
    assert np.allclose(src["f0"], g["harvest_d4c_f0_after_cheaptrick"], rtol=1e-9, atol=0)   # mutated in place
    out = d4c(x, 16000, src)
    assert out is src and np.allclose(src["f0"], g["harvest_d4c_f0"], rtol=1e-9, atol=0)
    assert np.array_equal(src["f0"] == 0, g["harvest_d4c_f0"] == 0)
    assert np.max(np.abs(src["aperiodicity"] - g["harvest_d4c_aperiodicity"])) < 1e-8
    with pytest.raises(Exception):
        main.World().encode(16000, x, f0_method="nope")


def test_batch_decode_device_noise(engine, syn16k):
    """encode_batch -> decode_batch (device noise): shapes, determinism per seed, finite output."""
    from world_b200 import main
    W = main.World()
    x = syn16k["x"]
    xs = np.stack([x, x[::-1].copy(), np.r_[x[:8000], np.zeros(8000)]])
    d = W.encode_batch(16000, xs, n_samples=[16000, 16000, 8000], f0_method="harvest")
    assert list(d["n_frames"].numpy()) == [201, 201, 101]
    dd = {k: (v.clone() if hasattr(v, "clone") else v) for k, v in d.items()}
    o1 = W.decode_batch(dd, seed=3)
    y1 = o1["out"].clone()
    o2 = W.decode_batch(dd, seed=3)
    assert list(o1["out_len"].numpy()) == [16001, 16001, 8001]
    assert np.isfinite(y1.numpy()).all()
    assert np.allclose(y1.numpy(), o2["out"].numpy(), atol=1e-12)


def test_48k_and_8k_full_path_vs_oracle(engine):
    """Other sampling rates through the whole facade, against the oracle on the same seeded input:
    48 kHz (config-5 shape: FFT 2048 / 4096, 5 aperiodicity bands, Harvest ratio 6, DIO ratio 12) and
    12 kHz (no Harvest decimation margin, 2 kHz band interval does not apply to requiem)."""
    from oracle import pipeline
    from oracle import synthesis as o_syn
    from world_b200 import main, synth_input
    W = main.World()
    for fs, secs, method, req in ((48000, 0.4, "harvest", False), (48000, 0.8, "dio", True), (12000, 0.5, "harvest", True)):
        x = synth_input.utterance(fs, secs, 5, 3)
        _reseed()
        dat = W.encode(fs, x, f0_method=method, is_requiem=req)
        _reseed()
        ref = pipeline.encode(fs, x, method, is_requiem=req)
        assert np.array_equal(dat["vuv"], ref["vuv"]), (fs, method)
        v = ref["vuv"] > 0
        if v.any():
            assert np.max(np.abs(dat["f0"][v] - ref["f0"][v]) / ref["f0"][v]) < 1e-6
        m = ref["spectrogram"] > 1e-10
        assert np.max(np.abs(np.log10(dat["spectrogram"][m]) - np.log10(ref["spectrogram"][m]))) < 1e-3
        assert dat["aperiodicity"].shape == ref["aperiodicity"].shape
        assert np.max(np.abs(dat["aperiodicity"] - ref["aperiodicity"])) < (1e-3 if req else 1e-5)
        _reseed()
        W.decode(dat)
        _reseed()
        y, _ = o_syn.decode(ref)
        assert len(dat["out"]) == len(y)
        assert rms(dat["out"], y) < 1e-4, (fs, method, req)


def test_prosody_edits_then_decode(engine, syn16k):
    """scale_pitch / scale_duration (host-side dict edits, main.py:154-177) feed decode like the reference's
    example/prosody.py; checked against the oracle decoding the same edited dict."""
    from oracle import synthesis as o_syn
    from world_b200 import main
    W = main.World()
    g = syn16k
    _reseed()
    dat = W.encode(16000, g["x"], f0_method="harvest", is_requiem=True)
    W.scale_pitch(dat, 1.5)
    W.scale_duration(dat, 2)
    import copy
    ref = copy.deepcopy(dat)
    _reseed()
    W.decode(dat)
    _reseed()
    y, _ = o_syn.decode(ref)
    assert len(dat["out"]) == len(y) and abs(len(y) - 32001) <= 1
    assert rms(dat["out"], y) < 1e-6


def test_device_resident_edit_decode(engine, syn16k):
    """encode_batch(device_resident) -> scale_pitch on the CUDA tensors -> decode_batch, no host round trip."""
    import torch
    from world_b200 import main
    W = main.World()
    x = syn16k["x"]
    d = W.encode_batch(16000, np.stack([x, x]), f0_method="harvest", is_requiem=True, device_resident=True)
    assert d["f0"].is_cuda and d["_d2h_bytes"] == 0
    f0_before = d["f0"].clone()
    W.scale_pitch(d, 1.25)
    assert torch.allclose(d["f0"], f0_before * 1.25)
    o = W.decode_batch(d)
    y = o["out"].numpy()
    assert list(o["out_len"].numpy()) == [16001, 16001]
    assert np.isfinite(y).all() and np.allclose(y[0], y[1])
