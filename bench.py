"""bench.py -- WORLD analysis frames/sec (16 kHz, 5 ms hop) on B200, BASELINE.json configs 2-5.

  python bench.py [--config 2|3|4|5] [--gpus N] [--steps K] [--warmup W]     our arm (CUDA, sm_100a)
  python bench.py --impl reference [...]                                       CPU arm on the host cores

  config 2 (default)  batch 256 x 16 kHz x 4 s, Harvest + CheapTrick + D4C            (the metric's configuration)
  config 3            256 per GPU x 16 kHz x 4 s, full encode with D4C-Requiem        (BASELINE: 2048 over 8 GPUs)
  config 4            decode only from precomputed features, batch 4096 (--flavour synthesis|requiem)
  config 5            128 per GPU x 48 kHz x 4 s, Harvest + CheapTrick (FFT 2048)     (BASELINE: 512 over 4 GPUs)

One JSON line on stdout (rank 0).  A "step" = one pass of the path over one batch (utterance-sharded for N > 1:
weak scaling, no data-path collective; `--gather` adds the NCCL gather of the decoded audio to config 4).
  value      frames/s with the inputs resident in HBM (CUDA events, L2 flushed between steps, max over ranks)
  e2e        frames/s through World().encode_batch() / decode_batch() with HOST buffers, H2D + D2H inside the timing
  roofline   slowest kernel of the step: FP64 flop/s (flop model of profiles/, measured duration, DFMA peak measured
             in this run) and algorithmic HBM GB/s (SURVEY 8d bytes per frame) -- `bound` names the binding one
  cpu_baseline  the unmodified reference (baseline/_ref, when shipped) on a bounded sample, plus the oracle port

The reference arm times the reference's own code (kind "reference": world/*.py copied unmodified to the git-ignored
baseline/_ref by __graft_entry__.build(), imported through the three shims of SURVEY 8c; Harvest fans out to all
host cores through its own multiprocessing.Pool) and falls back to the oracle port (kind "port") without it.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "python-world_b200"))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FRAME_PERIOD = 5.0
METRIC = "WORLD analysis frames/sec (16 kHz, 5 ms hop)"
REF_DIR = os.path.join(ROOT, "baseline", "_ref")

# bytes per frame: SURVEY 8d (inputs read once + API outputs written once at the reference's dtypes)
CONFIGS = {
    2: dict(fs=16000, seconds=4.0, batch=256, seed_config=2, kind="encode", requiem=False, aperiodicity="full",
            bytes=8872, bytes_ps=25256,
            workload="config2: batch=256 x 16 kHz 4 s synthetic, Harvest+CheapTrick+D4C"),
    3: dict(fs=16000, seconds=4.0, batch=256, seed_config=3, kind="encode", requiem=True, aperiodicity="full",
            bytes=4792, bytes_ps=21176,
            workload="config3: 256 per GPU x 16 kHz 4 s synthetic, Harvest+CheapTrick+D4C-Requiem"),
    4: dict(fs=16000, seconds=4.0, batch=4096, seed_config=2, kind="decode", bytes=8872, bytes_requiem=4792,
            workload="config4: decode only, batch=4096 x 801 frames (16 kHz), precomputed features"),
    5: dict(fs=48000, seconds=4.0, batch=128, seed_config=5, kind="encode", requiem=False, aperiodicity="none",
            bytes=10144, bytes_ps=42912,
            workload="config5: 128 per GPU x 48 kHz 4 s synthetic (FFT 2048), Harvest+CheapTrick"),
}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def config_dict(cfg, args, world, batch):
    """Identical for both arms (the driver compares it)."""
    c = {"workload": cfg["workload"], "config": args.config, "batch_per_gpu": batch,
         "frames_per_gpu": batch * frames_of(cfg), "l2": "256 MiB flush buffer written between timed iterations",
         "parallelism": "utterance-sharded, %d GPU(s), no data-path collective" % world}
    if cfg["kind"] == "decode":
        c["flavour"] = args.flavour
    return c


def frames_of(cfg):
    return int(1000 * int(round(cfg["fs"] * cfg["seconds"])) / cfg["fs"] / FRAME_PERIOD + 1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop = False
        self.max_mhz = None
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [s.strip() for s in out.strip().split(",")]
                self.samples.append(float(parts[0]))
                self.max_mhz = float(parts[1])
                for n, v in zip(names, parts[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def make_inputs(cfg, rank, batch):
    from world_b200 import synth_input
    return synth_input.batch(cfg["fs"], cfg["seconds"], cfg["seed_config"], batch, first=rank * batch)


# ------------------------------------------------------------------------------------------- CPU arms
def load_reference():
    """The UNMODIFIED reference package from baseline/_ref (None when it was not shipped)."""
    if not os.path.isdir(os.path.join(REF_DIR, "world")):
        return None
    os.environ["WORLD_REFERENCE_ROOT"] = REF_DIR
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import importlib
    import refload
    refload.load()
    return importlib.import_module("refworld.main").World(), refload


def reference_step(W, refload, cfg, x, flavour, dat_cache):
    """One utterance through the reference's own public API for this config; returns (frames, seconds)."""
    fs = cfg["fs"]
    refload.reseed(0)
    if cfg["kind"] == "decode":
        import copy
        key = flavour
        if key not in dat_cache:
            dat_cache[key] = W.encode(fs, np.array(x), f0_method="harvest", is_requiem=(flavour == "requiem"))
        dat = copy.deepcopy(dat_cache[key])
        refload.reseed(0)
        t0 = time.perf_counter()
        W.decode(dat)
        return len(dat["f0"]), time.perf_counter() - t0
    t0 = time.perf_counter()
    if cfg["aperiodicity"] == "none":
        d = W.get_spectrum(fs, np.array(x), f0_method="harvest")
    else:
        d = W.encode(fs, np.array(x), f0_method="harvest", is_requiem=cfg["requiem"])
    return len(d["f0"]), time.perf_counter() - t0


def port_step(cfg, x, flavour, dat_cache):
    """The same through the oracle port (NumPy restatement of the reference's algorithm), one process."""
    from oracle import pipeline
    fs = cfg["fs"]
    np.random.seed(0)
    if cfg["kind"] == "decode":
        from oracle import synthesis as o_syn
        key = "port_" + flavour
        if key not in dat_cache:
            dat_cache[key] = pipeline.encode(fs, x, "harvest", is_requiem=(flavour == "requiem"))
        import random
        np.random.seed(0)
        random.seed(0)
        t0 = time.perf_counter()
        o_syn.decode(dict(dat_cache[key]))
        return len(dat_cache[key]["f0"]), time.perf_counter() - t0
    t0 = time.perf_counter()
    if cfg["aperiodicity"] == "none":
        from oracle import cheaptrick as o_ct, harvest as o_hv
        src = o_hv.harvest(x, fs, 71, 800, 5)
        o_ct.cheaptrick(x, fs, src["temporal_positions"], src["f0"], src["vuv"])
        n = len(src["f0"])
    else:
        n = len(pipeline.encode(fs, x, "harvest", is_requiem=cfg["requiem"])["f0"])
    return n, time.perf_counter() - t0


def cpu_baseline(cfg, x_one, flavour):
    """Bounded sample on the host cores: one utterance through the reference itself (when shipped) and through
    the oracle port."""
    cache = {}
    cores = os.cpu_count() or 1
    out = None
    ref = load_reference()
    n, dt = port_step(cfg, x_one, flavour, cache)
    port = {"value": n / dt, "unit": "frames/s", "cores": 1, "kind": "port",
            "sample": "1 utterance (%d frames) of the workload, oracle port of the reference, 1 process" % n}
    if ref is not None:
        W, refload = ref
        reference_step(W, refload, cfg, x_one, flavour, cache)  # numba warm-up
        n, dt = reference_step(W, refload, cfg, x_one, flavour, cache)
        out = {"value": n / dt, "unit": "frames/s", "cores": cores if cfg["kind"] == "encode" else 1, "kind": "reference",
               "sample": "1 utterance (%d frames) of the workload through the unmodified reference (baseline/_ref, "
                         "main.World() public API; Harvest's own Pool uses all %d host cores, the other stages one), "
                         "numba warmed by one untimed run" % (n, cores),
               "port": port}
    return out or port


def run_reference(args, cfg):
    """--impl reference: K timed steps of one utterance each through the reference's own code on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    batch = args.batch or cfg["batch"]
    warm = max(3, args.warmup)
    xs = make_inputs(cfg, 0, min(batch, warm + args.steps))
    ref = load_reference()
    cache = {}
    cores = os.cpu_count() or 1
    if ref is not None:
        W, refload = ref
        step = lambda i: reference_step(W, refload, cfg, xs[i % len(xs)], args.flavour, cache)
        kind = "reference"
        used = cores if cfg["kind"] == "encode" else 1
        how = ("the unmodified reference (baseline/_ref) through main.World(); Harvest's own multiprocessing.Pool "
               "uses all %d host cores, every other stage one" % cores)
    else:
        step = lambda i: port_step(cfg, xs[i % len(xs)], args.flavour, cache)
        kind, used = "port", 1
        how = "oracle port of the reference (baseline/_ref not shipped), 1 process"
    for i in range(warm):
        step(i)
    frames = 0
    t0 = time.perf_counter()
    for i in range(args.steps):
        n, _ = step(warm + i)
        frames += n
    dt = time.perf_counter() - t0
    val = frames / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": warm, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(cfg, args, world, batch),
            "cpu_baseline": {"value": val, "unit": "frames/s", "cores": used, "kind": kind,
                             "sample": "each step = 1 utterance (%d frames) of the workload; %s" % (frames // args.steps, how)},
            "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- rooflines
def flop_model():
    """FP64 operations per frame of every kernel, from the committed ncu capture of this workload
    (profiles/r02_kernels.json: dadd + dmul + 2 dfma thread instructions per launch / frames of the launch)."""
    for name in ("r02_kernels.json", "r01_kernels.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return json.load(f), "profiles/" + name
        except Exception:
            continue
    return {}, None


def dfma_peak(E, torch):
    """FP64 peak of this GPU, measured now: best of 5 launches of the library's DFMA probe, CUDA events."""
    import ctypes
    out = E.empty(8)
    threads = 148 * 8 * 256 * 4
    iters = 4096
    flops = ctypes.c_double()
    best = 0.0
    for i in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        E._check(E.L.wb_probe_dfma(E.h, E._stream(), threads, iters, E.ptr(out), ctypes.byref(flops)))
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            best = max(best, flops.value / (a.elapsed_time(b) * 1e-3) / 1e12)
    return best


def roofline(stage_ms, frames, cfg_bytes, model, model_src, tag_map, peak_fp64):
    peak_hbm, peak_kind = peaks()
    top = max(stage_ms, key=lambda k: stage_ms[k])
    ms = stage_ms[top]
    hbm = frames * cfg_bytes / (ms / 1e3) / 1e9
    m = model.get(tag_map.get(top, top)) or {}
    fpf = m.get("fp64_flop_per_frame")
    r = {"kernel": top, "kernel_ms": ms, "stage_ms": stage_ms,
         "hbm": {"achieved": hbm, "peak": peak_hbm, "unit": "GB/s", "frac": hbm / peak_hbm, "peak_kind": peak_kind,
                 "note": "algorithmic bytes = %d B/frame (SURVEY 8d) x frames / duration of the slowest kernel" % cfg_bytes},
         "traffic": m.get("dram_bytes"),
         "traffic_source": (model_src + " (ncu --set full of the same workload and batch; not re-measured in this run)") if m else None}
    if fpf and peak_fp64:
        tf = fpf * frames / (ms / 1e3) / 1e12
        r["fp64"] = {"achieved": tf, "peak": peak_fp64, "unit": "TFLOP/s", "frac": tf / peak_fp64,
                     "flop_per_frame": fpf, "flop_source": model_src + " (dadd + dmul + 2 dfma per launch / frames)",
                     "peak_kind": "measured in this run (wb_probe_dfma: 8 independent DFMA chains per thread)"}
        r.update({"bound": "fp64", "achieved": tf, "peak": peak_fp64, "unit": "TFLOP/s", "frac": tf / peak_fp64,
                  "fp64_frac": tf / peak_fp64, "hbm_frac": hbm / peak_hbm})
    else:
        r.update({"bound": "hbm", "achieved": hbm, "peak": peak_hbm, "unit": "GB/s", "frac": hbm / peak_hbm,
                  "hbm_frac": hbm / peak_hbm})
    return r


# ------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--flavour", default="synthesis", choices=["synthesis", "requiem"], help="config 4: which decoder")
    ap.add_argument("--batch", type=int, default=0, help="utterances per GPU (default: the config's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e-variants", action="store_true", help="skip the extra e2e legs (full aperiodicity, ps)")
    ap.add_argument("--gather", action="store_true", help="config 4: add the NCCL gather of the decoded audio")
    ap.add_argument("--streams", type=int, default=2, help="CUDA streams the batch is split over inside encode()")
    ap.add_argument("--pipeline", type=int, default=16, help="parts encode_batch() pipelines H2D / kernels / D2H over")
    ap.add_argument("--e2e-chunk", type=int, default=512, help="config 4: utterances per decode_batch() call")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        return run_reference(args, cfg)

    import torch
    import torch.distributed as dist
    from world_b200 import engine as eng, main as wmain, numa

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa_note = numa.bind_to_gpu(local)  # pinned staging buffers land on the GPU's NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    E = eng.default_engine(local)
    W = wmain.World()
    warm = max(3, args.warmup)
    batch = args.batch or cfg["batch"]
    fs = cfg["fs"]
    F = frames_of(cfg)
    frames_rank = batch * F
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=E.device)  # > 126 MB of L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    extra = {}
    if cfg["kind"] == "encode":
        xs = make_inputs(cfg, rank, batch)
        X = E.f64(xs)
        ns = E.i32([xs.shape[1]] * batch)
        enc_kw = dict(f0_method="harvest", is_requiem=cfg["requiem"], aperiodicity=cfg["aperiodicity"])

        def step_resident():
            return E.encode(X, ns, fs, streams=args.streams, **enc_kw)

        launches = E.launches_per_encode("harvest", cfg["requiem"]) - (1 if cfg["aperiodicity"] == "none" else 0)
        gpu_launches = int(launches) * max(1, args.streams)
        x_one = xs[0]
    else:
        # precomputed features: 32 synthetic utterances encoded once, tiled to the batch
        src = make_inputs(cfg, rank, 32)
        req = args.flavour == "requiem"
        d0 = E.encode(E.f64(src), E.i32([src.shape[1]] * 32), fs, f0_method="harvest", is_requiem=req)
        torch.cuda.synchronize()
        reps = (batch + 31) // 32
        feats = {k: d0[k].repeat((reps,) + (1,) * (d0[k].dim() - 1))[:batch].contiguous()
                 for k in ("temporal_positions", "f0", "vuv", "spectrogram", "aperiodicity", "n_frames")}
        del d0
        ylen = E.synthesis_length(0.0, (F - 1) * FRAME_PERIOD / 1000.0, fs)
        seeds = None
        if req:
            from world_b200.get_seeds_signals import get_seeds_signals
            np.random.seed(0)
            sd = get_seeds_signals(fs)
            seeds = (E.f64(sd["pulse"]), E.f64(sd["noise"]))

        def step_resident():
            return E.decode(feats["temporal_positions"], feats["f0"], feats["vuv"], feats["spectrogram"],
                            feats["aperiodicity"], feats["n_frames"], fs, ylen, is_requiem=req, seeds=seeds, seed=1)

        gpu_launches = E.launches_per_decode(req)
        x_one = src[0]

    for _ in range(warm):
        step_resident()
    barrier()
    with ClockSampler(local) as clk:
        evs = []
        barrier()
        for _ in range(args.steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step_resident()
            b.record()
            evs.append((a, b))
        barrier()
        total_ms = sum(a.elapsed_time(b) for a, b in evs)

        # per-kernel durations (same stream, same inputs) for the roofline
        if cfg["kind"] == "encode":
            stage_ms = E.profile_stages(X, ns, fs, f0_method="harvest", is_requiem=cfg["requiem"],
                                        iters=max(2, min(args.steps, 5)), with_d4c=cfg["aperiodicity"] != "none")
        else:
            stage_ms = E.profile_decode(feats, fs, ylen, is_requiem=req, seeds=seeds, iters=max(2, min(args.steps, 5)))
        peak_fp64 = dfma_peak(E, torch)

        # end to end through the public batch API with host buffers
        e2e = {}
        if cfg["kind"] == "encode":
            xs_pinned = torch.from_numpy(xs).pin_memory()
            legs = [("default", dict(aperiodicity="coarse" if cfg["aperiodicity"] == "full" and not cfg["requiem"] else cfg["aperiodicity"]))]
            if not args.no_e2e_variants:
                # opt-in lossy transport: the spectrogram crosses PCIe as float32 (6e-8 relative rounding)
                legs.append(("f32_spectrogram", dict(aperiodicity=legs[0][1]["aperiodicity"], spectrogram_dtype=torch.float32)))
            if not args.no_e2e_variants and world == 1:
                if cfg["aperiodicity"] == "full" and not cfg["requiem"]:
                    legs.append(("full_aperiodicity", dict(aperiodicity="full")))
                legs.append(("with_ps_spectrogram", dict(aperiodicity=legs[0][1]["aperiodicity"], want_ps=True)))
            for name, kw in legs:
                call = lambda: W.encode_batch(fs, xs_pinned, f0_method="harvest", is_requiem=cfg["requiem"],
                                              pipeline=args.pipeline, **kw)
                n_steps = args.steps if name == "default" else min(args.steps, 2)
                for _ in range(2 if name == "default" else 1):
                    call()
                barrier()
                t0 = time.perf_counter()
                for _ in range(n_steps):
                    out = call()
                barrier()
                e2e[name] = ((time.perf_counter() - t0) / n_steps, out["_h2d_bytes"], out["_d2h_bytes"])
                W._pinned.clear()  # the legs use differently shaped staging buffers
        else:
            chunk = min(args.e2e_chunk, batch)
            keys = ("temporal_positions", "f0", "vuv", "spectrogram", "n_frames") + (("aperiodicity",) if req else ("coarse_ap",))
            if not req:  # the compact transport form of the aperiodicity (what encode_batch hands back by default)
                c0 = E.encode(E.f64(src), E.i32([src.shape[1]] * 32), fs, f0_method="harvest", aperiodicity="coarse")
                feats["coarse_ap"] = c0["coarse_ap"].repeat((reps, 1, 1))[:batch].contiguous()
            host = {k: feats[k][:chunk].cpu().pin_memory() for k in keys}
            host.update(fs=fs, is_requiem=req)
            n_calls = (batch + chunk - 1) // chunk

            def call():
                h2d = d2h = 0
                for _ in range(n_calls):
                    o = W.decode_batch(dict(host), seed=1)
                    h2d += o["_h2d_bytes"]
                    d2h += o["_d2h_bytes"]
                return h2d, d2h
            call()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                h2d, d2h = call()
            barrier()
            e2e["default"] = ((time.perf_counter() - t0) / args.steps, h2d, d2h)

        # config 4 --gather: the one collective north_star names (decoded audio to every rank)
        if args.gather and cfg["kind"] == "decode" and world > 1:
            from world_b200 import distributed as wdist
            y, out_len, _ = step_resident()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for i in range(3):
                if i == 1:
                    a.record()
                g, gl = wdist.gather_padded(y, out_len)
            b.record()
            barrier()
            g_ms = a.elapsed_time(b) / 2
            recv = (world - 1) * y.numel() * y.element_size()
            extra["gather"] = {"ms": g_ms, "bytes_received_per_rank": recv, "bus_gbs": recv / (g_ms / 1e3) / 1e9,
                               "collective": "all_gather of padded [B, S] float64 rows + int32 lengths (NCCL)"}

    vals = [total_ms] + [e2e[k][0] * 1e3 for k in sorted(e2e)]
    tmax = torch.tensor(vals, dtype=torch.float64, device=E.device)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms = float(tmax[0])
    e2e_ms = {k: float(tmax[1 + i]) for i, k in enumerate(sorted(e2e))}
    ms_per_step = total_ms / args.steps
    value = frames_rank * world / (ms_per_step / 1e3)

    if rank == 0:
        model, model_src = flop_model()
        cfg_bytes = cfg["bytes_requiem"] if cfg["kind"] == "decode" and args.flavour == "requiem" else cfg["bytes"]
        tag_map = {}  # stage keys of profile_stages / profile_decode are the keys of the model file
        model_key = "config%d%s" % (args.config, "r" if cfg["kind"] == "decode" and args.flavour == "requiem" else "")
        model_cfg = model.get(model_key, {})
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(cfg, args, world, batch),
            "clocks": clk.summary(),
            "e2e": {"value": frames_rank * world / (e2e_ms["default"] / 1e3), "unit": "frames/s",
                    "h2d_bytes_per_step": int(e2e["default"][1]), "d2h_bytes_per_step": int(e2e["default"][2]),
                    "api": "World.encode_batch (pinned host buffers, %d pipelined parts; aperiodicity travels as "
                           "'coarse_ap' and is rebuilt on first access)" % args.pipeline
                    if cfg["kind"] == "encode" else "World.decode_batch (pinned host buffers, %d utterances per call)" % min(args.e2e_chunk, batch)},
            "gpu_launches": gpu_launches * args.steps,
            "roofline": roofline(stage_ms, frames_rank, cfg_bytes, model_cfg, model_src, tag_map, peak_fp64),
            "host": {"cores": os.cpu_count(), "numa": numa_note},
        }
        for k in e2e_ms:
            if k != "default":
                line.setdefault("e2e_variants", {})[k] = {
                    "value": frames_rank * world / (e2e_ms[k] / 1e3), "unit": "frames/s",
                    "h2d_bytes_per_step": int(e2e[k][1]), "d2h_bytes_per_step": int(e2e[k][2])}
        line.update(extra)
        if not args.no_cpu_baseline and world == 1:  # a reported baseline, timed at N = 1 only
            line["cpu_baseline"] = cpu_baseline(cfg, x_one, args.flavour)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
