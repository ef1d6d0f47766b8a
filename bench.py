"""bench.py -- WORLD analysis frames/sec (16 kHz, 5 ms hop), BASELINE.json config 2:
batch=256 synthetic 16 kHz 4 s utterances, Harvest + CheapTrick + D4C on 1 x B200
(utterance-sharded, 256 per GPU, for --gpus N > 1: weak scaling, no data-path collective).

  python bench.py [--gpus N] [--steps K] [--warmup W]           our arm (CUDA, sm_100a)
  python bench.py --impl reference [...]                        CPU arm: the oracle port of the
        reference algorithm on the host cores (the reference is pure Python and /root/reference
        does not exist on the GPU box; kind="port")

One JSON line on stdout (rank 0).  A "step" = one full analysis pass over one batch.
  value      frames/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e        frames/s through World().encode_batch() with HOST buffers, H2D + D2H inside the timing
  roofline   dominant kernel (hv_refine ... see DESIGN.md) algorithmic bytes / measured duration
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

FS = 16000
SECONDS = 4.0
BATCH = 256
FRAME_PERIOD = 5.0
BYTES_PER_FRAME = 25256          # SURVEY 8d, config 2 (with the 'ps spectrogram' key)
BYTES_PER_FRAME_NO_PS = 8872     # what encode_batch moves by default (ps spectrogram is optional)
WORKLOAD = "config2: batch=256 x 16 kHz 4 s synthetic, Harvest+CheapTrick+D4C"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


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


def kernel_profile(top, batch):
    """What the committed `ncu --set full` capture of this workload (profiles/r01_kernels.json, tools/ncu_kernels_json.py)
    says about the dominant kernel: DRAM bytes per launch (roofline.traffic) and how busy the FP64 pipe / issue slots
    were -- the path is FP64-latency / issue bound, not HBM bound."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_kernels.json")) as f:
            prof = json.load(f)
    except Exception:
        return None, None
    k = prof.get(top) or prof.get({"hv_channels": "hv_channels_fft"}.get(top, top))
    if not k or batch != BATCH:
        return None, None
    fp64 = {"kernel": k.get("kernel"), "fp64_pipe_busy_pct": k.get("fp64_pipe_pct"), "issue_slots_busy_pct": k.get("issue_pct"),
            "source": "profiles/r01_kernels.json (ncu --set full of the same workload; not measured in this run)"}
    return k.get("dram_bytes"), fp64


def make_inputs(rank, batch):
    from world_b200 import synth_input
    return synth_input.batch(FS, SECONDS, 2, batch, first=rank * batch)


def cpu_baseline(x_one, cores_hint=None):
    """Oracle port of the reference on ONE synthetic utterance (bounded sample), single process."""
    from oracle import pipeline
    t0 = time.perf_counter()
    dat = pipeline.encode(FS, x_one, f0_method="harvest", is_requiem=False)
    dt = time.perf_counter() - t0
    frames = len(dat["f0"])
    return {"value": frames / dt, "unit": "frames/s", "cores": 1, "kind": "port",
            "sample": "1 utterance (16 kHz, 4 s, %d frames) of the config-2 workload, oracle/pipeline.encode "
                      "(NumPy port of the reference algorithm), 1 process" % frames}


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores (oracle port, all host cores via
    one process per core, each on its own utterance)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import pipeline
    cores = os.cpu_count() or 1
    n_utt = max(1, min(cores, 64))
    xs = make_inputs(0, n_utt)
    frames = int(1000 * xs.shape[1] / FS / FRAME_PERIOD + 1) * n_utt

    def one_step():
        t0 = time.perf_counter()
        with mp.Pool(n_utt) as pool:
            pool.starmap(pipeline.encode_quiet, [(FS, xs[i]) for i in range(n_utt)])
        return time.perf_counter() - t0

    for _ in range(max(0, min(args.warmup, 1))):
        one_step()
    steps = max(1, min(args.steps, 3))
    dt = sum(one_step() for _ in range(steps)) / steps
    val = frames / dt
    line = {"impl": "reference", "metric": "WORLD analysis frames/sec (16 kHz, 5 ms hop)", "value": val,
            "unit": "frames/s", "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1),
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": "%d utterances per step" % n_utt},
            "cpu_baseline": {"value": val, "unit": "frames/s", "cores": n_utt, "kind": "port",
                             "sample": "%d utterances (16 kHz 4 s) per step, one process per core, "
                                       "oracle port of the reference" % n_utt},
            "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=2, help="CUDA streams the batch is split over inside encode()")
    ap.add_argument("--pipeline", type=int, default=16, help="parts encode_batch() pipelines H2D / kernels / D2H over")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from world_b200 import engine as eng, main as wmain

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    E = eng.default_engine(local)
    W = wmain.World()
    warm = max(3, args.warmup)

    xs = make_inputs(rank, args.batch)
    B, S = xs.shape
    F = int(1000 * S / FS / FRAME_PERIOD + 1)
    frames_rank = B * F
    X = E.f64(xs)
    ns = E.i32([S] * B)
    # L2 flush buffer (> 126 MB) written between timed iterations
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=E.device)

    def step_resident():
        return E.encode(X, ns, FS, f0_method="harvest", is_requiem=False, streams=args.streams)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warm):
        step_resident()
    barrier()
    with ClockSampler(local) as clk:
        evs = []
        barrier()
        t_wall = time.perf_counter()
        total_ms = 0.0
        for _ in range(args.steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step_resident()
            b.record()
            evs.append((a, b))
        barrier()
        total_ms = sum(a.elapsed_time(b) for a, b in evs)
        # per-kernel timing of the dominant kernel on the same stream, same inputs
        stage_ms = E.profile_stages(X, ns, FS, f0_method="harvest", is_requiem=False, iters=max(2, args.steps))
        # end to end through the public batch API with host buffers
        xs_pinned = torch.from_numpy(xs).pin_memory()
        for _ in range(2):
            W.encode_batch(FS, xs_pinned, f0_method="harvest", is_requiem=False, pipeline=args.pipeline)
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(1, args.steps)
        h2d = d2h = 0
        for _ in range(e2e_steps):
            out = W.encode_batch(FS, xs_pinned, f0_method="harvest", is_requiem=False, pipeline=args.pipeline)
            h2d, d2h = out["_h2d_bytes"], out["_d2h_bytes"]
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
    tmax = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device=E.device)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(tmax[0]), float(tmax[1])
    ms_per_step = total_ms / args.steps
    value = frames_rank * world / (ms_per_step / 1e3)
    e2e_val = frames_rank * world / (e2e_ms / 1e3)

    if rank == 0:
        peak, peak_kind = peaks()
        top = max(stage_ms, key=lambda k: stage_ms[k])
        kern_ms = stage_ms[top]
        traffic, fp64 = kernel_profile(top, B)
        achieved = frames_rank * BYTES_PER_FRAME_NO_PS / (kern_ms / 1e3) / 1e9
        line = {
            "metric": "WORLD analysis frames/sec (16 kHz, 5 ms hop)", "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "frames_per_gpu": frames_rank,
                       "l2": "256 MiB flush buffer written between timed iterations",
                       "parallelism": "utterance-sharded, %d GPU(s), no data-path collective" % world,
                       "streams_per_gpu": args.streams},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_val, "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(E.launches_per_encode("harvest", False)) * args.steps * max(1, args.streams),
            "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_kind": peak_kind, "traffic": traffic,
                         "stage_ms": stage_ms,
                         "fp64": fp64,
                         "note": "algorithmic bytes = %d B/frame (SURVEY 8d config 2 without the optional "
                                 "'ps spectrogram' key) x frames / duration of the slowest stage" % BYTES_PER_FRAME_NO_PS},
        }
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(xs[0])
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
