#!/usr/bin/env python
"""Extract-stage throughput benchmark (BASELINE.json: audio-seconds per second; % of tensor / HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--songs-per-gpu S] [--impl reference]

One *step* = the whole hot path over one batch of synthetic songs: fused log-mel front-end -> every 512-frame window
through the hFT-Transformer -> piano-rolls stitched on the device (-> notes for the e2e leg).  Workload (N = 1 and
N > 1 alike, weak scaling): S synthetic 4-minute 16 kHz songs per GPU (default 32: at N = 8 this is BASELINE config 4,
"256 synthetic 4-min songs sharded by window across 8 B200"), random-init weights of the named architecture, window
batch 32.  Songs are sharded over ranks with no collective on the hot path; one all_gather of per-song note counts at
the end of each e2e step stands for the final gather.

Printed JSON (rank 0): `value` = audio-s/s with the waves resident in HBM (log-mel + model, rolls left on the device);
`e2e` = the same through the public API `AMTAPC_Extractor.extract_many` with pinned HOST waves in and note lists out;
`roofline` = the dominant kernel class (CUDA-event timed inside the timed region) against the measured bf16 peak;
`frontend` = the log-mel kernel against the measured HBM peak; `cpu_baseline` = the oracle port (torch CPU restatement of
the reference) on this box's host cores over a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 16000
SONG_SECONDS = 240
SONG_SAMPLES = SR * SONG_SECONDS
FLOP_PER_WINDOW_REF = 1050.9e9       # reference formulation; executed count comes from the library
METRIC = "extract_throughput_audio_seconds_per_second"
UNIT = "audio-s/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_burst": p["bf16_tflops"], "bf16_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = sorted(sm)[len(sm) // 4:] if sm else []   # drop idle samples before the first launch
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_waves(n_songs, rank, n_samples=SONG_SAMPLES):
    from etude_b200 import synth
    return [synth.noise(n_samples, seed=1234 + rank * 1000 + i) for i in range(n_songs)]


def cpu_baseline_sample(threads, windows=4, repeats=1):
    """Oracle port (torch fp32 CPU restatement of the reference extractor) on a bounded sample: log-mel + `windows`
    windows + note decoding of one synthetic clip.  Returns (audio-s/s, description)."""
    from etude_b200 import synth
    from oracle import logmel as ologmel
    from oracle import model as omodel
    from oracle import notes as onotes
    torch.set_num_threads(threads)
    sd = omodel.init_state_dict(0)
    n = 256 * (512 * windows - 1)
    wave = synth.noise(n, 1234)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        feat = ologmel.logmel(wave, dtype=np.float32)
        outs = omodel.transcript(sd, feat, batch=windows)
        onotes.mpe2note(outs[4], outs[5], outs[6], outs[7], 0.5, 1.0, 0.5)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    secs = n / SR
    return secs / best, f"{windows} windows ({secs:.1f} s of 16 kHz noise): log-mel + model (batch {windows}) + notes, torch fp32 CPU, {best:.2f} s wall"


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path, as restated by the oracle port (the reference
    itself is not on the GPU box), on all host threads; each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    windows = 2
    vals = []
    for i in range(args.warmup + args.steps):
        v, desc = cpu_baseline_sample(threads, windows=windows)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
            "warmup": args.warmup, "ms_per_step": 1e3 * windows * 8.192 / value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_frontend_microbench(args):
    """BASELINE config 2: log-mel front-end over 150 four-minute songs (10 h, 576 M samples) resident in HBM; HBM roofline with
    the algorithmic bytes of SURVEY 8(d): 4 B per sample in + 4 B x 256 per frame out (intermediates never touch HBM)."""
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    from etude_b200 import AMTAPC_Extractor, ExtractorConfig, synth
    from etude_b200.weights import default_state_dict
    ckpt = os.path.join(tempfile.gettempdir(), "etude_bench_sd_fe.pth")
    torch.save(default_state_dict(seed=0), ckpt)
    ex = AMTAPC_Extractor(ExtractorConfig(), ckpt, device=dev, max_windows=1)
    n_songs = 150
    base = [torch.from_numpy(synth.noise(SONG_SAMPLES, 1234 + i)) for i in range(4)] + \
           [torch.from_numpy(synth.tones(SONG_SAMPLES, 4321 + i).astype(np.float32)) for i in range(2)]
    wave = torch.cat([base[i % len(base)] for i in range(n_songs)]).to(dev)      # both input classes of SURVEY 8(d)
    n_samples = [SONG_SAMPLES] * n_songs
    wave_off = (np.arange(n_songs, dtype=np.int64) * SONG_SAMPLES)
    frames = sum(1 + n // 256 for n in n_samples)
    alg_bytes = 4.0 * sum(n_samples) + 4.0 * 256 * frames
    for _ in range(max(3, args.warmup)):
        ex.engine.logmel(wave, wave_off, n_samples)
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = max(args.steps, 60)   # ~0.75 s: long enough for the 200 ms clock sampler
    e0.record()
    for _ in range(steps):
        ex.engine.logmel(wave, wave_off, n_samples)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / steps
    pk = peaks()
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    secs = sum(n_samples) / SR
    print(json.dumps({
        "metric": "logmel_frontend_audio_seconds_per_second", "value": secs / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "log-mel STFT front-end microbench over 10 h of synthetic 16 kHz audio (150 x 4 min, noise + tones), 1 B200",
                   "cache": "inputs larger than L2 (2.3 GB of samples, 2.4 GB of features vs 126 MB L2)"},
        "gpu_launches": steps, "clocks": clocks,
        "roofline": {"kernel": "logmel2_kernel", "bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                     "traffic": None, "peak_source": pk["source"], "bytes_per_launch": alg_bytes, "avg_launch_ms": ms,
                     "note": "fp32 issue-bound (~40 kFLOP and ~2 500 instructions per frame), see DESIGN.md section 4"}}))


def workload_config(args):
    return {"workload": f"full extractor (log-mel + hFT-Transformer + roll stitching) over {args.songs_per_gpu} synthetic 4-min 16 kHz "
                        f"songs per GPU, window batch {args.window_batch}, random-init weights",
            "songs_per_gpu": args.songs_per_gpu, "windows_per_song": 30, "window_batch": args.window_batch,
            "cache": "inputs larger than L2 (15.4 MB wave + 1 GB activations per window batch vs 126 MB L2)",
            "sharding": "songs over ranks, no hot-path collective"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--songs-per-gpu", type=int, default=32)
    ap.add_argument("--window-batch", type=int, default=32)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-windows", type=int, default=4)
    ap.add_argument("--frontend-microbench", action="store_true",
                    help="BASELINE config 2: the fused log-mel front-end alone over 10 h of synthetic audio on one GPU")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.frontend_microbench:
        return run_frontend_microbench(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    from etude_b200 import AMTAPC_Extractor, ExtractorConfig
    from etude_b200.weights import default_state_dict
    sd = default_state_dict(seed=0)
    ckpt = os.path.join(tempfile.gettempdir(), f"etude_bench_sd_{rank}.pth")
    torch.save(sd, ckpt)
    ex = AMTAPC_Extractor(ExtractorConfig(), ckpt, device=dev, max_windows=args.window_batch)
    eng = ex.engine

    waves = make_waves(args.songs_per_gpu, rank)
    n_samples = [len(w) for w in waves]
    wave_off = np.concatenate([[0], np.cumsum(n_samples)]).astype(np.int64)
    audio_seconds = sum(n_samples) / SR
    pinned = torch.empty(int(wave_off[-1]), dtype=torch.float32, pin_memory=True)
    pinned.numpy()[:] = np.concatenate(waves)
    wave_dev = pinned.to(dev)
    n_windows = sum((1 + n // 256 + 511) // 512 for n in n_samples)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident leg: log-mel + model, K steps timed with CUDA events on the launching stream
    def step_device():
        return ex.transcribe_device(wave_dev, wave_off[:-1], n_samples)

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.profile_reset(timing=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    prof = eng.profile_read()
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms_total / args.steps
    value = world * audio_seconds / (ms_per_step / 1e3)

    # ---------------- end-to-end leg: pinned host waves -> notes on the host, through the public API
    def step_e2e():
        recs = ex.extract_many(waves, as_dicts=False, pinned=pinned)
        counts = torch.tensor([len(r) for r in recs], dtype=torch.int64, device=dev)
        if dist is not None:   # the one final gather (per-song note counts; rolls / notes stay sharded)
            out = [torch.empty_like(counts) for _ in range(world)]
            dist.all_gather(out, counts)
        return recs

    eng.profile_reset(timing=False)
    recs = step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 2))
    for _ in range(e2e_steps):
        recs = step_e2e()
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    e2e_value = world * audio_seconds / e2e_s
    if rank == 0 and os.environ.get("ETUDE_E2E_TRACE"):   # diagnostic: where an e2e step spends its wall time
        import time as _t
        def lap(fn):
            torch.cuda.synchronize(); t = _t.perf_counter(); r = fn(); torch.cuda.synchronize(); return r, 1e3 * (_t.perf_counter() - t)
        def stage():
            hv = pinned.numpy()
            for w, o, n in zip(waves, wave_off[:-1], n_samples):
                hv[o:o + n] = w
            return pinned.to(dev, non_blocking=True)
        wd, t_stage = lap(stage)
        (rolls, sro, srows), t_dev = lap(lambda: ex.transcribe_device(wd, wave_off[:-1], n_samples))
        cfg = ex.config.infer
        _, t_notes = lap(lambda: eng.notes(rolls[0], rolls[1], rolls[2], rolls[3], sro, srows, cfg.onset_threshold, cfg.offset_threshold,
                                           cfg.frame_threshold))
        print(f"[e2e trace] stage+H2D {t_stage:.1f} ms, log-mel+model {t_dev:.1f} ms, notes (kernels + D2H + host) {t_notes:.1f} ms",
              file=sys.stderr)
    d2h = int(sum(r.nbytes for r in recs)) + 8 * len(recs) * 88
    h2d = int(4 * wave_off[-1])

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    pk = peaks()
    model_classes = ("embed", "gemm_bias", "gemm_ln", "gemm_heads", "attention", "chain")
    kernels = {}
    for name, p in prof.items():
        if p["launches"] == 0:
            continue
        avg_ms = p["ms"] / p["launches"]
        k = {"launches_per_step": p["launches"] / args.steps, "ms_per_step": p["ms"] / args.steps, "avg_launch_ms": avg_ms,
             "share_of_step": p["ms"] / ms_total}
        if p["flops"] > 0:
            k["tflops"] = p["flops"] / (p["ms"] * 1e-3) / 1e12 if p["ms"] > 0 else None
        if p["bytes"] > 0:
            k["gbs"] = p["bytes"] / (p["ms"] * 1e-3) / 1e9 if p["ms"] > 0 else None
        kernels[name] = k
    dom = max((n for n in kernels if n in model_classes), key=lambda n: kernels[n]["ms_per_step"])
    dom_p = prof[dom]
    achieved = dom_p["flops"] / (dom_p["ms"] * 1e-3) / 1e12
    traffic = None   # DRAM bytes per launch of the dominant kernel: measured once with `ncu --set full` (profiles/chain_traffic.json)
    tpath = os.path.join(ROOT, "profiles", "chain_traffic.json")
    if dom == "chain" and os.path.exists(tpath):
        t = json.load(open(tpath))
        traffic = (t["dram_bytes_read"] + t["dram_bytes_write"]) / t["flop"] * (dom_p["flops"] / dom_p["launches"])
    roofline = {"kernel": dom, "bound": "tensor", "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["bf16_sustained"], "frac_of_burst_peak": achieved / pk["bf16_burst"], "traffic": traffic,
                "traffic_note": "ncu-measured DRAM bytes per FLOP of the FFN chain launch x FLOP per launch here (algorithmic: 1 536 B/token)",
                "peak_source": pk["source"] + ", sustained bf16 (kernel timed inside a long step)",
                "flop_per_launch": dom_p["flops"] / dom_p["launches"], "avg_launch_ms": dom_p["ms"] / dom_p["launches"]}
    flops_exec = sum(prof[n]["flops"] for n in model_classes) / args.steps
    model_ms = sum(prof[n]["ms"] for n in model_classes) / args.steps
    fe = prof["logmel"]
    frontend = {"kernel": "logmel", "bound": "hbm", "achieved": fe["bytes"] / (fe["ms"] * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": fe["bytes"] / (fe["ms"] * 1e-3) / 1e9 / pk["hbm_gbs"], "traffic": None,
                "audio_s_per_s": args.steps * audio_seconds / (fe["ms"] * 1e-3),
                "bytes_per_launch": fe["bytes"] / fe["launches"], "avg_launch_ms": fe["ms"] / fe["launches"]}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic", "config": workload_config(args),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "api": "AMTAPC_Extractor.extract_many(host waves) -> note records on the host"},
        "gpu_launches": int(sum(p["launches"] for p in prof.values())),
        "clocks": clocks,
        "roofline": roofline,
        "model": {"tflops_executed": flops_exec / (model_ms * 1e-3) / 1e12, "gflop_per_window_executed": flops_exec / n_windows / 1e9,
                  "gflop_per_window_reference": FLOP_PER_WINDOW_REF / 1e9,
                  "tensor_util_of_sustained_peak": flops_exec / (ms_per_step * 1e-3) / 1e12 / pk["bf16_sustained"],
                  "tensor_util_of_burst_peak": flops_exec / (ms_per_step * 1e-3) / 1e12 / pk["bf16_burst"]},
        "frontend": frontend,
        "kernels": kernels,
    }
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, desc = cpu_baseline_sample(threads, windows=args.cpu_windows)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
