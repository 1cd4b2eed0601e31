#!/usr/bin/env python
"""bench.py — BigVGAN decode throughput (audio-seconds / wall-second) on N B200s.

    python bench.py --gpus N --steps K --warmup W             # our arm  (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU implementation of the path

A "step" is one pass of the hot path — `wav, _ = bigvgan(latent, mel_ref)` (indextts/infer.py:748,888)
— over one batch of synthetic latents.  Workloads (BASELINE.json configs):
    b16x10s   16 utterances x 234 frames (9.98 s) bf16 on one GPU      [default; configs[2], the
              configuration the throughput / roofline metric is quoted on]
    utt6p7s   1 utterance x 157 frames (6.70 s)                          [configs[1], latency case]
    mixed256  256 utterances, 47..469 frames, sharded by utterance over the ranks (LPT)  [configs[3]]
    long60s   ONE 1406-frame (59.99 s) utterance split along time over the ranks; every rank decodes
              its frame range plus receptive-field halos (overlap-recompute, bvg_decode_shard) [configs[4]]
Multi-GPU: utterances are independent, so ranks share nothing — no data-path collective; NCCL is
used only for the barrier and the max-over-ranks of the device time.  `scaling` is "weak" for
b16x10s / utt6p7s (every rank decodes its own batch) and "strong" for mixed256.

One JSON line on stdout (rank 0).  `value` = device-resident throughput; `e2e` = the same through
the C-ABI host-buffer call (pinned host latents in, int16 PCM out, copies inside the timed region);
`roofline` = the fused tcgen05 AMP-layer kernel class timed live with CUDA events inside the timed
region; `cpu_baseline` = the oracle port on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

SR = 24000
UP = 1024
METRIC = "bigvgan_decode_audio_seconds_per_second"
UNIT = "audio-s/s"
# algorithmic dense-conv FLOPs per latent frame (all layers) and AMP share
FLOP_PER_FRAME = 2922.725e6


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def workload_lengths(name: str, rank: int, world: int):
    """Latent-frame lengths decoded by this rank, and the global total."""
    from index_tts_lora_b200 import synth
    from index_tts_lora_b200.sharding import lpt_assign

    if name == "b16x10s":
        return [234] * 16, 16 * 234 * world, "weak"
    if name == "utt6p7s":
        return [157], 157 * world, "weak"
    if name == "long60s":
        return [1406], 1406, "strong"
    if name == "mixed256":
        lens = synth.synth_lengths(256, 47, 469, seed=2)
        shards = lpt_assign(lens, world)
        mine = sorted((lens[i] for i in shards[rank]), reverse=True)
        return mine, sum(lens), "strong"
    raise SystemExit(f"unknown workload {name}")


def batches_of(lengths, max_frames_per_batch=16 * 512):
    """Greedy length-sorted batching so padding stays small (ragged decode handles the rest)."""
    out, cur = [], []
    for L in lengths:
        if cur and (len(cur) + 1) * max(cur[0], L) > max_frames_per_batch:
            out.append(cur)
            cur = []
        cur.append(L)
    if cur:
        out.append(cur)
    return out


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        try:
            p = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                  "--format=csv,noheader,nounits", "-lms", "100"],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        while not self.stop_flag.is_set():
            line = p.stdout.readline()
            if not line:
                break
            self.rows.append([c.strip() for c in line.split(",")])
        p.terminate()

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ----------------------------------------------------------------------------- CPU arm
def cpu_oracle_run(frames: int, repeats: int, threads: int):
    """Oracle port of the reference path on the host cores: one utterance of `frames` frames."""
    import torch

    from index_tts_lora_b200 import synth
    from index_tts_lora_b200.config import default_config
    from index_tts_lora_b200.models import BigVGAN
    from oracle import bigvgan_oracle as O

    torch.set_grad_enabled(False)
    torch.set_num_threads(threads)
    h = default_config()
    m = BigVGAN(h)
    sd = synth.synth_state_dict(m.state_dict(), seed=1234, profile="init")
    m.load_state_dict(sd)
    m.eval()
    sdf = O.fold_state_dict(sd)
    lat = synth.synth_latent(1, frames, h.gpt_dim, seed=0)
    mel = synth.synth_mel(1, 300, h.num_mels, seed=1)
    times = []
    for _ in range(repeats):
        t = time.perf_counter()
        emb = m.speaker_encoder(mel)            # ECAPA is part of BigVGAN.forward (models.py:204)
        O.generator_forward(sdf, h, lat, emb)
        times.append(time.perf_counter() - t)
    return times


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path (the oracle port:
    the reference is Python and does not travel to the GPU box) with all host threads."""
    if rank != 0:
        return
    import torch

    threads = os.cpu_count() or 1
    frames = {"b16x10s": 234, "utt6p7s": 157, "mixed256": 234, "long60s": 234}[args.workload]
    steps = max(1, args.steps)
    warm = max(0, min(args.warmup, 1))
    times = cpu_oracle_run(frames, warm + steps, threads)[warm:]
    sec = frames * UP / SR
    val = sec * len(times) / sum(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "gpu_launches": 0,
        "config": {"workload": args.workload,
                   "note": "CPU oracle port of indextts.BigVGAN.models.BigVGAN.forward, torch CPU ops"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"1 utterance x {frames} frames ({sec:.2f} s audio) per step, "
                                   f"{len(times)} steps, torch {torch.__version__} threads={threads}"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def main():
    # stdout carries exactly ONE JSON line: everything else (NCCL banners, torch warnings, library
    # prints) is diverted to stderr for the whole run
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--workload", default="b16x10s", choices=["b16x10s", "utt6p7s", "mixed256", "long60s"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-l2-flush", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    from index_tts_lora_b200 import _lib, synth
    from index_tts_lora_b200.config import default_config
    from index_tts_lora_b200.models import BigVGAN

    torch.set_grad_enabled(False)
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    if world != args.gpus:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    W = max(3, args.warmup)
    K = max(1, args.steps)

    # ---- model: random-init weights of the configured architecture (config.yaml:88-146)
    h = default_config()
    m = BigVGAN(h)
    m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=1234, profile="init"))
    m = m.to(dev)
    m.remove_weight_norm()
    m.eval()
    m.precision = args.precision
    lib = _lib.load()
    plan = m._ensure_plan(dev)
    prec = _lib.PREC_BF16 if args.precision == "bf16" else _lib.PREC_F32

    # ---- inputs
    lengths, total_frames, scaling = workload_lengths(args.workload, rank, world)
    batches = batches_of(lengths)
    lat_dtype = torch.bfloat16 if args.precision == "bf16" else torch.float32
    mel = synth.synth_mel(1, 300, h.num_mels, seed=1).to(dev)          # cached prompt mel (infer.py:605-617)
    dev_lat, host_lat, host_wav, lens_c = [], [], [], []
    for bi, bl in enumerate(batches):
        x = synth.synth_latent(len(bl), max(bl), h.gpt_dim, seed=100 * rank + bi).to(lat_dtype)
        for b, L in enumerate(bl):
            x[b, L:] = 0
        host_lat.append(x.pin_memory())
        dev_lat.append(x.to(dev))
        host_wav.append(torch.empty(len(bl), 1, max(bl) * UP, dtype=torch.int16).pin_memory())
        lens_c.append((C.c_int32 * len(bl))(*bl))
    my_frames = sum(lengths)
    flush_buf = None if args.no_l2_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)

    shard = None
    p2p = None
    if args.workload == "long60s" and world > 1:
        # per-stage NVLink P2P halo exchange (CUDA-IPC peer stores + flags); torch.distributed only swaps
        # the IPC handles once and provides the host barrier between decodes
        from index_tts_lora_b200.longform import TimeSplitP2P
        full = synth.synth_latent(1, 1406, h.gpt_dim, seed=0).to(lat_dtype)
        p2p = TimeSplitP2P(m, 1406, rank, world).setup(dev).connect_distributed()
        p2p_win_host = p2p.window(full).pin_memory()
        p2p_win = p2p_win_host.to(dev)
        p2p_out_host = torch.empty((p2p.fe - p2p.fb) * UP, dtype=torch.int16).pin_memory()
        my_frames = p2p.fe - p2p.fb
    elif args.workload == "long60s":
        from index_tts_lora_b200.sharding import time_shards
        rf = lib.bvg_receptive_field_frames(plan)
        fb, fe, hl, hr = time_shards(1406, world, rf)[rank]
        full = synth.synth_latent(1, 1406, h.gpt_dim, seed=0).to(lat_dtype)       # same utterance on every rank
        sh_dev = full[0, fb - hl: fe + hr].contiguous().to(dev)
        sh_host = full[0, fb - hl: fe + hr].contiguous().pin_memory()
        sh_out = torch.empty((fe - fb) * UP, dtype=torch.int16, device=dev)
        sh_out_host = torch.empty((fe - fb) * UP, dtype=torch.int16).pin_memory()
        shard = (fb, fe, hl, hr)
        my_frames = fe - fb

    def step_shard(host: bool):
        fb, fe, hl, hr = shard
        emb = m.speaker_embedding(mel).reshape(1, -1).float().contiguous()
        src = sh_dev
        if host:
            src = sh_host.to(dev, non_blocking=True)
        _lib.check(lib.bvg_decode_shard(plan, src.data_ptr(), _lib.torch_dtype_code(lat_dtype), fb, fe, 1406, hl, hr,
                                        emb.data_ptr(), sh_out.data_ptr(), _lib.BVG_I16, prec, stream.cuda_stream),
                   "bvg_decode_shard")
        if host:
            sh_out_host.copy_(sh_out, non_blocking=True)
            stream.synchronize()

    def step_p2p(host: bool):
        dist.barrier()                                   # no rank may start decode n+1 while a neighbour is in n
        emb = m.speaker_embedding(mel)
        win = p2p_win_host.to(dev, non_blocking=True) if host else p2p_win
        out = p2p.decode(win, emb, out_dtype=torch.int16)
        if host:
            p2p_out_host.copy_(out, non_blocking=True)
            stream.synchronize()

    def step_device():
        if p2p is not None:
            return step_p2p(False)
        if shard is not None:
            return step_shard(False)
        out = None
        for bi, bl in enumerate(batches):
            emb = m.speaker_embedding(mel)                      # ECAPA, part of forward (models.py:204)
            out = m.decode(dev_lat[bi], emb, lengths=bl if len(set(bl)) > 1 else None)
        return out

    def step_host():
        if p2p is not None:
            return step_p2p(True)
        if shard is not None:
            return step_shard(True)
        for bi, bl in enumerate(batches):
            emb = m.speaker_embedding(mel).reshape(1, -1).float().expand(len(bl), -1).contiguous()
            _lib.check(lib.bvg_decode_host(plan, host_lat[bi].data_ptr(), _lib.torch_dtype_code(lat_dtype),
                                           lens_c[bi], len(bl), max(bl), emb.data_ptr(),
                                           host_wav[bi].data_ptr(), _lib.BVG_I16, prec, stream.cuda_stream),
                       "bvg_decode_host")

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, n, profile=False):
        """n steps, each bracketed by CUDA events on the launching stream; L2 flushed in between."""
        evs = []
        barrier()
        if profile:
            lib.bvg_plan_set_profiling(plan, 1)
        t_wall = time.perf_counter()
        for _ in range(n):
            if flush_buf is not None:
                flush_buf.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            evs.append((e0, e1))
        barrier()
        wall = time.perf_counter() - t_wall
        if profile:
            lib.bvg_plan_set_profiling(plan, 0)
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall

    # ---- warm-up, then the timed region (device-resident inputs)
    timed(step_device, W)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_dev, wall_dev = timed(step_device, K)
    launches = lib.bvg_plan_last_launches(plan) * len(batches)
    # second pass of K steps with per-launch CUDA events (AMP blocks serialised, see bvg.h) for the
    # roofline of the dominant kernel; the headline numbers above come from the un-instrumented pass
    ms_prof, _ = timed(step_device, K, profile=True)
    prof = _lib.BvgProfile()
    _lib.check(lib.bvg_plan_read_profile(plan, C.byref(prof)), "bvg_plan_read_profile")
    # ---- end-to-end through the C ABI with host buffers
    timed(step_host, 1)
    ms_e2e, _ = timed(step_host, K)
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    audio_s_total = total_frames * UP / SR
    value = audio_s_total / (ms_dev / K / 1e3)
    e2e_value = audio_s_total / (ms_e2e / K / 1e3)

    if rank == 0:
        peaks, peaks_kind = measured_peaks()
        # dominant kernel class: the fused tcgen05 AMP layers of the tensor-bound stages (class 0)
        cls_ms = [prof.ms[i] for i in range(4)]
        cls_fl = [prof.flops[i] for i in range(4)]
        cls_by = [prof.bytes[i] for i in range(4)]
        cls_n = [prof.launches[i] for i in range(4)]
        peak_tf = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
        if args.precision == "fp32":
            peak_tf = 74.5  # fp32 FFMA nominal (148 SMs x 128 lanes x 2 x 1.965 GHz): not tensor work
        amp_ms = cls_ms[0] + cls_ms[1]
        amp_fl = cls_fl[0] + cls_fl[1]
        ach = (amp_fl / (amp_ms * 1e-3) / 1e12) if amp_ms > 0 else 0.0
        roofline = {"bound": "tensor", "kernel": "k_amp_tc (fused Activation1d + dilated Conv1d, all 108 launches/step)",
                    "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf if peak_tf else None,
                    "peak_source": f"{peaks_kind}:bf16_tflops_sustained", "traffic": None,
                    "avg_launch_ms": amp_ms / max(1, cls_n[0] + cls_n[1]),
                    "share_of_step": amp_ms / ms_prof if ms_prof else None,
                    "how": "per-launch CUDA events on the launching stream over a second pass of K steps with "
                           "the three AMP blocks of a stage serialised (they overlap on 3 streams in the timed pass)",
                    "serialised_ms_per_step": ms_prof / K}
        # DRAM traffic of the same kernel class from the committed ncu capture (b16x10s only):
        # dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the 108 AMP launches
        tpath = os.path.join(ROOT, "profiles", "r01_amp_traffic_b16x10s.json")
        if args.workload == "b16x10s" and args.precision == "bf16" and os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            roofline["traffic"] = tj["amp_dram_bytes_per_launch"]
            roofline["traffic_source"] = "profiles/r01_launches_s3_b16x10s_time_dram.csv (ncu, per launch avg)"
            roofline["algorithmic_bytes_per_launch"] = tj["amp_alg_bytes_per_launch"]
            roofline["algorithmic_flops_per_launch"] = tj["amp_flops_per_launch"]
        by_class = {}
        for i, nm in enumerate(["amp_tensor_stages", "amp_small_stages", "pre_ups_cond", "post"]):
            if cls_n[i]:
                by_class[nm] = {"ms_per_step": cls_ms[i] / K, "launches_per_step": cls_n[i] / K,
                                "tflops": cls_fl[i] / (cls_ms[i] * 1e-3) / 1e12 if cls_ms[i] else 0.0,
                                "hbm_gbs": cls_by[i] / (cls_ms[i] * 1e-3) / 1e9 if cls_ms[i] else 0.0,
                                "hbm_frac": cls_by[i] / (cls_ms[i] * 1e-3) / 1e9 / peaks["hbm_gbs"] if cls_ms[i] else 0.0}
        h2d = sum(x.numel() * x.element_size() for x in host_lat)
        d2h = sum(x.numel() * x.element_size() for x in host_wav)
        if shard is not None:
            h2d, d2h = sh_host.numel() * sh_host.element_size(), sh_out_host.numel() * 2
        if p2p is not None:
            h2d, d2h = p2p_win_host.numel() * p2p_win_host.element_size(), p2p_out_host.numel() * 2
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": args.workload,
                       "utterances_per_rank": len(lengths), "frames_per_rank": my_frames,
                       "audio_seconds_per_step": audio_s_total, "weights": "random-init (synth profile 'init', seed 1234)",
                       "l2": ("working set >> 126 MB L2 per step" +
                              ("" if args.no_l2_flush else " + 256 MiB L2 flush between steps (outside the events)")),
                       "parallelism": (f"time-split x{world}, per-stage NVLink P2P halo exchange (CUDA-IPC peer stores + "
                                       "flags, 6 exchanges per decode, no NCCL on the data path)" if p2p is not None else
                                       f"time-split x{world} with {shard[2]}/{shard[3]}-frame halos on rank 0 "
                                       "(overlap-recompute, no exchange)" if shard is not None else
                                       f"utterance-sharded x{world}, no data-path collective")},
            "gpu_launches": launches * K,
            "tensor_frac_of_step": (FLOP_PER_FRAME * my_frames / (ms_dev / K * 1e-3) / 1e12) / peak_tf,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / K, "api": "bvg_decode_host (pinned latents in, int16 PCM out)"},
            "roofline": roofline, "kernel_classes": by_class,
            "clocks": sampler.summary(),
            "wall_s_timed_region": wall_dev,
        }
        if not args.no_cpu_baseline and world == 1:     # the CPU arm is a single-GPU-run figure (rank 0 at N = 1 only)
            threads = os.cpu_count() or 1
            t = cpu_oracle_run(157, 2, threads)
            best = min(t)
            line["cpu_baseline"] = {"value": 157 * UP / SR / best, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"1 utterance x 157 frames (6.70 s audio), best of {len(t)} runs "
                                              f"({best:.2f} s), oracle port on torch CPU ops"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
