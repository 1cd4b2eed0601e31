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
    long60s   ONE 1406-frame (59.99 s) utterance; N > 1: split along time over the ranks with per-stage NVLink P2P
              halo exchange (bvg_shard_*), N = 1: the whole decode on one GPU            [configs[4]]
The headline line is measured on --workload (default b16x10s); the other three, and the fp32 exactness path on
b16x10s, are measured right after it with the same protocol (fewer steps) and reported under "configs" of the SAME
JSON line at every N, so the driver's 1/2/4/8-GPU runs also carry configs[3] (strong scaling) and configs[4] (P2P).

Multi-GPU: utterances are independent, so ranks share nothing — no data-path collective; NCCL is
used only for the barrier and the max-over-ranks of the device time.  `scaling` is "weak" for
b16x10s / utt6p7s (every rank decodes its own batch) and "strong" for mixed256 / long60s.

One JSON line on stdout (rank 0).  `value` = device-resident throughput; `e2e` = the same through
the C-ABI host-buffer call (pinned host latents in, int16 PCM out, copies inside the timed region);
`roofline` = the fused tcgen05 AMP-layer kernel class timed live with CUDA events inside the timed
region; `cpu_baseline` = the oracle port on this box's host cores on a bounded sample; `parity` = a sampled
utterance of the benchmarked batch against the CPU oracle (outside the timed region).
The speaker encoder runs EVERY step (models.py:204 does): the module's embedding cache, on by default for callers,
is switched off here, and the rate with it on is reported separately as `value_prompt_cached`.
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
WORKLOADS = ["b16x10s", "utt6p7s", "mixed256", "long60s"]


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
def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_oracle_setup(threads: int):
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
    return h, m, O.fold_state_dict(sd), O


def cpu_oracle_run(frames: int, repeats: int, threads: int):
    """Oracle port of the reference path on the host cores: one utterance of `frames` frames."""
    from index_tts_lora_b200 import synth

    h, m, sdf, O = cpu_oracle_setup(threads)
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
    the reference is Python and does not travel to the GPU box) with all host threads this process may use.
    W warm-up and K timed steps exactly as asked; a step is a bounded sample of the workload (one 157-frame
    utterance) so the run ends within a few minutes.  The metric is normalised per audio-second."""
    if rank != 0:
        return
    import torch

    threads = host_threads()
    frames = 157
    steps = max(1, args.steps)
    warm = max(0, args.warmup)
    times = cpu_oracle_run(frames, warm + steps, threads)[warm:]
    sec = frames * UP / SR
    val = sec * len(times) / sum(times)
    srt = sorted(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * sum(times) / len(times),
        "ms_per_step_min": 1e3 * srt[0], "ms_per_step_median": 1e3 * srt[len(srt) // 2],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "gpu_launches": 0,
        "config": {"workload": args.workload,
                   "note": "CPU oracle port of indextts.BigVGAN.models.BigVGAN.forward, torch CPU ops; bounded "
                           "sample of the workload, metric normalised per audio-second"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"1 utterance x {frames} frames ({sec:.2f} s audio) per step, "
                                   f"{len(times)} steps after {warm} warm-up, torch {torch.__version__} "
                                   f"threads={threads} (sched_getaffinity)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
class Bench:
    """Everything a workload needs on one rank: the module, its native plan, the prompt, timing helpers."""

    def __init__(self, args, rank, world, local_rank):
        import torch

        from index_tts_lora_b200 import _lib, synth
        from index_tts_lora_b200.config import default_config
        from index_tts_lora_b200.models import BigVGAN

        self.torch, self.lib_mod, self.synth = torch, _lib, synth
        self.args, self.rank, self.world = args, rank, world
        self.dev = torch.device("cuda", local_rank)
        self.h = default_config()
        m = BigVGAN(self.h)
        self.sd = synth.synth_state_dict(m.state_dict(), seed=1234, profile="init")
        m.load_state_dict(self.sd)
        m = m.to(self.dev)
        m.remove_weight_norm()
        m.eval()
        m.cache_speaker_embedding = False          # the speaker encoder runs every step, as models.py:204 does
        self.m = m
        self.lib = _lib.load()
        self.plan = m._ensure_plan(self.dev)
        self.mel = synth.synth_mel(1, 300, self.h.num_mels, seed=1).to(self.dev)   # the prompt (infer.py:605-617)
        self.flush_buf = None if args.no_l2_flush else torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        self.stream = torch.cuda.current_stream(self.dev)

    # -- precision ----------------------------------------------------------------------
    def set_precision(self, precision: str):
        self.m.precision = precision
        self.prec = self.lib_mod.PREC_BF16 if precision == "bf16" else self.lib_mod.PREC_F32
        self.lat_dtype = self.torch.bfloat16 if precision == "bf16" else self.torch.float32

    # -- timing -------------------------------------------------------------------------
    def barrier(self):
        torch = self.torch
        torch.cuda.synchronize(self.dev)
        if self.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(self.dev)

    def timed(self, fn, n, profile=False):
        """n steps, each bracketed by CUDA events on the launching stream; L2 flushed in between.  Device time, max
        over the ranks."""
        torch = self.torch
        evs = []
        self.barrier()
        if profile:
            self.lib.bvg_plan_set_profiling(self.plan, 1)
        t_wall = time.perf_counter()
        for _ in range(n):
            if self.flush_buf is not None:
                self.flush_buf.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            fn()
            e1.record(self.stream)
            evs.append((e0, e1))
        self.barrier()
        wall = time.perf_counter() - t_wall
        if profile:
            self.lib.bvg_plan_set_profiling(self.plan, 0)
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall

    # -- workloads ----------------------------------------------------------------------
    def prepare(self, name: str):
        torch, synth, _lib = self.torch, self.synth, self.lib_mod
        lengths, total_frames, scaling = workload_lengths(name, self.rank, self.world)
        wl = {"name": name, "lengths": lengths, "total_frames": total_frames, "scaling": scaling,
              "my_frames": sum(lengths), "p2p": None}
        if name == "long60s" and self.world > 1:
            # per-stage NVLink P2P halo exchange (CUDA-IPC peer stores + flags); torch.distributed only swaps the IPC
            # handles once.  No host barrier between decodes: the flags order neighbouring ranks.
            from index_tts_lora_b200.longform import TimeSplitP2P
            full = synth.synth_latent(1, 1406, self.h.gpt_dim, seed=0).to(self.lat_dtype)
            wl["full"] = full
            # single-GPU decode of the whole utterance on rank 0: the parity reference of the split (fp32 samples)
            whole = None
            if self.rank == 0:
                emb = self.m.speaker_embedding(self.mel)
                whole = self.m.decode(full.to(self.dev), emb, out_dtype=torch.float32)[0, 0].clone()
            p2p = TimeSplitP2P(self.m, 1406, self.rank, self.world).setup(self.dev).connect_distributed()
            wl["p2p"] = p2p
            wl["win_host"] = p2p.window(full).pin_memory()
            wl["win"] = wl["win_host"].to(self.dev)
            wl["out_host"] = torch.empty((p2p.fe - p2p.fb) * UP, dtype=torch.int16).pin_memory()
            wl["my_frames"] = p2p.fe - p2p.fb
            wl["whole"] = whole
            return wl
        batches = batches_of(lengths)
        wl["batches"] = batches
        wl["dev_lat"], wl["host_lat"], wl["host_wav"], wl["lens_c"] = [], [], [], []
        for bi, bl in enumerate(batches):
            x = synth.synth_latent(len(bl), max(bl), self.h.gpt_dim, seed=100 * self.rank + bi).to(self.lat_dtype)
            for b, L in enumerate(bl):
                x[b, L:] = 0
            wl["host_lat"].append(x.pin_memory())
            wl["dev_lat"].append(x.to(self.dev))
            wl["host_wav"].append(torch.empty(len(bl), 1, max(bl) * UP, dtype=torch.int16).pin_memory())
            wl["lens_c"].append((C.c_int32 * len(bl))(*bl))
        return wl

    def step_device(self, wl):
        m = self.m
        if wl["p2p"] is not None:
            emb = m.speaker_embedding(self.mel)
            return wl["p2p"].decode(wl["win"], emb, out_dtype=self.torch.int16)
        out = None
        for bi, bl in enumerate(wl["batches"]):
            emb = m.speaker_embedding(self.mel)                   # ECAPA, part of forward (models.py:204)
            out = m.decode(wl["dev_lat"][bi], emb, lengths=bl if len(set(bl)) > 1 else None)
        return out

    def step_host(self, wl):
        m, _lib = self.m, self.lib_mod
        if wl["p2p"] is not None:
            emb = m.speaker_embedding(self.mel)
            win = wl["win_host"].to(self.dev, non_blocking=True)
            out = wl["p2p"].decode(win, emb, out_dtype=self.torch.int16)
            wl["out_host"].copy_(out, non_blocking=True)
            self.stream.synchronize()
            return
        for bi, bl in enumerate(wl["batches"]):
            emb = m.speaker_embedding(self.mel).reshape(1, -1).float().expand(len(bl), -1).contiguous()
            _lib.check(self.lib.bvg_decode_host(self.plan, wl["host_lat"][bi].data_ptr(),
                                                _lib.torch_dtype_code(self.lat_dtype), wl["lens_c"][bi], len(bl),
                                                max(bl), emb.data_ptr(), wl["host_wav"][bi].data_ptr(), _lib.BVG_I16,
                                                self.prec, self.stream.cuda_stream), "bvg_decode_host")

    def io_bytes(self, wl):
        if wl["p2p"] is not None:
            return wl["win_host"].numel() * wl["win_host"].element_size(), wl["out_host"].numel() * 2
        return (sum(x.numel() * x.element_size() for x in wl["host_lat"]),
                sum(x.numel() * x.element_size() for x in wl["host_wav"]))

    def measure(self, wl, K, W):
        """Warm-up + timed device-resident pass + timed host-buffer pass of one workload."""
        self.timed(lambda: self.step_device(wl), W)
        ms_dev, wall = self.timed(lambda: self.step_device(wl), K)
        launches = self.lib.bvg_plan_last_launches(self.plan) * (len(wl["batches"]) if wl["p2p"] is None else 1)
        self.timed(lambda: self.step_host(wl), W)          # the host-buffer call has its own launch-graph key: warm it up too
        ms_e2e, _ = self.timed(lambda: self.step_host(wl), K)
        if wl["p2p"] is not None:
            self.torch.cuda.synchronize(self.dev)
            wl["p2p"].check()                     # a halo wait that timed out would have produced garbage
        audio_s = wl["total_frames"] * UP / SR
        h2d, d2h = self.io_bytes(wl)
        return {"ms_dev": ms_dev, "ms_e2e": ms_e2e, "wall": wall, "launches": launches, "audio_s": audio_s,
                "value": audio_s / (ms_dev / K / 1e3), "e2e_value": audio_s / (ms_e2e / K / 1e3),
                "h2d": h2d, "d2h": d2h}

    def parallelism(self, wl):
        if wl["p2p"] is not None:
            return (f"time-split x{self.world}, per-stage NVLink P2P halo exchange (CUDA-IPC peer stores + flags, "
                    "6 exchanges per decode, no NCCL and no host barrier on the data path)")
        if wl["name"] == "long60s":
            return "whole utterance on one GPU"
        return f"utterance-sharded x{self.world}, no data-path collective"

    def p2p_parity(self, wl):
        """max |split - whole| over all samples: every rank's fp32 output gathered on rank 0 (outside any timing)."""
        torch = self.torch
        p2p = wl["p2p"]
        emb = self.m.speaker_embedding(self.mel)
        mine = p2p.decode(wl["win"], emb, out_dtype=torch.float32)
        torch.cuda.synchronize(self.dev)
        p2p.check()
        n_max = max(s[1] - s[0] for s in p2p.shards) * UP
        pad = torch.zeros(n_max, device=self.dev, dtype=torch.float32)
        pad[: mine.numel()] = mine
        parts = [torch.empty_like(pad) for _ in range(self.world)]
        torch.distributed.all_gather(parts, pad)
        if self.rank != 0:
            return None
        split = torch.cat([parts[r][: (s[1] - s[0]) * UP] for r, s in enumerate(p2p.shards)])
        return float((split - wl["whole"]).abs().max().item())


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
    ap.add_argument("--workload", default="b16x10s", choices=WORKLOADS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-l2-flush", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true", help="only the headline workload")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    from index_tts_lora_b200 import _lib

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
    bn = Bench(args, rank, world, local_rank)
    bn.set_precision(args.precision)
    lib, plan, m = bn.lib, bn.plan, bn.m

    # the time split pins the plan's workspace (it is IPC-exported): a headline long60s run measures only itself
    extras = [] if (args.no_extra_configs or (args.workload == "long60s" and world > 1)) else \
        [w for w in WORKLOADS if w != args.workload]

    # ---- headline workload: warm-up, timed region (device-resident), profiled pass, end-to-end pass
    wl = bn.prepare(args.workload)
    bn.timed(lambda: bn.step_device(wl), W)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_dev, wall_dev = bn.timed(lambda: bn.step_device(wl), K)
    launches = lib.bvg_plan_last_launches(plan) * (len(wl["batches"]) if wl["p2p"] is None else 1)
    # second pass of K steps with per-launch CUDA events (AMP blocks serialised, see bvg.h) for the
    # roofline of the dominant kernel; the headline numbers above come from the un-instrumented pass
    ms_prof, _ = bn.timed(lambda: bn.step_device(wl), K, profile=True)
    prof = _lib.BvgProfile()
    _lib.check(lib.bvg_plan_read_profile(plan, C.byref(prof)), "bvg_plan_read_profile")
    # ---- end-to-end through the C ABI with host buffers
    bn.timed(lambda: bn.step_host(wl), W)
    ms_e2e, _ = bn.timed(lambda: bn.step_host(wl), K)
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    # ---- the same device-resident pass with the prompt's embedding cached (the module's default for callers)
    m.cache_speaker_embedding = True
    bn.timed(lambda: bn.step_device(wl), 2)
    ms_cached, _ = bn.timed(lambda: bn.step_device(wl), K)
    m.cache_speaker_embedding = False
    p2p_err = bn.p2p_parity(wl) if wl["p2p"] is not None else None

    audio_s_total = wl["total_frames"] * UP / SR
    value = audio_s_total / (ms_dev / K / 1e3)
    e2e_value = audio_s_total / (ms_e2e / K / 1e3)

    # ---- parity of the benchmarked batch: one sampled utterance against the CPU oracle (outside the timed region)
    parity = None
    if wl["p2p"] is None:
        bl = wl["batches"][0]
        b = len(bl) - 1
        emb = m.speaker_embedding(bn.mel)
        wav = m.decode(wl["dev_lat"][0], emb, lengths=bl if len(set(bl)) > 1 else None, out_dtype=torch.float32)
        got = wav[b, 0, : bl[b] * UP].cpu()
        if rank == 0:
            from oracle import bigvgan_oracle as O
            from index_tts_lora_b200.models import BigVGAN
            sdf = O.fold_state_dict(bn.sd)
            mc = BigVGAN(bn.h)
            mc.load_state_dict(bn.sd)
            mc.eval()
            emb_cpu = mc.speaker_encoder(bn.mel.cpu())
            t0 = time.perf_counter()
            ref = O.generator_forward(sdf, bn.h, wl["host_lat"][0][b:b + 1, : bl[b]].float(), emb_cpu)[0, 0]
            parity = {"workload": args.workload, "utterance": b, "frames": bl[b],
                      "snr_db": O.snr_db(ref, got), "max_abs": float((got - ref).abs().max().item()),
                      "ref_abs_max": float(ref.abs().max().item()),
                      "gate": "bf16: SNR >= 40 dB; fp32: max-abs <= 1e-4 (north_star)",
                      "oracle_s": time.perf_counter() - t0}

    # ---- the other BASELINE configs with the same protocol, fewer steps
    Ks = max(1, min(K, 3))
    configs = {}
    for name in extras:
        if name == "long60s":
            continue                                  # last: its set-up pins the workspace
        w2 = bn.prepare(name)
        r = bn.measure(w2, Ks, W)
        configs[name] = {"value": r["value"], "unit": UNIT, "ms_per_step": r["ms_dev"] / Ks, "steps": Ks, "warmup": W,
                         "scaling": w2["scaling"], "n_gpus": world, "frames_total": w2["total_frames"],
                         "frames_this_rank": w2["my_frames"], "batches_this_rank": len(w2["batches"]),
                         "e2e": {"value": r["e2e_value"], "unit": UNIT, "ms_per_step": r["ms_e2e"] / Ks,
                                 "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"]},
                         "gpu_launches_per_step": r["launches"], "parallelism": bn.parallelism(w2), "dtype": "bf16"}
        del w2
    if extras and args.precision == "bf16":
        # fp32 exactness path on the headline shape (one batch per rank)
        bn.set_precision("fp32")
        w2 = bn.prepare("b16x10s")
        r = bn.measure(w2, 1, 3)
        configs["b16x10s_fp32"] = {"value": r["value"], "unit": UNIT, "ms_per_step": r["ms_dev"], "steps": 1,
                                   "warmup": 3, "scaling": "weak", "n_gpus": world, "dtype": "f32",
                                   "e2e": {"value": r["e2e_value"], "unit": UNIT, "ms_per_step": r["ms_e2e"],
                                           "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"]},
                                   "note": "SIMT FFMA exactness path (parity anchor, not a throughput path)"}
        del w2
        bn.set_precision(args.precision)
    if "long60s" in extras:
        w2 = bn.prepare("long60s")
        r = bn.measure(w2, Ks, W)
        configs["long60s"] = {"value": r["value"], "unit": UNIT, "ms_per_step": r["ms_dev"] / Ks, "steps": Ks,
                              "warmup": W, "scaling": "strong", "n_gpus": world, "frames_total": 1406,
                              "frames_this_rank": w2["my_frames"],
                              "e2e": {"value": r["e2e_value"], "unit": UNIT, "ms_per_step": r["ms_e2e"] / Ks,
                                      "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"]},
                              "gpu_launches_per_step": r["launches"], "parallelism": bn.parallelism(w2),
                              "dtype": "bf16"}
        if w2["p2p"] is not None:
            err = bn.p2p_parity(w2)
            configs["long60s"]["max_abs_vs_single_gpu"] = err

    if rank == 0:
        peaks, peaks_kind = measured_peaks()
        # dominant kernel class: the fused tcgen05 AMP layers (classes 0 = wide, 1 = narrow stages)
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

        def rate(i, arr):
            return arr[i] / (cls_ms[i] * 1e-3) if cls_ms[i] else 0.0
        per_class = {
            "wide": {"stages": "C >= 192 (tensor-bound)", "ms_per_step": cls_ms[0] / K, "tflops": rate(0, cls_fl) / 1e12,
                     "frac": rate(0, cls_fl) / 1e12 / peak_tf if peak_tf else None, "tensor_pipe_active": None},
            "narrow": {"stages": "C <= 96 (activation-bound)", "ms_per_step": cls_ms[1] / K,
                       "tflops": rate(1, cls_fl) / 1e12, "hbm_gbs": rate(1, cls_by) / 1e9,
                       "frac": rate(1, cls_by) / 1e9 / peaks["hbm_gbs"], "issue_active": None}}
        roofline = {"bound": "tensor", "kernel": "k_amp_tc (fused Activation1d + dilated Conv1d, all 108 launches/step)",
                    "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf if peak_tf else None,
                    "peak_source": f"{peaks_kind}:bf16_tflops_sustained", "traffic": None,
                    "avg_launch_ms": amp_ms / max(1, cls_n[0] + cls_n[1]),
                    "share_of_step": amp_ms / ms_prof if ms_prof else None,
                    "how": "per-launch CUDA events on the launching stream over a second pass of K steps with "
                           "the three AMP blocks of a stage serialised (they overlap on 3 streams in the timed pass)",
                    "serialised_ms_per_step": ms_prof / K, "per_class": per_class}
        # DRAM traffic and pipe activity of the same kernel class from the committed ncu captures of THIS round
        # (profiles/r02_amp_ncu.json names the commit they were taken at)
        tpath = os.path.join(ROOT, "profiles", "r02_amp_ncu.json")
        if args.workload == "b16x10s" and args.precision == "bf16" and os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            roofline["traffic"] = tj.get("amp_dram_bytes_per_launch")
            roofline["traffic_source"] = tj.get("source")
            roofline["ncu_commit"] = tj.get("commit")
            roofline["algorithmic_bytes_per_launch"] = tj.get("amp_alg_bytes_per_launch")
            roofline["algorithmic_flops_per_launch"] = tj.get("amp_flops_per_launch")
            per_class["wide"]["tensor_pipe_active"] = tj.get("wide_tensor_pipe_active")
            per_class["narrow"]["issue_active"] = tj.get("narrow_issue_active")
            per_class["narrow"]["dram_throughput_pct"] = tj.get("narrow_dram_throughput_pct")
        by_class = {}
        for i, nm in enumerate(["amp_tensor_stages", "amp_small_stages", "pre_ups_cond", "post"]):
            if cls_n[i]:
                by_class[nm] = {"ms_per_step": cls_ms[i] / K, "launches_per_step": cls_n[i] / K,
                                "tflops": rate(i, cls_fl) / 1e12, "hbm_gbs": rate(i, cls_by) / 1e9,
                                "hbm_frac": rate(i, cls_by) / 1e9 / peaks["hbm_gbs"]}
        h2d, d2h = bn.io_bytes(wl)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": args.workload,
                       "utterances_per_rank": len(wl["lengths"]), "frames_per_rank": wl["my_frames"],
                       "audio_seconds_per_step": audio_s_total, "weights": "random-init (synth profile 'init', seed 1234)",
                       "l2": ("working set >> 126 MB L2 per step" +
                              ("" if args.no_l2_flush else " + 256 MiB L2 flush between steps (outside the events)")),
                       "speaker_encoder": "runs every step (embedding cache off for the measurement)",
                       "parallelism": bn.parallelism(wl)},
            "gpu_launches": launches * K,
            "tensor_frac_of_step": (FLOP_PER_FRAME * wl["my_frames"] / (ms_dev / K * 1e-3) / 1e12) / peak_tf,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / K, "api": "bvg_decode_host (pinned latents in, int16 PCM out)"},
            "value_prompt_cached": audio_s_total / (ms_cached / K / 1e3),
            "roofline": roofline, "kernel_classes": by_class,
            "clocks": sampler.summary(),
            "wall_s_timed_region": wall_dev,
            "parity": parity, "configs": configs,
        }
        if p2p_err is not None:
            line["max_abs_vs_single_gpu"] = p2p_err
        if not args.no_cpu_baseline and world == 1:     # the CPU arm is a single-GPU-run figure (rank 0 at N = 1 only)
            threads = host_threads()
            t = cpu_oracle_run(157, 3, threads)[1:]
            best = min(t)
            line["cpu_baseline"] = {"value": 157 * UP / SR / best, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"1 utterance x 157 frames (6.70 s audio), best of {len(t)} runs after one "
                                              f"warm-up ({best:.2f} s), oracle port on torch CPU ops"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
