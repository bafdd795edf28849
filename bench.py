#!/usr/bin/env python
"""Benchmark of the LineTR hot path: image-pairs/s of (encode side 0 + encode side 1 + match)
on synthetic 640x480 frames, cfg[1] of BASELINE.json (64 pairs x 128 lines x 21 tokens x d256
per GPU), weak scaling over 1..8 GPUs (pairs are independent; one all-gather of match counts).

    python bench.py [--gpus N --steps K --warmup W]        # this repo's CUDA path
    python bench.py --impl reference [...]                 # CPU baseline (oracle port), rank 0

Prints ONE JSON line (contract in the task statement): `value` = pairs/s with inputs resident
in HBM, `e2e` = the same through pinned-host buffers (H2D of every input tensor and D2H of the
match indices inside the timed region), `roofline` for the dominant kernel class (a second pass of
the same K steps with a CUDA-event pair around every launch on the launching stream; separate so
that the events do not serialise the kernels of the timed pass), `cpu_baseline` (oracle port on
the host cores).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "image-pairs/s (line-descriptor forward x2 + mutual-NN match)"
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel class from the
# committed `ncu --set full` capture (profiles/), bytes; None until a capture exists
TRAFFIC_NCU = {"linear": 38.7e6}   # profiles/r1_final_kernels.md (prof_v6), average over one signature layer
UNIT = "pairs/s"


def env_int(name, default):
    return int(os.environ.get(name, default))


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ---------------------------------------------------------------- algorithmic work
def flops_per_image(L, T, n_sig=7):
    """Useful FLOPs (2*MAC) of one image under reference semantics (CLS row of the last layer)."""
    N = T + 1
    f_wpe = 217280 * L * T
    f_lpe = 217408 * L
    f_desc_cls = 4096 * L * N + 1310720 * L
    f_sig = (1310720 * L + 1024 * L * L) * n_sig
    f_fin = 131072 * L
    return f_wpe + f_lpe + f_desc_cls + f_sig + f_fin


def gemm_flops_per_image(L, T, n_sig=7):
    """Useful FLOPs the `linear` kernel class (gemm_img_kernel launches) is responsible for, one
    image: line stage (per-head V projection, fc, FFN, wide line-positional layers), signature
    layers (qkv, MLP with the merge projection folded in, MLP out) and final_proj.  The 3x
    split-bf16 products are NOT counted."""
    line = 2 * (4 * 256 * 64 + 256 * 256 + 2 * 256 * 1024 + 128 * 256 + 256 * 256) * L
    sig = 2 * (256 * 768 + 512 * 512 + 512 * 256) * L * n_sig
    fin = 2 * 256 * 256 * L
    return line + sig + fin


def token_flops_per_image(L, T):
    """Useful FLOPs of token_fused_kernel: narrow MLP 3-32-64, 64-128-256-256 on tensor cores, CLS pooling."""
    return (2 * (3 * 32 + 32 * 64 + 64 * 128 + 128 * 256 + 256 * 256) + 4096) * L * T


def bytes_per_image(L, T):
    return 4 * (256 * L * T + 2 * L * T + L * T + L * (T + 1) + 4 * L + L + 2 * L) + 4 * 256 * L


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.004)

    def result(self):
        self.stop_flag = True
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def cpu_reference_pairs_per_s(L, T, budget_s, seed=900):
    """Times the CPU port of the reference path (B=1 per call, as Matching runs it)."""
    from linetr_b200 import synthetic as syn
    try:
        from oracle import linetr_oracle_torch as port
        kind_note = "torch-CPU functional port (same aten ops as the reference)"
    except ImportError:
        from oracle import linetr_oracle as port
        kind_note = "numpy port"
    sd = port.prepare(syn.make_state_dict(0, 1)) if hasattr(port, "prepare") else syn.make_state_dict(0, 1)
    pairs = [syn.make_pair_inputs(seed + i, L, T)[:2] for i in range(2)]
    port.match_pair(sd, *pairs[0])  # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        port.match_pair(sd, *pairs[n % len(pairs)])
        n += 1
        el = time.perf_counter() - t0
        if el > budget_s or n >= 256:
            break
    cores = os.cpu_count()
    try:
        import torch
        cores = torch.get_num_threads()
    except Exception:
        pass
    return n / el, n, cores, kind_note


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=64, help="image pairs per GPU per step")
    ap.add_argument("--lines", type=int, default=128)
    ap.add_argument("--tokens", type=int, default=21)
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU baseline sampling")
    ap.add_argument("--e2e-chunks", type=int, default=4, help="pair groups whose H2D copy overlaps compute in the e2e leg")
    ap.add_argument("--profile-only", action="store_true", help="resident steps only (for runs under ncu)")
    args = ap.parse_args()
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)
    P, L, T = args.pairs, args.lines, args.tokens
    workload = f"cfg1: {P} pairs/GPU x {L} lines x {T} tokens x d256, 1 descriptive + 7 signature layers"
    config = {"workload": workload, "pairs_per_gpu": P, "lines": L, "tokens": T, "sharding": f"pairs over {world} ranks",
              "l2_policy": f"inputs {2 * P * bytes_per_image(L, T) / 1e6:.0f} MB per step > 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return
        per_step = max(args.cpu_budget / max(args.steps + args.warmup, 1), 1.0)
        vals = []
        cores = n = 0
        note = ""
        for i in range(args.warmup + args.steps):
            v, n, cores, note = cpu_reference_pairs_per_s(L, T, per_step, seed=900 + i)
            if i >= args.warmup:
                vals.append(v)
        value = float(np.mean(vals))
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": 1e3 * per_step, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic (seeded inputs, random-init weights)",
               "config": config, "impl": "reference",
               "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{n} pairs of {L}x{T} per step, B=1 per call; {note}"},
               "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(out))
        return

    import torch
    import torch.distributed as dist
    from linetr_b200 import LineBatch, LineTransformer, PairEngine, _native, synthetic as syn
    from linetr_b200.engine import gather_counts

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sd = syn.make_state_dict(0, 1)
    model = LineTransformer({"mode": "train"})
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    model = model.eval().to(dev)
    eng = PairEngine(model, dev)

    # synthetic inputs, pinned on the host (e2e source) and a resident device copy
    pairs = [syn.make_pair_inputs(10_000 * rank + i, L, T)[:2] for i in range(P)]
    # one packed batch: images [0, P) = side 0, [P, 2P) = side 1 (one encode launch sequence per step)
    host = LineBatch.from_images([a for a, _ in pairs] + [b for _, b in pairs]).pin()
    resident = host.to(dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pending = []   # all-gathers in flight: the gather of step i overlaps the kernels of step i+1

    def drain():
        while pending:
            pending.pop(0)[1].wait()

    def step_resident():
        res = eng.match_packed(resident, P, 0.8)
        if world > 1:
            drain()
            pending.append(gather_counts(res.counts, P * world, async_op=True))
        return res.counts

    out_host = {"m": torch.empty(P * L, dtype=torch.int32).pin_memory(), "c": torch.empty(P, dtype=torch.int32).pin_memory()}

    def step_e2e():
        m0, cnt, _ = eng.match_packed_host(host, P, 0.8, n_chunks=args.e2e_chunks)
        out_host["m"].copy_(m0, non_blocking=True)
        out_host["c"].copy_(cnt, non_blocking=True)
        if world > 1:
            gather_counts(cnt, P * world)
        torch.cuda.current_stream().synchronize()   # the caller reads the result on the host

    for _ in range(args.warmup):
        step_resident()
    drain()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- timed region: K steps, inputs resident in HBM, CUDA events on the launching stream ----
    _native.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_resident()
    drain()            # the last step's all-gather is inside the timed region
    e1.record()
    barrier()
    launches = _native.launch_count()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.result()
    # ---- the same K steps again with a CUDA-event pair around EVERY kernel launch (per-class device
    #      time for the roofline).  Kept out of the region above because an event record between two
    #      kernels serialises them and would switch off the programmatic dependent launch overlap. ----
    _native.profile_begin()
    for _ in range(args.steps):
        step_resident()
    drain()
    barrier()
    prof = _native.profile_end()

    if args.profile_only:
        return
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0

    t = torch.tensor([ms_total, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(t[0]), float(t[1])
    if rank == 0:
        peaks = load_peaks()
        ms_step = ms_total / args.steps
        value = P * world / (ms_step / 1e3)
        # dominant kernel class by device time
        dom = max(prof.items(), key=lambda kv: kv[1][0]) if prof else ("none", (0.0, 0))
        dom_name, (dom_ms, dom_launches) = dom
        shares = {k: round(v[0] / max(sum(x[0] for x in prof.values()), 1e-9), 4) for k, v in prof.items()}
        class_flops = {"linear": 2 * P * gemm_flops_per_image(L, T), "token_fused": 2 * P * token_flops_per_image(L, T),
                       "sig_attention": 2 * P * 7 * 1024 * L * L}
        class_kernel = {"linear": "gemm_img_kernel (tcgen05, split-bf16 x3, TMA-fed tile images)",
                        "token_fused": "token_fused_kernel (tcgen05 + CUDA-core pooling)",
                        "sig_attention": "sig_attention_tc_kernel (tcgen05)"}
        useful_flops_step = 2 * P * flops_per_image(L, T) + P * 512 * L * L
        roof = None
        if dom_name in class_flops and dom_launches:
            per_launch_flops = class_flops[dom_name] * args.steps / dom_launches
            avg_ms = dom_ms / dom_launches
            ach = per_launch_flops / (avg_ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": class_kernel[dom_name], "achieved": ach,
                    "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                    "frac": ach / peaks["bf16_tflops_sustained"], "traffic": TRAFFIC_NCU.get(dom_name),
                    "peak_source": f"{peaks['source']} bf16 sustained (MEASURED_PEAKS.json); useful FLOPs only",
                    "launches_per_step": dom_launches / args.steps, "avg_launch_ms": avg_ms,
                    "algorithmic_flops_per_launch": per_launch_flops}
        # secondary rooflines (same live per-class times): the attention kernels the metric names
        # ("attn roofline %") against the tensor peak, and the token stage against HBM - its one
        # mandatory stream is the sampled descriptors, everything else stays on chip.
        by_class = {}
        for name in ("sig_attention", "token_fused", "linear"):
            if name in prof and prof[name][1]:
                ms_c, n_c = prof[name]
                tf = class_flops[name] * args.steps / (ms_c * 1e-3) / 1e12
                by_class[name] = {"kernel": class_kernel[name], "launches_per_step": n_c / args.steps,
                                  "avg_launch_ms": ms_c / n_c, "useful_tflops": tf,
                                  "frac_tensor_peak": tf / peaks["bf16_tflops_sustained"]}
        if "token_fused" in by_class:
            tok_bytes = 2 * P * bytes_per_image(L, T) - 2 * P * 4 * 256 * L + 2 * P * 4 * 1024 * L   # inputs + z image (hi/lo bf16)
            gbs = tok_bytes * args.steps / (prof["token_fused"][0] * 1e-3) / 1e9
            by_class["token_fused"].update({"algorithmic_bytes_per_launch": tok_bytes * args.steps / prof["token_fused"][1],
                                            "hbm_gbs": gbs, "frac_hbm_peak": gbs / peaks["hbm_gbs"]})
        cpu_v, cpu_n, cores, note = cpu_reference_pairs_per_s(L, T, args.cpu_budget)
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic (seeded inputs, random-init weights)",
               "config": config, "clocks": clocks, "gpu_launches": int(launches),
               "e2e": {"value": P * world / (e2e_ms / 1e3 / args.steps), "unit": UNIT,
                       "h2d_bytes_per_step": host.nbytes(),
                       "d2h_bytes_per_step": out_host["m"].numel() * 4 + out_host["c"].numel() * 4},
               "roofline": roof,
               "useful_tflops": useful_flops_step / (ms_step * 1e-3) / 1e12,
               "hbm_gbs_algorithmic": (2 * P * bytes_per_image(L, T) + 8 * P * L) / (ms_step * 1e-3) / 1e9,
               "kernel_time_shares": shares,
               "roofline_by_class": by_class,
               "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{cpu_n} pairs of {L}x{T}, B=1 per call; {note}"}}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
