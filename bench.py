#!/usr/bin/env python
"""Benchmark of WalkGPT's pixel-grounding forward path (BASELINE.json: grounding images/sec @448^2, batch 64 per GPU).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference's own fp32 PyTorch CPU path (oracle restatement)

One "step" = one pass of the whole hot path over one batch of 64 synthetic 448x448 images with 3 [SEG] hidden states per
image (config 2 of BASELINE.json): CLIP ViT-L/14@448 tower (23 layers) -> MSQP, out_mm_projector + neck -> CTP ->
prompt encoder + mask decoder -> bilinear upsample + threshold + score (+ depth extension).
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))

# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on the first communicator
# when NCCL_DEBUG is set), so file descriptor 1 is pointed at stderr for the whole run and the result line goes to the saved one.
_RESULT_FD = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_RESULT_FD, (json.dumps(line) + "\n").encode())

sys.path.insert(0, ROOT)

METRIC = "grounding_images_per_sec_448px_bs64"
UNIT = "images/s"
B_PER_GPU, S_PER_IMG, HIDDEN = 64, 3, 4096
WORKLOAD = "BASELINE.json configs[1]: ViT-L/14@448 tower + MSQP + out_mm_projector/neck + CTP + mask decoder + upsample/score, batch 64/GPU, 3 [SEG]/image, H=4096"
# algorithmic FLOPs per image
FLOPS_PER_IMAGE = 800.7e9


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_burst": d["bf16_tflops"], "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "hbm": d["hbm_gbs"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(sm)}


def cpu_path_a(n_images: int, threads: int):
    """Time the reference's fp32 CPU path (oracle restatement of the reference modules) on n_images images, 3 [SEG] each."""
    import torch
    from oracle import path_a
    from walkgpt_b200 import specs
    from walkgpt_b200.modules import clip_param_spec

    torch.set_num_threads(threads)
    W = {"clip": specs.make_state_dict(clip_param_spec(), 0), "msqp": specs.make_state_dict(specs.msqp_spec(1024, HIDDEN), 0),
         "proj": specs.make_state_dict(specs.out_mm_projector_spec(1024, HIDDEN), 0), "neck": specs.make_state_dict(specs.neck_spec(HIDDEN), 0),
         "ctp": specs.make_state_dict(specs.ctp_spec(HIDDEN, 256), 0), "prompt": specs.make_state_dict(specs.prompt_encoder_spec(), 0),
         "decoder": specs.make_state_dict(specs.mask_decoder_multiscale_spec(), 0)}
    g = torch.Generator().manual_seed(1234)
    px = torch.randn(n_images, 3, 448, 448, generator=g)
    seg = torch.randn(n_images * S_PER_IMG, HIDDEN, generator=g)
    offs = list(range(0, n_images * S_PER_IMG + 1, S_PER_IMG))

    def once():
        t0 = time.perf_counter()
        with torch.no_grad():
            path_a.path_a_forward(W, px, seg, offs)
        return time.perf_counter() - t0

    return once


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path (fp32 PyTorch eager, all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    n_img = 2  # bounded sample per step: 2 of the 64 images of the workload (the path is per-image independent)
    once = cpu_path_a(n_img, threads)
    for _ in range(max(1, min(args.warmup, 1))):
        once()
    steps = max(1, min(args.steps, 5))
    t = [once() for _ in range(steps)]
    sec = sum(t) / len(t)
    val = n_img / sec
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": 1,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": f"{n_img} images x {S_PER_IMG} [SEG] per step on the host CPU"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{n_img} images x {S_PER_IMG} [SEG] per step, fp32 PyTorch eager, torch {torch.__version__}"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seg", type=int, default=S_PER_IMG, help="[SEG] tokens per image (3 = config 2, 12 = config 3)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gather", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from walkgpt_b200 import _lib
    from walkgpt_b200.modules import GroundingPath

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    _lib.require_device(local)
    lib = _lib.lib()
    peaks = load_peaks()
    W = max(args.warmup, 3)
    K = max(args.steps, 1)
    B, S, H = B_PER_GPU, args.seg, HIDDEN

    model = GroundingPath(hidden_size=H, clip_layers=24, seed=0).to(dev)
    gen = torch.Generator().manual_seed(1234 + rank)
    NBUF = 3  # distinct input batches rotated between steps (plus per-step activations of several GB >> 126 MB L2)
    host_px = [torch.randn(B, 3, 448, 448, generator=gen).to(torch.bfloat16).pin_memory() for _ in range(NBUF)]
    host_seg = [torch.randn(B * S, H, generator=gen).to(torch.bfloat16).pin_memory() for _ in range(NBUF)]
    dev_px = [t.to(dev) for t in host_px]
    dev_seg = [t.to(dev) for t in host_seg]
    offs = list(range(0, B * S + 1, S))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(i):
        return model(dev_px[i % NBUF], dev_seg[i % NBUF], offs)

    # ---------------- device-resident throughput ("value") ----------------
    for i in range(W):
        step_resident(i)
    barrier()
    lib.wg_launch_count(1)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        step_resident(i)
    e1.record()
    barrier()
    launches = lib.wg_launch_count(0)
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = ms.item()
    value = world * B * K / (ms_total / 1e3)

    # ---------------- end-to-end through the public API with HOST buffers ("e2e") ----------------
    # Every step copies ITS inputs from pinned host memory and returns ITS results to pinned host memory inside the timed
    # region.  The loop is the double-buffered serving loop a caller would write around GroundingPath.forward: the H2D copy
    # of step i+1 and the D2H copy of step i-1 run on their own streams while step i computes; the host consumes (waits
    # for) the results of step i-1 before it launches step i+1, and the region ends when the last result is on the host.
    NB2 = 2
    h_masks = [torch.empty(B * S, 448, 448, dtype=torch.uint8).pin_memory() for _ in range(NB2)]
    h_scores = [torch.empty(B * S, dtype=torch.float32).pin_memory() for _ in range(NB2)]
    h_iou = [torch.empty(B * S, 1, dtype=torch.float32).pin_memory() for _ in range(NB2)]
    h_depth = [torch.empty(B * S, dtype=torch.float32).pin_memory() for _ in range(NB2)]
    h_vis = [torch.empty(B, 36, H, dtype=torch.bfloat16).pin_memory() for _ in range(NB2)]
    d_px = [torch.empty_like(dev_px[0]) for _ in range(NB2)]
    d_seg = [torch.empty_like(dev_seg[0]) for _ in range(NB2)]
    s_main = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in range(NB2)]        # inputs of the step using buffer b are on the device
    ev_used = [torch.cuda.Event() for _ in range(NB2)]      # the step using input buffer b has finished reading it
    ev_out = [torch.cuda.Event() for _ in range(NB2)]       # results of the step using host buffer b are on the host

    def issue_h2d(i):
        b = i % NB2
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_used[b])
            d_px[b].copy_(host_px[i % NBUF], non_blocking=True)
            d_seg[b].copy_(host_seg[i % NBUF], non_blocking=True)
            ev_in[b].record(s_in)

    def run_e2e(n):
        for b in range(NB2):
            ev_used[b].record(s_main)
        e0.record(s_in)  # device timestamp right before the first H2D copy
        issue_h2d(0)
        for i in range(n):
            b = i % NB2
            if i + 1 < n:
                issue_h2d(i + 1)
            s_main.wait_event(ev_in[b])
            out = model(d_px[b], d_seg[b], offs)
            ev_used[b].record(s_main)
            done = torch.cuda.Event()
            done.record(s_main)
            if i >= NB2:
                ev_out[b].synchronize()  # host buffers b are about to be overwritten: their previous results were consumed
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                for h, key in ((h_masks, "masks"), (h_scores, "scores"), (h_iou, "iou"), (h_depth, "depth"), (h_vis, "vis_tokens")):
                    out[key].record_stream(s_out)
                    h[b].copy_(out[key], non_blocking=True)
                ev_out[b].record(s_out)
            if i >= 1:
                ev_out[(i - 1) % NB2].synchronize()  # the caller consumes the result of step i-1 while step i runs
        e1.record(s_out)  # device timestamp right after the last D2H copy
        ev_out[(n - 1) % NB2].synchronize()

    h2d = d_px[0].numel() * 2 + d_seg[0].numel() * 2
    d2h = h_masks[0].numel() + 4 * (h_scores[0].numel() + h_iou[0].numel() + h_depth[0].numel()) + 2 * h_vis[0].numel()
    run_e2e(2)
    barrier()
    t0 = time.perf_counter()
    run_e2e(K)
    torch.cuda.synchronize()
    e2e_ms_host = (time.perf_counter() - t0) * 1e3   # host clock cross-check: first H2D issued -> last result on the host
    e2e_ms = e0.elapsed_time(e1)                     # CUDA events: before the first H2D copy -> after the last D2H copy
    barrier()
    ms2 = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = world * B * K / (ms2.item() / 1e3)

    # ---------------- NCCL gather of masks + scores for scoring (off the hot path) ----------------
    gather_ms = None
    if world > 1 and not args.no_gather:
        out = step_resident(0)
        bufs_m = [torch.empty_like(out["masks"]) for _ in range(world)]
        bufs_s = [torch.empty_like(out["scores"]) for _ in range(world)]
        for _ in range(2):
            dist.all_gather(bufs_m, out["masks"])
            dist.all_gather(bufs_s, out["scores"])
        barrier()
        e0.record()
        dist.all_gather(bufs_m, out["masks"])
        dist.all_gather(bufs_s, out["scores"])
        e1.record()
        barrier()
        g = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dist.all_reduce(g, op=dist.ReduceOp.MAX)
        gather_ms = g.item()

    # ---------------- per-kernel CUDA-event profile over the same step (roofline leg) ----------------
    lib.wg_profile_enable(1)
    PK = min(K, 3)
    for i in range(PK):
        step_resident(i)
    torch.cuda.synchronize()
    import ctypes
    buf = ctypes.create_string_buffer(1 << 16)
    lib.wg_profile_collect(buf, len(buf))
    lib.wg_profile_enable(0)
    kern = {}
    for ln in buf.value.decode().splitlines():
        name, n, tms, fl, by = ln.split()
        kern[name] = {"launches": int(n) // PK, "ms_per_step": float(tms) / PK, "flops_per_step": float(fl) / PK, "bytes_per_step": float(by) / PK}
    tot_ms = sum(k["ms_per_step"] for k in kern.values()) or 1.0
    # the dominant kernel family: all instantiations of the tcgen05 GEMM
    gemm = [k for n, k in kern.items() if n.startswith("gemm_") or n.startswith("gemm2_")]
    g_ms = sum(k["ms_per_step"] for k in gemm)
    g_fl = sum(k["flops_per_step"] for k in gemm)
    g_n = sum(k["launches"] for k in gemm)
    achieved = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    peak = peaks["bf16_sustained"]
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "r01c_ncu_full_summary.json")
    if os.path.exists(tpath):  # dram__bytes_read+write of the dominant launch (ViT fc1), one `ncu --set full` capture
        tj = json.load(open(tpath))["gemm2_fc1"]
        traffic = tj["dram_bytes"]
        traffic_note = f"ncu dram bytes of the fc1 launch ({tj['shape']}); algorithmic bytes of that launch {tj['algorithmic_bytes']}"
    roofline = {"kernel": "gemm2_bf16_kernel (CTA pair, tcgen05 cta_group::2) + gemm_bf16_kernel (1 CTA): tcgen05/TMEM + TMA, all epilogue variants", "bound": "tensor", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note, "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
                "flops_per_launch": g_fl / max(g_n, 1), "avg_launch_ms": g_ms / max(g_n, 1), "launches_per_step": g_n,
                "share_of_step": g_ms / tot_ms,
                "step_tflops": FLOPS_PER_IMAGE * B / (ms_total / K * 1e-3) / 1e12, "step_frac_of_peak": FLOPS_PER_IMAGE * B / (ms_total / K * 1e-3) / 1e12 / peak}
    att = kern.get("attention_d64")
    tail = kern.get("postprocess_bilinear_score")
    extra = {}
    if att:
        extra["attention_d64"] = {"tflops": att["flops_per_step"] / (att["ms_per_step"] * 1e-3) / 1e12, "ms_per_step": att["ms_per_step"],
                                  "share_of_step": att["ms_per_step"] / tot_ms}
    if tail:
        gbs = tail["bytes_per_step"] / (tail["ms_per_step"] * 1e-3) / 1e9
        extra["postprocess_bilinear_score"] = {"bound": "hbm", "achieved_gbs": gbs, "peak_gbs": peaks["hbm"], "frac": gbs / peaks["hbm"],
                                               "ms_per_step": tail["ms_per_step"]}
    extra["kernel_ms_per_step"] = {n: round(k["ms_per_step"], 4) for n, k in sorted(kern.items(), key=lambda kv: -kv[1]["ms_per_step"])}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        once = cpu_path_a(2, threads)
        once()
        ts = [once() for _ in range(3)]
        sec = min(ts)
        cpu_baseline = {"value": 2 / sec, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": "2 images x 3 [SEG] (of the 64-image step), fp32 PyTorch eager oracle of the reference modules, best of 3 after 1 warm-up"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": WORKLOAD if S == 3 else WORKLOAD.replace("3 [SEG]", f"{S} [SEG]"), "batch_per_gpu": B, "seg_per_image": S,
                           "hidden": H, "clip_layers_run": 23, "weights": "random-init, bf16",
                           "l2": f"{NBUF} input batches rotated; per-step activation traffic (several GB) exceeds the 126 MB L2",
                           "parallelism": f"dp{world} (images sharded, no collective on the hot path)"},
                "clocks": clocks, "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                          "ms_per_step": ms2.item() / K, "ms_per_step_host_clock": e2e_ms_host / K,
                                          "loop": "double-buffered: H2D of step i+1 and D2H of step i-1 on side streams overlap step i; results consumed one step late"},
                "gpu_launches": int(launches), "roofline": roofline, "kernels": extra}
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        if gather_ms is not None:
            line["gather_ms"] = gather_ms
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
