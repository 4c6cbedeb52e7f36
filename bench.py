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
PATH_DESC = "ViT-L/14@448 tower + MSQP + out_mm_projector/neck + CTP + mask decoder + upsample/score"
# BASELINE.json configs (index + 1): batch per GPU per step, micro-batches per step, [SEG] per image, LLM width, algorithmic FLOPs per
# image (ViT 693.5 + MSQP 15.55/15.62 + projector 85.9/128.8 + neck 3.36/3.89 GF + 0.792 GF and 4.46/5.51 MF per [SEG])
CONFIGS = {
    2: dict(B=64, micro=1, S=3, H=4096, flops=800.7e9, name=f"BASELINE.json configs[1]: {PATH_DESC}, batch 64/GPU, 3 [SEG]/image, H=4096"),
    3: dict(B=64, micro=1, S=12, H=4096, flops=807.9e9, name=f"BASELINE.json configs[2] (multi-target grounding): {PATH_DESC}, batch 64/GPU, 12 [SEG]/image, H=4096"),
    5: dict(B=64, micro=4, S=16, H=5120, flops=854.6e9,
            name=f"BASELINE.json configs[4]: {PATH_DESC}, batch 256/GPU as 4 micro-batches of 64, 16 [SEG]/image, H=5120 (LLaVA-13B width), "
                 "[SEG] hidden states produced by an LLM prefill that is fed this path's own MSQP visual tokens"),
}
WORKLOAD = CONFIGS[2]["name"]
FLOPS_PER_IMAGE = CONFIGS[2]["flops"]
SEG_TOKEN_ID, IMAGE_SLOT = 32003, 1  # [SEG] id as in the reference's tokenizer extension; position of the <image> token in a row


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


def cpu_path_a(n_images: int, threads: int, S_PER_IMG: int = S_PER_IMG, HIDDEN: int = HIDDEN):
    """Time the reference's fp32 CPU path on n_images images, S [SEG] each.  The CLIP tower -- 87 % of the path's FLOPs -- runs through
    the library the reference itself calls (HF transformers ``CLIPVisionModel``, eager attention, ``hidden_states[-2][:, 1:]`` as in
    clip_encoder.py:61-98) when transformers provides it; MSQP / projector / neck / CTP / prompt encoder / mask decoder / postprocess are
    the oracle's restatement of the reference modules (the reference's own Python sources do not exist on the GPU box)."""
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
    hf = None
    try:
        from transformers import CLIPVisionConfig, CLIPVisionModel
        cfg = CLIPVisionConfig(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16, image_size=448, patch_size=14)
        cfg._attn_implementation = "eager"
        hf = CLIPVisionModel(cfg).eval()
        hf.load_state_dict({k[len("vision_tower."):]: v for k, v in W["clip"].items()}, strict=True)
    except Exception:  # noqa: BLE001  (no transformers / incompatible version: the oracle's restatement of the same arithmetic)
        hf = None
    cpu_path_a.clip_impl = "HF transformers CLIPVisionModel (the reference's own dependency)" if hf is not None else "oracle restatement"

    def once():
        t0 = time.perf_counter()
        with torch.no_grad():
            if hf is not None:
                f_last = hf(px, output_hidden_states=True).hidden_states[-2][:, 1:]
                path_a.path_a_forward(W, px, seg, offs, f_last=f_last)
            else:
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
    cfg = CONFIGS[getattr(args, "config", 2)]
    S_PER_IMG, WORKLOAD = (args.seg if getattr(args, "seg", None) is not None else cfg["S"]), cfg["name"]
    once = cpu_path_a(n_img, threads, S_PER_IMG, cfg["H"])
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
                             "sample": f"{n_img} images x {S_PER_IMG} [SEG] per step, fp32 PyTorch eager, torch {torch.__version__}; CLIP tower: "
                                       f"{cpu_path_a.clip_impl}; other modules: oracle port of the reference modules"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


class LlmPrefill:
    """The LLM half of config 5.  The reference's LLM is HF transformers' Llama (``LlavaLlamaForCausalLM`` subclasses
    ``LlamaForCausalLM``, model/llava_walkgpt/model/language_model/llava_llama.py:64-143) and stays PyTorch: here a random-init
    ``LlamaModel`` of 13B width (hidden 5120, 40 heads, MLP 13824) at reduced depth, eager bf16 on the GPU.  It is fed the way
    ``prepare_inputs_labels_for_multimodal`` feeds it (llava_arch.py:252-259): the <image> token of every row is replaced by the 256
    visual tokens = this path's 36 MSQP tokens resampled 6x6 -> 16x16 (``wg_resample_tokens``); the [SEG] rows of the last hidden state
    are extracted with the reference's shifted mask (model/walkgpt.py:287-306, 406-420; ``wg_seg_gather``).  Not part of this repo's
    path: it only PRODUCES the path's [SEG] input, and it is timed separately (``with_llm_prefill``)."""

    def __init__(self, H, S, B, dev, depth=2, text_len=64, seed=0):
        import torch
        from transformers import LlamaConfig, LlamaModel

        heads = H // 128
        cfg = LlamaConfig(hidden_size=H, intermediate_size=13824 if H == 5120 else 11008, num_hidden_layers=depth, num_attention_heads=heads,
                          num_key_value_heads=heads, vocab_size=32008, max_position_embeddings=2048, attn_implementation="sdpa")
        torch.manual_seed(seed)
        self.llm = LlamaModel(cfg).to(dtype=torch.bfloat16, device=dev).eval()
        self.depth, self.text_len, self.S, self.B, self.H = depth, text_len, S, B, H
        g = torch.Generator().manual_seed(seed + 1)
        ids = torch.randint(1000, 30000, (B, text_len), generator=g)
        for r in range(B):  # S [SEG] tokens per row at distinct positions after the <image> slot
            pos = torch.randperm(text_len - IMAGE_SLOT - 2, generator=g)[:S] + IMAGE_SLOT + 2
            ids[r, pos] = SEG_TOKEN_ID
        ids[:, IMAGE_SLOT] = 0  # placeholder row of the embedding table; replaced by the visual tokens
        self.input_ids = ids.to(dev)
        self.rows = list(range(B + 1))

    def seg_states(self, vis_tokens):
        """vis_tokens bf16 [B, 36, H] -> ([SEG] hidden states bf16 [B*S, H], per-image offsets int32 [B+1] on the device)."""
        import torch
        from walkgpt_b200 import ops

        vis256 = ops.resample_tokens(vis_tokens, 16)
        with torch.no_grad():
            emb = self.llm.embed_tokens(self.input_ids)
            x = torch.cat([emb[:, :IMAGE_SLOT], vis256.to(emb.dtype), emb[:, IMAGE_SLOT + 1:]], dim=1)
            hidden = self.llm(inputs_embeds=x, use_cache=False).last_hidden_state.contiguous()
        rows, _, _, img_off = ops.seg_gather(hidden, self.input_ids, SEG_TOKEN_ID, offset=self.rows, shift=255, max_out=self.B * self.S)
        return rows, img_off


# algorithmic FLOPs per image of Path B: patch embed 8.05 + 32 blocks x 161.06 (linear layers over 4096 tokens) + 28 x 4.92
# (windowed attention, 25 windows of 196 tokens incl. the pad keys) + 4 x 85.90 (global attention) + neck 7.52 GF = 5650.8 GF for the
# encoder; MSQP(sam_dim 256, 4096 tokens) 50.8 GF; SAM mask decoder 3.61 GF per [SEG]
FLOPS_SAM_ENCODER = 5650.8e9
FLOPS_PATH_B_REST = 50.8e9


def run_path_b(args):
    """Path B from the pixels on ONE GPU: not BASELINE.json's headline (that is --path a), reported for SURVEY 8(f) row 1."""
    import ctypes

    import torch
    from walkgpt_b200 import _lib
    from walkgpt_b200.modules import GroundingPathB, ImageEncoderViT

    local = int(os.environ.get("LOCAL_RANK", "0"))
    if int(os.environ.get("RANK", "0")) != 0:
        return
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    _lib.require_device(local)
    lib = _lib.lib()
    peaks = load_peaks()
    B, S, H = args.batch or 16, 3, 4096
    W, K = max(args.warmup, 3), max(args.steps, 1)
    model = GroundingPathB(hidden_size=H, seed=0, image_encoder=ImageEncoderViT(seed=0)).to(dev)
    gen = torch.Generator().manual_seed(1234)
    NBUF = 3
    host_px = [torch.randn(B, 3, 1024, 1024, generator=gen).to(torch.bfloat16).pin_memory() for _ in range(NBUF)]
    host_seg = [torch.randn(B * S, H, generator=gen).to(torch.bfloat16).pin_memory() for _ in range(NBUF)]
    dev_px, dev_seg = [t.to(dev) for t in host_px], [t.to(dev) for t in host_seg]
    offs = list(range(0, B * S + 1, S))

    def step(i):
        return model.forward_from_pixels(dev_px[i % NBUF], dev_seg[i % NBUF], offs)

    for i in range(W):
        step(i)
    torch.cuda.synchronize()
    lib.wg_launch_count(1)
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    launches = lib.wg_launch_count(0)
    clocks = sampler.stop()
    value = B * K / (ms_total / 1e3)
    # end to end with host buffers (single-buffered: H2D of the step's pixels + [SEG] states, D2H of its masks / scores / IoU)
    h_masks = torch.empty(B * S, 1024, 1024, dtype=torch.uint8).pin_memory()
    h_scores = torch.empty(B * S, dtype=torch.float32).pin_memory()
    h_iou = torch.empty(B * S, 1, dtype=torch.float32).pin_memory()
    d_px, d_seg = torch.empty_like(dev_px[0]), torch.empty_like(dev_seg[0])

    def step_e2e(i):
        d_px.copy_(host_px[i % NBUF], non_blocking=True)
        d_seg.copy_(host_seg[i % NBUF], non_blocking=True)
        out = model.forward_from_pixels(d_px, d_seg, offs)
        h_masks.copy_(out["masks"], non_blocking=True)
        h_scores.copy_(out["scores"], non_blocking=True)
        h_iou.copy_(out["iou"], non_blocking=True)

    step_e2e(0)
    torch.cuda.synchronize()
    e0.record()
    for i in range(K):
        step_e2e(i)
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1)
    lib.wg_profile_enable(1)
    PK = min(K, 2)
    for i in range(PK):
        step(i)
    torch.cuda.synchronize()
    buf = ctypes.create_string_buffer(1 << 16)
    lib.wg_profile_collect(buf, len(buf))
    lib.wg_profile_enable(0)
    kern = {}
    for ln in buf.value.decode().splitlines():
        name, n, tms, fl, by = ln.split()
        kern[name] = {"launches": int(n) // PK, "ms_per_step": float(tms) / PK, "flops_per_step": float(fl) / PK, "bytes_per_step": float(by) / PK}
    tot_ms = sum(k["ms_per_step"] for k in kern.values()) or 1.0
    gemm = [k for n, k in kern.items() if n.startswith("gemm_") or n.startswith("gemm2_")]
    g_ms, g_fl, g_n = sum(k["ms_per_step"] for k in gemm), sum(k["flops_per_step"] for k in gemm), sum(k["launches"] for k in gemm)
    peak = peaks["bf16_sustained"]
    flops_img = FLOPS_SAM_ENCODER + FLOPS_PATH_B_REST + S * 3.61e9
    step_tf = flops_img * B / (ms_total / K * 1e-3) / 1e12
    extra = {}
    for name in ("sam_attention_window", "sam_attention_global"):
        if name in kern:
            k = kern[name]
            extra[name] = {"tflops": k["flops_per_step"] / (k["ms_per_step"] * 1e-3) / 1e12, "ms_per_step": k["ms_per_step"], "share_of_step": k["ms_per_step"] / tot_ms}
    extra["kernel_ms_per_step"] = {n: round(k["ms_per_step"], 4) for n, k in sorted(kern.items(), key=lambda kv: -kv[1]["ms_per_step"])}
    emit({"metric": "grounding_images_per_sec_1024px_path_b", "value": value, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms_total / K,
          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
          "config": {"workload": f"Path B (released SAM-1024 wiring) from the pixels: SAM ViT-H encoder (32 blocks, 1280 wide, 14 x 14 windows + 4 global blocks) + "
                                 f"MSQP(256) + CTP + SAM mask decoder + postprocess to 1024^2, batch {B}/GPU, {S} [SEG]/image, H={H}",
                     "batch_per_gpu": B, "seg_per_image": S, "hidden": H, "weights": "random-init, bf16", "flops_per_image": flops_img,
                     "l2": f"{NBUF} input batches rotated; per-step activations (GBs) exceed the 126 MB L2"},
          "clocks": clocks, "e2e": {"value": B * K / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": d_px.numel() * 2 + d_seg.numel() * 2,
                                    "d2h_bytes_per_step": h_masks.numel() + 8 * h_scores.numel(), "ms_per_step": e2e_ms / K, "loop": "single-buffered"},
          "gpu_launches": int(launches),
          "roofline": {"kernel": "tcgen05 GEMM family (gemm2_bf16_kernel / gemm_bf16_kernel)", "bound": "tensor", "achieved": g_fl / (g_ms * 1e-3) / 1e12 if g_ms else 0.0,
                       "peak": peak, "unit": "TFLOP/s", "frac": (g_fl / (g_ms * 1e-3) / 1e12 / peak) if g_ms else 0.0, "traffic": None,
                       "peak_source": peaks["source"] + ", sustained figure", "launches_per_step": g_n, "share_of_step": g_ms / tot_ms,
                       "step_tflops": step_tf, "step_frac_of_peak": step_tf / peak},
          "kernels": extra})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json config number (2 = the headline; 3; 5)")
    ap.add_argument("--seg", type=int, default=None, help="override the [SEG] tokens per image of the chosen config")
    ap.add_argument("--llm-depth", type=int, default=2, help="config 5: decoder layers of the stand-in LLM prefill")
    ap.add_argument("--path", default="a", choices=["a", "b"], help="a = the benchmarked CLIP-448 composition (BASELINE.json); b = the released "
                    "SAM-1024 wiring from the pixels (SAM ViT-H encoder + MSQP + CTP + SAM mask decoder), SURVEY 8(f) row 1")
    ap.add_argument("--batch", type=int, default=None, help="--path b: images per step (default 16)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gather", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.path == "b":
        return run_path_b(args)

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
    cfg = CONFIGS[args.config]
    B, MICRO, H = cfg["B"], cfg["micro"], cfg["H"]
    S = args.seg if args.seg is not None else cfg["S"]
    flops_per_image = cfg["flops"] + (S - cfg["S"]) * (0.792e9 + (4.46e6 if H == 4096 else 5.51e6))
    workload = cfg["name"] if S == cfg["S"] else cfg["name"].replace(f"{cfg['S']} [SEG]", f"{S} [SEG]")

    model = GroundingPath(hidden_size=H, clip_layers=24, seed=0).to(dev)
    gen = torch.Generator().manual_seed(1234 + rank)
    NBUF = max(3, MICRO)  # distinct input batches rotated between micro-steps (plus per-step activations of several GB >> 126 MB L2)
    host_px = [torch.randn(B, 3, 448, 448, generator=gen).to(torch.bfloat16).pin_memory() for _ in range(NBUF)]
    dev_px = [t.to(dev) for t in host_px]
    offs = list(range(0, B * S + 1, S))
    llm = None
    if args.config == 5:
        # the [SEG] hidden states are what the LLM prefill emits for THESE images (visual tokens from this path's tower + MSQP)
        llm = LlmPrefill(H, S, B, dev, depth=args.llm_depth)
        dev_seg = []
        for i in range(NBUF):
            rows, img_off = llm.seg_states(model.encode_images(dev_px[i])["vis_tokens"])
            assert img_off.tolist() == offs, "LLM prefill + [SEG] extraction did not give S rows per image"
            dev_seg.append(rows.clone())
        host_seg = [t.cpu().pin_memory() for t in dev_seg]
    else:
        host_seg = [torch.randn(B * S, H, generator=gen).to(torch.bfloat16).pin_memory() for _ in range(NBUF)]
        dev_seg = [t.to(dev) for t in host_seg]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(i):  # one micro-batch
        return model(dev_px[i % NBUF], dev_seg[i % NBUF], offs)

    def timed(fn, n_steps):
        """n_steps steps of MICRO micro-batches each, CUDA events on the launching stream, max over ranks -> total ms."""
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n_steps * MICRO):
            fn(i)
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ---------------- device-resident throughput ("value") ----------------
    for i in range(W * MICRO):
        step_resident(i)
    barrier()
    lib.wg_launch_count(1)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_resident, K)
    launches = lib.wg_launch_count(0)
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * MICRO * K / (ms_total / 1e3)

    # config 5: the same step with the LLM prefill + [SEG] extraction between the two halves of the path (reported beside `value`)
    with_llm = None
    if llm is not None:
        def step_llm(i):
            enc = model.encode_images(dev_px[i % NBUF])
            rows, img_off = llm.seg_states(enc["vis_tokens"])
            out = model.ground(enc["img_emb_split"], rows, img_off)  # device offsets: wg_prompt_index, no host synchronisation
            out.update(enc)
            return out
        for i in range(2 * MICRO):
            step_llm(i)
        ms_llm = timed(step_llm, K)
        with_llm = {"value": world * B * MICRO * K / (ms_llm / 1e3), "unit": UNIT, "ms_per_step": ms_llm / K,
                    "llm": f"HF transformers LlamaModel, random-init, hidden {H}, {H // 128} heads, MLP 13824, {args.llm_depth} layers (the real 13B has 40), "
                           f"bf16 eager + SDPA on the GPU, {llm.text_len + 255} positions per row",
                    "llm_ms_per_step": (ms_llm - ms_total) / K}

    # ---------------- end-to-end through the public API with HOST buffers ("e2e") ----------------
    # Every micro-step copies ITS inputs from pinned host memory and returns ITS results to pinned host memory inside the timed
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
    s_in, s_out, s_gather = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in range(NB2)]        # inputs of the step using buffer b are on the device
    ev_used = [torch.cuda.Event() for _ in range(NB2)]      # the step using input buffer b has finished reading it
    ev_out = [torch.cuda.Event() for _ in range(NB2)]       # results of the step using host buffer b are on the host
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # scoring gather: every rank receives all ranks' masks + scores + IoU + depths of the step
    GATHER_KEYS = ("masks", "scores", "iou", "depth")
    g_bufs = None
    if world > 1:
        g_bufs = [{"masks": torch.empty(world * B * S, 448, 448, dtype=torch.uint8, device=dev),
                   "scores": torch.empty(world * B * S, dtype=torch.float32, device=dev),
                   "iou": torch.empty(world * B * S, 1, dtype=torch.float32, device=dev),
                   "depth": torch.empty(world * B * S, dtype=torch.float32, device=dev)} for _ in range(NB2)]
    ev_gathered = [torch.cuda.Event() for _ in range(NB2)]

    def issue_h2d(i):
        b = i % NB2
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_used[b])
            d_px[b].copy_(host_px[i % NBUF], non_blocking=True)
            d_seg[b].copy_(host_seg[i % NBUF], non_blocking=True)
            ev_in[b].record(s_in)

    def run_e2e(n, gather=False):
        for b in range(NB2):
            ev_used[b].record(s_main)
            ev_gathered[b].record(s_main)
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
            if gather:  # NCCL all_gather over NVLink on its own stream, overlapping the next step's compute
                with torch.cuda.stream(s_gather):
                    s_gather.wait_event(done)
                    for key in GATHER_KEYS:
                        out[key].record_stream(s_gather)
                        dist.all_gather_into_tensor(g_bufs[b][key], out[key].contiguous())
                    ev_gathered[b].record(s_gather)
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                for h, key in ((h_masks, "masks"), (h_scores, "scores"), (h_iou, "iou"), (h_depth, "depth"), (h_vis, "vis_tokens")):
                    out[key].record_stream(s_out)
                    h[b].copy_(out[key], non_blocking=True)
                if gather:
                    s_out.wait_event(ev_gathered[b])  # a step counts as finished when its gather has landed, too
                ev_out[b].record(s_out)
            if i >= 1:
                ev_out[(i - 1) % NB2].synchronize()  # the caller consumes the result of step i-1 while step i runs
        e1.record(s_out)  # device timestamp right after the last D2H copy
        ev_out[(n - 1) % NB2].synchronize()

    h2d = MICRO * (d_px[0].numel() * 2 + d_seg[0].numel() * 2)
    d2h = MICRO * (h_masks[0].numel() + 4 * (h_scores[0].numel() + h_iou[0].numel() + h_depth[0].numel()) + 2 * h_vis[0].numel())

    def measure_e2e(gather):
        run_e2e(2, gather)
        barrier()
        t0 = time.perf_counter()
        run_e2e(K * MICRO, gather)
        torch.cuda.synchronize()
        host_ms = (time.perf_counter() - t0) * 1e3   # host clock cross-check: first H2D issued -> last result on the host
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)  # CUDA events: before first H2D -> after last D2H
        barrier()
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), host_ms

    e2e_ms, e2e_ms_host = measure_e2e(False)
    e2e_value = world * B * MICRO * K / (e2e_ms / 1e3)

    # ---------------- the same end-to-end loop WITH the NCCL scoring gather (masks + scores + IoU + depths) on a side stream -----
    gather_info = None
    if world > 1 and not args.no_gather:
        g_ms, _ = measure_e2e(True)
        out = step_resident(0)
        barrier()
        e0.record()
        for key in GATHER_KEYS:
            dist.all_gather_into_tensor(g_bufs[0][key], out[key].contiguous())
        e1.record()
        barrier()
        g = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dist.all_reduce(g, op=dist.ReduceOp.MAX)
        per_rank = B * S * (448 * 448 + 12)
        gather_info = {"e2e_with_gather": world * B * MICRO * K / (g_ms / 1e3), "e2e_without_gather": e2e_value, "unit": UNIT,
                       "gather_ms_isolated": g.item(), "bytes_per_rank_per_step": per_rank * MICRO, "what": "all_gather of u8 masks, scores, IoU, depths "
                       "(NCCL over NVLink, side stream, overlapping the next step)",
                       "gather_gbs_per_gpu_isolated": per_rank * (world - 1) / (g.item() * 1e-3) / 1e9}

    # ---------------- per-kernel CUDA-event profile over the same step (roofline leg) ----------------
    lib.wg_profile_enable(1)
    PK = min(K, 3)
    for i in range(PK * MICRO):
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
    step_tflops = flops_per_image * B * MICRO / (ms_total / K * 1e-3) / 1e12
    roofline = {"kernel": "gemm2_bf16_kernel (CTA pair, tcgen05 cta_group::2) + gemm_bf16_kernel (1 CTA): tcgen05/TMEM + TMA, all epilogue variants", "bound": "tensor", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note, "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
                "flops": "algorithmic (2*M*N*K of the product each launch stands for: split-bf16 operands count once)",
                "flops_per_launch": g_fl / max(g_n, 1), "avg_launch_ms": g_ms / max(g_n, 1), "launches_per_step": g_n,
                "share_of_step": g_ms / tot_ms, "step_tflops": step_tflops, "step_frac_of_peak": step_tflops / peak}
    att = kern.get("attention_d64")
    tail = kern.get("postprocess_bilinear_score")
    extra = {}
    if att:
        a_tf = att["flops_per_step"] / (att["ms_per_step"] * 1e-3) / 1e12
        extra["attention_d64"] = {"tflops": a_tf, "frac_of_tensor_peak": a_tf / peak, "ms_per_step": att["ms_per_step"], "share_of_step": att["ms_per_step"] / tot_ms}
        # at head_dim 64 the softmax's exponentials, not the tensor pipe, bound the kernel: one exponential per score (flops / (4 * 64)) against the
        # MUFU unit's measured 16 results per clock and SM (tools/ubench/xu_rates.cu) at the SM clock sampled under load
        if clocks and clocks.get("sm_mhz"):
            mufu_peak = 16.0 * torch.cuda.get_device_properties(dev).multi_processor_count * clocks["sm_mhz"] * 1e6
            exps = att["flops_per_step"] / 256.0 / (att["ms_per_step"] * 1e-3)
            extra["attention_d64"].update({"bound": "mufu", "exp_per_s": exps, "mufu_peak_exp_per_s": mufu_peak, "frac_of_mufu_peak": exps / mufu_peak})
        apath = os.path.join(ROOT, "profiles", "r02ab_attention_q4_summary.json")
        if os.path.exists(apath):  # one `ncu --set full` capture of attention_d64_q4_kernel at B=64, T=1025, 16 heads
            extra["attention_d64"]["traffic"] = json.load(open(apath))["dram_bytes"]
            extra["attention_d64"]["traffic_note"] = "ncu dram bytes of one launch (B=64, T=1025, 16 heads); algorithmic bytes 4 * B * T * 1024 * 2 = 537 MB" 
    if tail:
        gbs = tail["bytes_per_step"] / (tail["ms_per_step"] * 1e-3) / 1e9
        extra["postprocess_bilinear_score"] = {"bound": "hbm", "achieved_gbs": gbs, "peak_gbs": peaks["hbm"], "frac": gbs / peaks["hbm"],
                                               "ms_per_step": tail["ms_per_step"]}
    dec_ms = sum(k["ms_per_step"] for n, k in kern.items() if n.startswith("dec_") or n in ("depth_head", "ctp_tail", "prompt_index"))
    extra["decoder_per_prompt_kernels"] = {"ms_per_step": dec_ms, "share_of_step": dec_ms / tot_ms, "prompts_per_step": B * S * MICRO}
    extra["kernel_ms_per_step"] = {n: round(k["ms_per_step"], 4) for n, k in sorted(kern.items(), key=lambda kv: -kv[1]["ms_per_step"])}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        once = cpu_path_a(2, threads, S, H)
        once()
        ts = [once() for _ in range(3)]
        sec = min(ts)
        cpu_baseline = {"value": 2 / sec, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"2 images x {S} [SEG] (of the {B * MICRO}-image step), fp32 PyTorch eager, best of 3 after 1 warm-up; CLIP tower: "
                                  f"{cpu_path_a.clip_impl}; other modules: oracle port of the reference modules"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": workload, "baseline_config": args.config, "batch_per_gpu": B * MICRO, "micro_batches_per_step": MICRO, "seg_per_image": S,
                           "hidden": H, "clip_layers_run": 23, "weights": "random-init, bf16",
                           "l2": f"{NBUF} input batches rotated; per-step activation traffic (several GB) exceeds the 126 MB L2",
                           "parallelism": f"dp{world} (images sharded, no collective on the hot path)"},
                "clocks": clocks, "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                          "ms_per_step": e2e_ms / K, "ms_per_step_host_clock": e2e_ms_host / K,
                                          "loop": "double-buffered: H2D of step i+1 and D2H of step i-1 on side streams overlap step i; results consumed one step late"},
                "gpu_launches": int(launches), "roofline": roofline, "kernels": extra}
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        if gather_info is not None:
            line["gather"] = gather_info
            line["gather_ms"] = gather_info["gather_ms_isolated"]
        if with_llm is not None:
            line["with_llm_prefill"] = with_llm
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
