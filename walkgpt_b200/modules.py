"""Host-side mirror of the reference's nn.Module interface for the pixel-grounding path.

Each class keeps the reference's constructor arguments, parameter names/shapes (so the reference's
``load_state_dict(strict=True)`` checkpoints load unchanged) and ``forward`` signature, but its forward is a
single call into the C ABI (through the ``walkgpt_b200::*`` torch custom ops in :mod:`walkgpt_b200.torch_ops`).
Inference only; there is no PyTorch/CPU fallback.

Reference seams: ``model/walkgpt.py:59-146`` (initialize_walkgpt_modules) and
``model/llava_walkgpt/model/llava_arch.py:31-42``.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib
from .specs import Spec, init_tensor


# --------------------------------------------------------------------------------------------------
# parameter-tree helpers: modules are declared by a flat {dotted name: shape} spec that matches the
# reference state_dict; nested nn.Module containers are created on the fly.
# --------------------------------------------------------------------------------------------------
def _add_tensor(root: nn.Module, dotted: str, tensor: torch.Tensor, buffer: bool = False) -> None:
    parts = dotted.split(".")
    mod = root
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, nn.Module())
        mod = mod._modules[p]
    if buffer:
        mod.register_buffer(parts[-1], tensor)
    else:
        mod.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


class _SpecModule(nn.Module):
    """Base: parameters from ``self._spec()``; packed device weights cached until the parameters change."""

    def _build(self, spec: Dict[str, Tuple[Sequence[int], bool]], seed: int = 0) -> None:
        for name, (shape, is_buffer) in spec.items():
            _add_tensor(self, name, init_tensor(name, shape, seed), buffer=is_buffer)
        self._packed = None
        self.register_load_state_dict_post_hook(lambda m, k: m._invalidate())

    def _invalidate(self) -> None:
        self._packed = None

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._packed = None
        return out

    def train(self, mode: bool = True):
        if mode:
            pass  # inference-only modules: mode flags are accepted and ignored
        return super().train(False)

    def _sd(self) -> Dict[str, torch.Tensor]:
        return {k: v.detach() for k, v in self.state_dict().items()}


def _bf16(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.bfloat16).contiguous()


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


class _Workspace:
    """Grow-only device scratch buffer owned by a module (torch owns the memory, the C side only borrows it)."""

    def __init__(self):
        self.buf: Optional[torch.Tensor] = None

    def get(self, nbytes: int, device) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != torch.device(device):
            self.buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        return self.buf


# --------------------------------------------------------------------------------------------------
# A1  CLIP vision tower
# --------------------------------------------------------------------------------------------------
def clip_param_spec(hidden=1024, mlp=4096, layers=24, image=448, patch=14) -> Dict[str, Tuple[Sequence[int], bool]]:
    p = "vision_tower.vision_model."
    g = image // patch
    spec: Dict[str, Tuple[Sequence[int], bool]] = {
        p + "embeddings.class_embedding": ((hidden,), False),
        p + "embeddings.patch_embedding.weight": ((hidden, 3, patch, patch), False),
        p + "embeddings.position_embedding.weight": ((g * g + 1, hidden), False),
        p + "pre_layrnorm.weight": ((hidden,), False),
        p + "pre_layrnorm.bias": ((hidden,), False),
    }
    for i in range(layers):
        lp = p + f"encoder.layers.{i}."
        for proj in ("k_proj", "v_proj", "q_proj", "out_proj"):
            spec[lp + f"self_attn.{proj}.weight"] = ((hidden, hidden), False)
            spec[lp + f"self_attn.{proj}.bias"] = ((hidden,), False)
        spec[lp + "layer_norm1.weight"] = ((hidden,), False)
        spec[lp + "layer_norm1.bias"] = ((hidden,), False)
        spec[lp + "mlp.fc1.weight"] = ((mlp, hidden), False)
        spec[lp + "mlp.fc1.bias"] = ((mlp,), False)
        spec[lp + "mlp.fc2.weight"] = ((hidden, mlp), False)
        spec[lp + "mlp.fc2.bias"] = ((hidden,), False)
        spec[lp + "layer_norm2.weight"] = ((hidden,), False)
        spec[lp + "layer_norm2.bias"] = ((hidden,), False)
    spec[p + "post_layernorm.weight"] = ((hidden,), False)  # present in HF checkpoints; unused by feature_select
    spec[p + "post_layernorm.bias"] = ((hidden,), False)
    return spec


class _ClipArgs:
    mm_vision_select_layer = -2
    mm_vision_select_feature = "patch"


class CLIPVisionTower(_SpecModule):
    """Drop-in for ``CLIPVisionTower`` (model/llava_walkgpt/model/multimodal_encoder/clip_encoder.py:6-125).

    ``vision_tower`` may be a HF model directory (weights are read with ``transformers`` and the position table is
    resized to ``resize_vision_tower_size`` exactly like clip_encoder.py:38-55) or ``None`` for a random-init
    ViT-L/14@448 of the reference's dimensions."""

    def __init__(self, vision_tower=None, args=None, delay_load: bool = False, *, hidden=1024, mlp=4096, layers=24,
                 heads=16, image=448, patch=14, seed=0):
        super().__init__()
        args = args or _ClipArgs()
        self.vision_tower_name = vision_tower
        self.select_layer = getattr(args, "mm_vision_select_layer", -2)
        self.select_feature = getattr(args, "mm_vision_select_feature", "patch")
        if self.select_feature != "patch":
            raise ValueError(f"Unexpected select feature: {self.select_feature}")  # cls_patch is not on the hot path
        if getattr(args, "resize_vision_tower", False):
            image = getattr(args, "resize_vision_tower_size", image)
        self.geom = dict(hidden=hidden, mlp=mlp, layers=layers, heads=heads, image=image, patch=patch)
        self._build(clip_param_spec(hidden, mlp, layers, image, patch), seed)
        self.is_loaded = True
        self._ws = _Workspace()
        if isinstance(vision_tower, str) and not delay_load:
            self.load_model()

    def load_model(self):
        """Read HF CLIP weights (host-side plumbing) and resize the position table (clip_encoder.py:38-55)."""
        from transformers import CLIPVisionModel  # only used to read the checkpoint
        import torch.nn.functional as F

        hf = CLIPVisionModel.from_pretrained(self.vision_tower_name)
        sd = {"vision_tower." + k: v for k, v in hf.state_dict().items() if not k.endswith("position_ids")}
        key = "vision_tower.vision_model.embeddings.position_embedding.weight"
        pos = sd[key]
        g_new = self.geom["image"] // self.geom["patch"]
        g_old = int(math.isqrt(pos.shape[0] - 1))
        if g_old != g_new:
            # NOTE: the reference treats the LAST row as the CLS slot (clip_encoder.py:47-52); mirrored here.
            grid = pos[:-1].permute(1, 0).reshape(1, -1, g_old, g_old)
            grid = F.interpolate(grid, (g_new, g_new), mode="bilinear", align_corners=False)[0]
            sd[key] = torch.cat([grid.flatten(-2).permute(1, 0), pos[-1:]], dim=0)
        self.load_state_dict(sd, strict=False)
        self.is_loaded = True

    # -- properties of the reference class (clip_encoder.py:100-125)
    @property
    def dtype(self):
        return self.vision_tower.vision_model.pre_layrnorm.weight.dtype

    @property
    def device(self):
        return self.vision_tower.vision_model.pre_layrnorm.weight.device

    @property
    def hidden_size(self):
        return self.geom["hidden"]

    @property
    def num_patches(self):
        return (self.geom["image"] // self.geom["patch"]) ** 2

    @property
    def dummy_feature(self):
        return torch.zeros(1, self.hidden_size, device=self.device, dtype=self.dtype)

    # -- packing
    def _pack(self):
        g = self.geom
        sd = self._sd()
        p = "vision_tower.vision_model."
        keep: List[torch.Tensor] = []

        def hold(t):
            keep.append(t)
            return t.data_ptr()

        kvalid = 3 * g["patch"] ** 2
        kpad = (kvalid + 63) // 64 * 64
        pw = torch.zeros(g["hidden"], kpad, dtype=torch.bfloat16, device=self.device)
        pw[:, :kvalid] = sd[p + "embeddings.patch_embedding.weight"].reshape(g["hidden"], -1).to(torch.bfloat16)
        layers = (_lib.ClipLayer * g["layers"])()
        for i in range(g["layers"]):
            lp = p + f"encoder.layers.{i}."
            ly = layers[i]
            ly.ln1_g, ly.ln1_b = hold(_f32(sd[lp + "layer_norm1.weight"])), hold(_f32(sd[lp + "layer_norm1.bias"]))
            ly.w_qkv = hold(_bf16(torch.cat([sd[lp + f"self_attn.{n}.weight"] for n in ("q_proj", "k_proj", "v_proj")], 0)))
            ly.b_qkv = hold(_f32(torch.cat([sd[lp + f"self_attn.{n}.bias"] for n in ("q_proj", "k_proj", "v_proj")], 0)))
            ly.w_o, ly.b_o = hold(_bf16(sd[lp + "self_attn.out_proj.weight"])), hold(_f32(sd[lp + "self_attn.out_proj.bias"]))
            ly.ln2_g, ly.ln2_b = hold(_f32(sd[lp + "layer_norm2.weight"])), hold(_f32(sd[lp + "layer_norm2.bias"]))
            ly.w_fc1, ly.b_fc1 = hold(_bf16(sd[lp + "mlp.fc1.weight"])), hold(_f32(sd[lp + "mlp.fc1.bias"]))
            ly.w_fc2, ly.b_fc2 = hold(_bf16(sd[lp + "mlp.fc2.weight"])), hold(_f32(sd[lp + "mlp.fc2.bias"]))
        w = _lib.ClipWeights()
        w.hidden, w.heads, w.mlp, w.image, w.patch = g["hidden"], g["heads"], g["mlp"], g["image"], g["patch"]
        w.kpad, w.n_layers = kpad, g["layers"]
        w.patch_w = hold(pw)
        w.cls_emb = hold(_f32(sd[p + "embeddings.class_embedding"]))
        w.pos_emb = hold(_f32(sd[p + "embeddings.position_embedding.weight"]))
        w.pre_ln_g, w.pre_ln_b = hold(_f32(sd[p + "pre_layrnorm.weight"])), hold(_f32(sd[p + "pre_layrnorm.bias"]))
        w.layers = C.cast(layers, C.POINTER(_lib.ClipLayer))
        self._packed = (w, layers, keep)
        return self._packed

    def hidden_state_indices(self) -> Tuple[int, int]:
        n_states = self.geom["layers"] + 1
        return self.select_layer % n_states, (-11) % n_states

    @torch.no_grad()
    def forward(self, images, attention_mask=None):
        """images [B,3,S,S] (or a list of [3,S,S]); attention_mask [B,1+g*g] with 1 = valid key (or None).
        Returns (hidden_states[select_layer][:,1:], [hidden_states[-11][:,1:]]) cast to images.dtype."""
        if isinstance(images, (list, tuple)):
            images = torch.stack(list(images), dim=0)
        if not images.is_cuda:
            raise _lib.WalkGPTB200Error("CLIPVisionTower.forward needs CUDA tensors (no CPU fallback)")
        out_dtype = images.dtype
        if images.dtype not in (torch.float32, torch.bfloat16):
            images = images.float()
        images = images.contiguous()
        B = images.shape[0]
        g = self.geom
        assert images.shape[1:] == (3, g["image"], g["image"]), f"expected [B,3,{g['image']},{g['image']}], got {tuple(images.shape)}"
        w, _, _ = self._packed or self._pack()
        idx_last, idx_mid = self.hidden_state_indices()
        n_run = max(idx_last, idx_mid)
        L = self.num_patches
        kernel_dtype = torch.bfloat16 if out_dtype == torch.bfloat16 else torch.float32
        out_hi = torch.empty(B, L, g["hidden"], device=images.device, dtype=kernel_dtype)
        out_lo = torch.empty_like(out_hi) if idx_last != idx_mid else None
        kv = None
        if attention_mask is not None:
            kv = (attention_mask > 0.5).to(torch.uint8).contiguous()
            assert kv.shape == (B, L + 1)
        nbytes = _lib.lib().wg_clip_workspace_bytes(C.byref(w), B)
        ws = self._ws.get(nbytes, images.device)
        with torch.cuda.device(images.device):
            _lib.check(_lib.lib().wg_clip_forward(
                C.byref(w), images.data_ptr(), int(images.dtype == torch.bfloat16), None if kv is None else kv.data_ptr(), B,
                n_run, min(idx_last, idx_mid), out_hi.data_ptr(), None if out_lo is None else out_lo.data_ptr(),
                int(kernel_dtype == torch.bfloat16), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream),
                "wg_clip_forward")
        if out_lo is None:
            last = mid = out_hi
        elif idx_last >= idx_mid:
            last, mid = out_hi, out_lo
        else:
            last, mid = out_lo, out_hi
        return last.to(out_dtype), [mid.to(out_dtype)]
