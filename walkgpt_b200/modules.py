"""Host-side mirror of the reference's nn.Module interface for the pixel-grounding path.

Each class keeps the reference's constructor arguments, parameter names/shapes (so the reference's
``load_state_dict(strict=True)`` checkpoints load unchanged) and ``forward`` signature, but its forward is a
single call into the C ABI (ctypes, :mod:`walkgpt_b200._lib`; :mod:`walkgpt_b200.torch_ops` exposes the same calls as ``torch.library`` custom ops).
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
from . import specs
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
    """Grow-only device scratch owned by a module (torch owns the memory, the C side only borrows it), one buffer per
    (device, CUDA stream): kernels on one stream are ordered, so consecutive calls may share scratch; calls issued on different
    streams (or devices) never share it.  A buffer that has to grow is replaced through torch's stream-aware caching allocator,
    which does not hand the old block to another stream while this stream's kernels may still be using it."""

    def __init__(self):
        self.bufs: Dict[Tuple[int, int], torch.Tensor] = {}

    def get(self, nbytes: int, device) -> torch.Tensor:
        device = torch.device(device)
        key = (device.index if device.index is not None else torch.cuda.current_device(), torch.cuda.current_stream(device).cuda_stream)
        buf = self.bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = self.bufs[key] = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        return buf


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
        if self.select_feature not in ("patch", "cls_patch"):
            raise ValueError(f"Unexpected select feature: {self.select_feature}")  # clip_encoder.py:61-69
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
        w.layers = C.cast(layers, C.c_void_p)
        self._packed = (w, layers, keep)
        return self._packed

    def hidden_state_indices(self) -> Tuple[int, int]:
        n_states = self.geom["layers"] + 1
        return self.select_layer % n_states, (-11) % n_states

    @torch.no_grad()
    def forward(self, images, attention_mask=None, want_mid: bool = True):
        """images [B,3,S,S] (or a list of [3,S,S]); attention_mask [B,1+g*g] with 1 = valid key (or None).
        Returns (hidden_states[select_layer][:,1:], [hidden_states[-11][:,1:]]) cast to images.dtype.  want_mid=False (not part
        of the reference signature) skips the copy of hidden_states[-11], which only image_feature_scale_num=2 consumes; the
        second element is then the first one again."""
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
        if not want_mid:
            idx_mid = idx_last
        n_run = max(idx_last, idx_mid)
        L = self.num_patches
        keep_cls = int(self.select_feature == "cls_patch")
        kernel_dtype = torch.bfloat16 if out_dtype == torch.bfloat16 else torch.float32
        out_hi = torch.empty(B, L + keep_cls, g["hidden"], device=images.device, dtype=kernel_dtype)
        out_lo = torch.empty_like(out_hi) if idx_last != idx_mid else None
        kv = None
        if attention_mask is not None:
            kv = (attention_mask > 0.5).to(torch.uint8).contiguous()
            assert kv.shape == (B, L + 1)
        nbytes = _lib.lib().wg_clip_workspace_bytes(C.byref(w), B)
        ws = self._ws.get(nbytes, images.device)
        with torch.cuda.device(images.device):
            _lib.check(_lib.lib().wg_clip_forward_ex(
                C.byref(w), images.data_ptr(), int(images.dtype == torch.bfloat16), None if kv is None else kv.data_ptr(), B,
                n_run, min(idx_last, idx_mid), out_hi.data_ptr(), None if out_lo is None else out_lo.data_ptr(),
                int(kernel_dtype == torch.bfloat16), keep_cls, ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream),
                "wg_clip_forward")
        if out_lo is None:
            last = mid = out_hi
        elif idx_last >= idx_mid:
            last, mid = out_hi, out_lo
        else:
            last, mid = out_lo, out_hi
        return last.to(out_dtype), [mid.to(out_dtype)]



# --------------------------------------------------------------------------------------------------
# weight re-layout helpers (pure re-indexing / constant folding done once at load time; unit-tested on CPU)
# --------------------------------------------------------------------------------------------------
def fold_kv_norm(in_proj_weight: torch.Tensor, in_proj_bias: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, d: int):
    """K|V rows of a packed nn.MultiheadAttention in_proj with the preceding LayerNorm's affine folded in:
    W (g*xhat + b) + c  ==  (W diag(g)) xhat + (c + W b).  Returns (W' [2d, d], b' [2d])."""
    W, b = in_proj_weight.float()[d:], in_proj_bias.float()[d:]
    return W * gamma.float()[None, :], b + W @ beta.float()


def pack_conv3x3(weight: torch.Tensor) -> torch.Tensor:
    """Conv2d weight [co, ci, 3, 3] -> GEMM weight [co, (ky, kx, ci)] matching the channels-last im2col column order."""
    return weight.permute(0, 2, 3, 1).reshape(weight.shape[0], -1)


def pack_conv_transpose2x2(weight: torch.Tensor, bias: torch.Tensor):
    """ConvTranspose2d(k=2, s=2) weight [ci, co, 2, 2] -> GEMM weight [(dy, dx, co), ci] and bias tiled over the 4 sub-pixels."""
    return weight.permute(2, 3, 1, 0).reshape(-1, weight.shape[0]), bias.float().repeat(4)


def split_terms_needed(weights: Iterable[torch.Tensor]) -> int:
    """2 if every weight is exactly representable in bf16 (the usual case: WalkGPT checkpoints are bf16), else 3."""
    for w in weights:
        wf = w.detach().float()
        if (wf - wf.to(torch.bfloat16).float()).abs().max().item() > 0.0:
            return 3
    return 2


def split_weight(w: torch.Tensor, terms: int) -> torch.Tensor:
    """nn.Linear weight [N, K] -> bf16 [N, terms*K] = [W_hi | W_hi | W_lo] (terms=3) or [W | W] (terms=2): the B operand that
    pairs with a split-bf16 activation [hi | lo | hi] (see wg_gemm_args.a_k_wrap)."""
    wf = w.detach().float()
    hi = wf.to(torch.bfloat16)
    parts = [hi, hi]
    if terms == 3:
        parts.append((wf - hi.float()).to(torch.bfloat16))
    return torch.cat(parts, dim=1).contiguous()


def to_split(x: torch.Tensor) -> torch.Tensor:
    """fp32 [..., C] -> split-bf16 [..., 2C] (hi | lo): a storage-format conversion, like .to(bfloat16) but ~16 mantissa bits."""
    xf = x.float()
    hi = xf.to(torch.bfloat16)
    return torch.cat([hi, (xf - hi.float()).to(torch.bfloat16)], dim=-1).contiguous()


def merge_split(x: torch.Tensor) -> torch.Tensor:
    """split-bf16 [..., 2C] -> fp32 [..., C]."""
    c = x.shape[-1] // 2
    return x[..., :c].float() + x[..., c:].float()


class _Holder:
    """Keeps repacked device tensors alive for as long as a ctypes weight struct points at them."""

    def __init__(self):
        self.keep: List[torch.Tensor] = []

    def __call__(self, t: torch.Tensor) -> int:
        self.keep.append(t)
        return t.data_ptr()

    def bf16(self, t):
        return self(_bf16(t))

    def f32(self, t):
        return self(_f32(t))


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(t: torch.Tensor, who: str) -> None:
    if not t.is_cuda:
        raise _lib.WalkGPTB200Error(f"{who} needs CUDA tensors (walkgpt_b200 has no CPU fallback)")


def _as_kernel_input(x: torch.Tensor) -> torch.Tensor:
    """fp32/bf16 contiguous (other float types are widened to fp32)."""
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    return x.contiguous()


# --------------------------------------------------------------------------------------------------
# A2  MSQP
# --------------------------------------------------------------------------------------------------
class MultiScaleQFormerProjector(_SpecModule):
    """Drop-in for ``MultiScaleQFormerProjector`` (utils/utils_walkgpt.py:220-300)."""

    _GROUPS = (("q_x1", "cross_x1"), ("q_x2", "cross_x2"), ("q_x4", "cross_x4"), ("q_global", "cross_glb"))

    def __init__(self, sam_dim, llama_dim, grid_size=None, num_heads=8, pad_to_square: bool = True,
                 target_square_side: Optional[int] = None, *, seed=0):
        super().__init__()
        if num_heads != 8:
            raise ValueError("MultiScaleQFormerProjector: the CUDA path is built for num_heads=8")
        self.grid_size = grid_size
        self.d_proj = 1024
        self.num_layers = 2
        self.pad_to_square = pad_to_square
        self.target_square_side = target_square_side
        self.sam_dim, self.llama_dim = sam_dim, llama_dim
        self._build(specs.msqp_spec(sam_dim, llama_dim), seed)
        self._ws = _Workspace()

    def n_tokens(self) -> int:
        q = 32
        if not self.pad_to_square:
            return q
        s = int(math.ceil(math.sqrt(q))) if self.target_square_side is None else self.target_square_side
        assert s * s >= q, "target_square_side too small"  # utils_walkgpt.py:294
        return s * s

    def _pack(self):
        sd, hold, d = self._sd(), _Holder(), self.d_proj
        w = _lib.MsqpWeights()
        w.sam_dim, w.llama_dim, w.d, w.heads, w.n_tokens = self.sam_dim, self.llama_dim, d, 8, self.n_tokens()
        w.w_in, w.b_in = hold.bf16(sd["sam_to_proj.weight"]), hold.f32(sd["sam_to_proj.bias"])
        w.gate_ln_g, w.gate_ln_b = hold.f32(sd["gate.net.0.weight"]), hold.f32(sd["gate.net.0.bias"])
        w.w_g1, w.b_g1 = hold.bf16(sd["gate.net.1.weight"]), hold.f32(sd["gate.net.1.bias"])
        w.w_g2, w.b_g2 = hold.f32(sd["gate.net.3.weight"].reshape(-1)), hold.f32(sd["gate.net.3.bias"])
        w.pad_token = hold.f32(sd["pad_token"].reshape(-1))
        w.w_out, w.b_out = hold.bf16(sd["to_llama.weight"]), hold.f32(sd["to_llama.bias"])
        for si, (qn, cn) in enumerate(self._GROUPS):
            S = w.scales[si]
            S.queries, S.nq = hold.f32(sd[qn][0]), sd[qn].shape[1]
            wk, bk = [], []
            for l in range(2):
                pre = f"{cn}.{l}."
                W, b = sd[pre + "attn.in_proj_weight"].float(), sd[pre + "attn.in_proj_bias"].float()
                wf, bf = fold_kv_norm(W, b, sd[pre + "kv_norm.weight"], sd[pre + "kv_norm.bias"], d)
                wk.append(wf)
                bk.append(bf)
                Bk = S.blocks[l]
                Bk.qn_g, Bk.qn_b = hold.f32(sd[pre + "q_norm.weight"]), hold.f32(sd[pre + "q_norm.bias"])
                Bk.w_q, Bk.b_q = hold.bf16(W[:d]), hold.f32(b[:d])
                Bk.w_o, Bk.b_o = hold.bf16(sd[pre + "attn.out_proj.weight"]), hold.f32(sd[pre + "attn.out_proj.bias"])
                Bk.ffn_ln_g, Bk.ffn_ln_b = hold.f32(sd[pre + "ffn.0.weight"]), hold.f32(sd[pre + "ffn.0.bias"])
                Bk.w_f1, Bk.b_f1 = hold.bf16(sd[pre + "ffn.1.weight"]), hold.f32(sd[pre + "ffn.1.bias"])
                Bk.w_f2, Bk.b_f2 = hold.bf16(sd[pre + "ffn.3.weight"]), hold.f32(sd[pre + "ffn.3.bias"])
            S.w_kv, S.b_kv = hold.bf16(torch.cat(wk, 0)), hold.f32(torch.cat(bk, 0))
        self._packed = (w, hold)
        return self._packed

    def run(self, feats_bf16: torch.Tensor, out_dtype=torch.bfloat16) -> torch.Tensor:
        """feats bf16 [B, L, sam_dim] contiguous -> [B, n_tokens, llama_dim]."""
        B, L, _ = feats_bf16.shape
        w, _ = self._packed or self._pack()
        out = torch.empty(B, self.n_tokens(), self.llama_dim, device=feats_bf16.device, dtype=out_dtype)
        ws = self._ws.get(_lib.lib().wg_msqp_workspace_bytes(C.byref(w), B, L), feats_bf16.device)
        _lib.check(_lib.lib().wg_msqp_forward(C.byref(w), feats_bf16.data_ptr(), B, L, out.data_ptr(), int(out_dtype == torch.bfloat16),
                                              ws.data_ptr(), ws.numel(), _stream()), "wg_msqp_forward")
        return out

    @torch.no_grad()
    def forward(self, sam_feats, grid_size=None):
        _need_cuda(sam_feats, "MultiScaleQFormerProjector.forward")
        B, L, _ = sam_feats.shape
        hw = grid_size or self.grid_size
        if hw is not None and (hw[0] != hw[1] or hw[0] * hw[1] != L):
            raise ValueError("MultiScaleQFormerProjector: only square token grids are built")
        if int(math.isqrt(L)) ** 2 != L:
            raise ValueError(f"Token length {L} is not a perfect square.")  # utils_walkgpt.py:188-192
        out_dtype = sam_feats.dtype if sam_feats.dtype in (torch.float32, torch.bfloat16) else torch.float32
        with torch.cuda.device(sam_feats.device):
            out = self.run(sam_feats.to(torch.bfloat16).contiguous(), out_dtype)
        return out.to(sam_feats.dtype)


# --------------------------------------------------------------------------------------------------
# A5  CTP
# --------------------------------------------------------------------------------------------------
class CalibratedTextProjector(_SpecModule):
    """Drop-in for ``CalibratedTextProjector`` (utils/utils_walkgpt.py:302-327)."""

    def __init__(self, in_dim: int, out_dim: int, widen: int = 2, use_residual: bool = False, *, seed=0):
        super().__init__()
        if out_dim != 256 or widen != 2:
            raise ValueError("CalibratedTextProjector: the CUDA path is built for out_dim=256, widen=2 (model/walkgpt.py:141)")
        self.use_residual = use_residual and (in_dim == out_dim)
        if self.use_residual:
            raise ValueError("CalibratedTextProjector: use_residual=True is not on the reference's path")
        self.in_dim, self.out_dim = in_dim, out_dim
        self._build(specs.ctp_spec(in_dim, out_dim, widen), seed)
        self._ws = _Workspace()

    def _pack(self):
        sd, hold = self._sd(), _Holder()
        w = _lib.CtpWeights()
        w.in_dim, w.mid_dim, w.out_dim = self.in_dim, 2 * self.out_dim, self.out_dim
        w.ln0_g, w.ln0_b = hold.f32(sd["net.0.weight"]), hold.f32(sd["net.0.bias"])
        w.w1, w.b1 = hold.bf16(sd["net.1.weight"]), hold.f32(sd["net.1.bias"])
        w.w2, w.b2 = hold.bf16(sd["net.3.weight"]), hold.f32(sd["net.3.bias"])
        w.ln4_g, w.ln4_b = hold.f32(sd["net.4.weight"]), hold.f32(sd["net.4.bias"])
        w.text_type, w.log_temp = hold.f32(sd["text_type"].reshape(-1)), hold.f32(sd["log_temp"])
        self._packed = (w, hold)
        return self._packed

    def run(self, x2d: torch.Tensor, out_dtype=torch.float32) -> torch.Tensor:
        """x [rows, in_dim] fp32/bf16 contiguous -> [rows, 256]."""
        rows = x2d.shape[0]
        w, _ = self._packed or self._pack()
        out = torch.empty(rows, self.out_dim, device=x2d.device, dtype=out_dtype)
        if rows == 0:
            return out
        ws = self._ws.get(_lib.lib().wg_ctp_workspace_bytes(rows, self.in_dim), x2d.device)
        _lib.check(_lib.lib().wg_ctp_forward(C.byref(w), x2d.data_ptr(), int(x2d.dtype == torch.bfloat16), rows, out.data_ptr(),
                                             int(out_dtype == torch.bfloat16), ws.data_ptr(), ws.numel(), _stream()), "wg_ctp_forward")
        return out

    @torch.no_grad()
    def forward(self, x):
        _need_cuda(x, "CalibratedTextProjector.forward")
        xin = _as_kernel_input(x)
        lead = xin.shape[:-1]
        with torch.cuda.device(x.device):
            y = self.run(xin.reshape(-1, self.in_dim), xin.dtype)
        # text_type is [1,1,out]: a 2-D input comes back 3-D, exactly like the reference
        shape = torch.broadcast_shapes(tuple(lead) + (self.out_dim,), (1, 1, self.out_dim))
        return y.reshape(shape).to(x.dtype)


# --------------------------------------------------------------------------------------------------
# A3 + A4  out_mm_projector MLP and image_feature_neck
# --------------------------------------------------------------------------------------------------
class ProjectorNeck(_SpecModule):
    """``out_mm_projector`` (llava_arch.py:38-42) + ``image_feature_neck`` (model/walkgpt.py:97-113) as one module.

    Parameters live under ``out_mm_projector.*`` and ``image_feature_neck.*`` with the reference's names."""

    def __init__(self, mm_hidden: int, hidden: int, out_chans: int = 256, *, seed=0):
        super().__init__()
        self.mm_hidden, self.hidden, self.out_chans = mm_hidden, hidden, out_chans
        spec: Spec = {}
        spec.update({"out_mm_projector." + k: v for k, v in specs.out_mm_projector_spec(mm_hidden, hidden).items()})
        spec.update({"image_feature_neck." + k: v for k, v in specs.neck_spec(hidden, out_chans).items()})
        self._build(spec, seed)
        self._ws = _Workspace()

    def _pack(self):
        sd, hold = self._sd(), _Holder()
        w = _lib.ProjNeckWeights()
        w.mm_hidden, w.hidden, w.out_chans = self.mm_hidden, self.hidden, self.out_chans
        w.w_fc1, w.b_fc1 = hold.bf16(sd["out_mm_projector.0.weight"]), hold.f32(sd["out_mm_projector.0.bias"])
        w.w_fc2, w.b_fc2 = hold.bf16(sd["out_mm_projector.2.weight"]), hold.f32(sd["out_mm_projector.2.bias"])
        w1 = sd["image_feature_neck.0.weight"].reshape(self.out_chans, self.hidden)
        w3 = pack_conv3x3(sd["image_feature_neck.2.weight"])  # [co, ci, ky, kx] -> [co, (ky, kx, ci)]: channels-last im2col column order
        terms = split_terms_needed([w1, w3])
        w.split_terms = terms
        w.w_conv1 = hold(split_weight(w1, terms))
        w.ln1_g, w.ln1_b = hold.f32(sd["image_feature_neck.1.weight"]), hold.f32(sd["image_feature_neck.1.bias"])
        w.w_conv3 = hold(split_weight(w3, terms))
        w.ln2_g, w.ln2_b = hold.f32(sd["image_feature_neck.3.weight"]), hold.f32(sd["image_feature_neck.3.bias"])
        self._packed = (w, hold)
        return self._packed

    def run(self, feats_bf16: torch.Tensor, want_proj: bool = False, want_emb: bool = True):
        """feats bf16 [B, L, mm_hidden] -> (proj split-bf16 [B,L,2H] | None, emb tokens split-bf16 [B,L,512] | None)."""
        B, L, _ = feats_bf16.shape
        g = int(math.isqrt(L))
        assert g * g == L
        w, _ = self._packed or self._pack()
        proj = torch.empty(B, L, 2 * self.hidden, device=feats_bf16.device, dtype=torch.bfloat16) if want_proj else None
        emb = torch.empty(B, L, 2 * self.out_chans, device=feats_bf16.device, dtype=torch.bfloat16) if want_emb else None
        ws = self._ws.get(_lib.lib().wg_proj_neck_workspace_bytes(C.byref(w), B * L), feats_bf16.device)
        _lib.check(_lib.lib().wg_proj_neck_forward(C.byref(w), feats_bf16.data_ptr(), B, g, None if proj is None else proj.data_ptr(),
                                                   None if emb is None else emb.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
                   "wg_proj_neck_forward")
        return proj, emb

    @torch.no_grad()
    def project(self, feats):
        """out_mm_projector(feats): [B, L, mm_hidden] -> [B, L, hidden] (input dtype)."""
        _need_cuda(feats, "ProjectorNeck.project")
        with torch.cuda.device(feats.device):
            proj, _ = self.run(feats.to(torch.bfloat16).contiguous(), want_proj=True, want_emb=False)
        return merge_split(proj).to(feats.dtype)

    @torch.no_grad()
    def neck(self, x_nchw):
        """image_feature_neck(x): [B, hidden, g, g] -> [B, 256, g, g] (input dtype)."""
        _need_cuda(x_nchw, "ProjectorNeck.neck")
        B, Hd, gh, gw = x_nchw.shape
        assert gh == gw and Hd == self.hidden
        tokens = to_split(x_nchw.permute(0, 2, 3, 1))  # layout / storage-format change only
        w, _ = self._packed or self._pack()
        emb = torch.empty(B, gh * gw, 2 * self.out_chans, device=x_nchw.device, dtype=torch.bfloat16)
        with torch.cuda.device(x_nchw.device):
            ws = self._ws.get(_lib.lib().wg_neck_workspace_bytes(B * gh * gw), x_nchw.device)
            _lib.check(_lib.lib().wg_neck_forward(C.byref(w), tokens.data_ptr(), B, gh, emb.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
                       "wg_neck_forward")
            return tokens_to_nchw(emb, gh, gw, x_nchw.dtype, split=True)

    @torch.no_grad()
    def forward(self, feats):
        """feats [B, L, mm_hidden] -> image embedding [B, 256, g, g] (input dtype)."""
        _need_cuda(feats, "ProjectorNeck.forward")
        g = int(math.isqrt(feats.shape[1]))
        with torch.cuda.device(feats.device):
            _, emb = self.run(feats.to(torch.bfloat16).contiguous())
            return tokens_to_nchw(emb, g, g, feats.dtype, split=True)


def tokens_to_nchw(tokens_bf16: torch.Tensor, h: int, w: int, dtype=torch.float32, split: bool = False) -> torch.Tensor:
    """[B, h*w, C] bf16 (or split-bf16 [B, h*w, 2C]) -> [B, C, h, w]."""
    B, L, Cc = tokens_bf16.shape
    if split:
        Cc //= 2
    kd = torch.bfloat16 if dtype == torch.bfloat16 else torch.float32
    out = torch.empty(B, Cc, h, w, device=tokens_bf16.device, dtype=kd)
    _lib.check(_lib.lib().wg_tokens_to_nchw(tokens_bf16.data_ptr(), int(split), out.data_ptr(), int(kd == torch.bfloat16), B, L, Cc, _stream()),
               "wg_tokens_to_nchw")
    return out.to(dtype)


# --------------------------------------------------------------------------------------------------
# A6  Prompt encoder
# --------------------------------------------------------------------------------------------------
class PromptEncoder(_SpecModule):
    """Drop-in for ``PromptEncoder`` (segment_anything/modeling/prompt_encoder.py:16-186), text-prompt path only:
    the reference only ever calls it with ``points=boxes=masks=None`` (model/walkgpt.py:515-520, 724-729)."""

    def __init__(self, embed_dim: int, image_embedding_size: Tuple[int, int], input_image_size: Tuple[int, int], mask_in_chans: int,
                 activation=None, *, seed=0):
        super().__init__()
        self.embed_dim = embed_dim
        self.input_image_size = input_image_size
        self.image_embedding_size = image_embedding_size
        self._build(specs.prompt_encoder_spec(embed_dim, mask_in_chans), seed)
        self._pe_cache = None

    def _invalidate(self):
        super()._invalidate()
        self._pe_cache = None

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._pe_cache = None
        return out

    def dense_pe_tokens(self, size: Optional[Tuple[int, int]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """(pe [1, C, h, w] fp32, pe tokens [h*w, C] fp32), computed by the wg_dense_pe kernel and cached."""
        h, w = size or self.image_embedding_size
        gauss = self.pe_layer.positional_encoding_gaussian_matrix
        key = (gauss.data_ptr(), gauss._version, h, w)
        if self._pe_cache is None or self._pe_cache[0] != key:
            _need_cuda(gauss, "PromptEncoder.get_dense_pe")
            F_ = gauss.shape[1]
            g32 = _f32(gauss)
            chw = torch.empty(1, 2 * F_, h, w, device=gauss.device, dtype=torch.float32)
            tok = torch.empty(h * w, 2 * F_, device=gauss.device, dtype=torch.float32)
            with torch.cuda.device(gauss.device):
                _lib.check(_lib.lib().wg_dense_pe(g32.data_ptr(), F_, h, w, chw.data_ptr(), tok.data_ptr(), _stream()), "wg_dense_pe")
            self._pe_cache = (key, chw, tok)
        return self._pe_cache[1], self._pe_cache[2]

    def get_dense_pe(self) -> torch.Tensor:
        return self.dense_pe_tokens()[0]

    @torch.no_grad()
    def forward(self, points, boxes, masks, text_embeds):
        if points is not None or boxes is not None or masks is not None:
            raise NotImplementedError("PromptEncoder: point / box / mask prompts are not on WalkGPT's path (text_embeds only)")
        if text_embeds is None:
            raise ValueError("PromptEncoder.forward: text_embeds is required")
        bs = text_embeds.shape[0]
        sparse = text_embeds  # torch.cat([empty, text_embeds], dim=1) in the reference (:165-177)
        dense = self.no_mask_embed.weight.reshape(1, -1, 1, 1).expand(bs, -1, self.image_embedding_size[0], self.image_embedding_size[1])
        return sparse, dense


# --------------------------------------------------------------------------------------------------
# A7  Mask decoder (multi-scale variant, level 0)
# --------------------------------------------------------------------------------------------------
class MaskDecoderMultiScale(_SpecModule):
    """Drop-in for ``MaskDecoderMultiScale`` (segment_anything/modeling/mask_decoder_multi_scale.py:16-213).

    ``transformer`` is accepted for signature compatibility; the two-way transformer's dimensions are fixed to the
    reference's (depth 2, dim 256, 8 heads, mlp 2048; model/walkgpt.py:81-93).  Only ``level_num == 0`` is built."""

    def __init__(self, *, transformer_dim: int = 256, transformer=None, num_multimask_outputs: int = 3, activation=None,
                 iou_head_depth: int = 3, iou_head_hidden_dim: int = 256, image_feature_scale_num: int = 1, seed=0):
        super().__init__()
        if transformer_dim != 256 or num_multimask_outputs != 3 or iou_head_depth != 3 or iou_head_hidden_dim != 256:
            raise ValueError("MaskDecoderMultiScale: the CUDA path is built for the reference's dimensions (256 / 3 / 3 / 256)")
        self.transformer_dim = transformer_dim
        self.num_multimask_outputs = num_multimask_outputs
        self.num_mask_tokens = num_multimask_outputs + 1
        self.image_feature_scale_num = image_feature_scale_num
        self._build(specs.mask_decoder_multiscale_spec(transformer_dim, num_multimask_outputs, image_feature_scale_num), seed)
        self._ws = _Workspace()
        self._ws_up = _Workspace()
        self._pe_key = None
        self._levels = {}

    # hooks the SAM MaskDecoder subclass overrides
    _T = "transformer.0."   # state_dict prefix of the two-way transformer
    _UP_STAGES = 1          # ConvTranspose stages of output_upscaling
    _MULTIMASK_FIRST = 0    # mask_decoder_multi_scale.py:126-132 keeps mask 0 in both modes
    _MLP_GEMM = True        # token MLP as tensor-core GEMMs over all prompts' rows (False: inside the per-prompt token kernel, fp32)

    def _prefix(self, level: int) -> str:
        return self._T.replace("0", str(level)) if "0" in self._T else self._T

    def _pack_static(self, level: int = 0):
        sd, hold = self._sd(), _Holder()
        w = _lib.MaskDecoderWeights()
        w.n_mask_tokens, w.up_stages, w.multimask_first = self.num_mask_tokens, self._UP_STAGES, self._MULTIMASK_FIRST
        out_tok = torch.cat([sd["iou_token.weight"], sd["mask_tokens.weight"]], 0).float()
        if "level_embed.weight" in sd:
            lvl = sd["level_embed.weight"][level].float()
            w.out_tokens = hold.f32(out_tok + lvl[None])
            w.sparse_add = hold.f32(lvl)
        else:
            w.out_tokens = hold.f32(out_tok)
        T = self._prefix(level)

        def lin_t(name):  # transposed fp32 weight + fp32 bias (token side, CUDA cores)
            return hold.f32(sd[name + ".weight"].t()), hold.f32(sd[name + ".bias"])

        img_w = [sd[T + f"layers.{l}.{n}.weight"] for l in range(2) for n in
                 ("cross_attn_token_to_image.k_proj", "cross_attn_token_to_image.v_proj", "cross_attn_image_to_token.q_proj",
                  "cross_attn_image_to_token.out_proj")]
        img_w += [sd[T + "final_attn_token_to_image.k_proj.weight"], sd[T + "final_attn_token_to_image.v_proj.weight"], sd["output_upscaling.0.weight"]]
        if self._UP_STAGES == 2:
            img_w.append(sd["output_upscaling.3.weight"])
        img_w += [sd[T + f"layers.{l}.mlp.{n}.weight"] for l in range(2) for n in ("lin1", "lin2")]  # token MLP: tensor-core GEMMs too
        terms = split_terms_needed(img_w)
        w.split_terms = terms

        for l in range(2):
            L, lp = w.layers[l], T + f"layers.{l}."
            L.sa_wq_t, L.sa_bq = lin_t(lp + "self_attn.q_proj")
            L.sa_wk_t, L.sa_bk = lin_t(lp + "self_attn.k_proj")
            L.sa_wv_t, L.sa_bv = lin_t(lp + "self_attn.v_proj")
            L.sa_wo_t, L.sa_bo = lin_t(lp + "self_attn.out_proj")
            L.n1_g, L.n1_b = hold.f32(sd[lp + "norm1.weight"]), hold.f32(sd[lp + "norm1.bias"])
            L.t2i_wq_t, L.t2i_bq = lin_t(lp + "cross_attn_token_to_image.q_proj")
            L.t2i_wo_t, L.t2i_bo = lin_t(lp + "cross_attn_token_to_image.out_proj")
            L.n2_g, L.n2_b = hold.f32(sd[lp + "norm2.weight"]), hold.f32(sd[lp + "norm2.bias"])
            L.mlp_w1_t, L.mlp_b1 = lin_t(lp + "mlp.lin1")
            L.mlp_w2_t, L.mlp_b2 = lin_t(lp + "mlp.lin2")
            if self._MLP_GEMM:
                L.mlp_w1_split = hold(split_weight(sd[lp + "mlp.lin1.weight"], terms))
                L.mlp_w2_split = hold(split_weight(sd[lp + "mlp.lin2.weight"], terms))
            L.n3_g, L.n3_b = hold.f32(sd[lp + "norm3.weight"]), hold.f32(sd[lp + "norm3.bias"])
            L.i2t_wk_t, L.i2t_bk = lin_t(lp + "cross_attn_image_to_token.k_proj")
            L.i2t_wv_t, L.i2t_bv = lin_t(lp + "cross_attn_image_to_token.v_proj")
            L.w_img = hold(split_weight(torch.cat([sd[lp + "cross_attn_token_to_image.k_proj.weight"], sd[lp + "cross_attn_token_to_image.v_proj.weight"],
                                                   sd[lp + "cross_attn_image_to_token.q_proj.weight"]], 0), terms))
            L.i2t_wo = hold(split_weight(sd[lp + "cross_attn_image_to_token.out_proj.weight"], terms))
            L.i2t_bo = hold.f32(sd[lp + "cross_attn_image_to_token.out_proj.bias"])
            L.n4_g, L.n4_b = hold.f32(sd[lp + "norm4.weight"]), hold.f32(sd[lp + "norm4.bias"])
        w.fin_wq_t, w.fin_bq = lin_t(T + "final_attn_token_to_image.q_proj")
        w.fin_wo_t, w.fin_bo = lin_t(T + "final_attn_token_to_image.out_proj")
        w.nf_g, w.nf_b = hold.f32(sd[T + "norm_final_attn.weight"]), hold.f32(sd[T + "norm_final_attn.bias"])
        w.w_img_fin = hold(split_weight(torch.cat([sd[T + "final_attn_token_to_image.k_proj.weight"], sd[T + "final_attn_token_to_image.v_proj.weight"]], 0), terms))
        w_up, b_up = pack_conv_transpose2x2(sd["output_upscaling.0.weight"], sd["output_upscaling.0.bias"])
        w.w_up, w.b_up = hold(split_weight(w_up, terms)), hold.f32(b_up)
        w.up_ln_g, w.up_ln_b = hold.f32(sd["output_upscaling.1.weight"]), hold.f32(sd["output_upscaling.1.bias"])
        if self._UP_STAGES == 2:
            w_up2, b_up2 = pack_conv_transpose2x2(sd["output_upscaling.3.weight"], sd["output_upscaling.3.bias"])
            w.w_up2, w.b_up2 = hold(split_weight(w_up2, terms)), hold.f32(b_up2)
        for j, (wn, bn) in enumerate((("hyp_w0_t", "hyp_b0"), ("hyp_w1_t", "hyp_b1"), ("hyp_w2_t", "hyp_b2"))):
            ws_ = torch.stack([sd[f"output_hypernetworks_mlps.{i}.layers.{j}.weight"].t() for i in range(self.num_mask_tokens)], 0)
            bs_ = torch.stack([sd[f"output_hypernetworks_mlps.{i}.layers.{j}.bias"] for i in range(self.num_mask_tokens)], 0)
            setattr(w, wn, hold.f32(ws_))
            setattr(w, bn, hold.f32(bs_))
        for j, (wn, bn) in enumerate((("iou_w0_t", "iou_b0"), ("iou_w1_t", "iou_b1"), ("iou_w2_t", "iou_b2"))):
            setattr(w, wn, hold.f32(sd[f"iou_prediction_head.layers.{j}.weight"].t()))
            setattr(w, bn, hold.f32(sd[f"iou_prediction_head.layers.{j}.bias"]))
        if level == 0:
            self._packed = (w, hold)
            self._pe_key = None
            self._levels = {}
            return self._packed
        if "upsample_2x.0.weight" in sd:  # level > 0: the embedding up-sampler in front of the level's transformer
            w_u, b_u = pack_conv_transpose2x2(sd["upsample_2x.0.weight"], sd["upsample_2x.0.bias"])
            up_terms = split_terms_needed([sd["upsample_2x.0.weight"]])
            hold.up2x = (hold(split_weight(w_u, up_terms)), up_terms, hold.f32(b_u), hold.f32(sd["upsample_2x.1.weight"]), hold.f32(sd["upsample_2x.1.bias"]))
        self._levels[level] = {"packed": (w, hold), "pe_key": None, "pe_keep": []}
        return w, hold

    def _pack(self):
        return self._pack_static(0)

    def bind_prompt_constants(self, pe_tokens: torch.Tensor, no_mask: torch.Tensor, grid: Tuple[int, int], level: int = 0,
                              identity: Optional[Tuple[torch.Tensor, ...]] = None) -> None:
        """Fold the dense positional encoding into per-position bias tables (pe W^T + b; one-time fp32 constant folding, like
        the LayerNorm-affine folding in MSQP) and record the dense (no-mask) prompt embedding.  Cached until inputs/weights change.

        The cache is keyed by the address and version counter of the ``identity`` tensors (default: the two inputs), and the
        cache keeps those tensors ALIVE: a freed temporary's address could otherwise be handed to a different tensor with the
        same version counter and hit a stale entry."""
        if level == 0:
            w, hold = self._packed or self._pack()
        else:
            self._packed or self._pack()
            if level not in self._levels:
                self._pack_static(level)
            w, hold = self._levels[level]["packed"]
        identity = tuple(identity) if identity is not None else (pe_tokens, no_mask)
        key = tuple((t.data_ptr(), t._version, tuple(t.shape), tuple(t.stride())) for t in identity) + (tuple(grid),)
        if (self._pe_key if level == 0 else self._levels[level]["pe_key"]) == key:
            return
        sd = self._sd()
        pe32 = pe_tokens.float()
        hw = pe_tokens.shape[0]
        T = self._prefix(level)
        if level == 0:
            self._pe_keep = []
            keep = self._pe_keep
        else:
            keep = self._levels[level]["pe_keep"] = []

        def table(wk, bk, bv, wq=None, bq=None):
            parts = [pe32 @ wk.float().t() + bk.float(), bv.float()[None].expand(hw, -1)]
            if wq is not None:
                parts.append(pe32 @ wq.float().t() + bq.float())
            t = torch.cat(parts, dim=1).contiguous()
            keep.append(t)
            return t.data_ptr()

        for l in range(2):
            lp = T + f"layers.{l}."
            w.layers[l].b_img = table(sd[lp + "cross_attn_token_to_image.k_proj.weight"], sd[lp + "cross_attn_token_to_image.k_proj.bias"],
                                      sd[lp + "cross_attn_token_to_image.v_proj.bias"],
                                      sd[lp + "cross_attn_image_to_token.q_proj.weight"], sd[lp + "cross_attn_image_to_token.q_proj.bias"])
        w.b_img_fin = table(sd[T + "final_attn_token_to_image.k_proj.weight"], sd[T + "final_attn_token_to_image.k_proj.bias"],
                            sd[T + "final_attn_token_to_image.v_proj.bias"])
        nm = _f32(no_mask.reshape(-1))
        keep.append(nm)
        keep.append(identity)  # pins the keyed tensors' storage for as long as the key is cached
        w.no_mask = nm.data_ptr()
        w.grid_h, w.grid_w = grid
        if level == 0:
            self._pe_key = key
        else:
            self._levels[level]["pe_key"] = key

    def run(self, emb_tokens_bf16: torch.Tensor, txt_emb_f32: torch.Tensor, prompt_img_i32: torch.Tensor, multimask_output: bool = False,
            want_depth_pool: bool = False, level: int = 0, prev_masks: Optional[torch.Tensor] = None):
        """emb tokens split-bf16 [B, hw, 512], txt fp32 [P, 256], prompt_img int32 [P] ->
        (low_res fp32 [P, n, 2h, 2w], iou fp32 [P, n], depth_pool fp32 [P, 33] | None).  bind_prompt_constants first.
        level > 0: emb tokens are the up-sampled embeddings and prev_masks fp32 [P, M, h, w] gate them."""
        w, _ = self._packed if level == 0 else self._levels[level]["packed"]
        P = txt_emb_f32.shape[0]
        hw = w.grid_h * w.grid_w
        n_out = self.num_mask_tokens - self._MULTIMASK_FIRST if multimask_output else 1
        dev = emb_tokens_bf16.device
        up = 2 ** self._UP_STAGES
        low = torch.empty(P, n_out, up * w.grid_h, up * w.grid_w, device=dev, dtype=torch.float32)
        iou = torch.empty(P, n_out, device=dev, dtype=torch.float32)
        pool = torch.empty(P, 33, device=dev, dtype=torch.float32) if want_depth_pool else None
        if P == 0:
            return low, iou, pool
        ws = self._ws.get(_lib.lib().wg_mask_decoder_workspace_bytes_ex(P, hw, self._UP_STAGES), dev)
        n_prev = 0
        if prev_masks is not None:
            assert prev_masks.dtype == torch.float32 and prev_masks.is_contiguous() and prev_masks.shape[0] == P
            assert tuple(prev_masks.shape[-2:]) == (w.grid_h, w.grid_w), "previous masks must have this level's grid size"
            n_prev = prev_masks.shape[1]
        n_images = emb_tokens_bf16.numel() // (hw * emb_tokens_bf16.shape[-1])  # layer-0 projections are shared by an image's prompts
        _lib.check(_lib.lib().wg_mask_decoder_forward_images(C.byref(w), emb_tokens_bf16.data_ptr(), n_images, txt_emb_f32.data_ptr(),
                                                             prompt_img_i32.data_ptr(), P, int(multimask_output),
                                                             None if prev_masks is None else prev_masks.data_ptr(), n_prev,
                                                             low.data_ptr(), iou.data_ptr(), None if pool is None else pool.data_ptr(),
                                                             ws.data_ptr(), ws.numel(), _stream()), "wg_mask_decoder_forward_images")
        return low, iou, pool

    def upsample_embedding(self, emb_tokens_bf16: torch.Tensor, grid: Tuple[int, int]) -> torch.Tensor:
        """upsample_2x of mask_decoder_multi_scale.py:166 on split-bf16 tokens [B, h*w, 512] -> [B, 4*h*w, 512]."""
        self._packed or self._pack()
        if 1 not in self._levels:
            self._pack_static(1)
        _, hold = self._levels[1]["packed"]
        w_u, terms, b_u, g_u, be_u = hold.up2x
        B, hw, _ = emb_tokens_bf16.shape
        out = torch.empty(B, 4 * hw, 512, device=emb_tokens_bf16.device, dtype=torch.bfloat16)
        need = _lib.lib().wg_upsample2x_workspace_bytes(B, hw)
        ws = self._ws_up.get(need, emb_tokens_bf16.device)
        _lib.check(_lib.lib().wg_upsample2x_embedding(w_u, terms, b_u, g_u, be_u, emb_tokens_bf16.data_ptr(), B, grid[0], grid[1], out.data_ptr(),
                                                      ws.data_ptr(), ws.numel(), _stream()), "wg_upsample2x_embedding")
        return out

    @torch.no_grad()
    def forward(self, image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings, multimask_output: bool, level_num: int = 0,
                previous_masks=None):
        """Reference signature.  image_embeddings [1,256,h,w]; image_pe [1,256,h,w]; sparse [S,1,256]; dense [S,256,h,w]."""
        if level_num not in (0, 1) or level_num >= max(self.image_feature_scale_num, 1) or (level_num == 0) != (previous_masks is None):
            raise NotImplementedError("MaskDecoderMultiScale: level_num 0 (no previous_masks) and, with image_feature_scale_num=2, level_num 1 "
                                      "(with previous_masks) are built")
        _need_cuda(image_embeddings, "MaskDecoderMultiScale.forward")
        S = sparse_prompt_embeddings.shape[0]
        if sparse_prompt_embeddings.shape[1] != 1:
            raise NotImplementedError("MaskDecoderMultiScale: exactly one sparse (text) embedding per prompt is built")
        _, Cc, h, wd = image_embeddings.shape
        if image_embeddings.shape[0] != 1:
            raise ValueError("image_embeddings must be [1, C, h, w] (one image per call, as in the reference)")
        dt = image_embeddings.dtype
        with torch.cuda.device(image_embeddings.device):
            self._packed or self._pack()
            pe_tok = image_pe.reshape(Cc, h * wd).t().contiguous().float()  # layout change only
            # dense prompt embedding: the reference always passes no_mask_embed broadcast over (h, w)
            dense_vec = dense_prompt_embeddings[0, :, 0, 0].float().contiguous()
            emb_tok = to_split(image_embeddings.reshape(1, Cc, h * wd).permute(0, 2, 1))
            txt = sparse_prompt_embeddings.reshape(S, Cc).float().contiguous()
            pimg = torch.zeros(S, dtype=torch.int32, device=image_embeddings.device)
            # the cache identity is the CALLER's tensors (pe_tok / dense_vec are temporaries derived from them)
            ident = (image_pe, dense_prompt_embeddings)
            if level_num == 0:
                self.bind_prompt_constants(pe_tok, dense_vec, (h, wd), identity=ident)
                low, iou, _ = self.run(emb_tok, txt, pimg, multimask_output)
            else:
                # mask_decoder_multi_scale.py:165-171: up-sampled embedding gated by the previous masks, pe1 on the 2h x 2w grid;
                # the dense prompt embedding is no_mask_embed broadcast, so its bilinear resize is the same constant
                g_param = self.pe1.positional_encoding_gaussian_matrix
                pe1_key = (g_param.data_ptr(), g_param._version, h, wd)
                if getattr(self, "_pe1_cache", None) is None or self._pe1_cache[0] != pe1_key:
                    gauss = _f32(g_param)
                    F_ = gauss.shape[1]
                    pe1_tok = torch.empty(4 * h * wd, 2 * F_, device=image_embeddings.device, dtype=torch.float32)
                    _lib.check(_lib.lib().wg_dense_pe(gauss.data_ptr(), F_, 2 * h, 2 * wd, None, pe1_tok.data_ptr(), _stream()), "wg_dense_pe")
                    self._pe1_cache = (pe1_key, pe1_tok)
                pe1_tok = self._pe1_cache[1]
                self.bind_prompt_constants(pe1_tok, dense_vec, (2 * h, 2 * wd), level=1, identity=(pe1_tok, dense_prompt_embeddings))
                up_tok = self.upsample_embedding(emb_tok, (h, wd))
                low, iou, _ = self.run(up_tok, txt, pimg, multimask_output, level=1, prev_masks=previous_masks.float().contiguous())
        return low.to(dt), iou.to(dt)


class MaskDecoder(MaskDecoderMultiScale):
    """Drop-in for the standard SAM ``MaskDecoder`` (segment_anything/modeling/mask_decoder.py:16-164) -- the decoder of the
    released "SAM-1024" wiring (Path B): no level embedding, two ConvTranspose stages (256 -> 64 -> 32, masks
    at 4x the embedding grid), ``multimask_output`` returns masks 1..3.  Same CUDA kernels as the multi-scale decoder."""

    _T = "transformer."
    _UP_STAGES = 2
    _MULTIMASK_FIRST = 1  # mask_decoder.py:106-111

    def __init__(self, *, transformer_dim: int = 256, transformer=None, num_multimask_outputs: int = 3, activation=None,
                 iou_head_depth: int = 3, iou_head_hidden_dim: int = 256, seed=0):
        _SpecModule.__init__(self)
        if transformer_dim != 256 or num_multimask_outputs != 3 or iou_head_depth != 3 or iou_head_hidden_dim != 256:
            raise ValueError("MaskDecoder: the CUDA path is built for the reference's dimensions (256 / 3 / 3 / 256)")
        self.transformer_dim = transformer_dim
        self.num_multimask_outputs = num_multimask_outputs
        self.num_mask_tokens = num_multimask_outputs + 1
        self._build(specs.mask_decoder_sam_spec(transformer_dim, num_multimask_outputs), seed)
        self.image_feature_scale_num = 1
        self._ws = _Workspace()
        self._ws_up = _Workspace()
        self._pe_key = None
        self._levels = {}

    @torch.no_grad()
    def forward(self, image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings, multimask_output: bool):
        """Reference signature (mask_decoder.py:75-114)."""
        return MaskDecoderMultiScale.forward(self, image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings, multimask_output, 0, None)


# --------------------------------------------------------------------------------------------------
# A8 + A9  postprocess / threshold / score,  A10 depth extension
# --------------------------------------------------------------------------------------------------
_pp_ws = _Workspace()


@torch.no_grad()
def postprocess_masks_fused(low_res: torch.Tensor, input_size: Tuple[int, int], original_size: Tuple[int, int], target_size: Optional[int] = None,
                            want_mask: bool = True, want_score: bool = True):
    """low_res fp32 [n, Hm, Wm] -> (logits fp32 [n, H0, W0], mask uint8 | None, score fp32 [n] | None)."""
    _need_cuda(low_res, "postprocess_masks")
    assert low_res.dim() == 3 and low_res.dtype == torch.float32 and low_res.is_contiguous()
    n, Hm, Wm = low_res.shape
    T = max(input_size) if target_size is None else target_size
    H0, W0 = original_size
    dev = low_res.device
    logits = torch.empty(n, H0, W0, device=dev, dtype=torch.float32)
    mask = torch.empty(n, H0, W0, device=dev, dtype=torch.uint8) if want_mask else None
    score = torch.empty(n, device=dev, dtype=torch.float32) if want_score else None
    if n == 0:
        return logits, mask, score
    with torch.cuda.device(dev):
        ws = _pp_ws.get(_lib.lib().wg_postprocess_workspace_bytes(n, input_size[0], input_size[1]), dev)
        _lib.check(_lib.lib().wg_postprocess_masks(low_res.data_ptr(), n, Hm, Wm, T, input_size[0], input_size[1], H0, W0, logits.data_ptr(),
                                                   None if mask is None else mask.data_ptr(), None if score is None else score.data_ptr(),
                                                   ws.data_ptr(), ws.numel(), _stream()), "wg_postprocess_masks")
    return logits, mask, score


@torch.no_grad()
def postprocess_masks(masks: torch.Tensor, input_size: Tuple[int, ...], original_size: Tuple[int, ...]) -> torch.Tensor:
    """Drop-in for ``walkgptForCausalLM.postprocess_masks`` (model/walkgpt.py:749-790, vision_tower_for_mask=True):
    masks [B, C, Hm, Wm] -> [B, C, H0, W0], cast back to the input dtype."""
    b, c, Hm, Wm = masks.shape
    logits, _, _ = postprocess_masks_fused(masks.reshape(b * c, Hm, Wm).float().contiguous(), tuple(input_size[:2]), tuple(original_size[:2]),
                                           want_mask=False, want_score=False)
    return logits.reshape(b, c, *logits.shape[-2:]).to(masks.dtype)


@torch.no_grad()
def match_pred(out_mask: torch.Tensor, tgt_mask: torch.Tensor, point_coords: Optional[torch.Tensor] = None, num_points: int = 12544):
    """Drop-in for ``utils.matcher.match_pred`` (utils/matcher.py:93-128; call sites train_walkgpt.py:937,
    evaluation_walkgpt.py:751): out_mask [n_pred, H, W] logits, tgt_mask [n_tgt, H, W] -> (pred indices, target indices) of the
    minimum-cost assignment.  The cost matrix is built on the GPU (``ops.match_cost``); only the [n_pred, n_tgt] matrix crosses
    to the host for scipy's ``linear_sum_assignment``, as in the reference.  ``point_coords`` defaults to the reference's
    ``torch.rand(1, 12544, 2)`` drawn on the masks' device; pass it to make a call reproducible."""
    from scipy.optimize import linear_sum_assignment

    from . import ops

    _need_cuda(out_mask, "match_pred")
    if point_coords is None:
        point_coords = torch.rand(1, num_points, 2, device=out_mask.device)
    cost = ops.match_cost(out_mask, tgt_mask, point_coords)
    return linear_sum_assignment(cost.cpu())


class DepthHead(_SpecModule):
    """Relative-depth head.  THIS REPO'S EXTENSION -- the reference has no depth head (it emits depth as text);
    definition and checker: ``oracle/path_a.py:depth_head``.  Parity with the reference is therefore unpinned."""

    def __init__(self, in_ch: int = 32, hidden: int = 256, *, seed=0):
        super().__init__()
        assert in_ch == 32 and hidden == 256
        self._build(specs.depth_head_spec(in_ch, hidden), seed)

    def _pack(self):
        sd, hold = self._sd(), _Holder()
        self._packed = ((hold.f32(sd["0.weight"]), hold.f32(sd["0.bias"]), hold.f32(sd["2.weight"].reshape(-1)), hold.f32(sd["2.bias"])), hold)
        return self._packed

    @torch.no_grad()
    def forward(self, depth_pool: torch.Tensor, seg_offsets_i32: torch.Tensor, max_S: int) -> torch.Tensor:
        _need_cuda(depth_pool, "DepthHead.forward")
        (w1, b1, w2, b2), _ = self._packed or self._pack()
        P = depth_pool.shape[0]
        out = torch.empty(P, device=depth_pool.device, dtype=torch.float32)
        if P == 0:
            return out
        with torch.cuda.device(depth_pool.device):
            _lib.check(_lib.lib().wg_depth_head(depth_pool.data_ptr(), seg_offsets_i32.data_ptr(), seg_offsets_i32.numel() - 1, max_S, w1, b1, w2, b2,
                                                out.data_ptr(), _stream()), "wg_depth_head")
        return out


# --------------------------------------------------------------------------------------------------
# SAM ViT image encoder (Path B's pixel encoder)
# --------------------------------------------------------------------------------------------------
def sam_qkv_row_order(heads: int, head_dim: int = 80, main: int = 64) -> torch.Tensor:
    """Row permutation of the reference's ``attn.qkv`` weight ([3, heads, head_dim] rows) into the column order wg_sam_attention
    reads: [Q main | K main | V main | Q rem | K rem | V rem], main = the first 64 channels of every head, rem = the other 16."""
    hidden = heads * head_dim
    idx = []
    for lo, hi in ((0, main), (main, head_dim)):
        for part in range(3):
            for h in range(heads):
                idx.append(torch.arange(part * hidden + h * head_dim + lo, part * hidden + h * head_dim + hi))
    return torch.cat(idx)


def sam_rel_table(rel_pos_h: torch.Tensor, rel_pos_w: torch.Tensor, side: int) -> torch.Tensor:
    """[rel_pos_h ; rel_pos_w] stacked for the attention kernel's prologue MMA: windows (side 14) -> [64, 80] with the two tables at rows 0
    and 32, global (side 64) -> [256, 80] with them at rows 0 and 128.  The reference interpolates tables of another length
    (get_rel_pos, image_encoder.py:292-322); the released checkpoints never need that and it is not built."""
    n = 2 * side - 1
    if rel_pos_h.shape[0] != n or rel_pos_w.shape[0] != n:
        raise NotImplementedError(f"rel_pos tables must have 2 * {side} - 1 rows (got {rel_pos_h.shape[0]} / {rel_pos_w.shape[0]}): interpolated tables are not built")
    half = 32 if side == 14 else 128
    t = torch.zeros(2 * half, rel_pos_h.shape[1], dtype=torch.float32, device=rel_pos_h.device)
    t[:n] = rel_pos_h.float()
    t[half:half + n] = rel_pos_w.float()
    return t


class ImageEncoderViT(_SpecModule):
    """Drop-in for SAM's ``ImageEncoderViT`` (segment_anything/modeling/image_encoder.py:17-126; ViT-H as built by build_sam.py:15-22 is the
    default).  The CUDA path is built for the released geometry: 1024-pixel input, 16-pixel patches (64 x 64 map), head_dim 80, 14 x 14
    windows, relative position tables of 2 * side - 1 rows; embed_dim / depth / heads / global_attn_indexes are free."""

    def __init__(self, img_size: int = 1024, patch_size: int = 16, in_chans: int = 3, embed_dim: int = 1280, depth: int = 32, num_heads: int = 16,
                 mlp_ratio: float = 4.0, out_chans: int = 256, qkv_bias: bool = True, norm_layer=None, act_layer=None, use_abs_pos: bool = True,
                 use_rel_pos: bool = True, rel_pos_zero_init: bool = True, window_size: int = 14, global_attn_indexes: Sequence[int] = (7, 15, 23, 31),
                 *, seed=0):
        super().__init__()
        if (img_size, patch_size, in_chans, window_size, out_chans) != (1024, 16, 3, 14, 256) or not (qkv_bias and use_abs_pos and use_rel_pos):
            raise ValueError("ImageEncoderViT: the CUDA path is built for SAM's geometry (1024 / 16 / 3 channels / window 14 / 256 out, qkv bias, "
                             "absolute + relative position)")
        if embed_dim != num_heads * 80 or embed_dim % 128:
            raise ValueError("ImageEncoderViT: embed_dim must be num_heads * 80 and a multiple of 128 (ViT-H: 1280 = 16 x 80)")
        self.img_size, self.embed_dim, self.depth, self.num_heads = img_size, embed_dim, depth, num_heads
        self.mlp_dim = int(embed_dim * mlp_ratio)
        self.global_attn_indexes = tuple(global_attn_indexes)
        self._build(specs.sam_image_encoder_spec(img_size, patch_size, embed_dim, depth, num_heads, mlp_ratio, out_chans, window_size, self.global_attn_indexes), seed)
        self._ws = _Workspace()

    def _pack(self):
        sd, hold = self._sd(), _Holder()
        D, heads = self.embed_dim, self.num_heads
        order = sam_qkv_row_order(heads).to(sd["pos_embed"].device)
        blocks = (_lib.SamBlock * self.depth)()
        for i in range(self.depth):
            p, bk = f"blocks.{i}.", blocks[i]
            glob = i in self.global_attn_indexes
            bk.ln1_g, bk.ln1_b = hold.f32(sd[p + "norm1.weight"]), hold.f32(sd[p + "norm1.bias"])
            bk.w_qkv, bk.b_qkv = hold.bf16(sd[p + "attn.qkv.weight"][order]), hold.f32(sd[p + "attn.qkv.bias"][order])
            bk.rel_table = hold.bf16(sam_rel_table(sd[p + "attn.rel_pos_h"], sd[p + "attn.rel_pos_w"], 64 if glob else 14))
            bk.w_proj, bk.b_proj = hold.bf16(sd[p + "attn.proj.weight"]), hold.f32(sd[p + "attn.proj.bias"])
            bk.ln2_g, bk.ln2_b = hold.f32(sd[p + "norm2.weight"]), hold.f32(sd[p + "norm2.bias"])
            bk.w_fc1, bk.b_fc1 = hold.bf16(sd[p + "mlp.lin1.weight"]), hold.f32(sd[p + "mlp.lin1.bias"])
            bk.w_fc2, bk.b_fc2 = hold.bf16(sd[p + "mlp.lin2.weight"]), hold.f32(sd[p + "mlp.lin2.bias"])
            bk.is_global = int(glob)
        w = _lib.SamEncoderWeights()
        w.hidden, w.heads, w.mlp, w.image, w.patch, w.depth, w.out_chans = D, heads, self.mlp_dim, 1024, 16, self.depth, 256
        w.patch_w = hold.bf16(sd["patch_embed.proj.weight"].reshape(D, -1))
        w.pos_bias = hold.f32(sd["pos_embed"].reshape(-1, D).float() + sd["patch_embed.proj.bias"].float()[None])
        w.blocks = C.cast(blocks, C.c_void_p)
        w1 = sd["neck.0.weight"].reshape(256, D)
        w3 = pack_conv3x3(sd["neck.2.weight"])
        terms = split_terms_needed([w1, w3])
        w.neck.mm_hidden, w.neck.hidden, w.neck.out_chans, w.neck.split_terms = D, D, 256, terms
        w.neck.w_conv1 = hold(split_weight(w1, terms))
        w.neck.ln1_g, w.neck.ln1_b = hold.f32(sd["neck.1.weight"]), hold.f32(sd["neck.1.bias"])
        w.neck.w_conv3 = hold(split_weight(w3, terms))
        w.neck.ln2_g, w.neck.ln2_b = hold.f32(sd["neck.3.weight"]), hold.f32(sd["neck.3.bias"])
        self._packed = (w, blocks, hold)
        return self._packed

    def run(self, pixels: torch.Tensor, n_run: Optional[int] = None, want_emb: bool = True, want_x: bool = False):
        """pixels [B,3,1024,1024] fp32/bf16 contiguous -> (image embedding TOKENS split-bf16 [B, 4096, 512] | None, residual stream after
        ``n_run`` blocks fp32 [B, 4096, embed_dim] | None)."""
        B = pixels.shape[0]
        w, _, _ = self._packed or self._pack()
        n_run = self.depth if n_run is None else n_run
        dev = pixels.device
        emb = torch.empty(B, 4096, 512, device=dev, dtype=torch.bfloat16) if want_emb else None
        x = torch.empty(B, 4096, self.embed_dim, device=dev, dtype=torch.float32) if want_x else None
        ws = self._ws.get(_lib.lib().wg_sam_encoder_workspace_bytes(C.byref(w), B), dev)
        _lib.check(_lib.lib().wg_sam_encoder_forward(C.byref(w), pixels.data_ptr(), int(pixels.dtype == torch.bfloat16), B, n_run,
                                                     None if emb is None else emb.data_ptr(), None if x is None else x.data_ptr(),
                                                     ws.data_ptr(), ws.numel(), _stream()), "wg_sam_encoder_forward")
        return emb, x

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [B,3,1024,1024] -> image embeddings [B,256,64,64] in x's dtype (image_encoder.py:107-116)."""
        _need_cuda(x, "ImageEncoderViT.forward")
        assert tuple(x.shape[1:]) == (3, 1024, 1024), f"expected [B,3,1024,1024], got {tuple(x.shape)}"
        with torch.cuda.device(x.device):
            emb, _ = self.run(_as_kernel_input(x))
            return tokens_to_nchw(emb, 64, 64, x.dtype, split=True)


class _PromptIndex:
    """Per-image [SEG] offsets -> (offsets int32 [B+1] on the device, prompt -> image index int32 [P], P, max prompts per image).

    A host sequence is validated on the host and its device copies are cached by value (a fixed number of [SEG] per image -- the
    benchmark, most evaluation batches -- then costs no launch and no copy at all); a device tensor goes through the
    ``wg_prompt_index`` kernel, which clamps inconsistent offsets instead of indexing out of bounds (no host synchronisation, so
    the number of prompts per image is bounded by P for the depth head's shared memory)."""

    def __init__(self, capacity: int = 32):
        self.cache: Dict[Tuple, Tuple[torch.Tensor, torch.Tensor, int, int]] = {}
        self.capacity = capacity

    def index(self, seg_offsets, B: int, n_rows: int, dev) -> Tuple[torch.Tensor, torch.Tensor, int, int]:
        dev = torch.device(dev)
        if torch.is_tensor(seg_offsets):
            if seg_offsets.numel() != B + 1:
                raise ValueError(f"seg_offsets must have B + 1 = {B + 1} entries, got {seg_offsets.numel()}")
            offs_dev = seg_offsets.to(device=dev, dtype=torch.int32).contiguous()
            P = n_rows
            prompt_img = torch.empty(P, dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                _lib.check(_lib.lib().wg_prompt_index(offs_dev.data_ptr(), B, P, prompt_img.data_ptr(), None, _stream()), "wg_prompt_index")
            return offs_dev, prompt_img, P, P
        offs = tuple(int(v) for v in seg_offsets)
        if len(offs) != B + 1 or offs[0] != 0 or offs[-1] != n_rows or any(offs[i + 1] < offs[i] for i in range(B)):
            raise ValueError(f"seg_offsets must be B + 1 = {B + 1} non-decreasing offsets from 0 to the number of [SEG] rows ({n_rows})")
        key = (offs, dev.index)
        hit = self.cache.get(key)
        if hit is None:
            if len(self.cache) >= self.capacity:
                self.cache.pop(next(iter(self.cache)))
            counts = [offs[i + 1] - offs[i] for i in range(B)]
            host = torch.tensor(offs, dtype=torch.int32)
            img = torch.repeat_interleave(torch.arange(B, dtype=torch.int32), torch.tensor(counts, dtype=torch.long))
            hit = self.cache[key] = (host.to(dev), img.to(dev), offs[-1], max(counts + [0]))
        return hit


# --------------------------------------------------------------------------------------------------
# Path A composition: the batched grounding forward a caller uses instead of the reference's
# per-image Python loops (model/walkgpt.py:511-541, 713-743).
# --------------------------------------------------------------------------------------------------
class GroundingPath(nn.Module):
    """CLIP tower -> {MSQP, out_mm_projector+neck} ; [SEG] hidden states -> CTP -> prompt encoder -> mask decoder ->
    postprocess / threshold / score (+ depth extension), batched over all images and prompts of a batch."""

    def __init__(self, hidden_size: int = 4096, clip_layers: int = 24, select_layer: int = -2, image: int = 448, with_depth: bool = True,
                 seed: int = 0):
        super().__init__()
        args = _ClipArgs()
        args.mm_vision_select_layer = select_layer
        self.vision_tower = CLIPVisionTower(None, args, layers=clip_layers, image=image, seed=seed)
        g = image // 14
        self.grid = g
        self.image = image
        self.msqp = MultiScaleQFormerProjector(sam_dim=1024, llama_dim=hidden_size, pad_to_square=True, target_square_side=6, seed=seed)
        self.proj_neck = ProjectorNeck(1024, hidden_size, 256, seed=seed)
        self.text_hidden_fcs = nn.ModuleList([CalibratedTextProjector(hidden_size, 256, widen=2, use_residual=False, seed=seed)])
        self.prompt_encoder = PromptEncoder(256, (g, g), (image, image), 16, seed=seed)
        self.mask_decoder = MaskDecoderMultiScale(transformer_dim=256, num_multimask_outputs=3, iou_head_depth=3, iou_head_hidden_dim=256,
                                                  image_feature_scale_num=1, seed=seed)
        self.depth_head = DepthHead(seed=seed) if with_depth else None
        self.hidden_size = hidden_size
        self._prompts = _PromptIndex()

    @torch.no_grad()
    def encode_images(self, images_clip: torch.Tensor, attention_mask: Optional[torch.Tensor] = None, want_vis_tokens: bool = True):
        """First half (before the LLM): CLIP tower -> MSQP visual tokens [B,36,H] (handed to the reference LLM, which resamples them to
        16 x 16, llava_arch.py:252-259) and out_mm_projector + neck -> image embedding, split-bf16 tokens [B, hw, 512]."""
        _need_cuda(images_clip, "GroundingPath.encode_images")
        out: Dict[str, torch.Tensor] = {}
        with torch.cuda.device(images_clip.device):
            feats, _ = self.vision_tower(images_clip.to(torch.bfloat16) if images_clip.dtype != torch.bfloat16 else images_clip, attention_mask,
                                         want_mid=False)
            if want_vis_tokens:
                out["vis_tokens"] = self.msqp.run(feats, torch.bfloat16)
            _, emb = self.proj_neck.run(feats)
            out["img_emb_split"] = emb  # split-bf16 [B, hw, 512]; merge_split() gives the fp32 [B, hw, 256] embedding
        return out

    @torch.no_grad()
    def ground(self, img_emb_split: torch.Tensor, seg_hidden: torch.Tensor, seg_offsets, input_size: Optional[Tuple[int, int]] = None,
               original_size: Optional[Tuple[int, int]] = None):
        """Second half (after the LLM): [SEG] hidden states [sum S, H] -> CTP -> prompt encoder + mask decoder on the image embeddings
        of ``encode_images`` -> postprocess / threshold / score (+ depth extension)."""
        _need_cuda(img_emb_split, "GroundingPath.ground")
        dev = img_emb_split.device
        B = img_emb_split.shape[0]
        input_size = input_size or (self.image, self.image)
        original_size = original_size or input_size
        offs_dev, prompt_img, P, max_S = self._prompts.index(seg_offsets, B, seg_hidden.shape[0], dev)
        out: Dict[str, torch.Tensor] = {}
        with torch.cuda.device(dev):
            txt = self.text_hidden_fcs[0].run(_as_kernel_input(seg_hidden), torch.float32)
            out["txt_emb"] = txt
            _, pe_tok = self.prompt_encoder.dense_pe_tokens()
            self.mask_decoder._packed or self.mask_decoder._pack()
            self.mask_decoder.bind_prompt_constants(pe_tok, self.prompt_encoder.no_mask_embed.weight, (self.grid, self.grid))
            low, iou, pool = self.mask_decoder.run(img_emb_split, txt, prompt_img, False, want_depth_pool=self.depth_head is not None)
            out["low_res"], out["iou"] = low, iou
            logits, mask, score = postprocess_masks_fused(low[:, 0], input_size, original_size)
            out["logits"], out["masks"], out["scores"] = logits, mask, score
            if self.depth_head is not None:
                out["depth"] = self.depth_head(pool, offs_dev, max_S)
        return out

    @torch.no_grad()
    def forward(self, images_clip: torch.Tensor, seg_hidden: torch.Tensor, seg_offsets, attention_mask: Optional[torch.Tensor] = None,
                input_size: Optional[Tuple[int, int]] = None, original_size: Optional[Tuple[int, int]] = None, want_vis_tokens: bool = True):
        """images_clip [B,3,448,448]; seg_hidden [sum S, H] (LLM hidden states at the [SEG] positions);
        seg_offsets: int sequence / tensor [B+1].  Returns a dict of device tensors (encode_images + ground)."""
        out = self.encode_images(images_clip, attention_mask, want_vis_tokens)
        out.update(self.ground(out["img_emb_split"], seg_hidden, seg_offsets, input_size, original_size))
        return out


class GraphedGroundingPath:
    """One batch shape of ``GroundingPath.forward`` (or ``GroundingPathB.forward_from_pixels``) captured as a CUDA graph: ~275 kernel
    launches become one ``cudaGraphLaunch``, which matters for small batches where the step is launch-bound (at batch 64 it is not).
    Inputs are copied into the graph's static buffers; the returned dict holds the graph's static outputs (valid until the next call).

        g = GraphedGroundingPath(path, images, seg_hidden, seg_offsets)     # seg_offsets: host list (fixed [SEG] layout per graph)
        out = g(images2, seg_hidden2)
    """

    def __init__(self, path: nn.Module, images: torch.Tensor, seg_hidden: torch.Tensor, seg_offsets, from_pixels_b: bool = False, **kw):
        _need_cuda(images, "GraphedGroundingPath")
        self.images, self.seg = images.clone(), seg_hidden.clone()
        fn = path.forward_from_pixels if from_pixels_b else path
        offs = [int(v) for v in seg_offsets]
        cur = torch.cuda.current_stream(images.device)
        side = torch.cuda.Stream(images.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):  # warm-up outside the capture: packs weights, caches offsets / PE tables, opts kernels in
            for _ in range(2):
                fn(self.images, self.seg, offs, **kw)
        cur.wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = fn(self.images, self.seg, offs, **kw)

    def __call__(self, images: torch.Tensor, seg_hidden: torch.Tensor):
        self.images.copy_(images)
        self.seg.copy_(seg_hidden)
        self.graph.replay()
        return self.out


class GroundingPathB(nn.Module):
    """Path B, the released wiring ("SAM-1024 branch"), from the SAM image embeddings on: MSQP(sam_dim=256) over the
    [B, 4096, 256] embedding tokens ; [SEG] hidden states -> CTP -> PromptEncoder(64 x 64 grid, 1024-pixel input) -> standard SAM
    MaskDecoder -> Sam.postprocess_masks (1024^2 -> crop -> original) / threshold / score, batched over all images and prompts.
    The SAM ViT-H image encoder that produces the embeddings is not built yet: the caller passes its output."""

    def __init__(self, hidden_size: int = 4096, grid: int = 64, image: int = 1024, seed: int = 0, image_encoder: Optional[nn.Module] = None):
        super().__init__()
        self.grid, self.image, self.hidden_size = grid, image, hidden_size
        self.image_encoder = image_encoder  # ImageEncoderViT: forward_from_pixels() then starts at the pixels
        self.msqp = MultiScaleQFormerProjector(sam_dim=256, llama_dim=hidden_size, pad_to_square=True, target_square_side=6, seed=seed)
        self.text_hidden_fcs = nn.ModuleList([CalibratedTextProjector(hidden_size, 256, widen=2, use_residual=False, seed=seed)])
        self.prompt_encoder = PromptEncoder(256, (grid, grid), (image, image), 16, seed=seed)
        self.mask_decoder = MaskDecoder(transformer_dim=256, num_multimask_outputs=3, iou_head_depth=3, iou_head_hidden_dim=256, seed=seed)
        self._prompts = _PromptIndex()

    @torch.no_grad()
    def forward_from_pixels(self, images: torch.Tensor, seg_hidden: torch.Tensor, seg_offsets, input_size: Optional[Tuple[int, int]] = None,
                            original_size: Optional[Tuple[int, int]] = None, want_vis_tokens: bool = True):
        """images [B,3,1024,1024] (SAM-preprocessed pixels, model/walkgpt.py:241-260) -> SAM ViT encoder -> the rest of Path B.  The image
        embedding stays in its split-bf16 token layout between the encoder's neck and the mask decoder (no NCHW round trip)."""
        if self.image_encoder is None:
            raise _lib.WalkGPTB200Error("GroundingPathB was built without an image_encoder")
        _need_cuda(images, "GroundingPathB.forward_from_pixels")
        with torch.cuda.device(images.device):
            emb_split, _ = self.image_encoder.run(_as_kernel_input(images))
        return self.forward(None, seg_hidden, seg_offsets, input_size, original_size, want_vis_tokens, _emb_split=emb_split)

    @torch.no_grad()
    def forward(self, image_embeddings: Optional[torch.Tensor], seg_hidden: torch.Tensor, seg_offsets, input_size: Optional[Tuple[int, int]] = None,
                original_size: Optional[Tuple[int, int]] = None, want_vis_tokens: bool = True, _emb_split: Optional[torch.Tensor] = None):
        """image_embeddings [B, 256, g, g] (output of the SAM image encoder, model/walkgpt.py:713-743); seg_hidden [sum S, H];
        seg_offsets int sequence [B+1].  Returns a dict of device tensors (low_res [P, 1, 4g, 4g], logits at original_size, ...)."""
        if _emb_split is not None:  # tokens straight from ImageEncoderViT.run: split-bf16 [B, hw, 512]
            B, Cc, h, wd = _emb_split.shape[0], 256, self.grid, self.grid
            dev = _emb_split.device
            image_embeddings = merge_split(_emb_split).permute(0, 2, 1).reshape(B, Cc, h, wd)  # view for MSQP's bf16 tokens below
        _need_cuda(image_embeddings, "GroundingPathB.forward")
        dev = image_embeddings.device
        B, Cc, h, wd = image_embeddings.shape
        assert Cc == 256 and (h, wd) == (self.grid, self.grid)
        input_size = input_size or (self.image, self.image)
        original_size = original_size or input_size
        _, prompt_img, P, _ = self._prompts.index(seg_offsets, B, seg_hidden.shape[0], dev)
        out: Dict[str, torch.Tensor] = {}
        with torch.cuda.device(dev):
            tokens = image_embeddings.reshape(B, Cc, h * wd).permute(0, 2, 1)       # [B, hw, 256] (layout change only)
            if want_vis_tokens:
                out["vis_tokens"] = self.msqp.run(tokens.to(torch.bfloat16).contiguous(), torch.bfloat16)
            txt = self.text_hidden_fcs[0].run(_as_kernel_input(seg_hidden), torch.float32)
            out["txt_emb"] = txt
            _, pe_tok = self.prompt_encoder.dense_pe_tokens()
            self.mask_decoder._packed or self.mask_decoder._pack()
            self.mask_decoder.bind_prompt_constants(pe_tok, self.prompt_encoder.no_mask_embed.weight, (h, wd))
            low, iou, _ = self.mask_decoder.run(_emb_split if _emb_split is not None else to_split(tokens), txt, prompt_img, False)
            out["low_res"], out["iou"] = low, iou
            out["img_emb_split"] = _emb_split
            logits, mask, score = postprocess_masks_fused(low[:, 0], input_size, original_size, target_size=self.image)
            out["logits"], out["masks"], out["scores"] = logits, mask, score
        return out
