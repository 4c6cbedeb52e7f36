"""Parameter name/shape tables of the reference modules on the hot path.

These mirror the reference ``state_dict`` keys exactly (verified against the reference modules with
``load_state_dict(strict=True)`` by ``oracle/make_golden.py``), so reference checkpoints load unchanged.
Value = (shape, is_buffer).
"""
from __future__ import annotations

import math
from typing import Dict, Sequence, Tuple

import torch

Spec = Dict[str, Tuple[Sequence[int], bool]]


def _lin(spec: Spec, name: str, n_out: int, n_in: int, bias: bool = True) -> None:
    spec[name + ".weight"] = ((n_out, n_in), False)
    if bias:
        spec[name + ".bias"] = ((n_out,), False)


def _ln(spec: Spec, name: str, n: int) -> None:
    spec[name + ".weight"] = ((n,), False)
    spec[name + ".bias"] = ((n,), False)


def msqp_spec(sam_dim: int, llama_dim: int, d: int = 1024) -> Spec:
    """MultiScaleQFormerProjector (utils/utils_walkgpt.py:220-254)."""
    s: Spec = {"pad_token": ((1, 1, d), False), "q_x1": ((1, 12, d), False), "q_x2": ((1, 8, d), False),
               "q_x4": ((1, 8, d), False), "q_global": ((1, 4, d), False)}
    _lin(s, "sam_to_proj", d, sam_dim)
    for grp in ("cross_x1", "cross_x2", "cross_x4", "cross_glb"):
        for i in range(2):
            p = f"{grp}.{i}."
            _ln(s, p + "q_norm", d)
            _ln(s, p + "kv_norm", d)
            s[p + "attn.in_proj_weight"] = ((3 * d, d), False)
            s[p + "attn.in_proj_bias"] = ((3 * d,), False)
            _lin(s, p + "attn.out_proj", d, d)
            _ln(s, p + "ffn.0", d)
            _lin(s, p + "ffn.1", 4 * d, d)
            _lin(s, p + "ffn.3", d, 4 * d)
    _ln(s, "gate.net.0", d)
    _lin(s, "gate.net.1", 128, d)
    _lin(s, "gate.net.3", 1, 128)
    _lin(s, "to_llama", llama_dim, d)
    return s


def ctp_spec(in_dim: int, out_dim: int, widen: int = 2) -> Spec:
    """CalibratedTextProjector (utils/utils_walkgpt.py:302-319)."""
    mid = max(out_dim * widen, out_dim)
    s: Spec = {"text_type": ((1, 1, out_dim), False), "log_temp": ((1,), False)}
    _ln(s, "net.0", in_dim)
    _lin(s, "net.1", mid, in_dim)
    _lin(s, "net.3", out_dim, mid)
    _ln(s, "net.4", out_dim)
    return s


def out_mm_projector_spec(mm_hidden: int, hidden: int) -> Spec:
    """nn.Sequential(Linear(mm_hidden, 2*hidden), GELU, Linear(2*hidden, hidden)) (llava_arch.py:38-42)."""
    s: Spec = {}
    _lin(s, "0", 2 * hidden, mm_hidden)
    _lin(s, "2", hidden, 2 * hidden)
    return s


def neck_spec(embed_dim: int, out_chans: int = 256) -> Spec:
    """image_feature_neck (model/walkgpt.py:97-113)."""
    s: Spec = {"0.weight": ((out_chans, embed_dim, 1, 1), False)}
    _ln(s, "1", out_chans)
    s["2.weight"] = ((out_chans, out_chans, 3, 3), False)
    _ln(s, "3", out_chans)
    return s


def prompt_encoder_spec(embed_dim: int = 256, mask_in_chans: int = 16) -> Spec:
    """PromptEncoder (segment_anything/modeling/prompt_encoder.py:16-66)."""
    s: Spec = {"pe_layer.positional_encoding_gaussian_matrix": ((2, embed_dim // 2), True)}
    for i in range(4):
        s[f"point_embeddings.{i}.weight"] = ((1, embed_dim), False)
    s["not_a_point_embed.weight"] = ((1, embed_dim), False)
    c4 = mask_in_chans // 4
    s["mask_downscaling.0.weight"] = ((c4, 1, 2, 2), False)
    s["mask_downscaling.0.bias"] = ((c4,), False)
    _ln(s, "mask_downscaling.1", c4)
    s["mask_downscaling.3.weight"] = ((mask_in_chans, c4, 2, 2), False)
    s["mask_downscaling.3.bias"] = ((mask_in_chans,), False)
    _ln(s, "mask_downscaling.4", mask_in_chans)
    s["mask_downscaling.6.weight"] = ((embed_dim, mask_in_chans, 1, 1), False)
    s["mask_downscaling.6.bias"] = ((embed_dim,), False)
    s["no_mask_embed.weight"] = ((1, embed_dim), False)
    return s


def _attention_spec(s: Spec, p: str, dim: int, internal: int) -> None:
    for n in ("q_proj", "k_proj", "v_proj"):
        _lin(s, p + n, internal, dim)
    _lin(s, p + "out_proj", dim, internal)


def two_way_transformer_spec(prefix: str, depth: int = 2, dim: int = 256, mlp_dim: int = 2048, downsample: int = 2) -> Spec:
    """TwoWayTransformer (segment_anything/modeling/transformer.py:16-60,109-149,185-218)."""
    s: Spec = {}
    for i in range(depth):
        lp = f"{prefix}layers.{i}."
        _attention_spec(s, lp + "self_attn.", dim, dim)
        _ln(s, lp + "norm1", dim)
        _attention_spec(s, lp + "cross_attn_token_to_image.", dim, dim // downsample)
        _ln(s, lp + "norm2", dim)
        _lin(s, lp + "mlp.lin1", mlp_dim, dim)
        _lin(s, lp + "mlp.lin2", dim, mlp_dim)
        _ln(s, lp + "norm3", dim)
        _ln(s, lp + "norm4", dim)
        _attention_spec(s, lp + "cross_attn_image_to_token.", dim, dim // downsample)
    _attention_spec(s, prefix + "final_attn_token_to_image.", dim, dim // downsample)
    _ln(s, prefix + "norm_final_attn", dim)
    return s


def _decoder_heads(s: Spec, dim: int, n_mask: int, iou_hidden: int, iou_depth: int) -> None:
    for i in range(n_mask):
        p = f"output_hypernetworks_mlps.{i}.layers."
        _lin(s, p + "0", dim, dim)
        _lin(s, p + "1", dim, dim)
        _lin(s, p + "2", dim // 8, dim)
    dims = [dim] + [iou_hidden] * (iou_depth - 1) + [n_mask]
    for j in range(iou_depth):
        _lin(s, f"iou_prediction_head.layers.{j}", dims[j + 1], dims[j])


def mask_decoder_multiscale_spec(dim: int = 256, num_multimask_outputs: int = 3, scale_num: int = 1, depth: int = 2,
                                 mlp_dim: int = 2048, iou_hidden: int = 256, iou_depth: int = 3) -> Spec:
    """MaskDecoderMultiScale (segment_anything/modeling/mask_decoder_multi_scale.py:16-85)."""
    s: Spec = {}
    for lvl in range(scale_num):
        s.update(two_way_transformer_spec(f"transformer.{lvl}.", depth, dim, mlp_dim))
    n_mask = num_multimask_outputs + 1
    s["iou_token.weight"] = ((1, dim), False)
    s["mask_tokens.weight"] = ((n_mask, dim), False)
    s["output_upscaling.0.weight"] = ((dim, dim // 8, 2, 2), False)
    s["output_upscaling.0.bias"] = ((dim // 8,), False)
    _ln(s, "output_upscaling.1", dim // 8)
    s["upsample_2x.0.weight"] = ((dim, dim, 2, 2), False)
    s["upsample_2x.0.bias"] = ((dim,), False)
    _ln(s, "upsample_2x.1", dim)
    s["pe1.positional_encoding_gaussian_matrix"] = ((2, dim // 2), True)
    _decoder_heads(s, dim, n_mask, iou_hidden, iou_depth)
    s["level_embed.weight"] = ((scale_num, dim), False)
    return s


def mask_decoder_sam_spec(dim: int = 256, num_multimask_outputs: int = 3, depth: int = 2, mlp_dim: int = 2048,
                          iou_hidden: int = 256, iou_depth: int = 3) -> Spec:
    """MaskDecoder (segment_anything/modeling/mask_decoder.py:16-73)."""
    s: Spec = two_way_transformer_spec("transformer.", depth, dim, mlp_dim)
    n_mask = num_multimask_outputs + 1
    s["iou_token.weight"] = ((1, dim), False)
    s["mask_tokens.weight"] = ((n_mask, dim), False)
    s["output_upscaling.0.weight"] = ((dim, dim // 4, 2, 2), False)
    s["output_upscaling.0.bias"] = ((dim // 4,), False)
    _ln(s, "output_upscaling.1", dim // 4)
    s["output_upscaling.3.weight"] = ((dim // 4, dim // 8, 2, 2), False)
    s["output_upscaling.3.bias"] = ((dim // 8,), False)
    _decoder_heads(s, dim, n_mask, iou_hidden, iou_depth)
    return s


def sam_image_encoder_spec(img_size: int = 1024, patch: int = 16, embed: int = 1280, depth: int = 32, heads: int = 16, mlp_ratio: float = 4.0,
                           out_chans: int = 256, window_size: int = 14, global_attn_indexes: Sequence[int] = (7, 15, 23, 31)) -> Spec:
    """ImageEncoderViT (segment_anything/modeling/image_encoder.py:17-116; defaults = SAM ViT-H, build_sam.py:15-22).  SURVEY 8(f) row 1:
    only the parameter table and the oracle exist so far."""
    g = img_size // patch
    hd = embed // heads
    s: Spec = {"pos_embed": ((1, g, g, embed), False), "patch_embed.proj.weight": ((embed, 3, patch, patch), False),
               "patch_embed.proj.bias": ((embed,), False)}
    for i in range(depth):
        p = f"blocks.{i}."
        side = g if i in global_attn_indexes else window_size
        _ln(s, p + "norm1", embed)
        s[p + "attn.rel_pos_h"] = ((2 * side - 1, hd), False)
        s[p + "attn.rel_pos_w"] = ((2 * side - 1, hd), False)
        _lin(s, p + "attn.qkv", 3 * embed, embed)
        _lin(s, p + "attn.proj", embed, embed)
        _ln(s, p + "norm2", embed)
        _lin(s, p + "mlp.lin1", int(embed * mlp_ratio), embed)
        _lin(s, p + "mlp.lin2", embed, int(embed * mlp_ratio))
    s["neck.0.weight"] = ((out_chans, embed, 1, 1), False)
    _ln(s, "neck.1", out_chans)
    s["neck.2.weight"] = ((out_chans, out_chans, 3, 3), False)
    _ln(s, "neck.3", out_chans)
    return s


def depth_head_spec(in_ch: int = 32, hidden: int = 256) -> Spec:
    """Relative-depth head: THIS REPO'S EXTENSION (no reference counterpart; see oracle.path_a.depth_head)."""
    s: Spec = {}
    _lin(s, "0", hidden, in_ch)
    _lin(s, "2", 1, hidden)
    return s


# --------------------------------------------------------------------------------------------------
# deterministic synthetic weights (benchmarks / tests run without checkpoints: there is no network)
# --------------------------------------------------------------------------------------------------
def hash_name(name: str) -> int:
    h = 2166136261
    for ch in name.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h



def init_tensor(name: str, shape: Sequence[int], seed: int = 0) -> torch.Tensor:
    """Deterministic random init keyed by (parameter name, seed); independent of creation order."""
    gen = torch.Generator().manual_seed((hash_name(name) + 1000003 * seed) % (2 ** 31))
    leaf = name.split(".")[-1]
    n = 1
    for s in shape:
        n *= s
    r = torch.randn(n, generator=gen, dtype=torch.float32).reshape(tuple(shape))
    if name.endswith("positional_encoding_gaussian_matrix"):
        return r
    if leaf == "log_temp":
        return 0.1 * r
    if len(shape) == 1 and leaf == "weight":  # LayerNorm / LayerNorm2d scale
        return 1.0 + 0.1 * r
    if len(shape) == 1:  # biases, class_embedding
        return 0.02 * r if leaf == "bias" else 0.5 * r
    fan_in = n // shape[0]
    if leaf == "pos_embed":  # absolute position table of the SAM encoder [1, g, g, C]
        return 0.1 * r
    if len(shape) == 4:  # convolutions
        if "upscaling" in name or "upsample_2x" in name:
            fan_in = shape[0]  # ConvTranspose2d weight is [Cin, Cout, kh, kw]
        return r / math.sqrt(max(fan_in, 1))
    if "position_embedding" in name:
        return 0.1 * r
    parent = name.split(".")[-2] if "." in name else name
    if leaf in ("pad_token", "q_x1", "q_x2", "q_x4", "q_global", "text_type") or (
            leaf == "weight" and (parent.endswith("_token") or parent.endswith("_tokens") or parent.endswith("_embed")
                                  or name.startswith("point_embeddings."))):
        return 0.5 * r  # embedding tables / learned tokens
    return r / math.sqrt(max(fan_in, 1))




def make_state_dict(spec: Spec, seed: int = 0, prefix: str = "") -> Dict[str, torch.Tensor]:
    return {prefix + k: init_tensor(k, shape, seed) for k, (shape, _) in spec.items()}
