"""fp32 CPU restatement of the SAM ViT image encoder, the pixel encoder of the released WalkGPT wiring (Path B).
TEST INFRASTRUCTURE ONLY.  Pinned by tests/golden/sam_encoder_small.pt and sam_encoder_1024.pt, both produced by the reference
ImageEncoderViT (oracle/make_golden.py); the CUDA path is walkgpt_b200/csrc/sam_encoder.cu + sam_attention.cu.  Works on a flat state_dict with the reference parameter names; every function cites the reference lines
(segment_anything/modeling/image_encoder.py) it follows.
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def rel_pos_table(q_size: int, k_size: int, rel_pos: torch.Tensor) -> torch.Tensor:
    """get_rel_pos (image_encoder.py:292-322): [q_size, k_size, C] rows of the (optionally linearly resized) table picked by the
    relative coordinate q - k + (k_size - 1), with the coarser axis stretched when the sizes differ."""
    span = int(2 * max(q_size, k_size) - 1)
    if rel_pos.shape[0] != span:
        rel_pos = F.interpolate(rel_pos.t()[None], size=span, mode="linear")[0].t()
    qs = torch.arange(q_size)[:, None] * max(k_size / q_size, 1.0)
    ks = torch.arange(k_size)[None, :] * max(q_size / k_size, 1.0)
    idx = (qs - ks) + (k_size - 1) * max(q_size / k_size, 1.0)
    return rel_pos[idx.long()]


def attention_core_rel_pos(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, rel_pos_h: torch.Tensor, rel_pos_w: torch.Tensor, side: int) -> torch.Tensor:
    """The attention proper (image_encoder.py:231-241 + add_decomposed_rel_pos :325-361) on given q / k / v [N, side*side, d]:
    softmax(q d^-0.5 k^T + rel_h + rel_w) v with the position terms from the UNSCALED query."""
    N, L, d = q.shape
    s = (q * d ** -0.5) @ k.transpose(-1, -2)
    Rh, Rw = rel_pos_table(side, side, rel_pos_h), rel_pos_table(side, side, rel_pos_w)
    qg = q.reshape(N, side, side, d)
    bias_h = torch.einsum("nhwc,hkc->nhwk", qg, Rh)
    bias_w = torch.einsum("nhwc,wkc->nhwk", qg, Rw)
    s = (s.view(N, side, side, side, side) + bias_h[..., :, None] + bias_w[..., None, :]).view(N, L, L)
    return s.softmax(-1) @ v


def attention_rel_pos(sd: SD, p: str, x: torch.Tensor, heads: int) -> torch.Tensor:
    """Attention.forward + add_decomposed_rel_pos (image_encoder.py:222-247, 325-361): x [N, H, W, C] (a window batch or the whole
    map) -> [N, H, W, C].  The position term uses the UNSCALED query."""
    N, H, W, C = x.shape
    d = C // heads
    qkv = F.linear(x, sd[p + "qkv.weight"], sd[p + "qkv.bias"]).reshape(N, H * W, 3, heads, d)
    q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))        # [N, heads, HW, d]
    s = (q * d ** -0.5) @ k.transpose(-1, -2)                             # [N, heads, HW, HW]
    Rh = rel_pos_table(H, H, sd[p + "rel_pos_h"])                         # [H, H, d]
    Rw = rel_pos_table(W, W, sd[p + "rel_pos_w"])
    qg = q.reshape(N, heads, H, W, d)
    bias_h = torch.einsum("nahwc,hkc->nahwk", qg, Rh)                     # [N, heads, H, W, kh]
    bias_w = torch.einsum("nahwc,wkc->nahwk", qg, Rw)                     # [N, heads, H, W, kw]
    s = (s.view(N, heads, H, W, H, W) + bias_h[..., :, None] + bias_w[..., None, :]).view(N, heads, H * W, H * W)
    o = s.softmax(-1) @ v                                                 # [N, heads, HW, d]
    o = o.permute(0, 2, 1, 3).reshape(N, H, W, C)
    return F.linear(o, sd[p + "proj.weight"], sd[p + "proj.bias"])


def window_partition(x: torch.Tensor, ws: int):
    """image_encoder.py:250-270: [B, H, W, C] -> [B * nW, ws, ws, C] with zero padding at the bottom / right."""
    B, H, W, C = x.shape
    ph, pw = (ws - H % ws) % ws, (ws - W % ws) % ws
    x = F.pad(x, (0, 0, 0, pw, 0, ph))
    Hp, Wp = H + ph, W + pw
    x = x.view(B, Hp // ws, ws, Wp // ws, ws, C).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(-1, ws, ws, C), (Hp, Wp)


def window_unpartition(w: torch.Tensor, ws: int, pad_hw, hw) -> torch.Tensor:
    """image_encoder.py:273-289."""
    (Hp, Wp), (H, W) = pad_hw, hw
    B = w.shape[0] // ((Hp // ws) * (Wp // ws))
    x = w.view(B, Hp // ws, Wp // ws, ws, ws, -1).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, -1)
    return x[:, :H, :W]


def image_encoder_vit(sd: SD, pixels: torch.Tensor, heads: int, window_size: int, global_attn_indexes: Sequence[int], prefix: str = "",
                      ln_eps: float = 1e-6):
    """ImageEncoderViT.forward (image_encoder.py:107-116) with Block.forward (:178-192), MLPBlock (common.py:13-26: GELU erf) and the
    neck (:88-105: 1x1 conv, LayerNorm2d, 3x3 conv pad 1, LayerNorm2d, both eps 1e-6): pixels [B, 3, S, S] -> [B, out_chans, g, g].
    ln_eps: build_sam.py:73 constructs the blocks with LayerNorm(eps=1e-6)."""
    p = prefix
    w = sd[p + "patch_embed.proj.weight"]
    x = F.conv2d(pixels, w, sd[p + "patch_embed.proj.bias"], stride=w.shape[-1]).permute(0, 2, 3, 1)  # [B, g, g, C]
    x = x + sd[p + "pos_embed"]
    C = x.shape[-1]
    depth = 1 + max(int(k[len(p) + 7:].split(".")[0]) for k in sd if k.startswith(p + "blocks."))
    for i in range(depth):
        b = f"{p}blocks.{i}."
        h = F.layer_norm(x, (C,), sd[b + "norm1.weight"], sd[b + "norm1.bias"], ln_eps)
        if i in global_attn_indexes:
            h = attention_rel_pos(sd, b + "attn.", h, heads)
        else:
            H, W = h.shape[1:3]
            win, pad_hw = window_partition(h, window_size)
            h = window_unpartition(attention_rel_pos(sd, b + "attn.", win, heads), window_size, pad_hw, (H, W))
        x = x + h
        h = F.layer_norm(x, (C,), sd[b + "norm2.weight"], sd[b + "norm2.bias"], ln_eps)
        h = F.linear(F.gelu(F.linear(h, sd[b + "mlp.lin1.weight"], sd[b + "mlp.lin1.bias"])), sd[b + "mlp.lin2.weight"], sd[b + "mlp.lin2.bias"])
        x = x + h

    def ln2d(t, wt, bs):
        u = t.mean(-1, keepdim=True)
        v = (t - u).pow(2).mean(-1, keepdim=True)
        return (t - u) / torch.sqrt(v + 1e-6) * wt + bs

    y = ln2d(F.linear(x, sd[p + "neck.0.weight"].flatten(1)), sd[p + "neck.1.weight"], sd[p + "neck.1.bias"])
    y = F.conv2d(y.permute(0, 3, 1, 2), sd[p + "neck.2.weight"], padding=1).permute(0, 2, 3, 1)
    y = ln2d(y, sd[p + "neck.3.weight"], sd[p + "neck.3.bias"])
    return y.permute(0, 3, 1, 2).contiguous()


def path_b_from_embeddings(weights, image_embeddings: torch.Tensor, seg_hidden: torch.Tensor, seg_offsets, input_size=(1024, 1024),
                           original_size=(1024, 1024), image: int = 1024):
    """The released wiring after the image encoder (SURVEY section 8 "Path B"): MSQP(sam_dim=256) over the embedding tokens; CTP ->
    PromptEncoder -> SAM MaskDecoder (mask_decoder.py:75-164, multimask_output=False) -> Sam.postprocess_masks (sam.py:137-172) ->
    score (model/walkgpt.py:541).  weights: {"msqp","ctp","prompt","decoder"} state dicts.  Composition of oracle.path_a functions."""
    from . import path_a

    B, C, h, w = image_embeddings.shape
    out = {"vis_tokens": path_a.msqp_forward(weights["msqp"], image_embeddings.flatten(2).permute(0, 2, 1), target_square_side=6)}
    txt = path_a.ctp_forward(weights["ctp"], seg_hidden[None])[0]
    out["txt_emb"] = txt
    pe = path_a.dense_pe(weights["prompt"]["pe_layer.positional_encoding_gaussian_matrix"], h, w)[None]
    low, iou, logits, scores = [], [], [], []
    for b in range(B):
        t = txt[seg_offsets[b]:seg_offsets[b + 1]]
        if t.shape[0] == 0:
            continue
        sparse, dense = path_a.prompt_encoder(weights["prompt"], t[:, None], (h, w))
        m, i = path_a.mask_decoder_sam(weights["decoder"], image_embeddings[b:b + 1], pe, sparse, dense, multimask_output=False)
        low.append(m)
        iou.append(i)
        pm = path_a.postprocess_masks(m, input_size, original_size, target_size=image)
        logits.append(pm[:, 0])
        scores.append(path_a.mask_score(pm[:, 0]))
    out["low_res"], out["iou"] = torch.cat(low), torch.cat(iou)
    out["logits"], out["scores"] = torch.cat(logits), torch.cat(scores)
    return out
