"""fp32 CPU restatement of WalkGPT's pixel-grounding forward path (TEST INFRASTRUCTURE ONLY).

Every function works on a flat ``state_dict`` (reference parameter names) and cites the reference
lines it follows (paths relative to the reference repo root).  Nothing here is a product path.

The optional ``num`` argument (a :class:`Numerics`) lets the tests build a *storage-faithful* model of
the CUDA path (bf16 operands / bf16 activations, fp32 accumulation) so that per-kernel comparisons can
use tight tolerances.  The default ``FP32`` numerics is the oracle proper.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


class Numerics:
    """Where the CUDA path rounds.  ``q`` rounds a tensor to the storage type of activations / weights."""

    def __init__(self, q: Optional[Callable[[torch.Tensor], torch.Tensor]] = None):
        self.q = q or (lambda t: t)

    def linear(self, x, w, b=None):
        y = F.linear(self.q(x), self.q(w))
        return y if b is None else y + b

    def store(self, x):
        return self.q(x)


FP32 = Numerics()
BF16 = Numerics(lambda t: t.to(torch.bfloat16).to(torch.float32))


def _ln(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def _gelu(x):  # nn.GELU() default = exact erf form
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


# --------------------------------------------------------------------------------------------------
# A1  CLIP ViT-L/14 tower.  Reference: model/llava_walkgpt/model/multimodal_encoder/clip_encoder.py:61-98
# (feature_select + forward); the arithmetic is HF transformers' CLIPVisionModel (pinned 4.31.0 in
# requirements.txt:195): CLIPVisionEmbeddings -> pre_layrnorm -> N x CLIPEncoderLayer (pre-LN, quick_gelu).
# Key-padding mask: custom_clip.py:27-38 (_expand_mask, additive finfo.min on key columns) built in
# llava_arch.py:160-193 from a nearest-downsampled pixel-valid map with the CLS key always valid.
# --------------------------------------------------------------------------------------------------
def clip_key_valid_from_sizes(sizes: Sequence[Tuple[int, int]], image: int = 448, patch: int = 14) -> torch.Tensor:
    """[B, 1+g*g] float {0,1}: restates llava_arch.py:160-193 (valid rectangle -> nearest resize -> CLS=1)."""
    g = image // patch
    m = torch.zeros(len(sizes), image, image)
    for i, (h, w) in enumerate(sizes):
        m[i, :h, :w] = 1
    m = F.interpolate(m[:, None], size=(g, g), mode="nearest")[:, 0].flatten(1)
    return torch.cat([torch.ones(len(sizes), 1), m], dim=1)


def clip_hidden_states(sd: SD, pixels: torch.Tensor, key_valid: Optional[torch.Tensor] = None,
                       prefix: str = "vision_tower.vision_model.", num: Numerics = FP32,
                       n_layers: Optional[int] = None, heads: int = 16) -> List[torch.Tensor]:
    """All hidden states [embeddings-after-pre-LN, layer1, ...] like HF ``output_hidden_states=True``."""
    p = prefix
    wconv = sd[p + "embeddings.patch_embedding.weight"]
    D, _, ps, _ = wconv.shape
    B = pixels.shape[0]
    # conv k=stride=patch, no bias  ==  per-patch linear over (c, dy, dx)
    patches = F.unfold(pixels.float(), kernel_size=ps, stride=ps).transpose(1, 2)  # [B, g*g, 3*ps*ps]
    x = num.linear(patches, wconv.reshape(D, -1))
    cls = sd[p + "embeddings.class_embedding"].reshape(1, 1, D).expand(B, 1, D)
    x = torch.cat([cls, x], dim=1) + sd[p + "embeddings.position_embedding.weight"][None, : x.shape[1] + 1]
    x = _ln(x, sd[p + "pre_layrnorm.weight"], sd[p + "pre_layrnorm.bias"], 1e-5)
    T = x.shape[1]
    hd = D // heads
    add_mask = None
    if key_valid is not None:  # custom_clip.py:27-38
        inv = 1.0 - key_valid.float()[:, None, None, :]
        add_mask = inv.masked_fill(inv.bool(), torch.finfo(torch.float32).min)
    if n_layers is None:
        n_layers = 0
        while (p + f"encoder.layers.{n_layers}.layer_norm1.weight") in sd:
            n_layers += 1
    hs = [x]
    for i in range(n_layers):
        lp = p + f"encoder.layers.{i}."
        h = _ln(x, sd[lp + "layer_norm1.weight"], sd[lp + "layer_norm1.bias"], 1e-5)
        h = num.store(h)
        q = num.linear(h, sd[lp + "self_attn.q_proj.weight"], sd[lp + "self_attn.q_proj.bias"])
        k = num.linear(h, sd[lp + "self_attn.k_proj.weight"], sd[lp + "self_attn.k_proj.bias"])
        v = num.linear(h, sd[lp + "self_attn.v_proj.weight"], sd[lp + "self_attn.v_proj.bias"])
        q, k, v = (num.store(t).view(B, T, heads, hd).transpose(1, 2) for t in (q, k, v))
        s = (q @ k.transpose(-1, -2)) * (hd ** -0.5)
        if add_mask is not None:
            s = s + add_mask
        a = torch.softmax(s, dim=-1)
        o = (num.store(a) @ v).transpose(1, 2).reshape(B, T, D)
        o = num.linear(num.store(o), sd[lp + "self_attn.out_proj.weight"], sd[lp + "self_attn.out_proj.bias"])
        x = x + o
        h = num.store(_ln(x, sd[lp + "layer_norm2.weight"], sd[lp + "layer_norm2.bias"], 1e-5))
        h = num.linear(h, sd[lp + "mlp.fc1.weight"], sd[lp + "mlp.fc1.bias"])
        h = num.store(h * torch.sigmoid(1.702 * h))  # quick_gelu
        h = num.linear(h, sd[lp + "mlp.fc2.weight"], sd[lp + "mlp.fc2.bias"])
        x = x + h
        hs.append(x)
    return hs


def clip_tower(sd: SD, pixels, key_valid=None, select_layer: int = -2, prefix="vision_tower.vision_model.",
               num: Numerics = FP32, total_layers: int = 24):
    """clip_encoder.py:61-69,71-98: returns (hs[select_layer][:,1:], [hs[-11][:,1:]]).

    Only the layers that are needed are evaluated (hidden_states has ``total_layers+1`` entries)."""
    n_states = total_layers + 1
    idx_last = select_layer % n_states
    idx_mid = (-11) % n_states
    hs = clip_hidden_states(sd, pixels, key_valid, prefix, num, n_layers=max(idx_last, idx_mid))
    return hs[idx_last][:, 1:], [hs[idx_mid][:, 1:]]


# --------------------------------------------------------------------------------------------------
# A2  Multi-Scale Query Projector.  Reference: utils/utils_walkgpt.py:163-185 (CrossAttnBlock),
# :195-201 (_pool_grid_tokens), :204-217 (SegAwareGate), :220-300 (MultiScaleQFormerProjector).
# --------------------------------------------------------------------------------------------------
def _mha(sd: SD, p: str, q, kv, heads: int, num: Numerics):
    """nn.MultiheadAttention(batch_first=True) with packed in_proj (q from ``q``; k,v from ``kv``)."""
    E = q.shape[-1]
    w, b = sd[p + "in_proj_weight"], sd[p + "in_proj_bias"]
    qq = num.linear(q, w[:E], b[:E])
    kk = num.linear(kv, w[E:2 * E], b[E:2 * E])
    vv = num.linear(kv, w[2 * E:], b[2 * E:])
    B, Nq, _ = qq.shape
    Nk = kk.shape[1]
    hd = E // heads
    qq = qq.view(B, Nq, heads, hd).transpose(1, 2)
    kk = num.store(kk).view(B, Nk, heads, hd).transpose(1, 2)
    vv = num.store(vv).view(B, Nk, heads, hd).transpose(1, 2)
    a = torch.softmax((qq @ kk.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    o = (a @ vv).transpose(1, 2).reshape(B, Nq, E)
    return num.linear(o, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])


def _cross_attn_block(sd: SD, p: str, queries, kv, heads: int, num: Numerics):
    """utils_walkgpt.py:179-185."""
    q = _ln(queries, sd[p + "q_norm.weight"], sd[p + "q_norm.bias"], 1e-5)
    kvn = num.store(_ln(kv, sd[p + "kv_norm.weight"], sd[p + "kv_norm.bias"], 1e-5))
    out = queries + _mha(sd, p + "attn.", q, kvn, heads, num)
    h = _ln(out, sd[p + "ffn.0.weight"], sd[p + "ffn.0.bias"], 1e-5)
    h = _gelu(num.linear(h, sd[p + "ffn.1.weight"], sd[p + "ffn.1.bias"]))
    return out + num.linear(h, sd[p + "ffn.3.weight"], sd[p + "ffn.3.bias"])


def _seg_gate(sd: SD, p: str, x, num: Numerics):
    """utils_walkgpt.py:214-217."""
    h = _ln(x, sd[p + "net.0.weight"], sd[p + "net.0.bias"], 1e-5)
    h = _gelu(num.linear(h, sd[p + "net.1.weight"], sd[p + "net.1.bias"]))
    logit = F.linear(h, sd[p + "net.3.weight"], sd[p + "net.3.bias"])
    return x * torch.sigmoid(logit)


def _pool_grid(tokens, H, W, s):
    """utils_walkgpt.py:195-201 -- average pooling over the token grid."""
    B, L, C = tokens.shape
    g = tokens.view(B, H // s, s, W // s, s, C).mean(dim=(2, 4))
    return g.reshape(B, (H // s) * (W // s), C)


def msqp_forward(sd: SD, feats: torch.Tensor, prefix: str = "", heads: int = 8, pad_to_square: bool = True,
                 target_square_side: Optional[int] = None, grid_size=None, num: Numerics = FP32) -> torch.Tensor:
    """utils_walkgpt.py:259-300.  feats [B, L, sam_dim] -> [B, s*s, llama_dim]."""
    p = prefix
    B, L, _ = feats.shape
    if grid_size is None:
        H = int(math.sqrt(L))
        if H * H != L:
            raise ValueError(f"Token length {L} is not a perfect square.")  # utils_walkgpt.py:188-192
        W = H
    else:
        H, W = grid_size
    x1 = num.linear(feats, sd[p + "sam_to_proj.weight"], sd[p + "sam_to_proj.bias"])
    x1 = num.store(x1)
    x2 = _pool_grid(x1, H, W, 2)
    x4 = _pool_grid(x1, H, W, 4)
    xg = x1.mean(dim=1, keepdim=True)
    outs = []
    for qn, cn, kv in (("q_x1", "cross_x1", x1), ("q_x2", "cross_x2", x2), ("q_x4", "cross_x4", x4),
                       ("q_global", "cross_glb", xg)):
        kvg = _seg_gate(sd, p + "gate.", kv, num)
        q = sd[p + qn].expand(B, -1, -1)
        for li in range(2):
            q = _cross_attn_block(sd, p + f"{cn}.{li}.", q, kvg, heads, num)
        outs.append(q)
    vis = torch.cat(outs, dim=1)
    if pad_to_square:
        Q = vis.shape[1]
        s = int(math.ceil(math.sqrt(Q))) if target_square_side is None else target_square_side
        assert s * s >= Q, "target_square_side too small"
        if s * s > Q:
            vis = torch.cat([vis, sd[p + "pad_token"].expand(B, s * s - Q, -1)], dim=1)
    return num.linear(vis, sd[p + "to_llama.weight"], sd[p + "to_llama.bias"])


# --------------------------------------------------------------------------------------------------
# A3 out_mm_projector MLP (llava_arch.py:38-42,197-208) and A4 image_feature_neck (model/walkgpt.py:97-113,
# LayerNorm2d = segment_anything/modeling/common.py:31-43).
# --------------------------------------------------------------------------------------------------
def out_mm_projector_mlp(sd: SD, x, prefix: str = "", num: Numerics = FP32):
    h = num.store(_gelu(num.linear(x, sd[prefix + "0.weight"], sd[prefix + "0.bias"])))
    return num.linear(h, sd[prefix + "2.weight"], sd[prefix + "2.bias"])


def _ln2d_tokens(x, w, b, eps=1e-6):
    """LayerNorm2d over channels, applied on a channels-last [..., C] view (common.py:38-43)."""
    u = x.mean(-1, keepdim=True)
    s = (x - u).pow(2).mean(-1, keepdim=True)
    return (x - u) / torch.sqrt(s + eps) * w + b


def image_feature_neck(sd: SD, tokens: torch.Tensor, grid: int, prefix: str = "", num: Numerics = FP32):
    """tokens [B, g*g, H] (row-major grid) -> image embedding [B, 256, g, g]  (walkgpt.py:97-113)."""
    p = prefix
    B, L, Hd = tokens.shape
    w0 = sd[p + "0.weight"]
    C = w0.shape[0]
    x = num.linear(num.store(tokens), w0.reshape(C, Hd))
    x = num.store(_ln2d_tokens(x, sd[p + "1.weight"], sd[p + "1.bias"]))
    x = x.view(B, grid, grid, C).permute(0, 3, 1, 2)
    x = F.conv2d(num.q(x), num.q(sd[p + "2.weight"]), padding=1)
    x = x.permute(0, 2, 3, 1)
    x = _ln2d_tokens(x, sd[p + "3.weight"], sd[p + "3.bias"])
    return x.permute(0, 3, 1, 2).contiguous()


# --------------------------------------------------------------------------------------------------
# A5  Calibrated Text Projector.  Reference: utils/utils_walkgpt.py:302-327.
# --------------------------------------------------------------------------------------------------
def ctp_forward(sd: SD, x: torch.Tensor, prefix: str = "", num: Numerics = FP32):
    p = prefix
    h = num.store(_ln(x, sd[p + "net.0.weight"], sd[p + "net.0.bias"], 1e-5))
    h = num.store(_gelu(num.linear(h, sd[p + "net.1.weight"], sd[p + "net.1.bias"])))
    h = num.linear(h, sd[p + "net.3.weight"], sd[p + "net.3.bias"])
    h = _ln(h, sd[p + "net.4.weight"], sd[p + "net.4.bias"], 1e-5)
    y = F.normalize(h + sd[p + "text_type"], dim=-1)  # text_type is [1,1,out]: 2-D input becomes 3-D
    return y * sd[p + "log_temp"].exp()


# --------------------------------------------------------------------------------------------------
# A6  Prompt encoder (text_embeds passthrough) + dense positional encoding.
# Reference: segment_anything/modeling/prompt_encoder.py:140-186, :67-76, :203-229.
# --------------------------------------------------------------------------------------------------
def dense_pe(gauss: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """PositionEmbeddingRandom.forward((h,w)) -> [C, h, w]; gauss is [2, C/2]."""
    ys = (torch.arange(h, dtype=gauss.dtype) + 0.5) / h
    xs = (torch.arange(w, dtype=gauss.dtype) + 0.5) / w
    coords = torch.stack([xs[None, :].expand(h, w), ys[:, None].expand(h, w)], dim=-1)
    c = (2 * coords - 1) @ gauss
    c = 2 * math.pi * c
    return torch.cat([torch.sin(c), torch.cos(c)], dim=-1).permute(2, 0, 1)


def prompt_encoder(sd: SD, text_embeds: torch.Tensor, grid: Tuple[int, int], prefix: str = ""):
    """Only the path the reference uses: points=boxes=masks=None, text_embeds [S,1,256]."""
    S = text_embeds.shape[0]
    sparse = text_embeds
    dense = sd[prefix + "no_mask_embed.weight"].reshape(1, -1, 1, 1).expand(S, -1, grid[0], grid[1])
    return sparse, dense


# --------------------------------------------------------------------------------------------------
# A7  Two-way transformer + mask decoders.
# Reference: segment_anything/modeling/transformer.py:62-106,151-182,220-242;
# mask_decoder_multi_scale.py:137-213 (Path A); mask_decoder.py:116-164 (Path B).
# --------------------------------------------------------------------------------------------------
def _sam_attn(sd: SD, p: str, q, k, v, heads: int, num: Numerics):
    q = num.linear(q, sd[p + "q_proj.weight"], sd[p + "q_proj.bias"])
    k = num.linear(k, sd[p + "k_proj.weight"], sd[p + "k_proj.bias"])
    v = num.linear(v, sd[p + "v_proj.weight"], sd[p + "v_proj.bias"])
    B, Nq, C = q.shape
    hd = C // heads
    q = q.view(B, Nq, heads, hd).transpose(1, 2)
    k = k.view(B, -1, heads, hd).transpose(1, 2)
    v = v.view(B, -1, heads, hd).transpose(1, 2)
    a = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, Nq, C)
    return num.linear(o, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])


def two_way_transformer(sd: SD, p: str, src, pos, tokens, depth: int = 2, heads: int = 8, num: Numerics = FP32):
    """src,pos: [P, C, h, w]; tokens [P, Nt, C] -> (queries [P,Nt,C], keys [P,hw,C])."""
    keys = src.flatten(2).permute(0, 2, 1)
    kpe = pos.flatten(2).permute(0, 2, 1)
    queries, qpe = tokens, tokens
    for i in range(depth):
        lp = p + f"layers.{i}."
        if i == 0:  # skip_first_layer_pe: output REPLACES the queries (transformer.py:155-156)
            queries = _sam_attn(sd, lp + "self_attn.", queries, queries, queries, heads, num)
        else:
            qq = queries + qpe
            queries = queries + _sam_attn(sd, lp + "self_attn.", qq, qq, queries, heads, num)
        queries = _ln(queries, sd[lp + "norm1.weight"], sd[lp + "norm1.bias"], 1e-5)
        queries = queries + _sam_attn(sd, lp + "cross_attn_token_to_image.", queries + qpe, keys + kpe, keys, heads, num)
        queries = _ln(queries, sd[lp + "norm2.weight"], sd[lp + "norm2.bias"], 1e-5)
        h = torch.relu(num.linear(queries, sd[lp + "mlp.lin1.weight"], sd[lp + "mlp.lin1.bias"]))
        queries = queries + num.linear(h, sd[lp + "mlp.lin2.weight"], sd[lp + "mlp.lin2.bias"])
        queries = _ln(queries, sd[lp + "norm3.weight"], sd[lp + "norm3.bias"], 1e-5)
        keys = keys + _sam_attn(sd, lp + "cross_attn_image_to_token.", keys + kpe, queries + qpe, queries, heads, num)
        keys = num.store(_ln(keys, sd[lp + "norm4.weight"], sd[lp + "norm4.bias"], 1e-5))
    queries = queries + _sam_attn(sd, p + "final_attn_token_to_image.", queries + qpe, keys + kpe, keys, heads, num)
    queries = _ln(queries, sd[p + "norm_final_attn.weight"], sd[p + "norm_final_attn.bias"], 1e-5)
    return queries, keys


def _mlp_relu(sd: SD, p: str, x, n_layers: int = 3):
    for j in range(n_layers):
        x = F.linear(x, sd[p + f"layers.{j}.weight"], sd[p + f"layers.{j}.bias"])
        if j < n_layers - 1:
            x = torch.relu(x)
    return x


def _ln2d(x, w, b, eps=1e-6):
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    return (x - u) / torch.sqrt(s + eps) * w[:, None, None] + b[:, None, None]


def mask_decoder_multiscale(sd: SD, image_embedding, image_pe, sparse, dense, multimask_output: bool = False,
                            level_num: int = 0, previous_masks=None, prefix: str = "", num: Numerics = FP32,
                            return_aux: bool = False):
    """MaskDecoderMultiScale.forward (mask_decoder_multi_scale.py:87-213).
    image_embedding [1,256,h,w]; image_pe [1,256,h,w]; sparse [S,Ns,256]; dense [S,256,h,w]."""
    p = prefix
    S = sparse.shape[0]
    out_tok = torch.cat([sd[p + "iou_token.weight"], sd[p + "mask_tokens.weight"]], dim=0)
    n_mask = sd[p + "mask_tokens.weight"].shape[0]
    tokens = torch.cat([out_tok[None].expand(S, -1, -1), sparse], dim=1)
    tokens = tokens + sd[p + "level_embed.weight"][level_num][None, None]
    src = image_embedding.expand(S, -1, -1, -1)
    if level_num > 0:  # :165-171
        src = F.conv_transpose2d(src, sd[p + "upsample_2x.0.weight"], sd[p + "upsample_2x.0.bias"], stride=2)
        src = _gelu(_ln2d(src, sd[p + "upsample_2x.1.weight"], sd[p + "upsample_2x.1.bias"]))
        _, _, h, w = src.shape
        prev = previous_masks.mean(dim=1)
        src = (prev[:, None].sigmoid() + 1) * src
        image_pe = dense_pe(sd[p + "pe1.positional_encoding_gaussian_matrix"], h, w)[None]
        dense = F.interpolate(dense.float(), size=(h, w), mode="bilinear", align_corners=False)
    src = src + dense
    pos = image_pe.expand(S, -1, -1, -1)
    _, c, h, w = src.shape
    hs, keys = two_way_transformer(sd, p + f"transformer.{level_num}.", num.store(src), pos, tokens, num=num)
    iou_tok = hs[:, 0]
    mask_tok = hs[:, 1:1 + n_mask]
    feat = keys.transpose(1, 2).reshape(S, c, h, w)
    wu = sd[p + "output_upscaling.0.weight"]  # ConvTranspose2d weight [Cin, Cout, 2, 2]
    up = F.conv_transpose2d(num.q(feat), num.q(wu), sd[p + "output_upscaling.0.bias"], stride=2)
    up = _gelu(_ln2d(up, sd[p + "output_upscaling.1.weight"], sd[p + "output_upscaling.1.bias"]))
    hyper = torch.stack([_mlp_relu(sd, p + f"output_hypernetworks_mlps.{i}.", mask_tok[:, i]) for i in range(n_mask)], dim=1)
    S_, cu, hu, wu_ = up.shape
    masks = (hyper @ up.view(S_, cu, hu * wu_)).view(S_, n_mask, hu, wu_)
    iou = _mlp_relu(sd, p + "iou_prediction_head.", iou_tok)
    sl = slice(0, None) if multimask_output else slice(0, 1)  # :126-132 (note: MS keeps index 0 in both cases)
    if return_aux:
        return masks[:, sl], iou[:, sl], {"upscaled": up, "hyper_in": hyper, "hs": hs, "keys": keys}
    return masks[:, sl], iou[:, sl]


def mask_decoder_sam(sd: SD, image_embedding, image_pe, sparse, dense, multimask_output: bool = False,
                     prefix: str = "", num: Numerics = FP32):
    """Standard SAM MaskDecoder.forward (mask_decoder.py:75-164) -- Path B shapes."""
    p = prefix
    S = sparse.shape[0]
    out_tok = torch.cat([sd[p + "iou_token.weight"], sd[p + "mask_tokens.weight"]], dim=0)
    n_mask = sd[p + "mask_tokens.weight"].shape[0]
    tokens = torch.cat([out_tok[None].expand(S, -1, -1), sparse], dim=1)
    src = image_embedding.expand(S, -1, -1, -1) + dense
    pos = image_pe.expand(S, -1, -1, -1)
    _, c, h, w = src.shape
    hs, keys = two_way_transformer(sd, p + "transformer.", num.store(src), pos, tokens, num=num)
    iou_tok, mask_tok = hs[:, 0], hs[:, 1:1 + n_mask]
    feat = keys.transpose(1, 2).reshape(S, c, h, w)
    up = F.conv_transpose2d(num.q(feat), num.q(sd[p + "output_upscaling.0.weight"]), sd[p + "output_upscaling.0.bias"], stride=2)
    up = _gelu(_ln2d(up, sd[p + "output_upscaling.1.weight"], sd[p + "output_upscaling.1.bias"]))
    up = _gelu(F.conv_transpose2d(up, sd[p + "output_upscaling.3.weight"], sd[p + "output_upscaling.3.bias"], stride=2))
    hyper = torch.stack([_mlp_relu(sd, p + f"output_hypernetworks_mlps.{i}.", mask_tok[:, i]) for i in range(n_mask)], dim=1)
    S_, cu, hu, wu_ = up.shape
    masks = (hyper @ up.view(S_, cu, hu * wu_)).view(S_, n_mask, hu, wu_)
    iou = _mlp_relu(sd, p + "iou_prediction_head.", iou_tok)
    sl = slice(1, None) if multimask_output else slice(0, 1)  # mask_decoder.py:106-111
    return masks[:, sl], iou[:, sl]


# --------------------------------------------------------------------------------------------------
# A8 postprocess_masks (model/walkgpt.py:749-790, vision_tower_for_mask=True; sam.py:137-172 for Path B)
# A9 mask score (model/walkgpt.py:541 == :742) and the ``> 0`` threshold (evaluation_walkgpt.py:937).
# --------------------------------------------------------------------------------------------------
def postprocess_masks(masks: torch.Tensor, input_size: Tuple[int, int], original_size: Tuple[int, int],
                      target_size: Optional[int] = None, cast_back: bool = True) -> torch.Tensor:
    dtype = masks.dtype
    T = max(input_size) if target_size is None else target_size
    m = F.interpolate(masks.float(), (T, T), mode="bilinear", align_corners=False)
    m = m[..., : input_size[0], : input_size[1]]
    m = F.interpolate(m, tuple(original_size), mode="bilinear", align_corners=False)
    return m.to(dtype) if cast_back else m


def mask_score(pred: torch.Tensor) -> torch.Tensor:
    """pred [S, H, W] logits -> [S]."""
    pos = (pred > 0).flatten(1)
    return (pred.sigmoid().flatten(1) * pos).sum(1) / (pos.sum(1) + 1e-6)


def intersection_and_union(output: torch.Tensor, target: torch.Tensor, K: int = 2, ignore_index: int = 255):
    """utils/utils.py:192-204 (histogram IoU used by the reference's validate loop)."""
    output = output.reshape(-1).clone()
    target = target.reshape(-1)
    output[target == ignore_index] = ignore_index
    inter = output[output == target]
    ai = torch.histc(inter.float(), bins=K, min=0, max=K - 1)
    ao = torch.histc(output.float(), bins=K, min=0, max=K - 1)
    at = torch.histc(target.float(), bins=K, min=0, max=K - 1)
    return ai, ao + at - ai, at



# --------------------------------------------------------------------------------------------------
# SURVEY section 8(f) "next" rows 2 and 4: the steps on either side of the grounding path.  These live inline in large
# reference functions that cannot be imported under transformers 5.5, so they are restated from the cited
# lines; tests/test_oracle_golden.py re-derives them a second time with independent torch code.
# --------------------------------------------------------------------------------------------------
def resample_visual_tokens(tokens: torch.Tensor, t: int = 16) -> torch.Tensor:
    """model/llava_walkgpt/model/llava_arch.py:252-259: [n, p*p, c] -> [n, t*t, c], bilinear on the fp32 grid, cast back."""
    n, l, c = tokens.shape
    p = int(l ** 0.5)
    assert p * p == l, f"Token count {l} is not square."
    grid = tokens.permute(0, 2, 1).reshape(n, c, p, p).float()
    grid = F.interpolate(grid, size=(t, t), mode="bilinear", align_corners=False)
    return grid.flatten(2).permute(0, 2, 1).to(dtype=tokens.dtype)


def seg_token_mask(input_ids: torch.Tensor, seg_token_idx, shift: int = 255) -> torch.Tensor:
    """model/walkgpt.py:287-306: mask over the hidden-state positions [rows, Lin + shift]."""
    ids = seg_token_idx if isinstance(seg_token_idx, (list, tuple)) else [seg_token_idx]
    m = torch.zeros_like(input_ids[:, 1:]).bool()
    for i in ids:
        m = m | (input_ids[:, 1:] == i)
    m = torch.cat([m, torch.zeros((m.shape[0], 1)).bool()], dim=1)
    return torch.cat([torch.zeros((m.shape[0], shift)).bool(), m], dim=1)


def gather_seg_rows(hidden: torch.Tensor, input_ids: torch.Tensor, seg_token_idx, offset: Sequence[int], shift: int = 255):
    """model/walkgpt.py:406-420: hidden [rows, L, H] -> (pred_embeddings [sum S, H], counts [rows], per-image offsets [B+1])."""
    mask = seg_token_mask(input_ids, seg_token_idx, shift)
    pred = hidden[mask]
    counts = mask.int().sum(-1)
    off = torch.cat([torch.zeros(1).long(), counts.cumsum(-1)], dim=0)
    return pred, counts, off[torch.as_tensor(list(offset), dtype=torch.long)]


def point_sample(inp: torch.Tensor, point_coords: torch.Tensor) -> torch.Tensor:
    """utils/matcher.py:64-90 with align_corners=False: inp [N, C, H, W], point_coords [N, P, 2] in [0, 1]^2 -> [N, C, P]."""
    out = F.grid_sample(inp.float(), 2.0 * point_coords.float().unsqueeze(2) - 1.0, align_corners=False)
    return out.squeeze(3)


def match_cost(out_mask: torch.Tensor, tgt_mask: torch.Tensor, point_coords: torch.Tensor) -> torch.Tensor:
    """utils/matcher.py:93-124 up to the cost matrix: out_mask [n_pred, H, W] logits, tgt_mask [n_tgt, H, W], point_coords
    [1, P, 2] -> C [n_pred, n_tgt] = batch_sigmoid_ce_loss (matcher.py:33-58) + batch_dice_loss (matcher.py:10-26)."""
    t = point_sample(tgt_mask[:, None], point_coords.repeat(tgt_mask.shape[0], 1, 1)).squeeze(1).float()
    x = point_sample(out_mask[:, None], point_coords.repeat(out_mask.shape[0], 1, 1)).squeeze(1).float()
    pos = F.binary_cross_entropy_with_logits(x, torch.ones_like(x), reduction="none")
    neg = F.binary_cross_entropy_with_logits(x, torch.zeros_like(x), reduction="none")
    ce = (torch.einsum("nc,mc->nm", pos, t) + torch.einsum("nc,mc->nm", neg, 1 - t)) / x.shape[1]
    sig = x.sigmoid()
    dice = 1 - (2 * torch.einsum("nc,mc->nm", sig, t) + 1) / (sig.sum(-1)[:, None] + t.sum(-1)[None, :] + 1)
    return ce + dice


# --------------------------------------------------------------------------------------------------
# A10 relative-depth head.  NOT IN THE REFERENCE: the reference emits depth as LLM text.
# This is this repo's own extension, defined here so the CUDA tail has something to be checked against:
#   pooled[s, c]   = sum_xy sigmoid(logit[s]) * upscaled[s, c] / (sum_xy sigmoid(logit[s]) + 1e-6)
#   raw[s]         = W2 . gelu(W1 . pooled[s] + b1) + b2            (W1 [256,32], W2 [1,256])
#   depth[s]       = (raw[s] - min_s raw) / (max_s raw - min_s raw + 1e-6)   over the S masks of one image
# PARITY UNPINNED: there is no reference implementation to compare with.
# --------------------------------------------------------------------------------------------------
def depth_head(sd: SD, low_res_logits: torch.Tensor, upscaled: torch.Tensor, prefix: str = "depth_head."):
    """low_res_logits [S,Hm,Wm]; upscaled [S,32,Hm,Wm] -> relative depth [S] in [0,1]."""
    wgt = torch.sigmoid(low_res_logits)[:, None]
    pooled = (wgt * upscaled).sum(dim=(2, 3)) / (wgt.sum(dim=(2, 3)) + 1e-6)
    h = _gelu(F.linear(pooled, sd[prefix + "0.weight"], sd[prefix + "0.bias"]))
    raw = F.linear(h, sd[prefix + "2.weight"], sd[prefix + "2.bias"])[:, 0]
    lo, hi = raw.min(), raw.max()
    return (raw - lo) / (hi - lo + 1e-6)


# --------------------------------------------------------------------------------------------------
# Path A composition (SURVEY §8 "Canonical composition").
# --------------------------------------------------------------------------------------------------
def path_a_forward(weights: Dict[str, SD], pixels: torch.Tensor, seg_hidden: torch.Tensor, seg_offsets: Sequence[int],
                   key_valid=None, select_layer: int = -2, num: Numerics = FP32, clip_layers: int = 24,
                   input_size=(448, 448), original_size=(448, 448), f_last: Optional[torch.Tensor] = None):
    """weights: {"clip","msqp","proj","neck","ctp","prompt","decoder"} state dicts (module-local names).
    Returns dict of every stage output.  f_last: the tower's output computed elsewhere (bench.py's CPU arm runs HF CLIP itself)."""
    out = {}
    if f_last is None:
        f_last, f_mid = clip_tower(weights["clip"], pixels, key_valid, select_layer, num=num, total_layers=clip_layers)
    out["f_last"] = f_last
    B, L, _ = f_last.shape
    g = int(math.sqrt(L))
    out["vis_tokens"] = msqp_forward(weights["msqp"], f_last, target_square_side=6, num=num)
    proj = out_mm_projector_mlp(weights["proj"], f_last, num=num)
    out["proj"] = proj
    emb = image_feature_neck(weights["neck"], proj, g, num=num)
    out["img_emb"] = emb
    txt = ctp_forward(weights["ctp"], seg_hidden[None], num=num)[0]
    out["txt_emb"] = txt
    pe = dense_pe(weights["prompt"]["pe_layer.positional_encoding_gaussian_matrix"], g, g)[None]
    low, iou, logits, scores = [], [], [], []
    for b in range(B):
        t = txt[seg_offsets[b]:seg_offsets[b + 1]]
        if t.shape[0] == 0:
            continue
        sparse, dense = prompt_encoder(weights["prompt"], t[:, None], (g, g))
        m, i = mask_decoder_multiscale(weights["decoder"], emb[b:b + 1], pe, sparse, dense, False, 0, num=num)
        low.append(m)
        iou.append(i)
        pm = postprocess_masks(m, input_size, original_size)
        logits.append(pm[:, 0])
        scores.append(mask_score(pm[:, 0]))
    out["low_res"] = torch.cat(low) if low else torch.zeros(0, 1, 2 * g, 2 * g)
    out["iou"] = torch.cat(iou) if iou else torch.zeros(0, 1)
    out["logits"] = torch.cat(logits) if logits else torch.zeros(0, *original_size)
    out["scores"] = torch.cat(scores) if scores else torch.zeros(0)
    return out
