"""Generate tests/golden/*.pt by EXECUTING THE REFERENCE (run in the build container only; /root/reference and HF
transformers must be importable).  TEST INFRASTRUCTURE ONLY.

For every hot-path module: build the reference nn.Module, load this repo's seeded synthetic state_dict with
``load_state_dict(strict=True)`` (which also pins parameter names and shapes), run it on seeded inputs in fp32 on
the CPU and save inputs + outputs.  tests/test_oracle_golden.py then checks oracle/path_a.py against these files,
and the GPU tests check the CUDA path against the same files and against the oracle.

    python -m oracle.make_golden            # writes tests/golden/
"""
from __future__ import annotations

import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("WALKGPT_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from walkgpt_b200 import specs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def save(name, obj):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".pt")
    torch.save(obj, path)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.0f} KiB)")


@torch.no_grad()
def main():
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    from utils.utils_walkgpt import CalibratedTextProjector, MultiScaleQFormerProjector
    from model.segment_anything.modeling import MaskDecoder, MaskDecoderMultiScale, PromptEncoder, Sam, TwoWayTransformer
    from model.segment_anything.modeling.common import LayerNorm2d

    # ---- A5 CTP
    for in_dim in (256, 4096):
        ref = CalibratedTextProjector(in_dim, 256, widen=2, use_residual=False).eval()
        ref.load_state_dict(specs.make_state_dict(specs.ctp_spec(in_dim, 256), seed=11), strict=True)
        x3 = rnd((2, 5, in_dim), 101)
        x2 = rnd((3, in_dim), 102)
        save(f"ctp_{in_dim}", {"seed": 11, "in_dim": in_dim, "x3": x3, "y3": ref(x3), "x2": x2, "y2": ref(x2)})

    # ---- A2 MSQP (Path B width, small grid) and (Path A width, 32x32 grid)
    for tag, sam_dim, llama, B, L in (("b_small", 256, 64, 2, 64), ("a_32", 1024, 128, 1, 1024)):
        ref = MultiScaleQFormerProjector(sam_dim=sam_dim, llama_dim=llama, pad_to_square=True, target_square_side=6).eval()
        ref.load_state_dict(specs.make_state_dict(specs.msqp_spec(sam_dim, llama), seed=12), strict=True)
        x = rnd((B, L, sam_dim), 201)
        save(f"msqp_{tag}", {"seed": 12, "sam_dim": sam_dim, "llama_dim": llama, "x_shape": (B, L, sam_dim), "x_seed": 201, "y": ref(x)})

    # ---- A3 out_mm_projector MLP (llava_arch.py:38-42) and A4 neck (walkgpt.py:97-113)
    mm, hid = 128, 64
    proj = nn.Sequential(nn.Linear(mm, hid * 2), nn.GELU(), nn.Linear(hid * 2, hid)).eval()
    proj.load_state_dict(specs.make_state_dict(specs.out_mm_projector_spec(mm, hid), seed=13), strict=True)
    neck = nn.Sequential(nn.Conv2d(hid, 256, kernel_size=1, bias=False), LayerNorm2d(256),
                         nn.Conv2d(256, 256, kernel_size=3, padding=1, bias=False), LayerNorm2d(256)).eval()
    neck.load_state_dict(specs.make_state_dict(specs.neck_spec(hid, 256), seed=13), strict=True)
    x = rnd((2, 64, mm), 301)
    p = proj(x)
    e = neck(p.permute(0, 2, 1).reshape(2, hid, 8, 8))
    save("proj_neck_small", {"seed": 13, "mm": mm, "hidden": hid, "x": x, "proj": p, "emb": e})

    # ---- A6 prompt encoder + A7 decoders
    for g in (8, 32):
        pe_mod = PromptEncoder(embed_dim=256, image_embedding_size=(g, g), input_image_size=(g * 14, g * 14), mask_in_chans=16).eval()
        pe_mod.load_state_dict(specs.make_state_dict(specs.prompt_encoder_spec(256, 16), seed=14), strict=True)
        dec = MaskDecoderMultiScale(num_multimask_outputs=3, transformer=TwoWayTransformer(depth=2, embedding_dim=256, mlp_dim=2048, num_heads=8),
                                    transformer_dim=256, iou_head_depth=3, iou_head_hidden_dim=256, image_feature_scale_num=1).eval()
        dec.load_state_dict(specs.make_state_dict(specs.mask_decoder_multiscale_spec(), seed=15), strict=True)
        S = 3 if g == 8 else 2
        emb = rnd((1, 256, g, g), 401)
        txt = rnd((S, 1, 256), 402, 0.5)
        sparse, dense = pe_mod(points=None, boxes=None, masks=None, text_embeds=txt)
        pe = pe_mod.get_dense_pe()
        m1, i1 = dec(image_embeddings=emb, image_pe=pe, sparse_prompt_embeddings=sparse, dense_prompt_embeddings=dense,
                     multimask_output=False, level_num=0)
        m4, i4 = dec(image_embeddings=emb, image_pe=pe, sparse_prompt_embeddings=sparse, dense_prompt_embeddings=dense,
                     multimask_output=True, level_num=0)
        save(f"decoder_ms_g{g}", {"seed_prompt": 14, "seed_dec": 15, "grid": g, "emb_shape": (1, 256, g, g), "emb_seed": 401, "txt": txt,
                                  "dense_pe": pe if g == 8 else pe[:, ::8],
                                  "masks1": m1, "iou1": i1, "masks4": m4, "iou4": i4})
    # standard SAM decoder (Path B shapes of the shared two-way transformer)
    g = 8
    pe_mod = PromptEncoder(embed_dim=256, image_embedding_size=(g, g), input_image_size=(g * 16, g * 16), mask_in_chans=16).eval()
    pe_mod.load_state_dict(specs.make_state_dict(specs.prompt_encoder_spec(256, 16), seed=14), strict=True)
    dec = MaskDecoder(num_multimask_outputs=3, transformer=TwoWayTransformer(depth=2, embedding_dim=256, mlp_dim=2048, num_heads=8),
                      transformer_dim=256, iou_head_depth=3, iou_head_hidden_dim=256).eval()
    dec.load_state_dict(specs.make_state_dict(specs.mask_decoder_sam_spec(), seed=16), strict=True)
    emb = rnd((1, 256, g, g), 411)
    txt = rnd((2, 1, 256), 412, 0.5)
    sparse, dense = pe_mod(points=None, boxes=None, masks=None, text_embeds=txt)
    m1, i1 = dec(image_embeddings=emb, image_pe=pe_mod.get_dense_pe(), sparse_prompt_embeddings=sparse, dense_prompt_embeddings=dense,
                 multimask_output=False)
    m3, i3 = dec(image_embeddings=emb, image_pe=pe_mod.get_dense_pe(), sparse_prompt_embeddings=sparse, dense_prompt_embeddings=dense,
                 multimask_output=True)
    save("decoder_sam_g8", {"seed_prompt": 14, "seed_dec": 16, "grid": g, "emb": emb, "txt": txt, "masks1": m1, "iou1": i1, "masks3": m3, "iou3": i3})

    # ---- A8 postprocess (walkgpt.py:771-789 restated with the same two F.interpolate calls; Sam.postprocess_masks executed)
    low = rnd((3, 1, 64, 64), 501, 3.0)
    cases = {}
    for name, inp, orig in (("full", (448, 448), (448, 448)), ("pad169", (252, 448), (360, 640)), ("odd", (448, 301), (517, 347))):
        T = max(inp)
        m = F.interpolate(low.float(), (T, T), mode="bilinear", align_corners=False)
        m = m[..., : inp[0], : inp[1]]
        m = F.interpolate(m, orig, mode="bilinear", align_corners=False)
        score = (m[:, 0].sigmoid().flatten(1) * (m[:, 0] > 0).flatten(1)).sum(1) / ((m[:, 0] > 0).flatten(1).sum(1) + 1e-6)  # walkgpt.py:541
        cases[name] = {"input_size": inp, "original_size": orig, "logits_sub": m[:, 0, ::7, ::5].clone(), "score": score,
                       "pos_count": (m[:, 0] > 0).flatten(1).sum(1), "sum": m.double().sum(dim=(1, 2, 3))}
    sam_self = types.SimpleNamespace(image_encoder=types.SimpleNamespace(img_size=1024))
    low_b = rnd((2, 1, 256, 256), 502, 3.0)
    mb = Sam.postprocess_masks(sam_self, low_b, (768, 1024), (480, 640))
    cases["sam_1024"] = {"input_size": (768, 1024), "original_size": (480, 640), "logits_sub": mb[:, 0, ::7, ::5].clone(),
                         "sum": mb.double().sum(dim=(1, 2, 3))}
    save("postprocess", {"low": low, "low_b_seed": 502, "cases": cases})

    # ---- A1 CLIP tower (HF transformers is where the reference's arithmetic lives; clip_encoder.py:61-98)
    from transformers import CLIPVisionConfig, CLIPVisionModel
    L = 3
    cfg = CLIPVisionConfig(hidden_size=1024, intermediate_size=4096, num_hidden_layers=L, num_attention_heads=16, image_size=448, patch_size=14)
    cfg._attn_implementation = "eager"
    hf = CLIPVisionModel(cfg).eval()
    sd = specs.make_state_dict({k: v for k, v in __import__("walkgpt_b200.modules", fromlist=["x"]).clip_param_spec(layers=L).items()}, seed=17)
    hf.load_state_dict({k[len("vision_tower."):]: v for k, v in sd.items()}, strict=True)
    px = rnd((2, 3, 448, 448), 601)
    out = hf(px, output_hidden_states=True)
    hs = out.hidden_states
    assert len(hs) == L + 1
    # masked variant: drive the layers by hand with the additive key mask of custom_clip.py:27-38 / llava_arch.py:160-193
    sizes = [(252, 448), (448, 300)]
    valid = torch.zeros(2, 448, 448)
    for i, (h, w) in enumerate(sizes):
        valid[i, :h, :w] = 1
    kv = F.interpolate(valid[:, None], size=(32, 32), mode="nearest")[:, 0].flatten(1)
    kv = torch.cat([torch.ones(2, 1), kv], dim=-1)
    inv = 1.0 - kv[:, None, None, :].expand(2, 1, 1025, 1025)
    mask4d = inv.masked_fill(inv.bool(), torch.finfo(torch.float32).min)
    vm = hf.vision_model
    h = vm.pre_layrnorm(vm.embeddings(px))
    hs_m = [h]
    for layer in vm.encoder.layers:
        h = layer(h, mask4d)
        if isinstance(h, tuple):
            h = h[0]
        hs_m.append(h)
    save("clip_3layer", {"seed": 17, "layers": L, "px_seed": 601, "sizes": sizes, "key_valid": kv,
                         "hs_sub": [t[:, ::41, ::13].clone() for t in hs], "hs_mean": [t.double().mean() for t in hs],
                         "hs_std": [t.double().std() for t in hs],
                         "hs_masked_sub": [t[:, ::41, ::13].clone() for t in hs_m]})

    # ---- scoring helper (utils/utils.py:192-204), executed if importable
    try:
        from utils.utils import intersectionAndUnionGPU
        o = (rnd((64, 64), 701) > 0).float()
        t = (rnd((64, 64), 702) > 0).float()
        t[:4] = 255
        ai, au, at = intersectionAndUnionGPU(o.clone(), t.clone(), 2, ignore_index=255)
        save("iou_hist", {"output": o, "target": t, "inter": ai, "union": au, "target_area": at})
    except Exception as e:  # noqa: BLE001
        print("intersectionAndUnionGPU not importable here:", repr(e))


@torch.no_grad()
def make_decoder_two_level():
    """MaskDecoderMultiScale with image_feature_scale_num = 2 -- level 0, then level 1 fed with the level-0
    masks (mask_decoder_multi_scale.py:165-171: upsample_2x, previous-mask gating, pe1, transformer[1], level_embed[1]).
    Writes tests/golden/decoder_ms2_g8.pt only:  python -m oracle.make_golden two_level"""
    from model.segment_anything.modeling import MaskDecoderMultiScale, PromptEncoder, TwoWayTransformer

    g = 8
    pe_mod = PromptEncoder(embed_dim=256, image_embedding_size=(g, g), input_image_size=(g * 14, g * 14), mask_in_chans=16).eval()
    pe_mod.load_state_dict(specs.make_state_dict(specs.prompt_encoder_spec(256, 16), seed=14), strict=True)
    dec = MaskDecoderMultiScale(num_multimask_outputs=3, transformer=TwoWayTransformer(depth=2, embedding_dim=256, mlp_dim=2048, num_heads=8),
                                transformer_dim=256, iou_head_depth=3, iou_head_hidden_dim=256, image_feature_scale_num=2).eval()
    dec.load_state_dict(specs.make_state_dict(specs.mask_decoder_multiscale_spec(scale_num=2), seed=17), strict=True)
    emb = rnd((1, 256, g, g), 421)
    txt = rnd((3, 1, 256), 422, 0.5)
    sparse, dense = pe_mod(points=None, boxes=None, masks=None, text_embeds=txt)
    pe = pe_mod.get_dense_pe()
    m0, i0 = dec(image_embeddings=emb, image_pe=pe, sparse_prompt_embeddings=sparse, dense_prompt_embeddings=dense, multimask_output=True,
                 level_num=0)
    m1, i1 = dec(image_embeddings=emb, image_pe=pe, sparse_prompt_embeddings=sparse, dense_prompt_embeddings=dense, multimask_output=True,
                 level_num=1, previous_masks=m0)
    m1s, i1s = dec(image_embeddings=emb, image_pe=pe, sparse_prompt_embeddings=sparse, dense_prompt_embeddings=dense, multimask_output=False,
                   level_num=1, previous_masks=m0)
    save("decoder_ms2_g8", {"seed_prompt": 14, "seed_dec": 17, "grid": g, "emb": emb, "txt": txt, "masks_l0": m0, "iou_l0": i0,
                            "masks_l1": m1, "iou_l1": i1, "masks_l1_single": m1s, "iou_l1_single": i1s})


@torch.no_grad()
def make_match_cost():
    """the point-sampled cost matrix of match_pred (utils/matcher.py:93-128).  match_pred draws its points
    with torch.rand right after being called, so seeding the global generator before the call and before our own torch.rand
    gives the same points; the cost matrix is then rebuilt from the reference's own point_sample / batch_*_loss functions and
    the assignment is the reference's match_pred output.  Writes tests/golden/match_cost.pt:  python -m oracle.make_golden match"""
    from utils.matcher import batch_dice_loss, batch_sigmoid_ce_loss, match_pred, point_sample

    cases = []
    for k, (n_pred, n_tgt, H, W) in enumerate(((5, 4, 56, 56), (3, 6, 40, 72), (1, 1, 24, 24))):
        # blobby targets (discs) and logits that follow a shuffled, noisy version of them: the assignment is non-trivial
        yy, xx = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
        g = torch.Generator().manual_seed(500 + k)
        cx, cy, rr = torch.rand(n_tgt, generator=g) * W, torch.rand(n_tgt, generator=g) * H, 6 + torch.rand(n_tgt, generator=g) * 10
        tgt = (((xx[None] - cx[:, None, None]) ** 2 + (yy[None] - cy[:, None, None]) ** 2) < rr[:, None, None] ** 2).float()
        perm = torch.randperm(max(n_pred, n_tgt), generator=g)[:n_pred] % n_tgt
        out = (tgt[perm] * 2 - 1) * 4 + torch.randn(n_pred, H, W, generator=g) * 3
        out = out.half().float()  # values exact in fp16: keeps the fixture small
        torch.manual_seed(900 + k)
        idx = match_pred(out, tgt)
        torch.manual_seed(900 + k)
        pts = torch.rand(1, 12544, 2)
        t = point_sample(tgt[:, None], pts.repeat(n_tgt, 1, 1), align_corners=False).squeeze(1).float()
        x = point_sample(out[:, None], pts.repeat(n_pred, 1, 1), align_corners=False).squeeze(1).float()
        C = batch_sigmoid_ce_loss(x, t) + batch_dice_loss(x, t)
        cases.append({"out_mask": out.half(), "tgt_mask": tgt.to(torch.uint8), "seed": 900 + k, "cost": C, "pred_idx": torch.as_tensor(idx[0]),
                      "tgt_idx": torch.as_tensor(idx[1])})
    save("match_cost", {"cases": cases, "num_points": 12544})


@torch.no_grad()
def make_sam_encoder():
    """SURVEY 8(f) row 1 groundwork: the reference ImageEncoderViT at a small configuration that keeps every mechanism of SAM ViT-H
    (head_dim 80, windows that need padding: 12 -> 15 with window 5, one global block, decomposed relative position in both kinds of
    block, neck).  Writes tests/golden/sam_encoder_small.pt:  python -m oracle.make_golden sam_encoder"""
    from model.segment_anything.modeling.image_encoder import ImageEncoderViT

    cfg = dict(img_size=192, patch=16, embed=160, depth=3, heads=2, window_size=5, global_attn_indexes=(1,), out_chans=256)
    ref = ImageEncoderViT(img_size=cfg["img_size"], patch_size=cfg["patch"], embed_dim=cfg["embed"], depth=cfg["depth"], num_heads=cfg["heads"],
                          out_chans=cfg["out_chans"], use_rel_pos=True, window_size=cfg["window_size"],
                          global_attn_indexes=cfg["global_attn_indexes"], norm_layer=lambda d: nn.LayerNorm(d, eps=1e-6)).eval()
    # build_sam.py:73 builds the blocks with LayerNorm(eps=1e-6)
    spec = specs.sam_image_encoder_spec(cfg["img_size"], cfg["patch"], cfg["embed"], cfg["depth"], cfg["heads"], 4.0, cfg["out_chans"],
                                        cfg["window_size"], cfg["global_attn_indexes"])
    ref.load_state_dict(specs.make_state_dict(spec, seed=31), strict=True)
    x = rnd((2, 3, cfg["img_size"], cfg["img_size"]), 611)
    save("sam_encoder_small", {"cfg": cfg, "seed": 31, "pixels_seed": 611, "out": ref(x), "ln_eps": 1e-6})


def sam_encoder_full_geometry_cfg():
    return dict(img_size=1024, patch=16, embed=640, depth=2, heads=8, window_size=14, global_attn_indexes=(1,), out_chans=256)


@torch.no_grad()
def make_sam_encoder_full_geometry():
    """The reference ImageEncoderViT at SAM's REAL geometry -- 1024-pixel input, 64 x 64 map, head_dim 80, 14 x 14 windows padded 64 -> 70,
    one windowed and one global block, relative-position tables of 27 / 127 rows, neck -- at a reduced width (8 heads) so that the file
    stays small.  Weights are rounded to bf16 (what the CUDA path stores) before the reference runs in fp32.  Saves the output subsampled
    by 4 in both map axes plus whole-map statistics, and the attention of the two block kinds in isolation (qkv -> softmax(qk + rel) v,
    before proj) for the kernel-level test.  Writes tests/golden/sam_encoder_1024.pt:  python -m oracle.make_golden sam_encoder_1024"""
    from model.segment_anything.modeling.image_encoder import ImageEncoderViT, add_decomposed_rel_pos

    cfg = sam_encoder_full_geometry_cfg()
    ref = ImageEncoderViT(img_size=cfg["img_size"], patch_size=cfg["patch"], embed_dim=cfg["embed"], depth=cfg["depth"], num_heads=cfg["heads"],
                          out_chans=cfg["out_chans"], use_rel_pos=True, window_size=cfg["window_size"],
                          global_attn_indexes=cfg["global_attn_indexes"], norm_layer=lambda d: nn.LayerNorm(d, eps=1e-6)).eval()
    spec = specs.sam_image_encoder_spec(cfg["img_size"], cfg["patch"], cfg["embed"], cfg["depth"], cfg["heads"], 4.0, cfg["out_chans"],
                                        cfg["window_size"], cfg["global_attn_indexes"])
    sd = {k: (v.to(torch.bfloat16).float() if v.dim() >= 2 else v) for k, v in specs.make_state_dict(spec, seed=33).items()}
    ref.load_state_dict(sd, strict=True)
    x = rnd((1, 3, cfg["img_size"], cfg["img_size"]), 613).to(torch.bfloat16).float()
    out = ref(x)
    # attention of each block kind in isolation, on seeded bf16-representable q / k / v (the reference's own arithmetic, Attention.forward
    # lines 231-247 without the qkv / proj linears)
    att = {}
    for tag, n, side, blk in (("window", 3, 14, 0), ("global", 1, 64, 1)):
        a = ref.blocks[blk].attn
        heads, d = 2, 80
        q, k, v = (rnd((n * heads, side * side, d), 700 + i + 10 * blk).to(torch.bfloat16).float() for i in range(3))
        s = (q * a.scale) @ k.transpose(-2, -1)
        s = add_decomposed_rel_pos(s, q, a.rel_pos_h, a.rel_pos_w, (side, side), (side, side))
        o = s.softmax(dim=-1) @ v
        att[tag] = {"n": n, "heads": heads, "seeds": [700 + i + 10 * blk for i in range(3)], "out_sub": o[:, ::7].clone(), "out_absmax": o.abs().max(),
                    "block": blk}
    save("sam_encoder_1024", {"cfg": cfg, "seed": 33, "pixels_seed": 613, "out_sub": out[:, :, ::4, ::4].clone(), "out_mean": out.mean(), "out_std": out.std(),
                              "out_absmax": out.abs().max(), "attention": att, "ln_eps": 1e-6})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "sam_encoder_1024":
        make_sam_encoder_full_geometry()
    elif len(sys.argv) > 1 and sys.argv[1] == "two_level":
        make_decoder_two_level()
    elif len(sys.argv) > 1 and sys.argv[1] == "match":
        make_match_cost()
    elif len(sys.argv) > 1 and sys.argv[1] == "sam_encoder":
        make_sam_encoder()
    else:
        main()
        make_decoder_two_level()
        make_match_cost()
        make_sam_encoder()
        make_sam_encoder_full_geometry()
