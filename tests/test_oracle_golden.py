"""CPU: the oracle restatement (oracle/path_a.py) against fixtures produced by EXECUTING the reference modules
(oracle/make_golden.py).  fp32 vs fp32 on the same weights/inputs: differences are summation-order only."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import path_a
from walkgpt_b200 import specs
from walkgpt_b200.modules import clip_param_spec

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def close(a, b, tol=2e-4):
    a, b = a.float(), b.float()
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs().max().item()
    ref = b.abs().max().item() + 1e-9
    assert err <= tol * max(ref, 1.0), f"max abs err {err:.3e} (ref absmax {ref:.3e})"


@pytest.mark.parametrize("in_dim", [256, 4096])
def test_ctp(in_dim):
    g = load(f"ctp_{in_dim}")
    sd = specs.make_state_dict(specs.ctp_spec(in_dim, 256), seed=g["seed"])
    close(path_a.ctp_forward(sd, g["x3"]), g["y3"], 1e-5)
    y2 = path_a.ctp_forward(sd, g["x2"])
    assert y2.shape == g["y2"].shape == (1, 3, 256)  # the reference's 2-D -> 3-D quirk
    close(y2, g["y2"], 1e-5)
    # unit norm times exp(log_temp)
    n = g["y3"].norm(dim=-1)
    assert torch.allclose(n, sd["log_temp"].exp().expand_as(n), atol=1e-5)


@pytest.mark.parametrize("tag", ["b_small", "a_32"])
def test_msqp(tag):
    g = load(f"msqp_{tag}")
    sd = specs.make_state_dict(specs.msqp_spec(g["sam_dim"], g["llama_dim"]), seed=g["seed"])
    x = rnd(g["x_shape"], g["x_seed"])
    y = path_a.msqp_forward(sd, x, target_square_side=6)
    assert y.shape == g["y"].shape and y.shape[1] == 36
    close(y, g["y"], 1e-4)


def test_msqp_rejects_non_square():
    sd = specs.make_state_dict(specs.msqp_spec(256, 64), seed=1)
    with pytest.raises(ValueError, match="perfect square"):
        path_a.msqp_forward(sd, torch.zeros(1, 60, 256))


def test_projector_and_neck():
    g = load("proj_neck_small")
    sdp = specs.make_state_dict(specs.out_mm_projector_spec(g["mm"], g["hidden"]), seed=g["seed"])
    sdn = specs.make_state_dict(specs.neck_spec(g["hidden"], 256), seed=g["seed"])
    p = path_a.out_mm_projector_mlp(sdp, g["x"])
    close(p, g["proj"], 1e-5)
    e = path_a.image_feature_neck(sdn, p, 8)
    close(e, g["emb"], 1e-4)


@pytest.mark.parametrize("grid", [8, 32])
def test_decoder_multiscale(grid):
    g = load(f"decoder_ms_g{grid}")
    sdp = specs.make_state_dict(specs.prompt_encoder_spec(256, 16), seed=g["seed_prompt"])
    sdd = specs.make_state_dict(specs.mask_decoder_multiscale_spec(), seed=g["seed_dec"])
    emb = rnd(g["emb_shape"], g["emb_seed"])
    pe = path_a.dense_pe(sdp["pe_layer.positional_encoding_gaussian_matrix"], grid, grid)[None]
    close(pe if grid == 8 else pe[:, ::8], g["dense_pe"], 1e-5)
    sparse, dense = path_a.prompt_encoder(sdp, g["txt"], (grid, grid))
    m1, i1 = path_a.mask_decoder_multiscale(sdd, emb, pe, sparse, dense, multimask_output=False)
    m4, i4 = path_a.mask_decoder_multiscale(sdd, emb, pe, sparse, dense, multimask_output=True)
    assert m1.shape == g["masks1"].shape and m4.shape == g["masks4"].shape
    close(m1, g["masks1"], 2e-4)
    close(i1, g["iou1"], 2e-4)
    close(m4, g["masks4"], 2e-4)
    close(i4, g["iou4"], 2e-4)


def test_decoder_multiscale_two_levels():
    """level 1 of MaskDecoderMultiScale (upsample_2x, previous-mask gating, pe1, transformer[1]) against the
    fixture produced by the reference module with image_feature_scale_num=2."""
    g = load("decoder_ms2_g8")
    sdp = specs.make_state_dict(specs.prompt_encoder_spec(256, 16), seed=g["seed_prompt"])
    sdd = specs.make_state_dict(specs.mask_decoder_multiscale_spec(scale_num=2), seed=g["seed_dec"])
    pe = path_a.dense_pe(sdp["pe_layer.positional_encoding_gaussian_matrix"], 8, 8)[None]
    sparse, dense = path_a.prompt_encoder(sdp, g["txt"], (8, 8))
    m0, i0 = path_a.mask_decoder_multiscale(sdd, g["emb"], pe, sparse, dense, multimask_output=True, level_num=0)
    close(m0, g["masks_l0"], 2e-4)
    m1, i1 = path_a.mask_decoder_multiscale(sdd, g["emb"], pe, sparse, dense, multimask_output=True, level_num=1, previous_masks=g["masks_l0"])
    assert m1.shape == (3, 4, 32, 32)
    close(m1, g["masks_l1"], 2e-4)
    close(i1, g["iou_l1"], 2e-4)
    m1s, _ = path_a.mask_decoder_multiscale(sdd, g["emb"], pe, sparse, dense, multimask_output=False, level_num=1, previous_masks=g["masks_l0"])
    close(m1s, g["masks_l1_single"], 2e-4)


def test_decoder_sam():
    g = load("decoder_sam_g8")
    sdp = specs.make_state_dict(specs.prompt_encoder_spec(256, 16), seed=g["seed_prompt"])
    sdd = specs.make_state_dict(specs.mask_decoder_sam_spec(), seed=g["seed_dec"])
    pe = path_a.dense_pe(sdp["pe_layer.positional_encoding_gaussian_matrix"], 8, 8)[None]
    sparse, dense = path_a.prompt_encoder(sdp, g["txt"], (8, 8))
    m1, i1 = path_a.mask_decoder_sam(sdd, g["emb"], pe, sparse, dense, multimask_output=False)
    m3, i3 = path_a.mask_decoder_sam(sdd, g["emb"], pe, sparse, dense, multimask_output=True)
    close(m1, g["masks1"], 2e-4)
    close(i1, g["iou1"], 2e-4)
    assert m3.shape[1] == 3
    close(m3, g["masks3"], 2e-4)
    close(i3, g["iou3"], 2e-4)


def test_postprocess_and_score():
    g = load("postprocess")
    for name, c in g["cases"].items():
        if name == "sam_1024":
            low = rnd((2, 1, 256, 256), g["low_b_seed"], 3.0)
            m = path_a.postprocess_masks(low, c["input_size"], c["original_size"], target_size=1024, cast_back=False)
        else:
            m = path_a.postprocess_masks(g["low"], c["input_size"], c["original_size"])
        assert tuple(m.shape[-2:]) == tuple(c["original_size"])
        close(m[:, 0, ::7, ::5], c["logits_sub"], 1e-6)
        assert torch.allclose(m.double().sum(dim=(1, 2, 3)), c["sum"], rtol=1e-6)
        if "score" in c:
            close(path_a.mask_score(m[:, 0]), c["score"], 1e-6)
            assert torch.equal((m[:, 0] > 0).flatten(1).sum(1), c["pos_count"])


def test_clip_tower():
    g = load("clip_3layer")
    sd = specs.make_state_dict(clip_param_spec(layers=g["layers"]), seed=g["seed"])
    px = rnd((2, 3, 448, 448), g["px_seed"])
    hs = path_a.clip_hidden_states(sd, px, None)
    assert len(hs) == g["layers"] + 1
    for i, t in enumerate(hs):
        close(t[:, ::41, ::13], g["hs_sub"][i], 2e-4)
        assert abs(t.double().mean().item() - g["hs_mean"][i].item()) < 1e-4
        assert abs(t.double().std().item() - g["hs_std"][i].item()) < 1e-3
    kv = path_a.clip_key_valid_from_sizes(g["sizes"])
    assert torch.equal(kv, g["key_valid"])
    hs_m = path_a.clip_hidden_states(sd, px, kv)
    for i, t in enumerate(hs_m):
        close(t[:, ::41, ::13], g["hs_masked_sub"][i], 2e-4)
    # feature_select semantics (clip_encoder.py:61-69): hs[select][:,1:], [hs[-11][:,1:]]
    last, mid = path_a.clip_tower(sd, px, None, select_layer=-2, total_layers=g["layers"])
    assert torch.equal(last, hs[-2][:, 1:]) and torch.equal(mid[0], hs[(-11) % (g["layers"] + 1)][:, 1:])


def test_iou_histogram():
    g = load("iou_hist")
    ai, au, at = path_a.intersection_and_union(g["output"], g["target"], 2, 255)
    assert torch.equal(ai, g["inter"]) and torch.equal(au, g["union"]) and torch.equal(at, g["target_area"])


def test_match_cost_against_reference_golden():
    """oracle.match_cost (utils/matcher.py:93-124) reproduces the cost matrix built from the reference's own
    point_sample / batch_sigmoid_ce_loss / batch_dice_loss, and scipy's assignment on it is match_pred's output."""
    from scipy.optimize import linear_sum_assignment

    g = load("match_cost")
    for case in g["cases"]:
        torch.manual_seed(case["seed"])
        pts = torch.rand(1, g["num_points"], 2)
        C = path_a.match_cost(case["out_mask"].float(), case["tgt_mask"].float(), pts)
        assert C.shape == case["cost"].shape and (C - case["cost"]).abs().max().item() < 1e-6
        r, c = linear_sum_assignment(C)
        assert list(r) == case["pred_idx"].tolist() and list(c) == case["tgt_idx"].tolist()


def test_sam_image_encoder_oracle_against_reference_golden():
    """SURVEY 8(f) row 1 groundwork (no CUDA path yet): oracle.path_b.image_encoder_vit reproduces the reference
    ImageEncoderViT -- head_dim 80, padded 5 x 5 windows on a 12 x 12 map, one global block, decomposed relative position, neck --
    and the parameter table loads into the reference module strictly (make_golden.make_sam_encoder)."""
    from oracle import path_b

    g = load("sam_encoder_small")
    cfg = g["cfg"]
    spec = specs.sam_image_encoder_spec(cfg["img_size"], cfg["patch"], cfg["embed"], cfg["depth"], cfg["heads"], 4.0, cfg["out_chans"],
                                        cfg["window_size"], cfg["global_attn_indexes"])
    sd = specs.make_state_dict(spec, seed=g["seed"])
    x = rnd((2, 3, cfg["img_size"], cfg["img_size"]), g["pixels_seed"])
    out = path_b.image_encoder_vit(sd, x, cfg["heads"], cfg["window_size"], cfg["global_attn_indexes"], ln_eps=g["ln_eps"])
    assert out.shape == g["out"].shape == (2, 256, 12, 12)
    close(out, g["out"], 2e-5)
    # the full-size table has the released checkpoint's shapes (SAM ViT-H, build_sam.py:15-22)
    full = specs.sam_image_encoder_spec()
    assert full["blocks.7.attn.rel_pos_h"][0] == (127, 80) and full["blocks.0.attn.rel_pos_h"][0] == (27, 80)
    assert full["pos_embed"][0] == (1, 64, 64, 1280) and full["neck.2.weight"][0] == (256, 256, 3, 3)


def test_depth_extension_is_self_consistent():
    """No reference exists for depth (parity unpinned): only check the in-repo definition's invariants."""
    sd = specs.make_state_dict(specs.depth_head_spec(), seed=3, prefix="depth_head.")
    low = rnd((4, 64, 64), 1, 2.0)
    up = rnd((4, 32, 64, 64), 2)
    d = path_a.depth_head(sd, low, up)
    assert d.shape == (4,) and d.min().item() == 0.0 and abs(d.max().item() - 1.0) < 1e-4


def test_seg_row_extraction_restatement():
    """oracle.gather_seg_rows (model/walkgpt.py:287-306, 406-420) against an independent loop over the token ids."""
    g = torch.Generator().manual_seed(5)
    rows, Lin, H, seg, shift = 5, 40, 16, 32003, 255
    ids = torch.randint(0, 32000, (rows, Lin), generator=g)
    ids[0, [3, 17, 39]] = seg      # the last position counts (it is input_ids[:, 1:][-1])
    ids[1, 0] = seg                # position 0 never counts: the mask is built from input_ids[:, 1:]
    ids[3, [1, 2]] = seg
    hidden = torch.randn(rows, Lin + shift, H, generator=g)
    pred, counts, off = path_a.gather_seg_rows(hidden, ids, seg, [0, 2, 5], shift)
    want = [hidden[r, i - 1 + shift] for r in range(rows) for i in range(1, Lin) if ids[r, i] == seg]
    assert counts.tolist() == [3, 0, 0, 2, 0] and off.tolist() == [0, 3, 5]
    assert torch.equal(pred, torch.stack(want))
    pred2, counts2, _ = path_a.gather_seg_rows(hidden, ids, [seg, int(ids[2, 5])], [0, 5], shift)
    assert counts2[2] >= 1 and pred2.shape[0] == counts2.sum()


def test_visual_token_resample_restatement():
    """oracle.resample_visual_tokens (llava_arch.py:252-259): 36 -> 256 tokens, fp32 interpolation, cast back."""
    x = rnd((2, 36, 64), 6).bfloat16()
    y = path_a.resample_visual_tokens(x, 16)
    assert y.shape == (2, 256, 64) and y.dtype == torch.bfloat16
    grid = F.interpolate(x.float().transpose(1, 2).reshape(2, 64, 6, 6), (16, 16), mode="bilinear", align_corners=False)
    assert torch.equal(y, grid.reshape(2, 64, 256).transpose(1, 2).bfloat16())
    with pytest.raises(AssertionError, match="not square"):
        path_a.resample_visual_tokens(torch.zeros(1, 35, 8))


def test_sam_image_encoder_oracle_at_full_geometry_against_reference_golden():
    """SURVEY 8(f) row 1 at SAM's real geometry (1024-pixel input, 64 x 64 map, head_dim 80, 14 x 14 windows padded to 70, 27- and 127-row
    relative-position tables): the oracle restatement against a fixture produced by the reference ImageEncoderViT, including the attention
    of both block kinds in isolation (the fixture the CUDA attention kernel is tested against)."""
    from oracle import path_b

    g = load("sam_encoder_1024")
    cfg = g["cfg"]
    spec = specs.sam_image_encoder_spec(cfg["img_size"], cfg["patch"], cfg["embed"], cfg["depth"], cfg["heads"], 4.0, cfg["out_chans"],
                                        cfg["window_size"], cfg["global_attn_indexes"])
    sd = {k: (v.to(torch.bfloat16).float() if v.dim() >= 2 else v) for k, v in specs.make_state_dict(spec, seed=g["seed"]).items()}
    x = rnd((1, 3, cfg["img_size"], cfg["img_size"]), g["pixels_seed"]).to(torch.bfloat16).float()
    out = path_b.image_encoder_vit(sd, x, cfg["heads"], cfg["window_size"], cfg["global_attn_indexes"], ln_eps=g["ln_eps"])
    assert out.shape == (1, 256, 64, 64)
    assert (out[:, :, ::4, ::4] - g["out_sub"]).abs().max().item() < 2e-4 * g["out_absmax"].item()
    assert abs(out.std().item() - g["out_std"].item()) < 1e-4
    for tag, side in (("window", 14), ("global", 64)):
        a = g["attention"][tag]
        q, k, v = (rnd((a["n"] * a["heads"], side * side, 80), s).to(torch.bfloat16).float() for s in a["seeds"])
        p = f"blocks.{a['block']}.attn."
        o = path_b.attention_core_rel_pos(q, k, v, sd[p + "rel_pos_h"], sd[p + "rel_pos_w"], side)
        assert (o[:, ::7] - a["out_sub"]).abs().max().item() < 2e-5 * a["out_absmax"].item()
