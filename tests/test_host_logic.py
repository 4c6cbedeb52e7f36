"""CPU: host-side logic of the drop-in modules -- parameter names/shapes, weight re-layout and constant folding,
reference error behaviour, sharding helpers."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

from walkgpt_b200 import parallel, specs
from walkgpt_b200 import modules as M

REF = "/root/reference"


def test_state_dict_keys_follow_the_specs():
    m = M.GroundingPath(hidden_size=256, clip_layers=1)
    sd = m.state_dict()
    assert set(k for k in sd if k.startswith("msqp.")) == {"msqp." + k for k in specs.msqp_spec(1024, 256)}
    assert set(k for k in sd if k.startswith("mask_decoder.")) == {"mask_decoder." + k for k in specs.mask_decoder_multiscale_spec()}
    assert set(k for k in sd if k.startswith("text_hidden_fcs.0.")) == {"text_hidden_fcs.0." + k for k in specs.ctp_spec(256, 256)}
    assert "prompt_encoder.pe_layer.positional_encoding_gaussian_matrix" in sd
    # buffers stay buffers (the random Gaussian PE matrix is part of the checkpoint, prompt_encoder.py:198-201)
    assert "pe_layer.positional_encoding_gaussian_matrix" in dict(m.prompt_encoder.named_buffers())


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (build container only)")
def test_reference_modules_load_our_state_dicts_strictly():
    import sys
    sys.path.insert(0, REF)
    from utils.utils_walkgpt import CalibratedTextProjector, MultiScaleQFormerProjector
    from model.segment_anything.modeling import MaskDecoder, MaskDecoderMultiScale, PromptEncoder, TwoWayTransformer

    MultiScaleQFormerProjector(256, 64).load_state_dict(M.MultiScaleQFormerProjector(256, 64).state_dict(), strict=True)
    CalibratedTextProjector(512, 256).load_state_dict(M.CalibratedTextProjector(512, 256).state_dict(), strict=True)
    PromptEncoder(256, (32, 32), (448, 448), 16).load_state_dict(M.PromptEncoder(256, (32, 32), (448, 448), 16).state_dict(), strict=True)
    MaskDecoderMultiScale(num_multimask_outputs=3, transformer=TwoWayTransformer(depth=2, embedding_dim=256, mlp_dim=2048, num_heads=8),
                          transformer_dim=256, image_feature_scale_num=1).load_state_dict(M.MaskDecoderMultiScale().state_dict(), strict=True)
    MaskDecoder(num_multimask_outputs=3, transformer=TwoWayTransformer(depth=2, embedding_dim=256, mlp_dim=2048, num_heads=8),
                transformer_dim=256).load_state_dict(M.MaskDecoder().state_dict(), strict=True)  # Path B (released SAM-1024 wiring)


def test_synthetic_init_is_deterministic_and_order_independent():
    a = specs.make_state_dict(specs.ctp_spec(256, 256), seed=5)
    b = specs.make_state_dict(dict(reversed(list(specs.ctp_spec(256, 256).items()))), seed=5)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert not torch.equal(a["net.1.weight"], specs.make_state_dict(specs.ctp_spec(256, 256), seed=6)["net.1.weight"])


def test_kv_norm_folding_is_exact_algebra():
    d = 64
    g = torch.Generator().manual_seed(0)
    W, b = torch.randn(3 * d, d, generator=g), torch.randn(3 * d, generator=g)
    gamma, beta = torch.randn(d, generator=g), torch.randn(d, generator=g)
    x = torch.randn(10, d, generator=g)
    xhat = F.layer_norm(x, (d,))
    Wf, bf = M.fold_kv_norm(W, b, gamma, beta, d)
    ref = F.linear(F.layer_norm(x, (d,), gamma, beta), W[d:], b[d:])
    assert torch.allclose(F.linear(xhat, Wf, bf), ref, atol=1e-4)


def test_conv_weight_relayouts_match_torch_convolutions():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 16, 5, 5, generator=g)
    w = torch.randn(8, 16, 3, 3, generator=g)
    ref = F.conv2d(x, w, padding=1)
    xt = F.pad(x.permute(0, 2, 3, 1), (0, 0, 1, 1, 1, 1))  # channels-last, zero padded
    cols = torch.stack([xt[:, ky:ky + 5, kx:kx + 5] for ky in range(3) for kx in range(3)], dim=3).reshape(2, 5, 5, -1)
    got = (cols @ M.pack_conv3x3(w).t()).permute(0, 3, 1, 2)
    assert torch.allclose(got, ref, atol=1e-4)
    wt, bt = torch.randn(16, 4, 2, 2, generator=g), torch.randn(4, generator=g)
    ref = F.conv_transpose2d(x, wt, bt, stride=2)
    wg, bg = M.pack_conv_transpose2x2(wt, bt)
    u = x.permute(0, 2, 3, 1) @ wg.t() + bg  # [B, y, x, (dy, dx, co)]
    got = u.reshape(2, 5, 5, 2, 2, 4).permute(0, 5, 1, 3, 2, 4).reshape(2, 4, 10, 10)
    assert torch.allclose(got, ref, atol=1e-4)


def test_reference_error_behaviour_is_preserved():
    msqp = M.MultiScaleQFormerProjector(256, 64, pad_to_square=True, target_square_side=5)
    with pytest.raises(AssertionError, match="target_square_side too small"):
        msqp.n_tokens()
    assert M.MultiScaleQFormerProjector(256, 64, target_square_side=6).n_tokens() == 36
    assert M.MultiScaleQFormerProjector(256, 64, pad_to_square=False).n_tokens() == 32
    assert M.MultiScaleQFormerProjector(256, 64).n_tokens() == 36  # ceil(sqrt(32))^2
    with pytest.raises(ValueError, match="Unexpected select feature"):
        args = M._ClipArgs()
        args.mm_vision_select_feature = "bogus"
        M.CLIPVisionTower(None, args, layers=1)
    pe = M.PromptEncoder(256, (8, 8), (112, 112), 16)
    with pytest.raises(NotImplementedError):
        pe((torch.zeros(1, 1, 2), torch.zeros(1, 1)), None, None, None)
    sparse, dense = pe(None, None, None, torch.zeros(3, 1, 256))
    assert sparse.shape == (3, 1, 256) and dense.shape == (3, 256, 8, 8)


def test_clip_hidden_state_selection():
    t = M.CLIPVisionTower(layers=24)
    assert t.hidden_state_indices() == (23, 14)  # hidden_states[-2], hidden_states[-11] of 25 states
    a = M._ClipArgs()
    a.mm_vision_select_layer = -1
    assert M.CLIPVisionTower(None, a, layers=24).hidden_state_indices() == (24, 14)
    assert t.num_patches == 1024 and t.hidden_size == 1024


def test_shard_helpers():
    assert [parallel.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [parallel.shard_range(2, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
    imgs = torch.arange(5)[:, None].float()
    seg = torch.arange(9)[:, None].float()
    offs = [0, 2, 2, 5, 8, 9]  # ragged, including an image with zero [SEG] tokens
    got = [parallel.shard_batch(imgs, seg, offs, r, 2) for r in range(2)]
    assert got[0][0].flatten().tolist() == [0, 1, 2] and got[0][1].flatten().tolist() == [0, 1, 2, 3, 4] and got[0][2] == [0, 2, 2, 5]
    assert got[1][0].flatten().tolist() == [3, 4] and got[1][1].flatten().tolist() == [5, 6, 7, 8] and got[1][2] == [0, 3, 4]


def test_torch_custom_ops_have_shape_inference():
    from torch._subclasses.fake_tensor import FakeTensorMode

    from walkgpt_b200 import torch_ops

    h_ctp = torch_ops.register(M.CalibratedTextProjector(256, 256))
    h_msqp = torch_ops.register(M.MultiScaleQFormerProjector(256, 64, target_square_side=6))
    h_clip = torch_ops.register(M.CLIPVisionTower(layers=1))
    with FakeTensorMode():
        assert torch.ops.walkgpt_b200.ctp_forward(torch.empty(2, 5, 256), h_ctp).shape == (2, 5, 256)
        assert torch.ops.walkgpt_b200.msqp_forward(torch.empty(3, 64, 256), h_msqp).shape == (3, 36, 64)
        last, mid = torch.ops.walkgpt_b200.clip_forward(torch.empty(2, 3, 448, 448), h_clip)
        assert last.shape == mid.shape == (2, 1024, 1024)
        lg, mk, sc = torch.ops.walkgpt_b200.postprocess_masks(torch.empty(4, 64, 64), 448, 448, 360, 640)
        assert lg.shape == (4, 360, 640) and mk.dtype == torch.uint8 and sc.shape == (4,)
        h_ms, h_sam = torch_ops.register(M.MaskDecoderMultiScale()), torch_ops.register(M.MaskDecoder())
        dec_args = (torch.empty(1, 256, 32, 32), torch.empty(1, 256, 32, 32), torch.empty(5, 1, 256), torch.empty(5, 256, 32, 32))
        m, i = torch.ops.walkgpt_b200.mask_decoder_forward(*dec_args, False, h_ms)
        assert m.shape == (5, 1, 64, 64) and i.shape == (5, 1)
        m, i = torch.ops.walkgpt_b200.mask_decoder_forward(*dec_args, True, h_sam)  # SAM decoder: masks 1..3 at 4x the grid
        assert m.shape == (5, 3, 128, 128) and i.shape == (5, 3)
        h_pn = torch_ops.register(M.ProjectorNeck(1024, 512))
        assert torch.ops.walkgpt_b200.proj_neck_forward(torch.empty(2, 1024, 1024), h_pn).shape == (2, 256, 32, 32)
    torch_ops.unregister(h_pn)
    with pytest.raises(KeyError):
        torch_ops._get(h_pn)


def test_prompt_index_validates_host_offsets_and_caches_by_value():
    pi = M._PromptIndex(capacity=2)
    offs_dev, img, P, max_S = pi.index([0, 2, 2, 5], 3, 5, "cpu")
    assert offs_dev.tolist() == [0, 2, 2, 5] and img.tolist() == [0, 0, 2, 2, 2] and (P, max_S) == (5, 3)
    assert pi.index((0, 2, 2, 5), 3, 5, "cpu")[1] is img  # same offsets: the cached device tensors
    for bad in ([0, 2, 5], [1, 2, 2, 5], [0, 3, 2, 5], [0, 2, 2, 4]):
        with pytest.raises(ValueError):
            pi.index(bad, 3, 5, "cpu")
    pi.index([0, 1, 2, 5], 3, 5, "cpu")
    pi.index([0, 0, 0, 5], 3, 5, "cpu")
    assert len(pi.cache) == 2  # bounded


def test_sam_encoder_host_side_packing():
    """ImageEncoderViT drop-in: reference parameter names / shapes, the qkv row permutation (head slices of 64 + 16 channels) and the
    relative-position tables the attention kernel's prologue MMA reads."""
    enc = M.ImageEncoderViT(embed_dim=640, depth=2, num_heads=8, global_attn_indexes=(1,))
    sd = enc.state_dict()
    assert sd["pos_embed"].shape == (1, 64, 64, 640) and sd["blocks.0.attn.rel_pos_h"].shape == (27, 80) and sd["blocks.1.attn.rel_pos_w"].shape == (127, 80)
    assert sd["blocks.0.attn.qkv.weight"].shape == (1920, 640) and sd["neck.2.weight"].shape == (256, 256, 3, 3)
    order = M.sam_qkv_row_order(8)
    assert sorted(order.tolist()) == list(range(1920))                       # a permutation
    # head 3 of K: main slice then rem slice
    k3 = 640 + 3 * 80
    assert order[8 * 64 + 3 * 64: 8 * 64 + 4 * 64].tolist() == list(range(k3, k3 + 64))
    assert order[3 * 8 * 64 + 8 * 16 + 3 * 16: 3 * 8 * 64 + 8 * 16 + 4 * 16].tolist() == list(range(k3 + 64, k3 + 80))
    # a linear layer with permuted rows gives permuted outputs: the attention kernel sees q_h = [main | rem] of the reference's q_h
    x = torch.randn(5, 640)
    W, b = sd["blocks.0.attn.qkv.weight"], sd["blocks.0.attn.qkv.bias"]
    ref = F.linear(x, W, b).view(5, 3, 8, 80)
    got = F.linear(x, W[order], b[order])
    main, rem = got[:, :3 * 8 * 64].view(5, 3, 8, 64), got[:, 3 * 8 * 64:].view(5, 3, 8, 16)
    assert torch.equal(torch.cat([main, rem], dim=-1), ref)
    t = M.sam_rel_table(sd["blocks.0.attn.rel_pos_h"], sd["blocks.0.attn.rel_pos_w"], 14)
    assert t.shape == (64, 80) and torch.equal(t[:27], sd["blocks.0.attn.rel_pos_h"]) and torch.equal(t[32:59], sd["blocks.0.attn.rel_pos_w"])
    assert t[27:32].abs().sum() == 0 and t[59:].abs().sum() == 0
    t = M.sam_rel_table(sd["blocks.1.attn.rel_pos_h"], sd["blocks.1.attn.rel_pos_w"], 64)
    assert t.shape == (256, 80) and torch.equal(t[128:255], sd["blocks.1.attn.rel_pos_w"])
    with pytest.raises(NotImplementedError):
        M.sam_rel_table(torch.zeros(13, 80), torch.zeros(13, 80), 14)         # tables that would need get_rel_pos's interpolation
    with pytest.raises(ValueError):
        M.ImageEncoderViT(img_size=512)
    with pytest.raises(ValueError):
        M.ImageEncoderViT(embed_dim=768, num_heads=12)                        # head_dim 64: ViT-B is not SAM ViT-H's geometry


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (build container only)")
def test_reference_image_encoder_loads_our_state_dict_strictly():
    import sys
    import torch.nn as nn
    sys.path.insert(0, REF)
    from model.segment_anything.modeling.image_encoder import ImageEncoderViT
    ref = ImageEncoderViT(img_size=1024, patch_size=16, embed_dim=640, depth=2, num_heads=8, out_chans=256, use_rel_pos=True, window_size=14,
                          global_attn_indexes=(1,), norm_layer=lambda d: nn.LayerNorm(d, eps=1e-6))
    ref.load_state_dict(M.ImageEncoderViT(embed_dim=640, depth=2, num_heads=8, global_attn_indexes=(1,)).state_dict(), strict=True)
