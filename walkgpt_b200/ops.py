"""Tensor-level wrappers over the C ABI kernels (device tensors in, device tensors out).

These are plumbing: they validate shapes/dtypes, allocate outputs with torch, and pass raw device
pointers plus the current CUDA stream to ``libwalkgpt_b200.so``.  No arithmetic happens in Python.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import ACT_GELU_ERF, ACT_NONE, ACT_QUICK_GELU, ACT_RELU, OUT_BF16, OUT_BF16_LN, OUT_F32  # noqa: F401


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.WalkGPTB200Error("walkgpt_b200 ops need CUDA tensors (there is no CPU fallback)")


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE,
         out_mode: int = OUT_BF16, out: Optional[torch.Tensor] = None, resid: Optional[torch.Tensor] = None,
         bias_period: int = 1, ln_gamma: Optional[torch.Tensor] = None, ln_beta: Optional[torch.Tensor] = None,
         ln_eps: float = 1e-5) -> torch.Tensor:
    """out[M,N] = epilogue(a[M,K] @ w[N,K]^T).  a, w bf16 (row stride may exceed K); see include/walkgpt_b200.h."""
    _need_cuda(a, w, bias, out, resid)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.dim() == 2 and w.dim() == 2
    assert a.stride(1) == 1 and w.stride(1) == 1 and a.shape[1] == w.shape[1]
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.float32 if out_mode == OUT_F32 else torch.bfloat16)
    assert out.shape == (M, N) and out.stride(1) == 1
    assert out.dtype == (torch.float32 if out_mode == OUT_F32 else torch.bfloat16)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
        assert bias.numel() == (N if bias_period <= 1 else bias_period * N)
    if resid is not None:
        assert resid.shape == (M, N) and resid.stride(0) == out.stride(0) and resid.stride(1) == 1
        assert resid.dtype == (torch.float32 if out_mode == OUT_F32 else torch.bfloat16)
    args = _lib.GemmArgs()
    args.A, args.lda, args.W, args.ldw = a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0)
    args.M, args.N, args.K = M, N, K
    args.bias, args.bias_period, args.act, args.out_mode = _ptr(bias), bias_period, act, out_mode
    args.out, args.ldo, args.resid = out.data_ptr(), out.stride(0), _ptr(resid)
    if out_mode == OUT_BF16_LN:
        assert ln_gamma is not None and ln_beta is not None and ln_gamma.dtype == torch.float32
    args.ln_gamma, args.ln_beta, args.ln_eps = _ptr(ln_gamma), _ptr(ln_beta), ln_eps
    if M == 0:
        return out
    _lib.check(_lib.lib().wg_gemm(C.byref(args), _stream()), "wg_gemm")
    return out


def layernorm(x: torch.Tensor, gamma: Optional[torch.Tensor], beta: Optional[torch.Tensor], eps: float = 1e-5,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bf16 LayerNorm over the last dim of a 2-D fp32/bf16 tensor."""
    _need_cuda(x, gamma, beta, out)
    assert x.dim() == 2 and x.stride(1) == 1 and x.dtype in (torch.float32, torch.bfloat16)
    rows, D = x.shape
    if out is None:
        out = torch.empty(rows, D, device=x.device, dtype=torch.bfloat16)
    assert out.dtype == torch.bfloat16 and out.shape == (rows, D) and out.stride(1) == 1
    _lib.check(_lib.lib().wg_layernorm(x.data_ptr(), int(x.dtype == torch.bfloat16), x.stride(0), _ptr(gamma), _ptr(beta), eps,
                                       out.data_ptr(), out.stride(0), rows, D, _stream()), "wg_layernorm")
    return out


def sam_attention(qkv: torch.Tensor, rel_table: torch.Tensor, n_images: int, mode: int, heads: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """SAM ViT attention with decomposed relative position (image_encoder.py:222-247, 325-361), head_dim 80.  qkv bf16 [rows, 3*heads*80] in
    the kernel's column order (modules.sam_qkv_row_order): mode 0 = window-major rows [n_images * 25 * 196], mode 1 = image-major
    [n_images * 4096]; rel_table bf16 (modules.sam_rel_table).  Returns bf16 [n_images * 4096, heads * 80], image-major."""
    _need_cuda(qkv, rel_table, out)
    rows = n_images * (25 * 196 if mode == 0 else 4096)
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and tuple(qkv.shape) == (rows, 3 * heads * 80)
    assert rel_table.dtype == torch.bfloat16 and rel_table.is_contiguous() and tuple(rel_table.shape) == (64 if mode == 0 else 256, 80)
    if out is None:
        out = torch.empty(n_images * 4096, heads * 80, device=qkv.device, dtype=torch.bfloat16)
    assert out.is_contiguous() and out.dtype == torch.bfloat16
    _lib.check(_lib.lib().wg_sam_attention(qkv.data_ptr(), rel_table.data_ptr(), out.data_ptr(), n_images, mode, heads, _stream()), "wg_sam_attention")
    return out


def attention_d64(qkv: torch.Tensor, heads: int, scale: float, key_valid: Optional[torch.Tensor] = None,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """qkv bf16 [B,T,3*heads*64] -> bf16 [B,T,heads*64]; key_valid uint8 [B,T] (0 = padded key)."""
    _need_cuda(qkv, key_valid, out)
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and qkv.dim() == 3 and qkv.shape[2] == 3 * heads * 64
    B, T, _ = qkv.shape
    if out is None:
        out = torch.empty(B, T, heads * 64, device=qkv.device, dtype=torch.bfloat16)
    assert out.is_contiguous() and out.dtype == torch.bfloat16
    if key_valid is not None:
        assert key_valid.dtype == torch.uint8 and key_valid.shape == (B, T) and key_valid.is_contiguous()
    _lib.check(_lib.lib().wg_attention_d64(qkv.data_ptr(), out.data_ptr(), _ptr(key_valid), B, T, heads, scale, _stream()),
               "wg_attention_d64")
    return out


# ------------------------------------------------------------------------------------------------ SURVEY 8(f) "next" rows
def resample_tokens(tokens: torch.Tensor, t: int = 16) -> torch.Tensor:
    """Visual-token resample handed to the LLM (llava_arch.py:252-259): [n, p*p, C] -> [n, t*t, C] (bf16 or fp32)."""
    _need_cuda(tokens)
    assert tokens.dim() == 3 and tokens.is_contiguous() and tokens.dtype in (torch.bfloat16, torch.float32)
    n, l, c = tokens.shape
    p = int(l ** 0.5)
    if p * p != l:
        raise AssertionError(f"Token count {l} is not square.")  # llava_arch.py:254
    out = torch.empty(n, t * t, c, device=tokens.device, dtype=tokens.dtype)
    _lib.check(_lib.lib().wg_resample_tokens(tokens.data_ptr(), int(tokens.dtype == torch.bfloat16), n, p, c, t, out.data_ptr(), _stream()),
               "wg_resample_tokens")
    return out


def seg_gather(hidden: torch.Tensor, input_ids: torch.Tensor, seg_token_idx, offset=None, shift: int = 255, max_out: Optional[int] = None):
    """[SEG]-row extraction (model/walkgpt.py:287-306, 406-420).  hidden [rows, Lin + shift, H]; input_ids int64 [rows, Lin].
    Returns (rows_out [max_out, H], counts int32 [rows], row_offsets int32 [rows+1], img_offsets int32 [len(offset)] or None).
    Without max_out the number of selected rows is read back from the device (one synchronisation, like the reference's
    boolean-mask indexing) and rows_out is trimmed to it."""
    _need_cuda(hidden, input_ids)
    assert hidden.dim() == 3 and hidden.is_contiguous() and hidden.dtype in (torch.bfloat16, torch.float32)
    assert input_ids.dim() == 2 and input_ids.dtype == torch.int64 and input_ids.is_contiguous()
    rows, L, H = hidden.shape
    Lin = input_ids.shape[1]
    ids = list(seg_token_idx) if isinstance(seg_token_idx, (list, tuple)) else [int(seg_token_idx)]
    seg_arr = (C.c_int64 * len(ids))(*ids)
    dev = hidden.device
    counts = torch.empty(rows, dtype=torch.int32, device=dev)
    row_off = torch.empty(rows + 1, dtype=torch.int32, device=dev)
    img_rows = img_off = None
    if offset is not None:
        img_rows = torch.as_tensor(list(offset), dtype=torch.int32).to(dev)
        img_off = torch.empty(img_rows.numel(), dtype=torch.int32, device=dev)
    cap = rows * (Lin - 1) if max_out is None else max_out
    out = torch.empty(max(cap, 1), H, device=dev, dtype=hidden.dtype)
    _lib.check(_lib.lib().wg_seg_gather(input_ids.data_ptr(), rows, Lin, hidden.data_ptr(), int(hidden.dtype == torch.bfloat16), L, H,
                                        C.cast(seg_arr, C.c_void_p), len(ids), shift, _ptr(img_rows), 0 if img_rows is None else img_rows.numel(),
                                        out.data_ptr(), cap, counts.data_ptr(), row_off.data_ptr(), _ptr(img_off), _stream()), "wg_seg_gather")
    if max_out is None:
        out = out[: int(row_off[-1].item())]
    return out, counts, row_off, img_off


def intersection_and_union(output: torch.Tensor, target: torch.Tensor, K: int = 2, ignore_index: int = 255) -> torch.Tensor:
    """intersectionAndUnionGPU (utils/utils.py:192-204) for a batch of mask pairs: uint8 [n, ...] -> fp32 [n, 3, K]
    (area_intersection, area_union, area_target per class)."""
    _need_cuda(output, target)
    assert output.shape == target.shape and output.dtype == torch.uint8 and target.dtype == torch.uint8
    assert output.is_contiguous() and target.is_contiguous()
    n = output.shape[0]
    pixels = output[0].numel() if n else 0
    out = torch.empty(n, 3, K, device=output.device, dtype=torch.float32)
    if n == 0:
        return out
    need = _lib.lib().wg_intersection_and_union_workspace_bytes(n, K)
    ws = torch.empty(need, dtype=torch.uint8, device=output.device)
    _lib.check(_lib.lib().wg_intersection_and_union(output.data_ptr(), target.data_ptr(), n, pixels, K, ignore_index, out.data_ptr(), ws.data_ptr(), need,
                                                    _stream()), "wg_intersection_and_union")
    return out


def match_cost(out_mask: torch.Tensor, tgt_mask: torch.Tensor, point_coords: torch.Tensor) -> torch.Tensor:
    """Cost matrix of match_pred (utils/matcher.py:93-128) before the host-side linear_sum_assignment:
    out_mask fp32 [n_pred, H, W] logits, tgt_mask fp32 or uint8 [n_tgt, H, W], point_coords fp32 [P, 2] (or [1, P, 2]) in
    [0, 1]^2 -> fp32 [n_pred, n_tgt] = batch_sigmoid_ce_loss + batch_dice_loss on the point-sampled masks."""
    _need_cuda(out_mask, tgt_mask, point_coords)
    assert out_mask.dim() == 3 and tgt_mask.dim() == 3 and out_mask.shape[1:] == tgt_mask.shape[1:]
    pts = point_coords.reshape(-1, 2).float().contiguous()
    out_mask = out_mask.float().contiguous()
    if tgt_mask.dtype != torch.uint8:
        tgt_mask = tgt_mask.float()
    tgt_mask = tgt_mask.contiguous()
    n_pred, H, W = out_mask.shape
    n_tgt, P = tgt_mask.shape[0], pts.shape[0]
    cost = torch.empty(n_pred, n_tgt, device=out_mask.device, dtype=torch.float32)
    if n_pred == 0 or n_tgt == 0:
        return cost
    need = _lib.lib().wg_match_cost_workspace_bytes(n_pred, n_tgt, P)
    ws = torch.empty(need, dtype=torch.uint8, device=out_mask.device)
    _lib.check(_lib.lib().wg_match_cost(out_mask.data_ptr(), tgt_mask.data_ptr(), int(tgt_mask.dtype == torch.uint8), pts.data_ptr(), n_pred, n_tgt, H, W, P,
                                        cost.data_ptr(), ws.data_ptr(), need, _stream()), "wg_match_cost")
    return cost
