"""``torch.library`` custom ops over the C ABI (namespace ``walkgpt_b200::``), so the hot-path calls are opaque, shape-inferable
nodes for torch tooling (graph capture, export) instead of Python.  Weights are addressed by an integer handle into a
registry of packed modules (``register``), because custom ops only carry tensors and scalars.

    h = torch_ops.register(module)                      # module: a walkgpt_b200.modules.* instance
    y = torch.ops.walkgpt_b200.ctp_forward(x, h)        # == module(x) (without the reference's extra leading dim for 2-D x)
"""
from __future__ import annotations

import itertools
from typing import Dict, Tuple

import torch

from . import modules as M

# handle -> module.  Handles come from a counter (never reused for a different module, unlike id()); a registered module stays
# alive -- a captured graph may still name its handle -- until ``unregister`` releases it.
_REGISTRY: Dict[int, torch.nn.Module] = {}
_NEXT = itertools.count(1)


def register(module: torch.nn.Module) -> int:
    h = getattr(module, "_wg_op_handle", None)
    if h is not None and _REGISTRY.get(h) is module:
        return h
    h = next(_NEXT)
    _REGISTRY[h] = module
    module._wg_op_handle = h
    return h


def unregister(handle: int) -> None:
    """Release a handle (and the registry's reference to its module and packed device weights)."""
    _REGISTRY.pop(handle, None)


def _get(h: int):
    m = _REGISTRY.get(h)
    if m is None:
        raise KeyError(f"walkgpt_b200.torch_ops: handle {h} is not registered")
    return m


@torch.library.custom_op("walkgpt_b200::clip_forward", mutates_args=())
def clip_forward(images: torch.Tensor, handle: int) -> Tuple[torch.Tensor, torch.Tensor]:
    last, mid = _get(handle)(images)
    return last, mid[0]


@clip_forward.register_fake
def _(images, handle):
    m = _get(handle)
    shape = (images.shape[0], m.num_patches, m.hidden_size)
    return images.new_empty(shape), images.new_empty(shape)


@torch.library.custom_op("walkgpt_b200::msqp_forward", mutates_args=())
def msqp_forward(feats: torch.Tensor, handle: int) -> torch.Tensor:
    return _get(handle)(feats)


@msqp_forward.register_fake
def _(feats, handle):
    m = _get(handle)
    return feats.new_empty((feats.shape[0], m.n_tokens(), m.llama_dim))


@torch.library.custom_op("walkgpt_b200::ctp_forward", mutates_args=())
def ctp_forward(x: torch.Tensor, handle: int) -> torch.Tensor:
    m = _get(handle)
    xin = M._as_kernel_input(x)
    return m.run(xin.reshape(-1, m.in_dim), xin.dtype).reshape(tuple(x.shape[:-1]) + (m.out_dim,)).to(x.dtype)


@ctp_forward.register_fake
def _(x, handle):
    return x.new_empty(tuple(x.shape[:-1]) + (_get(handle).out_dim,))


@torch.library.custom_op("walkgpt_b200::postprocess_masks", mutates_args=())
def postprocess_masks(low_res: torch.Tensor, in_h: int, in_w: int, out_h: int, out_w: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    return M.postprocess_masks_fused(low_res.float().contiguous(), (in_h, in_w), (out_h, out_w))


@postprocess_masks.register_fake
def _(low_res, in_h, in_w, out_h, out_w):
    n = low_res.shape[0]
    return (low_res.new_empty((n, out_h, out_w), dtype=torch.float32), low_res.new_empty((n, out_h, out_w), dtype=torch.uint8),
            low_res.new_empty((n,), dtype=torch.float32))


@torch.library.custom_op("walkgpt_b200::proj_neck_forward", mutates_args=())
def proj_neck_forward(feats: torch.Tensor, handle: int) -> torch.Tensor:
    """out_mm_projector + image_feature_neck: feats [B, L, mm_hidden] -> image embedding [B, 256, g, g] (ProjectorNeck)."""
    return _get(handle)(feats)


@proj_neck_forward.register_fake
def _(feats, handle):
    m = _get(handle)
    g = int(round(feats.shape[1] ** 0.5))
    return feats.new_empty((feats.shape[0], m.out_chans, g, g))


@torch.library.custom_op("walkgpt_b200::mask_decoder_forward", mutates_args=())
def mask_decoder_forward(image_embeddings: torch.Tensor, image_pe: torch.Tensor, sparse_prompt_embeddings: torch.Tensor,
                         dense_prompt_embeddings: torch.Tensor, multimask_output: bool, handle: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """MaskDecoderMultiScale(level 0) / MaskDecoder .forward with the reference's argument order (mask_decoder_multi_scale.py:87,
    mask_decoder.py:75): ([1,256,h,w], [1,256,h,w], [S,1,256], [S,256,h,w]) -> (masks [S,n,u*h,u*w], iou [S,n])."""
    m = _get(handle)
    if isinstance(m, M.MaskDecoder):
        return m(image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings, multimask_output)
    return m(image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings, multimask_output, 0, None)


@mask_decoder_forward.register_fake
def _(image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings, multimask_output, handle):
    m = _get(handle)
    S = sparse_prompt_embeddings.shape[0]
    n = (m.num_mask_tokens - m._MULTIMASK_FIRST) if multimask_output else 1
    up = 2 ** m._UP_STAGES
    h, w = image_embeddings.shape[-2:]
    return image_embeddings.new_empty((S, n, up * h, up * w)), image_embeddings.new_empty((S, n))
