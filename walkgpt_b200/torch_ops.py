"""``torch.library`` custom ops over the C ABI (namespace ``walkgpt_b200::``), so the hot-path calls are opaque, shape-inferable
nodes for torch tooling (graph capture, export) instead of Python.  Weights are addressed by an integer handle into a
registry of packed modules (``register``), because custom ops only carry tensors and scalars.

    h = torch_ops.register(module)                      # module: a walkgpt_b200.modules.* instance
    y = torch.ops.walkgpt_b200.ctp_forward(x, h)        # == module(x) (without the reference's extra leading dim for 2-D x)
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from . import modules as M

_REGISTRY: Dict[int, torch.nn.Module] = {}


def register(module: torch.nn.Module) -> int:
    h = id(module)
    _REGISTRY[h] = module
    return h


def _get(h: int):
    return _REGISTRY[h]


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
