"""walkgpt_b200 -- B200-native (sm_100a) implementation of WalkGPT's pixel-grounding forward path.

Python here is host-side plumbing only (tensor allocation, streams, weight repacking, torch.distributed);
all arithmetic runs in hand-written CUDA kernels behind the C ABI of ``include/walkgpt_b200.h``.
"""
from ._lib import WalkGPTB200Error, exported_symbols, lib, require_device  # noqa: F401

__version__ = "0.1.0"
