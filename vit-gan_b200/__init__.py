"""vitgan_b200 -- B200-native (sm_100a) implementation of the ViT-GAN train-step hot path of
krzkro4122/vit-gan, behind the reference's own nn.Module interface.

Layout
  csrc/            hand-written CUDA kernels + the C ABI (include/vitgan_b200.h) -> libvitgan_b200.so
  lib.py, ops.py   ctypes binding and tensor-level wrappers (torch = memory/stream plumbing only)
  functional.py    torch.autograd.Function per reference block
  v2.py, v1.py     mirrors of the reference module classes (same names / parameters / state_dict keys)
  patch.py         swap the forwards of *reference* module instances in place (drop-in)
  train.py         the G+D step (reference call sites), data-parallel gradient buckets, CUDA-graph step

Importing this package requires the built CUDA library; there is no fallback path.
"""
from . import lib  # noqa: F401  (raises ImportError loudly if the .so is missing)
from . import ops, functional, v2, v1, patch, train  # noqa: F401
from .functional import set_precision, get_precision, skip_param_grads, set_operand_cache, set_dropout_policy  # noqa: F401

__all__ = ["lib", "ops", "functional", "v2", "v1", "patch", "train", "set_precision", "get_precision",
           "skip_param_grads", "set_operand_cache", "set_dropout_policy"]
