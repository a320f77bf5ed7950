"""Straight-through estimator (reference: src/networks/vq/grad_approximation.py:7-29).

The B200 quantiser fuses the straight-through gradient into `vq_bwd` (see
functions/vq_function.py); this stand-alone version is kept because the reference exports it and
other modules may use it on their own.  It is device-agnostic autograd bookkeeping (no kernels)."""
from typing import Tuple

import torch
from torch.autograd import Function


class _CustomSTE(Function):
    """forward: value of `input_forward`; backward: the gradient goes to `input_backward`."""

    @staticmethod
    def forward(ctx, input_forward: torch.Tensor, input_backward: torch.Tensor) -> torch.Tensor:
        ctx.shape = input_backward.shape
        return input_forward.view_as(input_forward)

    @staticmethod
    def backward(ctx, grad_in: torch.Tensor) -> Tuple[None, torch.Tensor]:
        return None, grad_in.sum_to_size(ctx.shape)


def custom_straight_through_estimator(input_forward: torch.Tensor, input_backward: torch.Tensor) -> torch.Tensor:
    return _CustomSTE.apply(input_forward, input_backward)
