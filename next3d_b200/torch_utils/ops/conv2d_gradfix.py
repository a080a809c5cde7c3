"""`torch_utils.ops.conv2d_gradfix` API shim (reference: torch_utils/ops/conv2d_gradfix.py:37-45).

The reference module exists to customise *gradients*; its forward is torch.nn.functional.conv2d / conv_transpose2d (cuDNN).
This inference-only mirror keeps the names, signatures and module-level flags and forwards to the same library entry
points -- it is NOT on next3d_b200's hot path (the engine and `conv2d_resample` use the tcgen05 implicit-GEMM kernel).
"""
import contextlib

import torch

enabled = False
weight_gradients_disabled = False


@contextlib.contextmanager
def no_weight_gradients(disable=True):
    global weight_gradients_disabled
    old = weight_gradients_disabled
    if disable:
        weight_gradients_disabled = True
    yield
    weight_gradients_disabled = old


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    return torch.nn.functional.conv2d(input=input, weight=weight, bias=bias, stride=stride, padding=padding, dilation=dilation, groups=groups)


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    return torch.nn.functional.conv_transpose2d(input=input, weight=weight, bias=bias, stride=stride, padding=padding,
                                                output_padding=output_padding, groups=groups, dilation=dilation)
