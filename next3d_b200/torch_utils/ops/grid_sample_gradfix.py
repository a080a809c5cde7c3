"""`torch_utils.ops.grid_sample_gradfix.grid_sample(input, grid)` API shim (reference: grid_sample_gradfix.py:28):
bilinear, zeros padding, align_corners=False.  The reference customises gradients only; the engine's tri-plane and UV
lookups are fused into render_kernel / uv_sample_kernel and never call this."""
import torch

enabled = False


def grid_sample(input, grid):
    return torch.nn.functional.grid_sample(input=input, grid=grid, mode='bilinear', padding_mode='zeros', align_corners=False)
