"""`torch_utils.ops.bias_act` on libnext3d_b200 (reference: torch_utils/ops/bias_act.py:54-88, plugin bias_act.cpp:36-94).

Same signature and defaults.  impl='cuda' (default) launches n3d_bias_act and REQUIRES a CUDA tensor (no silent CPU
fallback, unlike bias_act.py:86-88); impl='ref' evaluates the same formula with plain torch ops on any device -- it is
the per-op reference implementation the upstream file ships as `_bias_act_ref` (:93-122).  Forward only.
"""
import numpy as np
import torch

from ... import _lib


class _Spec(dict):
    __getattr__ = dict.__getitem__


activation_funcs = {                                                        # bias_act.py:23-33
    'linear': _Spec(func=lambda x, **_: x, def_alpha=0, def_gain=1, cuda_idx=1, ref='', has_2nd_grad=False),
    'relu': _Spec(func=lambda x, **_: torch.nn.functional.relu(x), def_alpha=0, def_gain=np.sqrt(2), cuda_idx=2, ref='y', has_2nd_grad=False),
    'lrelu': _Spec(func=lambda x, alpha, **_: torch.nn.functional.leaky_relu(x, alpha), def_alpha=0.2, def_gain=np.sqrt(2), cuda_idx=3, ref='y', has_2nd_grad=False),
    'tanh': _Spec(func=lambda x, **_: torch.tanh(x), def_alpha=0, def_gain=1, cuda_idx=4, ref='y', has_2nd_grad=True),
    'sigmoid': _Spec(func=lambda x, **_: torch.sigmoid(x), def_alpha=0, def_gain=1, cuda_idx=5, ref='y', has_2nd_grad=True),
    'elu': _Spec(func=lambda x, **_: torch.nn.functional.elu(x), def_alpha=0, def_gain=1, cuda_idx=6, ref='y', has_2nd_grad=True),
    'selu': _Spec(func=lambda x, **_: torch.nn.functional.selu(x), def_alpha=0, def_gain=1, cuda_idx=7, ref='y', has_2nd_grad=True),
    'softplus': _Spec(func=lambda x, **_: torch.nn.functional.softplus(x), def_alpha=0, def_gain=1, cuda_idx=8, ref='y', has_2nd_grad=True),
    'swish': _Spec(func=lambda x, **_: torch.sigmoid(x) * x, def_alpha=0, def_gain=np.sqrt(2), cuda_idx=9, ref='x', has_2nd_grad=True),
}

_DTYPES = {torch.float32: 0, torch.float16: 1}


def _bias_act_ref(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None):
    spec = activation_funcs[act]
    alpha = float(alpha if alpha is not None else spec.def_alpha)
    gain = float(gain if gain is not None else spec.def_gain)
    clamp = float(clamp if clamp is not None else -1)
    if b is not None:
        assert b.ndim == 1 and 0 <= dim < x.ndim and b.shape[0] == x.shape[dim]
        x = x + b.reshape([-1 if i == dim else 1 for i in range(x.ndim)])
    x = spec.func(x, alpha=alpha)
    if gain != 1:
        x = x * gain
    if clamp >= 0:
        x = x.clamp(-clamp, clamp)
    return x


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None, impl='cuda'):
    assert isinstance(x, torch.Tensor)
    assert impl in ['ref', 'cuda']
    if impl == 'ref':
        return _bias_act_ref(x=x, b=b, dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp)
    _lib.require_cuda(x, 'bias_act(x)')
    if x.dtype not in _DTYPES:
        raise RuntimeError(f'bias_act: dtype {x.dtype} is only available with impl="ref"')
    spec = activation_funcs[act]
    alpha = float(alpha if alpha is not None else spec.def_alpha)
    gain = float(gain if gain is not None else spec.def_gain)
    clamp = float(clamp if clamp is not None else -1)
    assert clamp is None or clamp >= 0 or clamp == -1
    if not (x.is_contiguous() or x.is_contiguous(memory_format=torch.channels_last)):
        x = x.contiguous()                                                     # plugin requires dense x (bias_act.cpp:51)
    y = torch.empty_like(x)
    size_b = step_b = 0
    if b is not None:
        assert b.ndim == 1 and 0 <= dim < x.ndim and b.shape[0] == x.shape[dim]
        b = b.to(x.dtype).contiguous()
        size_b, step_b = b.shape[0], max(int(x.stride(dim)), 1)
    if x.numel() > 0:
        _lib.check(_lib.lib.n3d_bias_act(x.data_ptr(), _lib.ptr(b), y.data_ptr(), _DTYPES[x.dtype], x.numel(), size_b, step_b, spec.cuda_idx,
                                         alpha, gain, clamp, _lib.stream_ptr(x.device)), 'n3d_bias_act')
    return y
