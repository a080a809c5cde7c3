"""Mirror of the reference's `torch_utils` package for the hot path: only `torch_utils.ops` (the public op API, SURVEY.md
section 8b).  To run the reference's own model code on these kernels, alias the modules before importing it:

    import sys, next3d_b200.torch_utils.ops as ops
    for m in ('bias_act', 'upfirdn2d', 'conv2d_resample', 'conv2d_gradfix', 'fma', 'filtered_lrelu', 'grid_sample_gradfix'):
        sys.modules[f'torch_utils.ops.{m}'] = getattr(ops, m)
"""
