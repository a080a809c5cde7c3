from . import bias_act, upfirdn2d, conv2d_gradfix, conv2d_resample, fma, filtered_lrelu, grid_sample_gradfix  # noqa: F401
