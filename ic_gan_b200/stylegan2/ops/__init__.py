from . import bias_act, conv2d_gradfix, conv2d_resample, fma, upfirdn2d  # noqa: F401
