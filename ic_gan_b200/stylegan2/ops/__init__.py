from . import bias_act, upfirdn2d  # noqa: F401
