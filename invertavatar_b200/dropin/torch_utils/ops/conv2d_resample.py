from invertavatar_b200.ops import conv2d_resample  # noqa: F401
