from invertavatar_b200.ops import fma  # noqa: F401
