from invertavatar_b200.ops import grid_sample  # noqa: F401

enabled = False
