from invertavatar_b200.ops import filtered_lrelu  # noqa: F401
