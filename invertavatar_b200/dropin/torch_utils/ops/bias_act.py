from invertavatar_b200.ops import activation_funcs, bias_act  # noqa: F401
