from invertavatar_b200.glue import load_network_pkl  # noqa: F401
