from invertavatar_b200.rendering import RaySampler, RaySampler_zxc  # noqa: F401
