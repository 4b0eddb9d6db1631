from invertavatar_b200.rendering import MipRayMarcher2  # noqa: F401
