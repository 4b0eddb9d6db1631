from invertavatar_b200.superresolution import SuperresolutionHybrid8XDC, SuperresolutionHybrid8X  # noqa: F401
