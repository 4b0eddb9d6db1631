from invertavatar_b200.segformer import improved_os_unet_encoder, inversionNet  # noqa: F401
