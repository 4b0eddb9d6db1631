from invertavatar_b200.encoder import unet_encoder, inversionNet  # noqa: F401
