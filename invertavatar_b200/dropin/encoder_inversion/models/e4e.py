from invertavatar_b200.encoder import GradualStyleBlock, Encoder4Editing  # noqa: F401
