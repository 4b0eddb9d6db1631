from invertavatar_b200.encoder import Bottleneck, get_block, get_blocks, SEModule, bottleneck_IR_SE  # noqa: F401
