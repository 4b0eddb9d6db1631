from invertavatar_b200.segformer import (Mlp, Attention, Block, OverlapPatchEmbed, MixVisionTransformer, DWConv, MLP,  # noqa: F401
                                         transformer_block)
