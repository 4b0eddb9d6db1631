from invertavatar_b200.segformer import (UpLayer, TriPlanefeat_SegformerDecoder, TriPlaneSFTfeat_SegformerDecoder,  # noqa: F401
                                         MLP, MixVisionTransformer, transformer_block)
