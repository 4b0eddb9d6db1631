from invertavatar_b200.encoder import (ConvGRU, DoubleConv, Up, recurrent_Up, TriPlanefeat_Encoder,  # noqa: F401
                                       TriPlaneSFTfeat_Encoder)
