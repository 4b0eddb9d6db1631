from invertavatar_b200.stylegan2 import (FullyConnectedLayer, MappingNetwork, SynthesisLayer, ToRGBLayer, SynthesisBlock,  # noqa: F401
                                         SynthesisNetwork, Generator)
from invertavatar_b200.ops import modulated_conv2d  # noqa: F401
