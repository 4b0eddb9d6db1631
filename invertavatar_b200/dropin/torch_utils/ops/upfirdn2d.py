from invertavatar_b200.ops import (setup_filter, upfirdn2d, filter2d, upsample2d, downsample2d,  # noqa: F401
                                   _parse_scaling, _parse_padding, _get_filter_size)
