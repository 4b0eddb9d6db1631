from invertavatar_b200.glue import (constant, nan_to_num, suppress_tracer_warnings, assert_shape, profiled_function,  # noqa: F401
                                    params_and_buffers, named_params_and_buffers, copy_params_and_buffers, ddp_sync)
