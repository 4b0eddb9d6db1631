from invertavatar_b200.glue import (EasyDict, get_obj_by_name, call_func_by_name, construct_class_by_name, is_url, open_url)  # noqa: F401
