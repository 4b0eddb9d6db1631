from invertavatar_b200.persistence import *  # noqa: F401,F403
from invertavatar_b200.persistence import persistent_class, is_persistent, import_hook, _reconstruct_persistent_obj  # noqa: F401
