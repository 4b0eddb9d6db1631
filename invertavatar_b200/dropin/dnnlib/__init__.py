from . import util  # noqa: F401
from .util import EasyDict  # noqa: F401
