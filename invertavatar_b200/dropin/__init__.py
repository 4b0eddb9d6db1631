"""Drop-in import tree: the reference's module paths resolved onto the B200 engine.

Put this directory in front of ``sys.path`` (``invertavatar_b200.dropin.install()`` or ``PYTHONPATH=.../invertavatar_b200/dropin``)
and ``reenact_avatar_next3d.py`` / ``eval_seq.py`` import ``torch_utils``, ``dnnlib``, ``legacy``, ``training``,
``training_avatar_texture`` and ``encoder_inversion`` from here instead of from the reference tree, unchanged."""
import os
import sys


def install():
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    return here
