from invertavatar_b200.rendering import (ImportanceRenderer, ImportanceRenderer_bsMotion, fill_mouth, sample_from_planes,  # noqa: F401
                                         generate_planes)
