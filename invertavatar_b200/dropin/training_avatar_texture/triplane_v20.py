from invertavatar_b200.triplane import TriPlaneGenerator, OSGDecoder  # noqa: F401
