from invertavatar_b200.faceverse import Faceverse_manager, FaceVerseModel  # noqa: F401
