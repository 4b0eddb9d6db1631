from invertavatar_b200.glue import (GaussianCameraPoseSampler, LookAtPoseSampler, UniformCameraPoseSampler,  # noqa: F401
                                    create_cam2world_matrix, FOV_to_intrinsics)
