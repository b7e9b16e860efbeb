from imp_release_b200.nets.layers import normalize_keypoints, SHARING_LAYERS  # noqa: F401
