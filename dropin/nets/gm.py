from imp_release_b200.nets.gm import GM, normalize_keypoints  # noqa: F401  (eval/matching.py:11)
