from imp_release_b200.nets.superpoint import SuperPoint  # noqa: F401  (components/extractors.py:11)
