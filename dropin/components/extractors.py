from imp_release_b200.extractors import ExtractSuperpoint, resize  # noqa: F401  (components/extractors.py:14,50)
