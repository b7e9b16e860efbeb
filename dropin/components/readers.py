from imp_release_b200.readers import standard_reader  # noqa: F401  (components/readers.py:8; used by components/evaluators.py)
