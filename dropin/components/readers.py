from imp_release_b200.readers import reader_set, standard_reader  # noqa: F401  (components/readers.py:8,41; used by components/evaluators.py)
