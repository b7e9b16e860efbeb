from imp_release_b200.nets.adgm import AdaGMN  # noqa: F401  (eval/eval_imp.py:19)
