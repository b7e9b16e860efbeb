from imp_release_b200.nets.gms import DGNNS  # noqa: F401  (eval/eval_imp.py:18)
