"""Launch the reference's UNMODIFIED eval/eval_imp.py against the B200 implementation:

    cd /path/to/imp-release && python /path/to/repo/dropin/run_eval_imp.py --matching_method IMP --dataset yfcc

(`eval_imp.py` reads configs/ and weights/ relative to the cwd, eval/eval_imp.py:240-248,332.)"""
import os
import runpy
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))      # our `nets` shadows the reference's
sys.path.insert(1, os.getcwd())
runpy.run_module('eval.eval_imp', run_name='__main__')
