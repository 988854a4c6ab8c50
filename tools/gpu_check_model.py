"""Quick GPU driver for the tiny-dims model parity (prints instead of asserting)."""
import sys, os, glob, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import llava_dpo_oracle as O
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gpu_model_parity as T
from rlaifv_b200.model import LlavaDPOPolicy
from rlaifv_b200 import ops
fails = 0
for path in T.GOLDEN:
    for fn in (T.test_forward_matches_reference_fixture, T.test_dpo_loss_and_grads_match_reference_fixture):
        try:
            fn(path)
            print("ok  ", fn.__name__, os.path.basename(path), flush=True)
        except Exception:
            fails += 1
            print("FAIL", fn.__name__, os.path.basename(path), flush=True)
            traceback.print_exc()
print("FAILS", fails)
