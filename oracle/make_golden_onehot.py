"""Generate tests/golden/onehot_*.npz from the UNMODIFIED reference `OneHotEncoder` (src/functions/onehot.py; it imports
nothing but torch).  Run in the build container only: python oracle/make_golden_onehot.py"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src/functions/onehot.py"


def load_reference_class():
    spec = importlib.util.spec_from_file_location("ref_onehot", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.OneHotEncoder


CASES = {"codes_k10": (2, 24, 24, 11, 0), "codes_k512": (1, 16, 16, 513, 1), "ragged": (3, 5, 7, 4, 2)}

if __name__ == "__main__":
    Ref = load_reference_class()
    for name, (B, H, W, C, seed) in CASES.items():
        g = torch.Generator().manual_seed(seed)
        t = torch.randint(0, C, (B, H, W), generator=g).int()       # the trainer passes .int() maps (single_window_trainer.py:93)
        out = Ref(C)(t)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"onehot_{name}.npz"), t=t.numpy(), n_classes=C,
                            out=out.numpy().astype(np.uint8))
        print(name, tuple(out.shape), out.dtype)
