import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def save(name, arr, save_dir):
    """np.save with the reference's file names and (6,TT)/(2,TT) float64 C-order layout (main_newton_method.py:184-186)."""
    import numpy as np
    os.makedirs(save_dir, exist_ok=True)
    np.save(os.path.join(save_dir, name), np.ascontiguousarray(arr, dtype=np.float64))
