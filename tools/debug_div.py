import ctypes, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rt-3d-reconstruction_b200"))
import torch
from rt3d import _lib
torch.zeros(1, device="cuda")
L = _lib.lib()
L.rt3d_debug_fast_division_mismatches.restype = ctypes.c_longlong
L.rt3d_debug_fast_division_mismatches.argtypes = [ctypes.c_double]
for b in (527.8, 527.6, 0.005, 0.002, 0.01, 791.7, 3.0):
    print(b, L.rt3d_debug_fast_division_mismatches(b))
