"""Run ONE C2-style adaptive solve (for ncu): python tools/gpu_one.py <fp> <sched> <N> [alg] [dtype]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
from gpu_probe import dg, time_asolve  # noqa: E402

fp, sched, N = sys.argv[1], sys.argv[2], int(float(sys.argv[3]))
alg = getattr(dg, sys.argv[4])() if len(sys.argv) > 4 else dg.GPUTsit5()
dtype = np.float64 if len(sys.argv) > 5 and sys.argv[5] == "f64" else np.float32
tol = float(sys.argv[6]) if len(sys.argv) > 6 else 1e-6
print(time_asolve(N, fp, sched, alg=alg, reps=4, dtype=dtype, tol=tol))
