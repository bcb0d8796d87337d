#!/usr/bin/env python
"""Single field at bw (default 2048) on all visible GPUs through s2kit_cuda_multi_*: ms per transform and, with
S2KIT_CUDA_MULTI_PROF=1, the per-stage CUDA-event times of every GPU (stderr)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import s2kit_b200 as s2  # noqa: E402

bw = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
G = int(sys.argv[2]) if len(sys.argv) > 2 else torch.cuda.device_count()
n = 2 * bw
rng = np.random.RandomState(2048)
rd, idt = rng.uniform(-1, 1, (n, n)), rng.uniform(-1, 1, (n, n))
M = s2.MultiPlan(bw, G)
got_r, got_i = M.forward(rd, idt)
out = {"bw": bw, "gpus": G}
for inv in (False, True):
    M.run(inverse=inv, iters=5)
    out["ms_inverse" if inv else "ms_forward"] = M.run(inverse=inv, iters=10)
M.close()
print(json.dumps(out))
