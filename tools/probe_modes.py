#!/usr/bin/env python
"""Random-access probe, all modes, at a few footprints."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
from lasv_b200 import _capi  # noqa: E402

lib = _capi.lib()
torch.cuda.set_device(0)
big = 100 * 10**9
buf = torch.zeros(big, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
n_ops = 1 << 27
names = {0: "read16+atomic", 1: "read16", 2: "atomic", 3: "read64+atomic"}
for F in (1 << 34, 50 * 10**9, 1 << 36, 80 * 10**9, big):
    for mode in (0, 1, 2, 3):
        best = None
        for it in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _capi.check(lib.lasv_random_access_probe_mode(buf.data_ptr(), F // 32, n_ops, 5 + it, mode, st))
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1) * 1e-3
            best = t if best is None else min(best, t)
        print(json.dumps({"footprint_bytes": F, "mode": names[mode], "ops_per_s": n_ops / best}), flush=True)
