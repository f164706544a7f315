#!/usr/bin/env python
"""Random-access ceiling of one B200 as a function of footprint: uniform random
32-byte read + 32-bit L2 atomic per op (lasv_random_access_probe) over the first
F bytes of one large buffer.  Prints one JSON line per footprint."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
from lasv_b200 import _capi  # noqa: E402

lib = _capi.lib()
torch.cuda.set_device(0)
big = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100 * 10**9
buf = torch.zeros(big, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
n_ops = 1 << 27
F = 1 << 24
sizes = []
while F < big:
    sizes.append(F)
    F *= 4
sizes.append(big)
for F in sizes:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = None
    for it in range(3):
        e0.record()
        _capi.check(lib.lasv_random_access_probe(buf.data_ptr(), F // 32, n_ops, 77 + it, st))
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e-3
        best = t if best is None else min(best, t)
    print(json.dumps({"footprint_bytes": F, "ops_per_s": n_ops / best, "GBps_64B_per_op": n_ops * 64 / best / 1e9}),
          flush=True)
