#!/usr/bin/env python
"""HBM bandwidth of a pure WRITE stream (torch fill of 4 GiB, best of 10, CUDA events) beside the copy figure of
MEASURED_PEAKS.json (read + write bytes): the roofline of a store-only kernel such as linquad_kernel."""
import json

import torch

x = torch.empty(1 << 29, dtype=torch.float64, device="cuda")  # 4 GiB
y = torch.empty_like(x)
best_w = best_c = 1e9
for _ in range(10):
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    x.fill_(1.0)
    e1.record()
    y.copy_(x)
    e2.record()
    torch.cuda.synchronize()
    best_w, best_c = min(best_w, e0.elapsed_time(e1)), min(best_c, e1.elapsed_time(e2))
print(json.dumps({"write_only_gbs": x.numel() * 8 / best_w / 1e6, "copy_read_plus_write_gbs": 2 * x.numel() * 8 / best_c / 1e6}))
