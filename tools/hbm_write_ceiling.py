"""GPU-box helper: measured PURE-WRITE and copy HBM bandwidth of this part, the two numbers that bracket the mask
einsum (it writes 2.04 GB and reads 0.54 GB per launch).  MEASURED_PEAKS.json's hbm_gbs is a copy (read + write
bytes); a write-dominated kernel cannot reach it if the part's write-only bandwidth is lower.  Prints one JSON line."""
import json
import torch

dev = torch.device('cuda', 0)
n = 1 << 30                       # 2 GiB of bf16
a = torch.empty(n, dtype=torch.bfloat16, device=dev)
b = torch.empty(n, dtype=torch.bfloat16, device=dev)
res = {}


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


res['write_only_fill_gbs'] = 2 * n / timed(lambda: a.fill_(1.0)) / 1e9          # cudaMemset-class store kernel
res['write_only_zero_gbs'] = 2 * n / timed(lambda: a.zero_()) / 1e9
res['copy_read_plus_write_gbs'] = 4 * n / timed(lambda: b.copy_(a)) / 1e9        # MEASURED_PEAKS.json's definition
# the einsum's own traffic (B=16, 1024^2, Q=100): 0.5455 GB read + 2.0403 GB written per launch.  Its floor is the slower of
# "all bytes at the copy rate" and "the written bytes at the write-only rate"
res['einsum_floor_ms'] = 1e3 * max((0.5455 + 2.0403) / res['copy_read_plus_write_gbs'], 2.0403 / res['write_only_fill_gbs'])
print(json.dumps(res))
