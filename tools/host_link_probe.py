"""What the host links give when 1, 2, 4, 8 GPUs copy pinned host memory at the same time (the ceiling of the
end-to-end `e2e` number, which moves 19.3 GB in and 19.3 GB out per GPU per step).  Developer tool (GPU box)."""
import json
import sys
import time

import torch

n = torch.cuda.device_count()
GB = 2
out = {"gpus_visible": n, "buffer_gb": GB, "results": []}
host = [torch.empty(GB << 30, dtype=torch.uint8).pin_memory() for _ in range(2 * n)]
dev = [(torch.empty(GB << 30, dtype=torch.uint8, device="cuda:%d" % g), torch.empty(GB << 30, dtype=torch.uint8, device="cuda:%d" % g)) for g in range(n)]
s_in = [torch.cuda.Stream(device=g) for g in range(n)]
s_out = [torch.cuda.Stream(device=g) for g in range(n)]


def run(k, mode, reps=3):
    for g in range(k):
        torch.cuda.synchronize(g)
    t0 = time.perf_counter()
    for _ in range(reps):
        for g in range(k):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s_in[g]):
                    dev[g][0].copy_(host[2 * g], non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s_out[g]):
                    host[2 * g + 1].copy_(dev[g][1], non_blocking=True)
    for g in range(k):
        torch.cuda.synchronize(g)
    dt = time.perf_counter() - t0
    per_dir = GB * (1 << 30) * reps * k / dt / 1e9
    return per_dir


k = 1
while k <= n:
    run(k, "both", 1)
    out["results"].append({"gpus": k, "h2d_only_gbs_total": round(run(k, "h2d"), 1), "d2h_only_gbs_total": round(run(k, "d2h"), 1),
                           "both_gbs_total_each_direction": round(run(k, "both"), 1)})
    k *= 2
print(json.dumps(out))
