"""Measure read-only / write-only / copy HBM bandwidth on this GPU (context for a write-dominated kernel)."""
import torch, json
dev = torch.device("cuda:0")
n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device=dev)
b = torch.empty(n, dtype=torch.uint8, device=dev)
af = a.view(torch.float32)
def t(fn, reps=10):
    for _ in range(3): fn()
    best = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best
w = t(lambda: a.zero_())
wf = t(lambda: af.fill_(1.5))
c = t(lambda: b.copy_(a))
r = t(lambda: af.sum())
print(json.dumps({"write_only_GBs_zero": n / w / 1e6, "write_only_GBs_fill": n / wf / 1e6, "copy_GBs_rw": 2 * n / c / 1e6, "read_only_GBs_sum": n / r / 1e6}))
