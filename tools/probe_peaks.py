#!/usr/bin/env python
"""Device bandwidth ceilings that bound the tokenizer kernels: copy (read+write, = MEASURED_PEAKS.json hbm_gbs),
write-only (fill) and read-only (sum), torch ops over 2 GiB, CUDA events, best of 10."""
import torch
dev = torch.device("cuda:0")
n = 1 << 28                                        # 2^28 int64 = 2 GiB
a = torch.empty(n, dtype=torch.int64, device=dev)
b = torch.empty(n, dtype=torch.int64, device=dev)


def best(fn, byt):
    ts = []
    for _ in range(12):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    t = min(ts[2:])
    return byt / t / 1e6


print("copy  (r+w)  GB/s: %.0f" % best(lambda: b.copy_(a), 2 * n * 8))
print("fill  (w)    GB/s: %.0f" % best(lambda: a.fill_(7), n * 8))
print("sum   (r)    GB/s: %.0f" % best(lambda: a.sum(), n * 8))
x = torch.empty(n * 8, dtype=torch.uint8, device=dev)
print("fill u8 (w)  GB/s: %.0f" % best(lambda: x.fill_(1), n * 8))
