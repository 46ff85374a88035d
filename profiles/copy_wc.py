"""Does write-combined pinned memory (cudaHostAllocWriteCombined) for the INPUT buffers raise the host<->device ceiling?
    python profiles/copy_wc.py [clips] [chunk_clips]
Same traffic pattern as profiles/copy_ceiling.py (H2D and D2H at once, chunks on two streams, no kernels), input buffer
allocated with torch pin_memory (cudaHostAlloc default) vs cudaHostAlloc(WriteCombined); one JSON line.
"""
import ctypes
import json
import sys
import time

import torch
from cuda.bindings import runtime as cudart

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 128
n = 480000
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)


def wc_tensor():
    size = clips * n * 4
    err, ptr = cudart.cudaHostAlloc(size, cudart.cudaHostAllocWriteCombined)
    assert int(err) == 0, err
    buf = (ctypes.c_float * (clips * n)).from_address(int(ptr))
    t = torch.frombuffer(buf, dtype=torch.float32).view(clips, n)
    t.zero_()
    return t


def run(xh, yh, h2d, d2h, reps=3):
    dx = [torch.empty((chunk, n), device=dev) for _ in range(2)]
    dy = [torch.zeros((chunk, n), device=dev) for _ in range(2)]
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    best = 1e9
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i, b0 in enumerate(range(0, clips, chunk)):
            nb = min(chunk, clips - b0)
            if h2d:
                with torch.cuda.stream(s_in):
                    dx[i & 1][:nb].copy_(xh[b0:b0 + nb], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s_out):
                    yh[b0:b0 + nb].copy_(dy[i & 1][:nb], non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return clips * n * 4 / best / 1e9


yh = torch.empty((clips, n), dtype=torch.float32, pin_memory=True)
out = {}
for name, xh in (("default", torch.zeros((clips, n), dtype=torch.float32).pin_memory()), ("write_combined", wc_tensor())):
    out[name] = {"h2d_alone": round(run(xh, yh, True, False), 2), "d2h_alone": round(run(xh, yh, False, True), 2),
                 "both_each_way": round(run(xh, yh, True, True), 2)}
print(json.dumps({"clips": clips, "chunk_clips": chunk, "GB/s": out}))
