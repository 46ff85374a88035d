"""Kernel-free copy benchmark: what the box's host<->device links deliver with the e2e leg's traffic pattern, and what
limits it.

    python profiles/copy_ceiling.py [clips] [chunk_clips]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/copy_ceiling.py

Per rank: `clips` x 480 000 float32 samples of host memory copied H2D and (a same-sized result) D2H in chunks of
`chunk_clips` on two streams, exactly like qd_render_host but with no kernel in between.  Variants:
  pinned      cudaHostAlloc'd memory (torch pin_memory=True) -- what bench.py's e2e leg uses
  hugepage    2 MB-aligned anonymous memory with MADV_HUGEPAGE, faulted in, then cudaHostRegister'ed: 512x fewer IOMMU
              / page-table entries per byte (is the ceiling an address-translation limit?)
each as H2D alone, D2H alone and both directions at once (max over ranks), in GB/s per direction summed over the ranks.
One JSON line from rank 0.
"""
import ctypes
import json
import mmap
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

HUGE = 2 << 20


def alloc(kind: str, clips: int, n: int):
    if kind == "pinned":
        t = torch.empty((clips, n), dtype=torch.float32, pin_memory=True)
        t.zero_()
        return t, None
    size = clips * n * 4
    mm = mmap.mmap(-1, size + HUGE, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    base = ctypes.addressof(ctypes.c_char.from_buffer(mm))
    off = (-base) % HUGE
    try:
        mm.madvise(mmap.MADV_HUGEPAGE, off, size - size % HUGE)
    except (OSError, AttributeError, ValueError):
        pass
    t = torch.frombuffer(mm, dtype=torch.float32, count=clips * n, offset=off).view(clips, n)
    t.zero_()                                              # fault the pages in (as huge pages when THP allows)
    rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), size, 0)
    assert int(rc) == 0, f"cudaHostRegister failed: {rc}"
    return t, mm


def one_pass(xh, yh, dx, dy, s_in, s_out, chunk, h2d, d2h, dev):
    clips = xh.shape[0]
    for i, b0 in enumerate(range(0, clips, chunk)):
        nb = min(chunk, clips - b0)
        if h2d:
            with torch.cuda.stream(s_in):
                dx[i & 1][:nb].copy_(xh[b0:b0 + nb], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s_out):
                yh[b0:b0 + nb].copy_(dy[i & 1][:nb], non_blocking=True)
    torch.cuda.synchronize(dev)


def measure(kind, clips, n, chunk, dev, world, reps=3):
    xh, keep_x = alloc(kind, clips, n)
    yh, keep_y = alloc(kind, clips, n)
    dx = [torch.empty((chunk, n), dtype=torch.float32, device=dev) for _ in range(2)]
    dy = [torch.zeros((chunk, n), dtype=torch.float32, device=dev) for _ in range(2)]
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    out = {"is_pinned": bool(xh.is_pinned()),
           "anon_huge_pages": next((ln.split(":")[1].strip() for ln in open("/proc/meminfo") if ln.startswith("AnonHugePages")), None)}
    for tag, (a, b) in (("h2d", (True, False)), ("d2h", (False, True)), ("both", (True, True))):
        one_pass(xh, yh, dx, dy, s_in, s_out, chunk, a, b, dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(reps):
            one_pass(xh, yh, dx, dy, s_in, s_out, chunk, a, b, dev)
        t = (time.perf_counter() - t0) / reps
        if world > 1:
            tt = torch.tensor([t], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt.item())
        out[tag + "_gbs"] = clips * n * 4 * world / t / 1e9
    out["ceiling_audio_s_per_s"] = out["both_gbs"] * 1e9 / (4 * 48000)
    if keep_x is not None:
        torch.cuda.cudart().cudaHostUnregister(xh.data_ptr())
        torch.cuda.cudart().cudaHostUnregister(yh.data_ptr())
    return out


def main():
    clips = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    n = 480000
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    res = {k: measure(k, clips, n, chunk, dev, world) for k in ("pinned", "hugepage")}
    if rank == 0:
        thp = ""
        try:
            thp = open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip()
        except OSError:
            pass
        print(json.dumps({"copy_ceiling_gbs_each_way": res, "n_gpus": world, "clips_per_gpu": clips, "chunk_clips": chunk,
                          "cpus": os.cpu_count(), "transparent_hugepage": thp}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
