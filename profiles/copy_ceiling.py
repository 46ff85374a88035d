"""Kernel-free copy benchmark: what the box's host<->device links deliver with the e2e leg's traffic pattern.

    python profiles/copy_ceiling.py [clips] [chunk_clips]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/copy_ceiling.py

Per rank: `clips` x 480 000 float32 samples in pinned host memory, copied H2D and (a same-sized result) D2H in
chunks of `chunk_clips` on two streams, exactly like qd_render_host but with no kernel in between.  Reports
H2D alone, D2H alone and both directions at once (max over ranks), as GB/s per direction summed over the ranks
and as the audio-seconds/s an infinitely fast GPU would reach.  One JSON line from rank 0.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def copy_ceiling(clips: int, n: int, chunk: int, dev, world: int, reps: int = 3, elem_bytes: int = 4):
    """-> dict of seconds per pass (max over ranks) for 'h2d', 'd2h', 'both'."""
    dt = torch.float32 if elem_bytes == 4 else torch.int16
    xh = torch.empty((clips, n), dtype=dt, pin_memory=True)
    yh = torch.empty((clips, n), dtype=dt, pin_memory=True)
    xh.zero_()
    yh.zero_()
    dx = [torch.empty((chunk, n), dtype=dt, device=dev) for _ in range(2)]
    dy = [torch.zeros((chunk, n), dtype=dt, device=dev) for _ in range(2)]
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def one_pass(h2d: bool, d2h: bool):
        for i, b0 in enumerate(range(0, clips, chunk)):
            nb = min(chunk, clips - b0)
            if h2d:
                with torch.cuda.stream(s_in):
                    dx[i & 1][:nb].copy_(xh[b0:b0 + nb], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s_out):
                    yh[b0:b0 + nb].copy_(dy[i & 1][:nb], non_blocking=True)
        torch.cuda.synchronize(dev)

    out = {}
    for tag, (a, b) in (("h2d", (True, False)), ("d2h", (False, True)), ("both", (True, True))):
        one_pass(a, b)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(reps):
            one_pass(a, b)
        t = (time.perf_counter() - t0) / reps
        if world > 1:
            tt = torch.tensor([t], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt.item())
        out[tag] = t
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
    res = {}
    for eb in (4, 2):
        t = copy_ceiling(clips, n, chunk, dev, world, elem_bytes=eb)
        gb = clips * n * eb * world / 1e9
        res["float32" if eb == 4 else "pcm16"] = {
            "h2d_gbs": gb / t["h2d"], "d2h_gbs": gb / t["d2h"], "both_gbs_each_way": gb / t["both"],
            "ceiling_audio_s_per_s": clips * 10 * world / t["both"], "ms_both": 1e3 * t["both"]}
    if rank == 0:
        print(json.dumps({"copy_ceiling": res, "n_gpus": world, "clips_per_gpu": clips, "chunk_clips": chunk,
                          "cpus": os.cpu_count()}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
