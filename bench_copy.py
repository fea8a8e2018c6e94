#!/usr/bin/env python
"""bench_copy.py -- copy-only ceiling of the host-buffer path: N ranks (one per GPU, torchrun) move the
exact transfers of bench.py's e2e leg -- S*B*4 bytes per block and direction between pinned host memory and the device --
with NO kernels, all ranks at the same time, and rank 0 prints one JSON object.

For every kind of host memory (0: cudaMallocHost; 1: write-combined input; 2: transparent-huge-page arena, cudaHostRegister'ed)
and every direction (H2D only, D2H only, both at once) it reports the aggregate GB/s over all ranks (bytes of all ranks / the
slowest rank's wall time between two barriers), with whole-block transfers and with 16 MB pieces.  What the engine's
submit/wait pipeline reaches (bench.py: e2e.host_copy_used_gbs) is held against the "both" figure.

  python bench_copy.py                                   # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench_copy.py
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=65536)
    ap.add_argument("--block", type=int, default=512)
    ap.add_argument("--iters", type=int, default=12)
    ap.add_argument("--modes", default="0,1,2")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from irbaboon_b200 import engine as eng
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def vmax(v):
        if world == 1:
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    nbytes = a.streams * a.block * 4
    names = {0: "cudaMallocHost", 1: "write-combined input (cudaHostAllocWriteCombined), default output", 2: "mmap + MADV_HUGEPAGE + cudaHostRegister"}
    out = {"what": "copy-only host<->device ceiling, %d rank(s), %d bytes per block and direction (%d streams x %d samples x 4), %d blocks per measurement"
                   % (world, nbytes, a.streams, a.block, a.iters), "ranks": world, "bytes_per_block": nbytes, "modes": []}
    for mode in [int(m) for m in a.modes.split(",")]:
        row = {"host_mode": mode, "host_memory": names[mode]}
        try:
            p = eng.CopyProbe(nbytes, mode, device=local)
        except Exception as ex:
            row["error"] = str(ex)
            ok = 0.0
        else:
            ok = 1.0
        if world > 1:
            t = torch.tensor([ok], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            ok = float(t.item())
        if not ok:
            row.setdefault("error", "setup failed on another rank")
            out["modes"].append(row)
            continue
        p.run(2, 3)
        for chunk, cname in ((0, "whole_block"), (16 << 20, "pieces_16MB")):
            for direction, dname in ((1, "h2d"), (2, "d2h"), (3, "both")):
                barrier()
                secs = vmax(p.run(a.iters, direction, chunk))
                row["%s_%s_gbs" % (cname, dname)] = world * nbytes * a.iters * (2 if direction == 3 else 1) / secs / 1e9
        # one rank at a time, both directions: what a single GPU gets when the others are idle
        solo = []
        for r in range(world):
            barrier()
            s = p.run(4, 3) if r == rank else 0.0
            solo.append(vmax(s))
        row["solo_both_gbs_per_rank"] = [nbytes * 4 * 2 / s / 1e9 for s in solo]
        p.close()
        out["modes"].append(row)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
