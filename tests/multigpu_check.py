"""Run under torchrun on N GPUs (gpurun --gpus N): the map sharded over the ranks (CUDA-IPC peers over
NVLink) must be bit-identical to the same frames on one GPU.  Rank 0 runs the unsharded context too and
compares; prints one line per check and exits non-zero on any difference.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tests/multigpu_check.py [dim] [particles]
"""
import os
import sys
import time
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from soilmachine_b200 import capi, presets, host, sharded  # noqa: E402


def main():
    dim = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pre = presets.load("rockgravelpebblessand")
    sh = sharded.DistShard(dim, dim, pre["world"]["scale"], device=local, max_particles=n)
    sh.ctx.set_soils(pre["soils"])
    sh.ctx.initialize(42, pre["layers"])
    host.srand(42)
    frames = [(host.spawn_list(n, dim, dim), host.spawn_list(n, dim, dim)) for _ in range(2)]
    tot = np.zeros(4)
    t_ms = 0.0
    for xw, xd in frames:
        dw, dd = sh.ctx.device_spawn(xw), sh.ctx.device_spawn(xd)
        dist.barrier()
        a = sh.run("water", dw, n)
        b = sh.run("wind", dd, n)
        sh.ctx.frequency_update()
        tot += [a.steps, b.steps, a.sweeps, b.sweeps]
        t_ms += a.device_ms + b.device_ms
        sh.ctx.device_free(dw); sh.ctx.device_free(dd)
    h = torch.from_numpy(sh.ctx.heights().copy()).cuda()
    parts = [torch.empty((c1 - c0, dim), dtype=torch.float64, device="cuda") for c0, c1 in _ranges(dim, world)]
    dist.all_gather(parts, h) if len(set(p.shape for p in parts)) == 1 else _gather_ragged(parts, h, rank, world)
    steps = torch.tensor(tot[:2], dtype=torch.float64, device="cuda")
    dist.all_reduce(steps)
    tmax = torch.tensor([t_ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ok = True
    if rank == 0:
        full = torch.cat(parts, 0).cpu().numpy()
        one = capi.Context(dim, dim, pre["world"]["scale"], device=local, max_particles=n)
        one.set_soils(pre["soils"]); one.initialize(42, pre["layers"])
        s1 = np.zeros(2); t1 = 0.0
        for xw, xd in frames:
            a = one.water_run(xw); b = one.wind_run(xd); one.frequency_update()
            s1 += [a.steps, b.steps]; t1 += a.device_ms + b.device_ms
        same_h = np.array_equal(full.view(np.uint8), one.heights().view(np.uint8))
        same_s = np.array_equal(s1, steps.cpu().numpy())
        ok = same_h and same_s
        print("multigpu_check world=%d dim=%d n=%d: heights %s, steps %s (%s) | sharded %.1f ms vs one GPU %.1f ms"
              % (world, dim, n, "IDENTICAL" if same_h else "DIFFER", "IDENTICAL" if same_s else "DIFFER",
                 steps.cpu().numpy().tolist(), tmax.item(), t1), flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    sh.close()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


def _ranges(dim, world):
    w = ((((dim + world - 1) // world) + 15) // 16) * 16
    return [(q * w, min(dim, (q + 1) * w)) for q in range(world)]


def _gather_ragged(parts, h, rank, world):
    for q in range(world):
        if q == rank:
            parts[q].copy_(h)
        dist.broadcast(parts[q], q)


if __name__ == "__main__":
    main()
