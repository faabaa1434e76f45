"""CPU, world_size 2 over gloo: the host-side logic of bench.py's N>1 path.  On the GPUs the map is sharded
into x-strips and the data path is peer memory inside the sweep kernel; what the host
adds at N>1 is the aggregation - device time is the MAX over ranks, particle-steps the SUM, column checksums
add up modulo 2^64 - and the rank-0-only reference arm.  The sharded kernels themselves are covered on the GPU
tier (virtual ranks and two CUDA-IPC processes on one GPU, tests/test_gpu_parity.py)."""
import json
import os
import subprocess
import sys
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    ev, e2e, steps, esteps = bench.aggregate(100.0 + 50.0 * rank, 200.0 - 30.0 * rank, 1000 * (rank + 1), 10 * (rank + 1))
    # strip checksums add up modulo 2^64 (one of them above 2^63, the sum wraps)
    cs = bench.checksum_sum([0xF000000000000001, 0x2000000000000005][rank])
    if rank == 0:
        with open(out, "w") as f:
            json.dump([ev, e2e, steps, esteps, cs], f)
    dist.barrier()
    dist.destroy_process_group()


def test_aggregate_max_time_sum_steps(tmp_path):
    out = str(tmp_path / "agg.json")
    mp.spawn(_worker, args=(2, 29531, out), nprocs=2, join=True)
    ev, e2e, steps, esteps, cs = json.load(open(out))
    assert ev == 150.0 and e2e == 200.0          # max over ranks
    assert steps == 3000.0 and esteps == 30.0    # sum over ranks
    assert cs == (0xF000000000000001 + 0x2000000000000005) % (1 << 64)


def test_reference_arm_only_rank0_prints():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--dim", "64", "--particles", "50", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
