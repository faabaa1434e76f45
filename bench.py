#!/usr/bin/env python
"""bench.py -- particle-steps/s of the particle/terrain hot path on B200 (BASELINE.json metric).

Workload (config 3 of BASELINE.json, the one the metric is quoted on that fits one GPU):
  4096^2 map, rockgravelpebblessand preset, SEED 42, per step (= one frame, SoilMachine.cpp:287-320
  without flood/seep): 25 000 WaterParticles run to completion in lockstep sweeps, then 25 000
  WindParticles run to completion, then the frequency update.  particle-step = one move() followed by
  one interact() of one particle.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

* value  : whole-job particle-steps/s, spawn lists already resident in HBM, timed with CUDA events on
           the stream the kernels run on (max over ranks).
* e2e    : the same frames through the host-buffer C ABI calls (sm_water_run / sm_wind_run): the
           spawn lists are copied from pinned host memory and the stats are read back inside the
           timed region, every step.
* roofline: algorithmic bytes (290 B / water step, 600 B / wind step) / duration of the
           dominant kernel (CUDA events around each launch), against the measured HBM peak.
* cpu_baseline / --impl reference: the reference's own CPU loop (oracle/_ref, the reference headers
           compiled verbatim) on a bounded sample of the same workload, on the box's host cores
           (1 thread: the reference is single-threaded and keeps its state in globals).
N > 1: the SAME 4096^2 map is sharded into N x-strips, one rank (GPU) per strip; halo records, bins and
hand-off words of the neighbouring strips are read and written over NVLink peer memory (CUDA IPC) by the sweep
kernel itself, sweeps are synchronised by a cross-GPU flag barrier and particles that leave a strip are handed
to the new owner.  Strong scaling: total work is fixed, results are bit-identical to N = 1.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(soil="rockgravelpebblessand", dim=4096, nwater=25000, nwind=25000, seed=42)
BYTES_WATER, BYTES_WIND, BYTES_FREQ_CELL = 290, 600, 16   # SURVEY.md section 8d / DESIGN.md section 5
CPU_SAMPLE = dict(nwater=2500, nwind=750)                  # bounded sample for the CPU arm, per step


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.rows = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except Exception:
                pass
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for k, nm in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def aggregate(ev_ms, e2e_ms, steps, e_steps, device=None):
    """Whole-job numbers from per-rank ones: time = MAX over ranks, particle-steps = SUM over ranks.
    Works on any initialised torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([ev_ms, e2e_ms, float(steps), float(e_steps)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm_ = t.clone(); dist.all_reduce(sm_, op=dist.ReduceOp.SUM)
        return mx[0].item(), mx[1].item(), sm_[2].item(), sm_[3].item()
    return float(ev_ms), float(e2e_ms), float(steps), float(e_steps)


def cpu_arm(steps, warmup, sample):
    """The reference's CPU loop on a bounded sample of the workload.  Returns (value, info)."""
    from oracle import refapi
    W = WORKLOAD
    r = refapi.get().init(W["soil"], seed=W["seed"], dimx=W["dim"], dimy=W["dim"], poolsize=int(W["dim"] * W["dim"] * 2 + 2000000))
    r.lib.smref_srand(W["seed"])
    tot_steps, tot_s = 0, 0.0
    per = []
    for it in range(warmup + steps):
        sw = r.water_seq(sample["nwater"])          # exactly SoilMachine.cpp:288-298 minus flood
        sd = r.wind_seq(sample["nwind"])            # SoilMachine.cpp:304-307
        t0 = time.perf_counter()
        r.frequency_update()                        # SoilMachine.cpp:313-320
        tf = time.perf_counter() - t0
        if it >= warmup:
            tot_steps += sw.steps + sd.steps
            tot_s += sw.seconds + sd.seconds + tf
            per.append((sw.steps, sw.seconds, sd.steps, sd.seconds))
    val = tot_steps / tot_s if tot_s > 0 else 0.0
    info = {"value": val, "unit": "particle-steps/s", "cores": 1, "kind": "reference",
            "sample": "%d water + %d wind particles per step on the same %d^2 %s map, sequential reference "
                      "loop (SoilMachine.cpp:287-320 without flood/seep), %d steps" %
                      (sample["nwater"], sample["nwind"], W["dim"], W["soil"], steps),
            "water_steps_per_s": sum(p[0] for p in per) / max(sum(p[1] for p in per), 1e-12),
            "wind_steps_per_s": sum(p[2] for p in per) / max(sum(p[3] for p in per), 1e-12),
            "host_cpus": os.cpu_count()}
    return val, info, tot_s / max(steps, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--dim", type=int, default=0, help="debug: override the map edge")
    ap.add_argument("--particles", type=int, default=0, help="debug: override particles per kind")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    W = dict(WORKLOAD)
    if args.dim:
        W["dim"] = args.dim
    if args.particles:
        W["nwater"] = W["nwind"] = args.particles
    WORKLOAD.update(W)
    K, Wm = args.steps, max(args.warmup, 0)
    config = {"workload": "4096^2 rockgravelpebblessand.soil, 25k Water + 25k Wind particles per step "
                          "(BASELINE.json configs[2])" if not (args.dim or args.particles) else
                          "DEBUG %d^2 %s %d+%d" % (W["dim"], W["soil"], W["nwater"], W["nwind"]),
              "map": "%dx%d" % (W["dim"], W["dim"]), "soil": W["soil"], "seed": W["seed"],
              "water_per_step": W["nwater"], "wind_per_step": W["nwind"],
              "step": "one frame = water batch + wind batch + frequency update, lockstep sweeps",
              "l2": "inputs larger than L2 (column records 0.5 GB + pool)",
              "parallelism": "single GPU" if world == 1 else
              "map sharded into %d x-strips, one per GPU; peer-memory halo/hand-off over NVLink, cross-GPU flag barrier per sweep" % world}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        sample = dict(CPU_SAMPLE)
        if args.particles:
            sample = dict(nwater=max(args.particles // 10, 1), nwind=max(args.particles // 30, 1))
        val, info, s_per_step = cpu_arm(K, min(Wm, 1), sample)
        line = {"impl": "reference", "metric": "particle-steps/sec", "value": val, "unit": "particle-steps/s",
                "n_gpus": args.gpus, "steps": K, "warmup": min(Wm, 1), "ms_per_step": s_per_step * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config, "cpu_baseline": info,
                "e2e": {"value": val, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm (GPU)
    import numpy as np
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from soilmachine_b200 import host

    if world == 1:
        sim = host.Simulation(W["soil"], seed=W["seed"], dimx=W["dim"], dimy=W["dim"], device=local_rank,
                              max_particles=max(W["nwater"], W["nwind"]))
        ctx = sim.ctx
    else:
        from soilmachine_b200 import presets, sharded
        pre = presets.load(W["soil"])
        sim = sharded.DistShard(W["dim"], W["dim"], pre["world"]["scale"], device=local_rank,
                                max_particles=max(W["nwater"], W["nwind"]))
        ctx = sim.ctx
        ctx.set_soils(pre["soils"])
        ctx.initialize(W["seed"], pre["layers"])
    nframes = Wm + 2 * K                      # warm-up, K device-resident frames, K end-to-end frames
    host.srand(W["seed"])                     # every rank draws the same spawn lists
    lists = [(host.spawn_list(W["nwater"], W["dim"], W["dim"]), host.spawn_list(W["nwind"], W["dim"], W["dim"]))
             for _ in range(nframes)]
    pinned = [(torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()) for a, b in lists]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up frames (these also advance the simulation: every frame erodes the same map further)
    for f in range(Wm):
        ctx.water_run(lists[f][0]); ctx.wind_run(lists[f][1]); ctx.frequency_update()
    ctx.sync()

    # ---- value: spawn lists resident in HBM ----
    dev = [(ctx.device_spawn(a), ctx.device_spawn(b)) for a, b in lists[Wm:Wm + K]]
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = ctx.launch_count()
    steps_w = steps_d = 0
    ms_w = ms_d = 0.0
    sweeps_w = sweeps_d = 0
    ctx.timer_start()
    t0 = time.perf_counter()
    for k in range(K):
        ctx.water_run_device(dev[k][0], W["nwater"])
        sw = ctx.last_stats()
        ctx.wind_run_device(dev[k][1], W["nwind"])
        sd = ctx.last_stats()
        ctx.frequency_update()
        steps_w += sw.steps; steps_d += sd.steps
        ms_w += sw.device_ms; ms_d += sd.device_ms
        sweeps_w += sw.sweeps; sweeps_d += sd.sweeps
    ev_ms = ctx.timer_stop()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop()
    for a, b in dev:
        ctx.device_free(a); ctx.device_free(b)

    # ---- e2e: host buffers through the public C ABI calls ----
    barrier()
    t0 = time.perf_counter()
    e_steps = 0
    for k in range(K):
        a, b = pinned[Wm + K + k]
        sw = ctx.water_run(a.numpy()); sd = ctx.wind_run(b.numpy()); ctx.frequency_update()
        e_steps += sw.steps + sd.steps
    ctx.sync()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    h2d = (W["nwater"] + W["nwind"]) * 8
    d2h = 2 * 424                                  # two RunCtl read-backs per step (sizeof(RunCtl) = 424)

    # ---- reported beside the metric, outside both timed regions: the rest of the water part of the frame
    # (the batch's floods and the seep pass, SoilMachine.cpp:296,300-301) on the map the frames above left.
    # BASELINE.json's metric excludes flood, so none of this enters `value` or `e2e`.
    hydrology = None
    if world == 1:
        try:
            sw = ctx.water_run(lists[-1][0])
            hf = ctx.water_flood()
            hs = ctx.seep()
            hydrology = {"water_batch_ms": sw.device_ms, "stalled": sw.exit_stall,
                         "flood_ms": hf.device_ms, "floods": hf.floods, "nested_particles": hf.nested,
                         "flood_transfers": hf.transfers,
                         "seep_ms": hs.device_ms, "seep_cells_visited": hs.cells, "seep_transfers": hs.transfers,
                         "cells": W["dim"] * W["dim"]}
        except Exception as e:                     # never let the extra report break the bench line
            hydrology = {"error": str(e)[:200]}

    # max over ranks / sums over ranks
    ev_ms, e2e_ms, tot_steps, tot_e = aggregate(ev_ms, e2e_ms, steps_w + steps_d, e_steps, device="cuda")
    value = tot_steps / (ev_ms * 1e-3)
    e2e_val = tot_e / (e2e_ms * 1e-3)

    # ---- roofline of the dominant kernel (rank 0's launches) ----
    peak, peak_src = hbm_peak()
    kern = [("k_run<wind>", steps_d, ms_d, BYTES_WIND, sweeps_d), ("k_run<water>", steps_w, ms_w, BYTES_WATER, sweeps_w)]
    kern.sort(key=lambda x: -x[2])
    kname, ksteps, kms, kbytes, ksweeps = kern[0]
    achieved = (ksteps * kbytes / K) / (kms / K * 1e-3) / 1e9 if kms > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(kname)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ksteps * kbytes / K, "launch_ms": kms / K,
                "share_of_step": kms / max(ms_w + ms_d, 1e-9),
                "note": "latency/dependency-bound path: %d dependent sweeps per launch, ~%d B touched per particle-step"
                        % (ksweeps // max(K, 1), kbytes),
                "kernels": {n: {"ms_per_step": m / K, "particle_steps_per_step": s / K, "sweeps_per_step": sw / K}
                            for n, s, m, _, sw in kern}}

    line = {"metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world,
            "steps": K, "warmup": Wm, "ms_per_step": ev_ms / K, "higher_is_better": True,
            "scaling": "weak" if world == 1 else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks,
            "scaling_note": None if world == 1 else "strong: one 4096^2 simulation over %d GPUs" % world,
            "e2e": {"value": e2e_val, "unit": "particle-steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / K},
            "gpu_launches": launches, "roofline": roofline, "wall_ms_per_step": wall_ms / K}
    if hydrology is not None:
        line["hydrology"] = hydrology

    if rank == 0 and world == 1 and not args.no_cpu:
        sample = dict(CPU_SAMPLE)
        if args.particles:
            sample = dict(nwater=max(args.particles // 10, 1), nwind=max(args.particles // 30, 1))
        _, info, _ = cpu_arm(1, 0, sample)
        line["cpu_baseline"] = info
    if rank == 0:
        print(json.dumps(line))
    sim.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
