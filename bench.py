#!/usr/bin/env python
"""bench.py -- all-vs-all pairwise structure alignment throughput (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU implementation of the same path on the host cores

A "step" is one full all-vs-all pass (every pair i<j: RBF score tiles -> Smith-Waterman + traceback -> Kabsch ->
RBF -> Smith-Waterman score) over the synthetic chain set.  Workload at N GPUs: round(1000*sqrt(N)) chains x 300
residues (BASELINE config 3 per GPU; weak scaling: ~499 500 pairs per GPU), d = 10, fp32 production mode.
Pairs are sharded by cost over the ranks (no data-path collective); at N > 1 the step ends with the single NCCL
all-gather of the packed score / RMSD / TM vectors.

The JSON line: value = pairs/s with the chains resident in HBM, device-timed (CUDA events on the engine's stream,
max over ranks); e2e = the same through the reference-facing call (crt_set_chains + crt_pairwise_all behind
make_pairwise_matrix) with page-locked HOST buffers: H2D of the packed float64 chains and D2H of the dense float64
[N,N] score / RMSD / TM matrices inside the timed region; roofline = the dominant kernel (stage-1 fill) against the FP32 FFMA peak measured
in the same run; cpu_baseline = the oracle port on the host cores on a bounded sample of the same workload.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L_CHAIN = 300
D_TENSOR = 10
GT, GC = 7.0, 0.03


def n_chains_for(gpus: int) -> int:
    return int(round(1000 * math.sqrt(gpus)))


def workload(gpus: int, n_override=None, l_override=None):
    from caretta_b200 import synth
    n = n_override or n_chains_for(gpus)
    L = l_override or L_CHAIN
    return synth.make_chains(n, L, D_TENSOR, seed=3), n, L


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       power_w_max=float(max(pw)) if pw else None, samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------- CPU arm
def cpu_port_rate(ch, n_pairs_sample: int, nthreads: int = 0, seed: int = 0):
    """Times the oracle port (oracle/caretta_oracle.c, OpenMP over pairs) on a random sample of pairs."""
    from oracle import oracle as O
    N = ch.n
    rng = np.random.default_rng(seed)
    pi = rng.integers(0, N - 1, size=n_pairs_sample)
    pj = np.array([rng.integers(i + 1, N) for i in pi])
    O.pairwise_list(ch.coords, ch.tensors, ch.offsets, pi[:8], pj[:8], GT, GC, nthreads, extras=False)    # warm
    t = time.perf_counter()
    O.pairwise_list(ch.coords, ch.tensors, ch.offsets, pi, pj, GT, GC, nthreads, extras=False)
    dt = time.perf_counter() - t
    return n_pairs_sample / dt, dt, O.num_threads() if nthreads <= 0 else nthreads


def run_reference(args):
    """--impl reference: the reference is pure Python + numba and cannot travel to the GPU box, so this arm times the
    oracle port of its path (kind = "port") with all host threads on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ch, n, L = workload(args.gpus, args.chains, args.length)
    cores = os.cpu_count() or 1
    rate0, _, _ = cpu_port_rate(ch, 64, nthreads=cores)
    sample = int(max(64, min(20000, rate0 * 4.0)))          # ~4 s per step
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_port_rate(ch, max(64, sample // 4), nthreads=cores)
    rates, t0 = [], time.perf_counter()
    for k in range(args.steps):
        r, dt, thr = cpu_port_rate(ch, sample, nthreads=cores, seed=k + 1)
        rates.append(r)
    total = time.perf_counter() - t0
    value = float(np.mean(rates))
    cells = 2.0 * L * L
    line = {
        "impl": "reference", "metric": "all-vs-all pairs/sec", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "gcups": value * cells / 1e9,
        "config": {"workload": f"synthetic {n} chains x {L} residues all-vs-all, d={D_TENSOR}, random sample of {sample} pairs per step",
                   "gamma_tensor": GT, "gamma_coords": GC},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": thr, "kind": "port",
                         "sample": f"{sample} random pairs per step x {args.steps} steps, OpenMP over pairs, {cores} logical cores"},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chains", type=int, default=None, help="override the number of chains (debug)")
    ap.add_argument("--length", type=int, default=None, help="override the chain length (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    from caretta_b200 import engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    ch, n, L = workload(args.gpus, args.chains, args.length)
    eng = engine.Engine(local)
    prm = eng.params(GT, GC, engine.FP32)
    eng.set_chains(ch.coords, ch.tensors, ch.offsets)
    total_pairs = n * (n - 1) // 2
    my_pairs = eng.shard_size(rank, world)
    sizes = [eng.shard_size(r, world) for r in range(world)]
    pad = max(sizes)

    # NCCL all-gather buffers (float32 packed vectors: score | rmsd | tm), equal padded counts per rank
    d_mine = torch.zeros(3 * pad, dtype=torch.float32, device="cuda")
    d_all = torch.zeros(world * 3 * pad, dtype=torch.float32, device="cuda") if world > 1 else None
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")          # > 126 MB L2
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0

    def step():
        """One all-vs-all pass with resident inputs; returns device ms (engine events + gather events)."""
        nonlocal launches
        eng.pairwise_shard(prm, rank, world)
        ms = eng.last_elapsed_ms()
        launches += eng.last_launches()
        if world > 1:
            ev0.record()
            eng.fetch_device(d_mine.data_ptr(), d_mine.data_ptr() + 4 * pad, d_mine.data_ptr() + 8 * pad, pad)
            dist.all_gather_into_tensor(d_all, d_mine)
            ev1.record()
            torch.cuda.synchronize()
            ms += ev0.elapsed_time(ev1)
            launches += 3
        return ms

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local)
    sync_all()
    if rank == 0:
        sampler.start()
    launches = 0
    dev_ms = 0.0
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        l2_flush.zero_()                      # flush L2 between timed iterations (not part of the step's device time)
        torch.cuda.synchronize()
        dev_ms += step()
    sync_all()
    wall_ms = 1e3 * (time.perf_counter() - t_wall)
    clocks = sampler.stop() if rank == 0 else {}
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    value = total_pairs / (ms_per_step * 1e-3)
    cells_per_step = 2.0 * L * L * total_pairs          # DP cell updates: two SW fills per residue pair

    # ------------------------------------------------------------------ e2e: host buffers in, host results out
    # N = 1: the reference-facing call (make_pairwise_matrix -> crt_set_chains + crt_pairwise_all): packed float64
    # chains in page-locked host memory in, dense float64 [N,N] score / RMSD / TM matrices in page-locked host memory out.
    # N > 1: upload, shard, all-gather of the packed float32 vectors, scatter into the dense matrices on the GPU, D2H of
    # the three dense [N,N] float64 matrices on every rank (caretta_b200.distributed.all_vs_all).
    pin_c, pin_t, pin_o = engine.pinned_like(ch.coords), engine.pinned_like(ch.tensors), engine.pinned_like(ch.offsets)
    dense = tuple(engine.pinned_empty((n, n)) for _ in range(3)) if world == 1 else None
    h_dense = torch.empty(3, n, n, dtype=torch.float64).pin_memory() if world > 1 else None
    if world > 1:
        from caretta_b200 import distributed as D
    e2e_steps = max(1, min(args.steps, 3))

    def e2e_step():
        eng.set_chains(pin_c, pin_t, pin_o)                                     # H2D of this step's inputs
        if world > 1:
            D.all_vs_all(eng, prm, rank, world, out=h_dense)                    # shard, all-gather, device scatter, D2H
        else:
            eng.pairwise_all(prm, want_rmsd_tm=True, out=dense)                 # D2H: three dense [N,N] float64 matrices

    e2e_step()                                                                  # untimed: first touch of the new buffers
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    sync_all()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    h2d = ch.coords.nbytes + ch.tensors.nbytes + ch.offsets.nbytes
    d2h = 3 * n * n * 8                      # per rank: every rank ends with the dense score / RMSD / TM matrices on its host

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ roofline of the dominant kernel (rank 0)
    peak_ffma, _ = eng.fp32_peak()
    os.environ["CARETTA_B200_STREAMS"] = "1"          # serial pass: per-kernel event timing is only meaningful unoverlapped
    eng.pairwise_shard(prm, rank, world)
    eng.pairwise_shard(prm, rank, world)
    ph = eng.last_phase_ms()
    serial_ms = eng.last_elapsed_ms()
    n_batches = eng.last_launches() // 4
    os.environ.pop("CARETTA_B200_STREAMS")
    my_cells = eng.last_cell_updates() / 2.0          # (a,b) residue pairs of this rank's shard
    W1 = 2 * D_TENSOR + 1 + 4                        # stage-1 RBF (2d+1) + SW cell (4) lane-instr per (a,b)
    W_ALL = 2 * D_TENSOR + 16
    fill1_s = ph["fill1"] * 1e-3
    achieved = my_cells * W1 / fill1_s / 1e12
    peak = peak_ffma / 1e12
    sm_clock = clocks.get("sm_mhz") or 0.0
    roofline = {
        "bound": "fp32", "kernel": "k_fill1_v3 (stage-1 Smith-Waterman fill + traceback bits)",
        "achieved": achieved, "peak": peak, "unit": "T lane-instr/s", "frac": achieved / peak,
        "peak_source": "FFMA micro-benchmark in this run (crt_fp32_peak); MEASURED_PEAKS.json has no FP32-pipe entry; "
                       "nominal 148 SM x 128 lanes x 1.965 GHz = 37.2",
        "algorithmic_work": f"{W1} lane-instr per residue pair (2d+1 RBF + 4 SW), {my_cells:.4g} residue pairs over {n_batches} launches",
        "avg_launch_ms": ph["fill1"] / max(n_batches, 1),
        "serial_step_ms": serial_ms, "phase_ms": ph,
        "whole_step_frac": (my_cells * W_ALL / (ms_per_step * 1e-3)) / peak_ffma,
        # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel: 2.422 GB for a launch of
        # 5.740e9 residue pairs (profiles/r01_ncu_summary.md, r1d) = 0.422 B per residue pair, scaled to this run's launches
        "traffic": 0.422 * my_cells / max(n_batches, 1),
        "traffic_unit": "bytes per launch (ncu-measured bytes per residue pair x residue pairs per launch)",
        "hbm": {"peak_gbs": _measured_hbm(), "achieved_gbs": 0.422 * my_cells / fill1_s / 1e9,
                "note": "path is FP32-issue bound; HBM traffic is the 0.4 B/cell traceback stream, equal to the algorithmic bytes"},
    }

    # ------------------------------------------------------------------ CPU baseline (rank 0, bounded sample)
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        ncpu = os.cpu_count() or 1
        rate0, _, _ = cpu_port_rate(ch, 64, nthreads=ncpu)
        sample = int(max(64, min(40000, rate0 * 12.0)))      # ~12 s of all-core CPU work
        rate, dt, thr = cpu_port_rate(ch, sample, nthreads=ncpu, seed=7)
        cpu = {"value": rate, "unit": "pairs/s", "cores": thr, "kind": "port",
               "sample": f"{sample} random pairs of the same {n}x{L} workload in {dt:.1f} s, oracle/caretta_oracle.c with OpenMP over pairs"}

    line = {
        "metric": "all-vs-all pairs/sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "gcups": cells_per_step / (ms_per_step * 1e-3) / 1e9,
        "config": {"workload": f"synthetic {n} chains x {L} residues all-vs-all ({total_pairs} pairs), d={D_TENSOR} "
                               f"(BASELINE config 3 per GPU), fp32 production mode",
                   "gamma_tensor": GT, "gamma_coords": GC, "pairs_per_gpu": sizes,
                   "l2": "256 MiB buffer written between timed iterations (L2 flush); per-step traceback stream >> L2",
                   "timing": "CUDA events on the engine's stream per step (+ torch events around the all-gather), max over ranks"},
        "wall_ms_per_step": wall_ms / args.steps,
        "clocks": {k: clocks.get(k) for k in ("sm_mhz", "sm_max_mhz", "reasons", "power_w_max")},
        "e2e": {"value": total_pairs / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1e3 * e2e_s, "steps": e2e_steps},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _measured_hbm():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs")
    except Exception:
        return None


if __name__ == "__main__":
    main()
