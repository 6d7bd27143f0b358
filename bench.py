#!/usr/bin/env python
"""bench.py -- all-vs-all pairwise structure alignment throughput (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's own numba path on the host cores

A "step" is one full all-vs-all pass (every pair i<j: RBF score tiles -> Smith-Waterman + traceback -> Kabsch ->
RBF -> Smith-Waterman score) over the synthetic chain set, fp32 production mode WITH its tie detection: pairs whose fp32
traceback meets a decision the reference's float64 DP may take differently are recomputed by the float64 kernels inside the
step (crt_fill1_v4.cuh), so the timed matrix is the one that meets the 1e-4 parity bound on every pair.

Workload of `value` at N GPUs: round(1000*sqrt(N)) chains x 300 residues (BASELINE config 3 per GPU; weak scaling:
~499 500 pairs per GPU), d = 10.  Pairs are sharded by cost over the ranks (no data-path collective); at N > 1 the step ends
with the single NCCL all-gather of the packed score | rmsd | tm vectors.

The JSON line:
  value        pairs/s with the chains resident in HBM, device-timed (CUDA events on the engine's stream, max over ranks)
  e2e          the same through the reference-facing call with page-locked HOST buffers: H2D of the packed float64 chains and
               D2H of the dense float64 [N,N] score / RMSD / TM matrices inside the timed region (N > 1: on rank 0, which alone
               builds the dense matrices)
  roofline     the dominant kernel (stage-1 fill) against the FP32 FFMA peak measured in the same run
  parity       the matrices the e2e steps produced, checked in this run against the oracle port on a random sample of pairs
               and against the UNMODIFIED numba reference on the pairs the numba timing computed
  cpu_baseline the oracle port (C + OpenMP) on the host cores; cpu_baseline_reference: the unmodified numba reference
               (oracle/_ref, NUMBA_NUM_THREADS = cores) as shipped (one core) and fanned out over all cores
  configs      the other contract configurations, device-timed in the same run: N = 1: C5 (500 x 1500) and a 1000-chain subset
               of C4 (lengths 50-1000); N > 1: strong-scaled C3 (1000 chains on N GPUs); N = 8: the target T (5000 x 300)
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L_CHAIN = 300
D_TENSOR = 10
GT, GC = 7.0, 0.03
W1 = 2 * D_TENSOR + 1 + 4          # stage-1 RBF (2d+1) + SW cell (4) lane-instr per residue pair
W_ALL = 2 * D_TENSOR + 16          # both stages


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


# ---------------------------------------------------------------------------------------------- CPU legs
def random_pairs(n: int, k: int, seed: int):
    rng = np.random.default_rng(seed)
    pi = rng.integers(0, n - 1, size=k)
    pj = np.array([rng.integers(i + 1, n) for i in pi])
    return pi, pj


def cpu_port_rate(ch, n_pairs_sample: int, nthreads: int = 0, seed: int = 0):
    """Times the oracle port (oracle/caretta_oracle.c, OpenMP over pairs) on a random sample of pairs."""
    from oracle import oracle as O
    pi, pj = random_pairs(ch.n, n_pairs_sample, seed)
    O.pairwise_list(ch.coords, ch.tensors, ch.offsets, pi[:8], pj[:8], GT, GC, nthreads, extras=False)    # warm
    t = time.perf_counter()
    O.pairwise_list(ch.coords, ch.tensors, ch.offsets, pi, pj, GT, GC, nthreads, extras=False)
    dt = time.perf_counter() - t
    return n_pairs_sample / dt, dt, O.num_threads() if nthreads <= 0 else nthreads


def numba_reference(ch, fan_pairs: int, shipped_chains: int, seed: int = 11):
    """The unmodified reference (oracle/_ref): as shipped (its own serial driver, one core) and the loop body fanned out over all
    cores.  Returns (record, (pi, pj, scores)) or (None, None) when numba / the reference copy is absent."""
    try:
        from oracle import ref_timing as RT
        if not RT.available():
            return {"unavailable": "oracle/_ref (build output of oracle/build_ref.py) or numba is missing"}, None
        cores = os.cpu_count() or 1
        _, jit_s = RT.load(cores)
        S, dt1 = RT.as_shipped(ch, shipped_chains)
        n1 = shipped_chains * (shipped_chains - 1) // 2
        pi, pj = random_pairs(ch.n, fan_pairs, seed)
        sc, dtn, nproc = RT.pairs_fanout(ch, pi, pj, cores)
        rec = {"kind": "reference", "unit": "pairs/s", "value": fan_pairs / dtn, "cores": nproc, "numba_threads": RT.numba_threads(),
               "as_shipped_1core": n1 / dt1,
               "sample": f"unmodified TurtleTools/caretta (numba {_numba_version()}) from oracle/_ref: loop body of "
                         f"make_pairwise_matrix (multiple_alignment.py:164-169) on {fan_pairs} random pairs of the workload over "
                         f"{nproc} forked workers in {dtn:.1f} s; as shipped = MultipleAlignment.make_pairwise_matrix on the first "
                         f"{shipped_chains} chains ({n1} pairs, serial, {dt1:.1f} s); JIT warm-up {jit_s:.0f} s excluded; "
                         f"NUMBA_NUM_THREADS={RT.numba_threads()} (feeds no function of this path)"}
        return rec, (pi, pj, sc, S)
    except Exception as e:                                   # the reference leg must never take the GPU numbers down with it
        return {"unavailable": f"{type(e).__name__}: {e}"}, None


def _numba_version():
    try:
        import numba
        return numba.__version__
    except Exception:
        return "?"


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores.  With the unmodified reference
    present (oracle/_ref) that is its numba code, the loop body fanned out over all cores; otherwise the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ch, n, L = workload(args.gpus, args.chains, args.length)
    cores = os.cpu_count() or 1
    kind, extra = "port", {}
    rates, t0 = [], None
    use_numba = False
    try:
        from oracle import ref_timing as RT
        use_numba = RT.available() and not args.port
    except Exception:
        use_numba = False
    if use_numba:
        try:
            _, jit_s = RT.load(cores)
            sc, dt, nproc = RT.pairs_fanout(ch, *random_pairs(n, 2 * cores, 1), cores)
            sample = int(max(2 * cores, min(4000, (2 * cores / dt) * 3.0)))           # ~3 s per step
            for _ in range(1 if args.warmup else 0):
                RT.pairs_fanout(ch, *random_pairs(n, max(cores, sample // 4), 2), cores)
            t0 = time.perf_counter()
            for k in range(args.steps):
                sc, dt, nproc = RT.pairs_fanout(ch, *random_pairs(n, sample, 100 + k), cores)
                rates.append(sample / dt)
            kind = "reference"
            thr = nproc
            how = (f"unmodified TurtleTools/caretta numba path from oracle/_ref: {sample} random pairs per step x {args.steps} steps, loop "
                   f"body of make_pairwise_matrix over {nproc} forked workers, NUMBA_NUM_THREADS={RT.numba_threads()}, JIT warm-up "
                   f"{jit_s:.0f} s excluded")
        except Exception as e:
            rates, use_numba = [], False
            extra["numba_error"] = f"{type(e).__name__}: {e}"
    if not use_numba:
        rate0, _, _ = cpu_port_rate(ch, 64, nthreads=cores)
        sample = int(max(64, min(20000, rate0 * 4.0)))          # ~4 s per step
        for _ in range(1 if args.warmup else 0):
            cpu_port_rate(ch, max(64, sample // 4), nthreads=cores)
        t0 = time.perf_counter()
        for k in range(args.steps):
            r, dt, thr = cpu_port_rate(ch, sample, nthreads=cores, seed=k + 1)
            rates.append(r)
        how = f"oracle port (oracle/caretta_oracle.c): {sample} random pairs per step x {args.steps} steps, OpenMP over pairs, {cores} logical cores"
    total = time.perf_counter() - t0
    value = float(np.mean(rates))
    cells = 2.0 * L * L
    line = {
        "impl": "reference", "metric": "all-vs-all pairs/sec", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "gcups": value * cells / 1e9,
        "config": {"workload": f"synthetic {n} chains x {L} residues all-vs-all, d={D_TENSOR}, random sample of {sample} pairs per step "
                               f"(same_config: the same chain set, a bounded sample of its pairs)",
                   "gamma_tensor": GT, "gamma_coords": GC},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": thr, "kind": kind, "sample": how, **extra},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- parity gate
def parity_check(ch, S, R, T, n_sample: int, ref):
    """The matrices of the timed e2e steps against the oracle port on a random sample of pairs (score / RMSD / TM, 1e-4) and
    against the scores the unmodified numba reference computed in this run."""
    from oracle import oracle as O
    pi, pj = random_pairs(ch.n, n_sample, 12345)
    o = O.pairwise_list(ch.coords, ch.tensors, ch.offsets, pi, pj, GT, GC, 0, extras=True)
    rel = np.abs(S[pi, pj] - o["score"]) / np.maximum(np.abs(o["score"]), 1e-300)
    ok_r = np.isclose(R[pi, pj], o["rmsd"], rtol=1e-4, atol=1e-6)
    ok_t = np.isclose(T[pi, pj], o["tm"], rtol=1e-4, atol=1e-9)
    out = {"checked_against": "oracle port (pinned on the reference's golden vectors)", "pairs": int(n_sample), "tolerance": 1e-4,
           "max_rel_score": float(rel.max()), "score_outside": int((rel > 1e-4).sum()), "rmsd_outside": int((~ok_r).sum()),
           "tm_outside": int((~ok_t).sum()), "symmetric": bool(np.array_equal(S, S.T)), "zero_diagonal": bool(np.all(np.diag(S) == 0))}
    if ref is not None:
        rpi, rpj, rsc, Sship = ref
        rr = np.abs(S[rpi, rpj] - rsc) / np.maximum(np.abs(rsc), 1e-300)
        k = Sship.shape[0]
        iu = np.triu_indices(k, 1)
        rs = np.abs(S[:k, :k][iu] - Sship[iu]) / np.maximum(np.abs(Sship[iu]), 1e-300)
        out["reference_numba"] = {"pairs": int(len(rpi) + len(iu[0])), "max_rel_score": float(max(rr.max(), rs.max())),
                                  "score_outside": int((rr > 1e-4).sum() + (rs > 1e-4).sum())}
    out["ok"] = bool(out["score_outside"] == 0 and out["rmsd_outside"] == 0 and out["tm_outside"] == 0 and out["symmetric"]
                     and out["zero_diagonal"] and (ref is None or out["reference_numba"]["score_outside"] == 0))
    return out


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
    ap.add_argument("--no-configs", action="store_true", help="skip the sub-records of the other contract configurations")
    ap.add_argument("--port", action="store_true", help="--impl reference: time the oracle port even when the numba reference is present")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    from caretta_b200 import engine, synth

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
        from caretta_b200 import distributed as D

    ch, n, L = workload(args.gpus, args.chains, args.length)
    eng = engine.Engine(local)
    prm = eng.params(GT, GC, engine.FP32)
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")          # > 126 MB L2
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    class Resident:
        """One chain set resident in HBM: a step = this rank's shard (+ the all-gather at N > 1), device-timed."""

        def __init__(self, chains):
            eng.set_chains(chains.coords, chains.tensors, chains.offsets)
            self.sizes = [eng.shard_size(r, world) for r in range(world)]
            self.pad = max(max(self.sizes), 1)
            self.d_mine = torch.zeros(3 * self.pad, dtype=torch.float32, device="cuda")
            self.d_all = torch.zeros(world * 3 * self.pad, dtype=torch.float32, device="cuda") if world > 1 else None
            self.launches = 0
            self.rerun = 0

        def step(self) -> float:
            eng.pairwise_shard(prm, rank, world)
            ms = eng.last_elapsed_ms()
            self.launches += eng.last_launches()
            self.rerun = eng.last_rerun()[0]
            if world > 1:
                ev0.record()
                eng.pack_results(self.d_mine.data_ptr(), self.pad, False)
                dist.all_gather_into_tensor(self.d_all, self.d_mine)
                ev1.record()
                torch.cuda.synchronize()
                ms += ev0.elapsed_time(ev1)
                self.launches += 1
            return ms

        def timed(self, steps: int, warm: int):
            for _ in range(warm):
                self.step()
            sync_all()
            self.launches = 0
            dev_ms = 0.0
            t = time.perf_counter()
            for _ in range(steps):
                l2_flush.zero_()                      # flush L2 between timed iterations (not part of the step's device time)
                torch.cuda.synchronize()
                dev_ms += self.step()
            sync_all()
            wall = 1e3 * (time.perf_counter() - t)
            return max_over_ranks(dev_ms) / max(steps, 1), wall / max(steps, 1)

    # ------------------------------------------------------------------ headline: weak-scaled C3, resident inputs
    total_pairs = n * (n - 1) // 2
    res = Resident(ch)
    res.timed(0, max(args.warmup, 3))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_per_step, wall_ms = res.timed(args.steps, 0)
    clocks = sampler.stop() if rank == 0 else {}
    launches = res.launches
    rerun_pairs = int(sum_over_ranks(res.rerun))
    sizes = res.sizes
    value = total_pairs / (ms_per_step * 1e-3)
    cells_per_step = 2.0 * L * L * total_pairs          # DP cell updates: two SW fills per residue pair

    # ------------------------------------------------------------------ e2e: host buffers in, host results out
    # N = 1: the reference-facing call (make_pairwise_matrix -> crt_set_chains + crt_pairwise_all): packed float64 chains in
    # page-locked host memory in, dense float64 [N,N] score / RMSD / TM matrices in page-locked host memory out.
    # N > 1: every rank uploads the chains and computes its shard, one all-gather of the packed float32 vectors, then rank 0
    # alone scatters into the dense matrices on its GPU and copies them to the host (caretta_b200.distributed.all_vs_all).
    pin_c, pin_t, pin_o = engine.pinned_like(ch.coords), engine.pinned_like(ch.tensors), engine.pinned_like(ch.offsets)
    dense = tuple(engine.pinned_empty((n, n)) for _ in range(3)) if rank == 0 else None
    e2e_steps = max(10, min(args.steps, 20))

    def e2e_step():
        eng.set_chains(pin_c, pin_t, pin_o)                                     # H2D of this step's inputs
        if world > 1:
            D.all_vs_all(eng, prm, rank, world, out=dense, result_ranks=(0,))   # shard, all-gather; rank 0: scatter + D2H
        else:
            eng.pairwise_all(prm, want_rmsd_tm=True, out=dense)                 # D2H: three dense [N,N] float64 matrices

    e2e_step()                                                                  # untimed: first touch of the new buffers
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    sync_all()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    h2d = ch.coords.nbytes + ch.tensors.nbytes + ch.offsets.nbytes          # per rank (the chains are replicated)
    d2h = 3 * n * n * 8                                                      # rank 0 only

    # ------------------------------------------------------------------ sub-records: the other contract configurations
    peak_ffma, _ = eng.fp32_peak()
    configs = {}

    def sub_record(name, chains, steps=3, note=""):
        r = Resident(chains)
        ms, _ = r.timed(steps, 2)
        cells = float((np.diff(chains.offsets).sum() ** 2 - (np.diff(chains.offsets) ** 2).sum()) / 2.0)      # residue pairs
        npairs = chains.n * (chains.n - 1) // 2
        configs[name] = {"workload": note, "n_gpus": world, "pairs": int(npairs), "ms_per_step": ms, "pairs_per_s": npairs / (ms * 1e-3),
                         "gcups": 2.0 * cells / (ms * 1e-3) / 1e9, "fp32_roofline_frac_w36": cells * W_ALL / (ms * 1e-3) / (peak_ffma * world),
                         "rerun_pairs": int(sum_over_ranks(r.rerun)), "steps": steps, "timing": "device (CUDA events), max over ranks"}

    if not args.no_configs and not args.chains and not args.length:
        if world == 1:
            sub_record("c5", synth.config("C5"), 3, "BASELINE config 5: 500 chains x 1500 residues (124 750 pairs) on one GPU")
            c4 = synth.config("C4")
            e = int(c4.offsets[1000])
            sub_record("c4_subset", synth.Chains(c4.coords[:e], c4.tensors[:e], c4.offsets[:1001].copy()), 3,
                       "first 1000 chains of BASELINE config 4 (lengths 50-1000 mixed; 499 500 pairs) on one GPU")
            # experimental: stage 1 of the same C3 workload on the tensor-core kernel (k_fill1_tc, CARETTA_B200_TC=1)
            os.environ["CARETTA_B200_TC"] = "1"
            sub_record("c3_tensor_core_stage1", ch, 5, "the headline workload with stage 1 on tcgen05.mma (experimental, off by default: profiles/r02_tensor_core_stage1.md)")
            configs["c3_tensor_core_stage1"]["tc_pairs"] = int(eng.last_tc_pairs())
            os.environ["CARETTA_B200_TC"] = "0"
        else:
            sub_record("strong_c3", synth.config("C3"), 5, f"BASELINE config 3 strong-scaled: 1000 chains x 300 on {world} GPUs")
            if world >= 8:
                sub_record("target_T", synth.config("T"), 3, "north_star target: 5000 chains x 300 residues (12 497 500 pairs) on 8 GPUs")
                sub_record("c5", synth.config("C5"), 3, "BASELINE config 5: 500 chains x 1500 residues on 8 GPUs")

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ roofline of the dominant kernel (rank 0)
    eng.set_chains(ch.coords, ch.tensors, ch.offsets)
    os.environ["CARETTA_B200_STREAMS"] = "1"          # serial pass: per-kernel event timing is only meaningful unoverlapped
    eng.pairwise_shard(prm, rank, world)
    eng.pairwise_shard(prm, rank, world)
    ph = eng.last_phase_ms()
    serial_ms = eng.last_elapsed_ms()
    rr_pairs, rr_ms = eng.last_rerun()
    n_batches = max((eng.last_launches() - (6 if rr_pairs else 0)) // 4, 1)
    tb_bytes = eng.last_traceback_bytes()
    os.environ.pop("CARETTA_B200_STREAMS")
    my_cells = eng.last_cell_updates() / 2.0          # (a,b) residue pairs of this rank's shard
    fill1_s = ph["fill1"] * 1e-3
    achieved = my_cells * W1 / fill1_s / 1e12
    peak = peak_ffma / 1e12
    roofline = {
        "bound": "fp32", "kernel": "k_fill1_v4 (stage-1 Smith-Waterman fill + traceback codes + tie flags)",
        "achieved": achieved, "peak": peak, "unit": "T lane-instr/s", "frac": achieved / peak,
        "peak_source": "FFMA micro-benchmark in this run (crt_fp32_peak); MEASURED_PEAKS.json has no FP32-pipe entry; "
                       "nominal 148 SM x 128 lanes x 1.965 GHz = 37.2",
        "algorithmic_work": f"{W1} lane-instr per residue pair (2d+1 RBF + 4 SW; the tie flags are overhead, not counted), "
                            f"{my_cells:.4g} residue pairs over {n_batches} launches",
        "avg_launch_ms": ph["fill1"] / n_batches,
        "serial_step_ms": serial_ms, "phase_ms": ph, "rerun_ms": rr_ms, "rerun_pairs": rr_pairs,
        "whole_step_frac": (my_cells * W_ALL / (ms_per_step * 1e-3)) / peak_ffma,
        # HBM traffic of the kernel = its traceback stream, from the allocation: strips x chunks x 32 lanes x 16 B per unit
        # (3 bits per cell in 32-bit row words; row records and strip boundaries come from L2)
        "traffic": tb_bytes / n_batches,
        "traffic_unit": "bytes per launch, computed from the traceback allocation of the run (crt_last_traceback_bytes)",
        "traffic_per_residue_pair": tb_bytes / my_cells,
        "hbm": {"peak_gbs": _measured_hbm(), "achieved_gbs": tb_bytes / fill1_s / 1e9,
                "note": "path is FP32-issue bound; HBM traffic is the traceback stream"},
    }

    # ------------------------------------------------------------------ CPU baselines and the parity gate (rank 0, bounded samples)
    cpu = cpu_ref = None
    ref_scores = None
    if not args.no_cpu_baseline and world == 1:
        ncpu = os.cpu_count() or 1
        rate0, _, _ = cpu_port_rate(ch, 64, nthreads=ncpu)
        sample = int(max(64, min(40000, rate0 * 10.0)))      # ~10 s of all-core CPU work
        rate, dt, thr = cpu_port_rate(ch, sample, nthreads=ncpu, seed=7)
        cpu = {"value": rate, "unit": "pairs/s", "cores": thr, "kind": "port",
               "sample": f"{sample} random pairs of the same {n}x{L} workload in {dt:.1f} s, oracle/caretta_oracle.c with OpenMP over pairs"}
        cpu_ref, ref_scores = numba_reference(ch, fan_pairs=max(64, 8 * ncpu), shipped_chains=5)
    S, R, T = dense
    parity = parity_check(ch, S, R, T, 400 if world == 1 else 200, ref_scores)

    line = {
        "metric": "all-vs-all pairs/sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "gcups": cells_per_step / (ms_per_step * 1e-3) / 1e9,
        "config": {"workload": f"synthetic {n} chains x {L} residues all-vs-all ({total_pairs} pairs), d={D_TENSOR} "
                               f"(BASELINE config 3 per GPU), fp32 production mode with tie detection + float64 re-run of the marked pairs",
                   "gamma_tensor": GT, "gamma_coords": GC, "pairs_per_gpu": sizes, "fp64_rerun_pairs_per_step": rerun_pairs,
                   "l2": "256 MiB buffer written between timed iterations (L2 flush); per-step traceback stream >> L2",
                   "timing": "CUDA events on the engine's stream per step (+ torch events around the pack + all-gather), max over ranks",
                   "cpu_legs": "cpu_baseline / --impl reference time a bounded random sample of the pairs of this same chain set"},
        "wall_ms_per_step": wall_ms,
        "clocks": {k: clocks.get(k) for k in ("sm_mhz", "sm_max_mhz", "reasons", "power_w_max")},
        "e2e": {"value": total_pairs / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1e3 * e2e_s, "steps": e2e_steps,
                "note": "h2d per rank (chains replicated); d2h on rank 0 (the one rank that builds the dense matrices)"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "parity": parity,
        "configs": configs,
        "cpu_baseline": cpu,
        "cpu_baseline_reference": cpu_ref,
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
