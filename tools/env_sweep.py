"""A/B of run-schedule knobs on one configuration: every knob is an environment variable the library reads per run.
   python tools/env_sweep.py C3 CARETTA_B200_HOLD_FILL2=0,1,2,3 CARETTA_B200_RERUN_EVERY=0,1
prints the median device time of 5 steps (after 2 warm-up steps) for every combination."""
import itertools, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from caretta_b200 import engine, synth

name = sys.argv[1]
world = int(os.environ.get("SWEEP_WORLD", "1"))          # > 1: rank 0's shard of a world of that size (strong scaling on one GPU)
knobs = [(a.split("=")[0], a.split("=")[1].split(",")) for a in sys.argv[2:]]
if name == "C4sub":
    c4 = synth.config("C4"); e = int(c4.offsets[1000])
    ch = synth.Chains(c4.coords[:e], c4.tensors[:e], c4.offsets[:1001].copy())
else:
    ch = synth.config(name)
eng = engine.Engine(0)
eng.set_chains(ch.coords, ch.tensors, ch.offsets)
prm = eng.params(precision=engine.FP32)
for combo in itertools.product(*[v for _, v in knobs]):
    for (k, _), v in zip(knobs, combo):
        os.environ[k] = v
    ts = []
    for it in range(7):
        eng.pairwise_shard(prm, 0, world)
        if it >= 2:
            ts.append(eng.last_elapsed_ms())
    print(name, " ".join(f"{k.replace('CARETTA_B200_', '')}={v}" for (k, _), v in zip(knobs, combo)),
          f"median {np.median(ts):.3f} ms  min {min(ts):.3f}  rerun {eng.last_rerun()}  phases {({k: round(float(v), 3) for k, v in eng.last_phase_ms().items()} if os.environ.get('CARETTA_B200_STREAMS') == '1' else '')}", flush=True)
