"""Prints the measured schedule of one fp32 all-vs-all run (CARETTA_B200_TIMELINE=1: every stage kernel of every batch with its
variant -- columns per lane C, single / multi strip --, its cells and its start / end on the device clock), after three warm-up runs.
   python tools/timeline_run.py C3 | C5 | C4sub [serial]      (serial: CARETTA_B200_STREAMS=1, the kernels unoverlapped)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from caretta_b200 import engine, synth

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
if len(sys.argv) > 2 and sys.argv[2] == "serial":
    os.environ["CARETTA_B200_STREAMS"] = "1"
if name == "C4sub":
    c4 = synth.config("C4"); e = int(c4.offsets[1000])
    ch = synth.Chains(c4.coords[:e], c4.tensors[:e], c4.offsets[:1001].copy())
else:
    ch = synth.config(name)
eng = engine.Engine(0)
eng.set_chains(ch.coords, ch.tensors, ch.offsets)
prm = eng.params(precision=engine.FP32)
for _ in range(3):
    eng.pairwise_shard(prm, 0, 1)
print(f"{name}: {eng.last_elapsed_ms():.3f} ms per step, phases {eng.last_phase_ms()}", file=sys.stderr)
os.environ["CARETTA_B200_TIMELINE"] = "1"
eng.pairwise_shard(prm, 0, 1)
print(f"{name}: timeline run itself {eng.last_elapsed_ms():.3f} ms", file=sys.stderr)
