"""Where the first call's time goes (one-shot CLI use): context creation, first upload, first / second all-vs-all."""
import os, sys, time
t00 = time.perf_counter()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from caretta_b200 import engine, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
ch = synth.make_chains(n, 300, 10, seed=3, family_size=20)
t0 = time.perf_counter(); e = engine.Engine(); t1 = time.perf_counter()
e.set_chains(ch.coords, ch.tensors, ch.offsets); t2 = time.perf_counter()
S = e.pairwise_all(e.params()); t3 = time.perf_counter()
d1 = e.last_elapsed_ms()
S = e.pairwise_all(e.params()); t4 = time.perf_counter()
print(f"N={n}: imports+synth {1e3*(t0-t00):.0f} ms, Engine() {1e3*(t1-t0):.0f} ms, set_chains {1e3*(t2-t1):.0f} ms, "
      f"first pairwise_all {1e3*(t3-t2):.0f} ms (device {d1:.1f}), second {1e3*(t4-t3):.0f} ms (device {e.last_elapsed_ms():.1f})")
