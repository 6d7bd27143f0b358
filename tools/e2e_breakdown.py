import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from caretta_b200 import synth, engine
e = engine.Engine(); ch = synth.config("C3")
pc = torch.from_numpy(ch.coords).pin_memory().numpy(); pt = torch.from_numpy(ch.tensors).pin_memory().numpy(); po = ch.offsets
prm = e.params()
for rep in range(4):
    t0 = time.perf_counter(); e.set_chains(pc, pt, po); t1 = time.perf_counter()
    e.pairwise_shard(prm, 0, 1); t2 = time.perf_counter()
    r = e.fetch(499500); t3 = time.perf_counter()
    print(f"set_chains {1e3*(t1-t0):.2f} ms  shard {1e3*(t2-t1):.2f} ms (device {e.last_elapsed_ms():.2f})  fetch {1e3*(t3-t2):.2f} ms")
