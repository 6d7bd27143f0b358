"""Unit-length sweep (rows streamed per warp) on one GPU, optionally on rank 0's shard of an N-GPU run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caretta_b200 import synth, engine
e = engine.Engine()
for cfg, world in (("C3", 1), ("C5", 1), ("C5", 8), ("C2", 1), ("C4s", 1)):
    if cfg == "C4s":
        import numpy as np
        rng = np.random.default_rng(4)
        ch = synth.make_chains(1200, list(rng.integers(50, 1001, 1200)), 10, seed=4)
    else:
        ch = synth.config(cfg)
    e.set_chains(ch.coords, ch.tensors, ch.offsets)
    for rows in (768, 1536, 3072, 6144, 12288):
        os.environ["CARETTA_B200_UNIT_ROWS"] = str(rows)
        best = min((e.pairwise_shard(e.params(), 0, world), e.last_elapsed_ms())[1] for _ in range(3))
        print(f"{cfg} world={world} unit_rows={rows:6d}: {best:9.2f} ms  launches {e.last_launches()}", flush=True)
