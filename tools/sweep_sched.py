"""Schedule sweep on one GPU: pipeline vs stream-per-batch, number of batches, rows per unit (C3 by default)."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caretta_b200 import synth, engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 300
e = engine.Engine(); ch = synth.make_chains(n, L, 10, seed=3); e.set_chains(ch.coords, ch.tensors, ch.offsets)
def run(tag, **env):
    for k, v in env.items(): os.environ[k] = str(v)
    best = min((e.pairwise_shard(e.params(), 0, 1), e.last_elapsed_ms())[1] for _ in range(4))
    print(f"{tag:60s} {best:8.2f} ms  launches {e.last_launches()}", flush=True)
run("legacy streams=3", CARETTA_B200_PIPE=0, CARETTA_B200_STREAMS=3)
for nb in (4, 6, 8, 12, 16, 24):
    run(f"pipe batches={nb}", CARETTA_B200_PIPE=1, CARETTA_B200_STREAMS=3, CARETTA_B200_BATCHES=nb, CARETTA_B200_WORKSPACE_MB=20000)
for rows in (1536, 3072, 12288):
    run(f"pipe batches=8 unit_rows={rows}", CARETTA_B200_PIPE=1, CARETTA_B200_BATCHES=8, CARETTA_B200_UNIT_ROWS=rows)
for rows in (1536, 3072):
    run(f"legacy streams=3 unit_rows={rows}", CARETTA_B200_PIPE=0, CARETTA_B200_STREAMS=3, CARETTA_B200_UNIT_ROWS=rows)
