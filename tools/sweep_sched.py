"""Schedule sweep on one GPU: number of batches / workspace budget of the stage pipeline (config given on the command line)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caretta_b200 import synth, engine
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
e = engine.Engine(); ch = synth.config(cfg); e.set_chains(ch.coords, ch.tensors, ch.offsets)
def run(tag, reps=3, **env):
    for k, v in env.items(): os.environ[k] = str(v)
    best = min((e.pairwise_shard(e.params(), 0, 1), e.last_elapsed_ms())[1] for _ in range(reps))
    print(f"{cfg} {tag:50s} {best:9.2f} ms  launches {e.last_launches()}", flush=True)
for nb, mb in ((8, 3072), (8, 8192), (6, 8192), (12, 8192), (16, 8192), (8, 16384), (24, 16384)):
    run(f"pipe batches>={nb} workspace<={mb} MB", CARETTA_B200_BATCHES=nb, CARETTA_B200_WORKSPACE_MB=mb)
