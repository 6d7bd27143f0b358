#!/bin/bash
# A/B timing of library variants in build/*.so against the in-tree build: tools/ab_time.sh [n] [L]
for lib in caretta_b200/libcaretta_b200.so build/*.so; do
  echo "== $lib"
  CARETTA_B200_LIB=$PWD/$lib python - "$@" <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from caretta_b200 import synth, engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 300
e = engine.Engine(); ch = synth.make_chains(n, L, 10, seed=3); e.set_chains(ch.coords, ch.tensors, ch.offsets)
for streams in (1, 3):
    os.environ["CARETTA_B200_STREAMS"] = str(streams)
    best = min((e.pairwise_shard(e.params(), 0, 1), e.last_elapsed_ms())[1] for _ in range(4))
    print(f"  streams={streams}: {best:.2f} ms", e.last_phase_ms() if streams == 1 else "")
PY
done
