"""ADVICE (round 1): MultipleAlignment.final_sequences / final_consensus_weights are lazy views of the engine's device pool; a
second progressive alignment on the same engine replaces that pool.  The engine therefore fetches an outstanding view before it
replaces the pool (Engine.msa_begin / msa_track).  Host logic only: a stand-in engine records the calls."""
import gc

from caretta_b200 import multiple_alignment as ma
from caretta_b200 import engine as E


class _FakeEngine:
    msa_begin = E.Engine.msa_begin
    msa_track = E.Engine.msa_track

    def __init__(self):
        self.fetched = []
        self._offsets = [0, 3, 7]
        self._tensor_width = 10
        self.h = 1

        class _Lib:
            @staticmethod
            def crt_msa_begin(h, w, n):
                return 0
        self.lib = _Lib()

    def _check(self, rc, what):
        assert rc == 0

    def msa_fetch(self, ids):
        self.fetched.append(list(ids))
        return [("t", "c", "w")] * len(ids)


def test_outstanding_nodes_are_fetched_before_the_pool_is_replaced():
    eng = _FakeEngine()
    eng.msa_begin(1.0)
    first = ma._PoolNodes(eng, [2, 3], ["a", "b"])
    assert first.data is None
    eng.msa_begin(1.0)                       # the next alignment: the pool is about to be replaced
    assert eng.fetched == [[2, 3]] and first.data is not None
    assert first.fetch() == [("t", "c", "w")] * 2          # still readable afterwards, no error
    second = ma._PoolNodes(eng, [4], ["c"])
    del second
    gc.collect()
    eng.msa_begin(1.0)                       # a dropped view is not fetched
    assert eng.fetched == [[2, 3]]
