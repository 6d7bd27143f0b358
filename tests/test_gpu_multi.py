"""Multi-GPU inside the C ABI (crt_multi_*): every device of the box behind one call, bitwise the one-GPU
matrices in both precisions; the packed exchange format (crt_pack_results / crt_scatter_gathered) on one device; the mirror's
make_pairwise_matrix takes all GPUs without torchrun.  The >= 2 device cases are skipped on a one-GPU box
(run them with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`)."""
import os

import numpy as np
import pytest

from caretta_b200 import engine, synth

pytestmark = pytest.mark.gpu


def _n_devices():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def chains():
    return synth.make_chains(90, list(np.random.default_rng(5).integers(40, 260, 90)), 10, seed=12)


@pytest.fixture(scope="module")
def single(chains):
    eng = engine.Engine(0)
    eng.set_chains(chains.coords, chains.tensors, chains.offsets)
    out = {p: eng.pairwise_all(eng.params(precision=p), want_rmsd_tm=True) for p in (engine.FP64, engine.FP32)}
    eng.close()
    return out


def test_one_device_set_is_the_single_engine(chains, single):
    m = engine.MultiEngine([0])
    assert m.n_devices == 1
    m.set_chains(chains.coords, chains.tensors, chains.offsets)
    for prec in (engine.FP64, engine.FP32):
        S, R, T = m.pairwise_all(m.params(precision=prec), want_rmsd_tm=True)
        for a, b in zip((S, R, T), single[prec]):
            assert np.array_equal(a, b)
    m.close()


@pytest.mark.parametrize("world", [2, 3, 5])
def test_pack_and_scatter_reassemble_the_matrices(chains, single, world):
    """The exchange format on ONE device: every rank's shard packed (float64: exact, float32: rounded once), concatenated the way
    an all-gather would, scattered by crt_scatter_gathered."""
    import torch
    eng = engine.Engine(0)
    eng.set_chains(chains.coords, chains.tensors, chains.offsets)
    for prec, f64 in ((engine.FP64, True), (engine.FP32, False)):
        prm = eng.params(precision=prec)
        pad = max(eng.shard_size(r, world) for r in range(world))
        g = torch.zeros(world, 3 * pad, dtype=torch.float64 if f64 else torch.float32, device="cuda:0")
        for r in range(world):
            eng.pairwise_shard(prm, r, world)
            eng.pack_results(g[r].data_ptr(), pad, f64)
        torch.cuda.synchronize()
        S, R, T = eng.scatter_gathered(g.data_ptr(), world, pad, f64)
        want = single[prec]
        if f64:
            for a, b in zip((S, R, T), want):
                assert np.array_equal(a, b)
        else:
            for a, b in zip((S, R, T), want):
                assert np.array_equal(a, b.astype(np.float32).astype(np.float64))
        assert np.array_equal(S, S.T) and np.all(np.diag(S) == 0) and np.all(np.diag(T) == 1)
    eng.close()


@pytest.mark.skipif(_n_devices() < 2, reason="needs >= 2 GPUs")
def test_all_devices_bitwise_equal_to_one(chains, single):
    m = engine.MultiEngine()
    assert m.n_devices == _n_devices()
    m.set_chains(chains.coords, chains.tensors, chains.offsets)
    S, R, T = m.pairwise_all(m.params(precision=engine.FP64), want_rmsd_tm=True)
    for a, b in zip((S, R, T), single[engine.FP64]):
        assert np.array_equal(a, b)                                   # float64 travels as float64: bit-identical
    S, R, T = m.pairwise_all(m.params(precision=engine.FP32), want_rmsd_tm=True)
    for a, b in zip((S, R, T), single[engine.FP32]):
        assert np.array_equal(a, b.astype(np.float32).astype(np.float64))      # the production exchange is float32
    t = m.last_timing()
    assert t["shard_ms"] > 0 and t["wall_ms"] > 0
    m.close()


@pytest.mark.skipif(_n_devices() < 2, reason="needs >= 2 GPUs")
def test_mirror_make_pairwise_matrix_uses_every_gpu(chains, single, monkeypatch):
    from caretta_b200 import multiple_alignment as ma
    monkeypatch.setattr(ma, "MULTI_GPU_MIN_CELLS", 0.0)
    monkeypatch.setattr(ma, "_shared_multi", None)
    monkeypatch.delenv("LOCAL_RANK", raising=False)
    monkeypatch.delenv("CARETTA_B200_DEVICE", raising=False)
    prots = [ma.Protein(f"p{k}", *chains.chain(k)) for k in range(chains.n)]
    msa = ma.MultipleAlignment(prots, precision=engine.FP64)
    S = msa.make_pairwise_matrix(dict(gamma_tensor=7.0, gamma_coords=0.03))
    assert ma._shared_multi is not None and ma._shared_multi.n_devices == _n_devices()
    assert np.array_equal(S, single[engine.FP64][0])
