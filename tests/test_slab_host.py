"""CPU tests of the host-side multi-GPU logic (z-slab planning, rebalancing rule, layer assignment), including
a world_size-2 gloo run: each rank histograms its share of the scene, the histograms are all-reduced, and both
ranks must derive the same cut planes and a consistent partition (own + 3 ghost layers per side)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_plan_balanced_and_respects_min_thickness(sf):
    from simplefluid_b200 import binding
    p = sf.default_params(64, "Dambreak")
    pos = sf.scene_generate(p)
    layers = binding.cell_layers(p, pos)
    nz = 64
    hist = np.bincount(layers, minlength=nz).astype(np.uint64)
    for nranks in (1, 2, 4, 8):
        cuts = binding.slab_plan(hist, nranks)
        assert cuts[0] == 0 and cuts[-1] == nz and np.all(np.diff(cuts) >= 6)
        own = np.array([hist[cuts[r]:cuts[r + 1]].sum() for r in range(nranks)])
        assert own.sum() == len(pos)
        if nranks <= 2:  # the fluid spans 16 layers of 2-layer lattice planes: 2 slabs can balance within a plane
            assert own.max() - own.min() <= 2 * hist.max()
    with pytest.raises(sf.SFError):
        binding.slab_plan(hist[:10], 4)  # 10 layers cannot host 4 slabs of >= 6 layers


def test_cell_layers_match_oracle_binning(sf, ob):
    from simplefluid_b200 import binding
    p, po = sf.default_params(24, "DoubleDambreak"), ob.default_params(24, "DoubleDambreak")
    pos = sf.scene_generate(p)
    orc = ob.Oracle(po, pos, boundary_seed=0)
    orc.advance()
    assert np.array_equal(binding.cell_layers(p, pos), (orc.cell_index() // (24 * 24)).astype(np.int32))
    orc.close()


def test_rebalance_rule_moves_one_layer_with_hysteresis(sf):
    from simplefluid_b200 import binding
    nz = 64
    cuts = np.array([0, 16, 32, 48, 64], np.int32)

    def table(own, first, last):
        t = np.zeros((4, 8), np.uint32)
        t[:, 2], t[:, 3], t[:, 4] = own, first, last
        return t

    # balanced: nothing moves
    assert np.array_equal(binding.slab_rebalance(table([100, 100, 100, 100], [10] * 4, [10] * 4), nz, cuts), cuts)
    # rank 0 overloaded by more than twice its top layer: boundary 1 moves down by exactly one layer
    out = binding.slab_rebalance(table([200, 100, 100, 100], [10] * 4, [10] * 4), nz, cuts)
    assert list(out) == [0, 15, 32, 48, 64]
    # within the hysteresis band (difference <= 2 * layer population): no move, so planes cannot oscillate
    out = binding.slab_rebalance(table([115, 100, 100, 100], [10] * 4, [10] * 4), nz, cuts)
    assert np.array_equal(out, cuts)
    # a slab at the minimum thickness never shrinks
    thin = np.array([0, 6, 32, 48, 64], np.int32)
    out = binding.slab_rebalance(table([500, 100, 100, 100], [10] * 4, [10] * 4), nz, thin)
    assert out[1] == 6
    # a heavy middle slab between two light ones would give a layer to BOTH neighbours: it gives its bottom layer only
    # (cuts are decided bottom-up), so no slab loses two layers in one substep
    mid = np.array([0, 20, 28, 44, 64], np.int32)  # slab 1 is 8 layers thick (minimum 6)
    out = binding.slab_rebalance(table([100, 900, 100, 100], [10] * 4, [10] * 4), nz, mid)
    assert out.tolist() == [0, 21, 28, 44, 64]
    out2 = binding.slab_rebalance(table([100, 900, 100, 100], [10] * 4, [10] * 4), nz, out)
    assert out2.tolist() == out.tolist()  # 7 layers left: a slab only shrinks while it keeps more than the minimum + 1
    # every plane moves by at most one layer per substep, slabs stay >= 6 layers and never lose two layers at once
    rng = np.random.default_rng(0)
    c = cuts.copy()
    for _ in range(200):
        own = rng.integers(0, 1000, 4)
        c2 = binding.slab_rebalance(table(own, rng.integers(0, 50, 4), rng.integers(0, 50, 4)), nz, c)
        assert np.all(np.abs(c2 - c) <= 1) and np.all(np.diff(c2) >= 6) and c2[0] == 0 and c2[-1] == nz
        assert np.all(np.diff(c2) >= np.diff(c) - 1)
        c = c2


def _gloo_worker(rank, world, port, res, scene, out_dir):
    import sys
    import torch.distributed as dist
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import simplefluid_b200 as sfm
    from simplefluid_b200 import binding
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = sfm.default_params(res, scene)
    pos = sfm.scene_generate(p)
    nz = int(res)
    mine = pos[rank::world]  # this rank's share of the input
    hist = torch.from_numpy(np.bincount(binding.cell_layers(p, mine), minlength=nz).astype(np.int64))
    dist.all_reduce(hist)  # global per-layer histogram
    cuts = binding.slab_plan(hist.numpy().astype(np.uint64), world)
    # what sf_upload_particles_global keeps on this rank: own layers + 3 ghost layers per side
    layers = binding.cell_layers(p, pos)
    zb, ze = cuts[rank], cuts[rank + 1]
    own = np.flatnonzero((layers >= zb) & (layers < ze))
    local = np.flatnonzero((layers >= zb - 3) & (layers < ze + 3))
    gathered = [None] * world
    dist.all_gather_object(gathered, (cuts.tolist(), own.tolist(), len(local)))
    if rank == 0:
        np.save(os.path.join(out_dir, "result.npy"), np.array([1]))
        all_own = np.concatenate([np.array(g[1], np.int64) for g in gathered])
        ok = all(g[0] == gathered[0][0] for g in gathered)  # same planes everywhere
        ok &= len(all_own) == len(pos) and len(np.unique(all_own)) == len(pos)  # ownership is a partition
        ok &= all(g[2] >= len(g[1]) for g in gathered)
        np.save(os.path.join(out_dir, "ok.npy"), np.array([int(ok)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_partition(sf, tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_gloo_worker, args=(2, port, 32.0, "DoubleDambreak", str(tmp_path)), nprocs=2, join=True)
    assert np.load(tmp_path / "ok.npy")[0] == 1
