"""World-size-2 gloo test of the multi-GPU exchange step on CPU: list l lives on rank
l % world, every rank searches only the lists it owns (here with the CPU oracle standing in for
the per-GPU engine), one all-gather of the per-rank [nq x k] results, then the k-way merge of
src/hybrid/core.rs:482-483.  The merged result must equal the unsharded search bit for bit."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

import oracle as O  # noqa: E402
from fabstir_vectordb_b200 import synth  # noqa: E402
from fabstir_vectordb_b200.shard import gather_layout, merge_parts_reference, owner_of_list, pack_layout  # noqa: E402

N, D, NLIST, NQ, K, NPROBE = 4000, 32, 16, 24, 10, 6


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case():
    x = synth.rows(0, N, D, 64, 0.7, 1234)
    q = synth.queries(0, NQ, D, N, 64, 0.7, 1234, synth.default_qnoise(D, 0.7), 5678)
    cents = x[:: N // NLIST][:NLIST].copy()
    return x, q, cents


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x, q, cents = _case()
        full = O.IVF(cents, x, np.arange(N, dtype=np.uint32))
        # this rank's shard: rows of the lists it owns, same centroids (replicated)
        mine = np.array([owner_of_list(int(l), world) == rank for l in full.assign])
        shard = O.IVF(cents, x[mine], np.arange(N, dtype=np.uint32)[mine], assign_=full.assign[mine])
        ids, dst, cnt = O.hybrid_batch_search(shard, None, None, q, K, NPROBE, tiers=2)
        # the product's exchange step: ONE all-gather of the packed chunk [ids | dist | count]
        o_i, o_d, o_c, chunk = pack_layout(NQ, K)
        pack = np.empty(chunk, dtype=np.int32)
        pack[o_i:o_i + NQ * K] = ids.astype(np.uint32).view(np.int32).ravel()
        pack[o_d:o_d + NQ * K] = dst.astype(np.float32).view(np.int32).ravel()
        pack[o_c:o_c + NQ] = cnt.astype(np.uint32).view(np.int32)
        g_pack = torch.empty((world, chunk), dtype=torch.int32)
        dist.all_gather_into_tensor(g_pack.view(-1), torch.from_numpy(pack))
        g = g_pack.numpy()
        (gs, cs) = gather_layout(NQ, K, world)
        g_ids = g[:, o_i:o_i + NQ * K].copy().view(np.uint32).reshape(gs)
        g_dst = g[:, o_d:o_d + NQ * K].copy().view(np.float32).reshape(gs)
        g_cnt = g[:, o_c:o_c + NQ].copy().view(np.uint32).reshape(cs)
        m_ids, m_dst, m_cnt = merge_parts_reference(g_ids, g_dst, g_cnt, K)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), ids=m_ids, dst=m_dst, cnt=m_cnt)
    finally:
        dist.destroy_process_group()


def test_list_sharded_search_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    x, q, cents = _case()
    full = O.IVF(cents, x, np.arange(N, dtype=np.uint32))
    w_ids, w_dst, w_cnt = O.hybrid_batch_search(full, None, None, q, K, NPROBE, tiers=2)
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npz")
        assert got["cnt"].tolist() == w_cnt.tolist()
        for i in range(NQ):
            c = int(w_cnt[i])
            assert got["ids"][i, :c].tolist() == w_ids[i, :c].tolist()
            assert got["dst"][i, :c].view(np.uint32).tolist() == w_dst[i, :c].view(np.uint32).tolist()


def test_owner_rule_partitions_lists():
    for world in (1, 2, 4, 8):
        owners = [owner_of_list(l, world) for l in range(64)]
        assert set(owners) == set(range(world))
        assert all(o == l % world for l, o in enumerate(owners))


def test_place_lists_balances_rows():
    """Greedy size-balanced placement: every list gets exactly one owner, the heaviest and the
    lightest rank differ by at most the largest list, and the table is a pure function of the histogram."""
    from fabstir_vectordb_b200.shard import place_lists
    rng = np.random.default_rng(5)
    sizes = (rng.pareto(1.5, 4096) * 300).astype(np.int64) + 1          # a few hub lists, many small ones
    for world in (1, 2, 4, 8):
        owner = place_lists(sizes, world)
        assert owner.shape == sizes.shape and owner.max() == world - 1
        load = np.bincount(owner, weights=sizes, minlength=world)
        assert load.sum() == sizes.sum()
        assert load.max() - load.min() <= sizes.max()
        if world > 1:
            mod_load = np.bincount(np.arange(sizes.size) % world, weights=sizes, minlength=world)
            assert load.max() <= mod_load.max()                               # never worse than l % world
        assert np.array_equal(owner, place_lists(sizes.copy(), world))
    # ties go to the lower rank, equal sizes are placed in list order
    assert place_lists([5, 5, 5, 5], 2).tolist() == [0, 1, 0, 1]


def _kmeans_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fabstir_vectordb_b200.shard import kmeans_allreduce_step
        x, _, cents = _case()
        per = (N + world - 1) // world
        mine = x[rank * per:(rank + 1) * per]
        # this rank's share of one Lloyd iteration: assignment (oracle standing in for the per-GPU engine),
        # per-cluster sums / counts / squared error of ITS points
        a = O.assign(mine, cents)
        sums = np.zeros((NLIST, D), dtype=np.float32)
        np.add.at(sums, a, mine)
        counts = np.bincount(a, minlength=NLIST).astype(np.int32)
        d = O.l2_many(mine[0], cents[a[:1]])  # (shape check of the helper)
        assert d.shape == (1,)
        sq = float(sum(O.l2(mine[i], cents[a[i]]) ** 2 for i in range(mine.shape[0])))
        t_sums, t_counts = torch.from_numpy(sums), torch.from_numpy(counts)
        t_sq, t_ch = torch.tensor([sq], dtype=torch.float64), torch.tensor([1 if rank == 0 else 0], dtype=torch.int32)
        tot, changed, total = kmeans_allreduce_step(t_sums, t_counts, t_sq, t_ch)
        means = cents.copy()
        nz = t_counts.numpy() > 0
        means[nz] = t_sums.numpy()[nz] / t_counts.numpy()[nz, None].astype(np.float32)
        np.savez(os.path.join(out_dir, f"km{rank}.npz"), means=means, counts=t_counts.numpy(), tot=tot,
                 changed=changed, total=total)
    finally:
        dist.destroy_process_group()


def test_sharded_kmeans_exchange_step_world2(tmp_path):
    """Points sharded over two ranks, one all-reduce of sums / counts / error (kmeans_allreduce_step): both
    ranks end up with the same means, equal to the unsharded update_centroids (src/ivf/core.rs:388-417) up
    to the f32 summation order, with exact counts and the unsharded error."""
    world = 2
    mp.spawn(_kmeans_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    x, _, cents = _case()
    a = O.assign(x, cents)
    want = O.update_centroids(x, a, cents)
    r0, r1 = np.load(tmp_path / "km0.npz"), np.load(tmp_path / "km1.npz")
    assert np.array_equal(r0["means"].view(np.uint32), r1["means"].view(np.uint32))
    assert r0["counts"].tolist() == np.bincount(a, minlength=NLIST).tolist()
    assert int(r0["total"]) == N and int(r0["changed"]) == 1
    np.testing.assert_allclose(r0["means"], want, rtol=2e-5, atol=2e-6)
    assert abs(float(r0["tot"]) / N - O.compute_error(x, cents, a)) < 1e-4
