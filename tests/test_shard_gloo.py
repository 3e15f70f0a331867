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
