"""N-sharding host logic on CPU with world_size = 2 (gloo): every rank owns a contiguous block of the GLOBAL
sample index space, leaves a (min cost, sum w, sum w*u) record, the records are all-gathered in rank order and
merged -- which must equal the unsharded softmax update.  The per-rank record is produced here by
the oracle (this is a test of the exchange protocol and of the layout covo_step_partial_device /
covo_step_merge_device use, not of the kernels)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, H, out_dir):
    sys.path.insert(0, ROOT)
    from oracle import oracle_np as o
    from tests.util import scenario

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p, ns, a_mean, rng = scenario("tracking_zigzag", seed=5, H=H, warm_steps=4)
    n = 4 * H
    n_pad = (n + 7) // 8 * 8
    lam = 0.01
    # the production RNG field is a function of the GLOBAL sample index: every rank regenerates its own block
    n_local = N // world
    eps = o.philox_normals(seed=42, stream=0, n_samples=n_local, n_cols=n, sample_offset=rank * n_local)
    L = (0.5 * np.eye(n)).astype(np.float32)
    mean_s = o.shift_mean(a_mean)
    a_s = o.sample_actions(mean_s, L, eps)
    cost = o.rollout_costs(ns, a_s, p)
    m, s_, v = o.softmax_partials(a_s.astype(np.float64), cost.astype(np.float64), lam)
    rec = np.zeros(4 + n_pad, np.float32)  # record layout of csrc/common.cuh: (m, s, pad, pad, v[n_pad])
    rec[0], rec[1] = m, s_
    rec[4:4 + n] = v.reshape(-1)
    gathered = torch.zeros(world, 4 + n_pad)
    dist.all_gather_into_tensor(gathered.view(-1), torch.from_numpy(rec))
    g = gathered.numpy().astype(np.float64)
    # merge in rank order (overflow safe: each record carries its own min)
    parts = [(g[r, 0], g[r, 1], g[r, 4:4 + n]) for r in range(world)]
    M, S, V = o.merge_partials(parts, lam)
    new_mean = (V / S).reshape(H, 4)
    np.save(os.path.join(out_dir, f"mean_{rank}.npy"), new_mean)
    if rank == 0:
        # unsharded reference
        eps_all = o.philox_normals(seed=42, stream=0, n_samples=N, n_cols=n)
        a_all = o.sample_actions(mean_s, L, eps_all)
        cost_all = o.rollout_costs(ns, a_all, p)
        ref, _ = o.softmax_update(mean_s, a_all, cost_all, lam)
        np.save(os.path.join(out_dir, "ref.npy"), ref)
        # a plain sum-allreduce of raw numerator/denominator would overflow float32
        assert -cost_all.min() / lam > 88.0
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partial_merge_equals_unsharded(tmp_path):
    world, N, H = 2, 256, 10
    port = _free_port()
    mp.spawn(_worker, args=(world, port, N, H, str(tmp_path)), nprocs=world, join=True)
    ref = np.load(tmp_path / "ref.npy")
    m0 = np.load(tmp_path / "mean_0.npy")
    m1 = np.load(tmp_path / "mean_1.npy")
    assert np.array_equal(m0, m1)  # every rank ends with the same mean (rank-ordered merge)
    assert np.abs(m0 - ref).max() < 1e-5


def test_env_sharding_assignment():
    """Environment-batch sharding has no collective: env e -> rank e % G, every env exactly once."""
    E, G = 4096, 8
    owners = [[e for e in range(E) if e % G == r] for r in range(G)]
    assert sum(len(x) for x in owners) == E and all(len(x) == E // G for x in owners)
    assert len(set().union(*map(set, owners))) == E
