"""The N > 1 path on CPU: two `gloo` ranks each sweep their shard (test-only CPU backend), all-gather the bitmaps and
all-reduce the totals; the result must equal the single-process sweep of the whole range."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import HOSTEMU, ROOT

N_TOTAL = 600


def _setup_driver():
    from gatekeeper_b200 import driver as D
    from gatekeeper_b200 import workloads as W
    tm, cons = W.config2()
    drv = D.Driver(lib_path=HOSTEMU, threads=2)
    for k, r in tm:
        drv.add_template(k, r)
    for c in cons:
        drv.AddConstraint(c)
    for ns in W.synth_namespaces():
        drv.AddData("admission.k8s.gatekeeper.sh", ["cluster", "v1", "Namespace", ns["metadata"]["name"]], ns)
    return drv


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gatekeeper_b200 import driver as D
    from gatekeeper_b200 import workloads as W
    from gatekeeper_b200.sweep import ShardedSweep, shard_range
    drv = _setup_driver()
    lo, hi = shard_range(N_TOTAL, rank, world)
    blob = W.synth_objects(lo, hi - lo)
    rb = drv.upload_blob(blob)
    sw = ShardedSweep(drv, rb, hi - lo, len(drv.constraints()), torch.device("cpu"), world)
    gathered, tot = sw.step(D.AUDIT_EP)
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), gathered.contiguous().numpy().reshape(world * (hi - lo), -1))
        np.save(os.path.join(out_dir, "totals.npy"), tot.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sweep_equals_single_process(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    gathered = np.load(tmp_path / "gathered.npy")
    totals = np.load(tmp_path / "totals.npy")
    from gatekeeper_b200 import driver as D
    from gatekeeper_b200 import workloads as W
    drv = _setup_driver()
    whole = drv.upload_blob(W.synth_objects(0, N_TOTAL)).eval(D.AUDIT_EP)
    assert (gathered.view("uint32") == whole.viol_bits).all()
    assert totals[0, :len(whole.totals)].tolist() == whole.totals
    assert totals[1, :len(whole.totals)].tolist() == whole.err_totals


def test_shard_ranges_partition_the_batch():
    from gatekeeper_b200.sweep import shard_range
    for n in (0, 1, 7, 1000, 10**6 + 3):
        for g in (1, 2, 4, 8):
            rs = [shard_range(n, r, g) for r in range(g)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(g - 1))


def test_peer_addressing_of_the_fused_exchange():
    """gk_batch_eval_device_peers on the test backend: two "ranks" in one process write their bitmap shard and totals
    into each other's receive buffers (host arrays standing in for peer-mapped device memory); every buffer must end up
    holding both shards in rank order and totals that sum to the single-process sweep."""
    from gatekeeper_b200 import driver as D
    from gatekeeper_b200 import workloads as W
    from gatekeeper_b200.sweep import shard_range
    world, n_total = 2, 500
    drv = _setup_driver()
    C = len(drv.constraints())
    words = (C + 31) // 32
    whole = drv.upload_blob(W.synth_objects(0, n_total)).eval(D.AUDIT_EP)
    n_local = n_total // world
    tot_off = n_local * words + (n_local * words) % 2
    slot = tot_off + 4 * C
    recv = [np.zeros(world * slot, dtype=np.int32) for _ in range(world)]
    bases = [int(r.ctypes.data) for r in recv]
    keep = []
    for rank in range(world):
        lo, hi = shard_range(n_total, rank, world)
        rb = drv.upload_blob(W.synth_objects(lo, hi - lo))
        keep.append(rb)
        err = np.zeros(n_local * words, dtype=np.int32)
        tot = np.zeros((2, C), dtype=np.int64)
        rb.eval_device_peers(D.AUDIT_EP, bases, rank, slot, tot_off, C, int(err.ctypes.data), int(tot[0].ctypes.data), int(tot[1].ctypes.data), 0)
    for r in recv:
        g = r.reshape(world, slot)
        bitmap = g[:, :n_local * words].reshape(world * n_local, words).view(np.uint32)
        assert (bitmap == whole.viol_bits).all()
        totals = g[:, tot_off:].copy().view(np.int64).reshape(world, 2, C).sum(0)
        assert totals[0].tolist() == whole.totals and totals[1].tolist() == whole.err_totals
