"""Host-side logic of the multi-GPU partitioned sort (oclradixsort_b200/dist.py) on CPU:
the exchange plan, and the whole protocol over gloo with world_size 2 where the device kernels are replaced by
numpy / the oracle (test infrastructure standing in for the CUDA ops, which need a GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oclradixsort_b200.dist import NUM_BINS, DistributedPairSorter, plan_exchange
from oracle import pyoracle as po


def test_plan_uniform_is_balanced_and_contiguous():
    rng = np.random.default_rng(0)
    for P in (2, 4, 8):
        hist = rng.integers(900, 1100, size=(P, NUM_BINS))
        plan = plan_exchange(hist)
        b2r = plan["bin_to_rank"]
        assert (np.diff(b2r.astype(int)) >= 0).all() and b2r[0] == 0 and b2r[-1] == P - 1
        assert plan["send_counts"].sum() == hist.sum() and (plan["send_counts"].sum(axis=1) == hist.sum(axis=1)).all()
        assert plan["recv_total"].max() <= 1.05 * hist.sum() / P
        # receive layouts tile every destination buffer exactly once
        for d in range(P):
            spans = sorted((int(plan["recv_offset"][s, d]), int(plan["send_counts"][s, d])) for s in range(P))
            assert spans[0][0] == 0 and all(a + c == b for (a, c), (b, _) in zip(spans, spans[1:])) and sum(spans[-1]) == plan["recv_total"][d]
            lo, hi = plan["edges"][d], plan["edges"][d + 1]
            segs = sorted((int(plan["bin_offset"][s, b]), int(hist[s, b])) for s in range(P) for b in range(lo, hi))
            assert segs[0][0] == 0 and all(a + c == b for (a, c), (b, _) in zip(segs, segs[1:])) and sum(segs[-1]) == plan["recv_total"][d]


def test_plan_skewed_keeps_digit_ranges_whole():
    hist = np.zeros((4, NUM_BINS), dtype=np.int64)
    hist[:, 7] = 1000  # everything in one digit: cannot be split by key, one rank gets it all
    plan = plan_exchange(hist)
    assert sorted(plan["recv_total"].tolist()) == [0, 0, 0, 4000]
    hist[:, 200] = 3000
    plan = plan_exchange(hist)
    assert plan["recv_total"].sum() == 16000 and (np.diff(plan["bin_to_rank"].astype(int)) >= 0).all()
    assert plan["bin_to_rank"][7] != plan["bin_to_rank"][200]


class NumpyOps:
    """CPU stand-in for CudaLocalOps (tests only)."""
    cuda = torch.device("cpu")

    def empty(self, n):
        return torch.empty(max(int(n), 1), dtype=torch.int64)

    def histogram(self, pairs, n):
        keys = pairs[:n].numpy().view(np.uint32)[0::2]
        return torch.from_numpy(np.bincount(keys >> 24, minlength=NUM_BINS).astype(np.int64))

    def partition(self, src, dst, n, bin_to_part, part_counts):
        a = src[:n].numpy()
        part = bin_to_part[a.view(np.uint32)[0::2] >> 24]
        assert np.array_equal(np.bincount(part, minlength=len(part_counts))[: len(part_counts)], part_counts)
        dst[:n] = torch.from_numpy(a[np.argsort(part, kind="stable")])

    def local_sort(self, pairs, m):
        a = pairs[:m].numpy().view(np.uint32).reshape(m, 2)
        a[:] = po.sort_pairs(np.ascontiguousarray(a))

    def to_host_matrix(self, t):
        return t.numpy()

    def release(self):
        pass


def _make_input(kind, rank, n):
    rng = np.random.default_rng(100 + rank)
    keys = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    if kind == "lowentropy":
        keys = (keys & np.uint32(0x03000003)) | np.uint32(0x40000000)
    elif kind == "ragged":
        keys = keys >> np.uint32(rank * 3)  # different spread on every rank
    kv = np.empty((n, 2), dtype=np.uint32)
    kv[:, 0], kv[:, 1] = keys, np.arange(n, dtype=np.uint32) + rank * 10_000_000
    return kv


def _worker(rank, world, port, kind, ns, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = ns[rank]
    kv = _make_input(kind, rank, n)
    sorter = DistributedPairSorter(None, None, max(ns), dist, ops=NumpyOps(), slack=8.0, exchange="nccl")
    out, m = sorter.sort(torch.from_numpy(kv.view(np.int64).reshape(-1).copy()), n)
    np.save(os.path.join(out_dir, f"out{rank}.npy"), out.numpy().view(np.uint32).reshape(m, 2))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("kind,ns", [("uniform", (5000, 5000)), ("lowentropy", (4096, 3000)), ("ragged", (7001, 1))])
def test_protocol_world2_gloo_matches_single_stable_sort(tmp_path, kind, ns):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), kind, ns, str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / f"out{r}.npy") for r in range(world)])
    whole = np.concatenate([_make_input(kind, r, ns[r]) for r in range(world)])
    assert np.array_equal(got, po.sort_pairs(whole))  # rank-order concatenation == one stable sort of the whole input
