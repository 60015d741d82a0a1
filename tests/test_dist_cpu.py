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

from oclradixsort_b200.dist import NUM_BINS, DistributedPairSorter, SplitterPlan, plan_exchange
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

    def filtered_histograms(self, pairs, n, shift, prefixes):
        keys = pairs[:n].numpy().view(np.uint32)[0::2]
        high = (keys.astype(np.uint64) >> np.uint64(shift + 8)).astype(np.uint32)
        digit = (keys >> np.uint32(shift)) & np.uint32(255)
        out = np.stack([np.bincount(digit[high == p], minlength=NUM_BINS) for p in prefixes]) if len(prefixes) else np.zeros((0, NUM_BINS))
        return torch.from_numpy(out.astype(np.int64).reshape(-1))

    def partition_by_splitters(self, src, dst, n, thresholds, part_counts):
        a = src[:n].numpy()
        keys = a.view(np.uint32)[0::2].astype(np.uint64)
        part = np.zeros(n, dtype=np.int64)
        for t in thresholds:
            part += keys >= np.uint64(t)
        assert np.array_equal(np.bincount(part, minlength=len(part_counts)), part_counts)
        dst[:n] = torch.from_numpy(a[np.argsort(part, kind="stable")])

    def to_host_matrix(self, t):
        return t.numpy()

    def release(self):
        pass


def _make_input(kind, rank, n):
    rng = np.random.default_rng(100 + rank)
    keys = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    if kind == "allequal":
        keys[:] = 0xDEADBEEF
    elif kind == "hotkey":
        keys[: (3 * n) // 4] = 0x12345678
    elif kind == "hotdigit":
        keys = (keys & np.uint32(0x00FFFFFF)) | np.uint32(0x77000000)
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
    slack = 1.25 if kind in ("allequal", "hotkey", "hotdigit") else 8.0  # skewed inputs: the default slack, so the digit-range plan overflows
    sorter = DistributedPairSorter(None, None, max(ns), dist, ops=NumpyOps(), slack=slack, exchange="nccl")
    out, m = sorter.sort(torch.from_numpy(kv.view(np.int64).reshape(-1).copy()), n)
    np.save(os.path.join(out_dir, f"out{rank}.npy"), out.numpy().view(np.uint32).reshape(m, 2))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("kind,ns", [("uniform", (5000, 5000)), ("lowentropy", (4096, 3000)), ("ragged", (7001, 1)),
                                     ("allequal", (6000, 6000)), ("hotkey", (5000, 5000)), ("hotdigit", (5000, 4000))])
def test_protocol_world2_gloo_matches_single_stable_sort(tmp_path, kind, ns):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), kind, ns, str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / f"out{r}.npy") for r in range(world)])
    whole = np.concatenate([_make_input(kind, r, ns[r]) for r in range(world)])
    assert np.array_equal(got, po.sort_pairs(whole))  # rank-order concatenation == one stable sort of the whole input


def _brute_force_split(keys_per_src):
    """Runs SplitterPlan the way DistributedPairSorter.sort_with_splitters does, with numpy histograms."""
    P = len(keys_per_src)
    plan = SplitterPlan(np.stack([np.bincount(k >> np.uint32(24), minlength=NUM_BINS) for k in keys_per_src]))
    for level in (1, 2, 3):
        if P == 1:
            break
        shift = 24 - 8 * level
        pref = plan.prefixes()
        H = np.zeros((P, P - 1, NUM_BINS), dtype=np.int64)
        for s, k in enumerate(keys_per_src):
            high = (k.astype(np.uint64) >> np.uint64(shift + 8)).astype(np.uint32)
            d = (k >> np.uint32(shift)) & np.uint32(255)
            for j in range(P - 1):
                H[s, j] = np.bincount(d[high == pref[j]], minlength=NUM_BINS)
        plan.refine(H)
    return plan.finish()


@pytest.mark.parametrize("P", [1, 2, 3, 8])
@pytest.mark.parametrize("kind", ["uniform", "allequal", "two", "ragged", "and3", "hotkey"])
def test_splitter_plan_is_exact_stable_and_balanced(P, kind):
    rng = np.random.default_rng(P * 31 + len(kind))
    keys = []
    for s in range(P):
        n = 2000
        u = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
        if kind == "allequal":
            u[:] = 0xDEADBEEF
        elif kind == "two":
            u = np.where(u & 1, np.uint32(5), np.uint32(0xFFFFFFFF))
        elif kind == "ragged":
            u = u >> np.uint32(2 * s)
        elif kind == "and3":
            u = u & rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32) & rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
        elif kind == "hotkey":
            u[: n * 3 // 4] = 0x12345678
        keys.append(u)
    res = _brute_force_split(keys)
    got = [[] for _ in range(P)]
    for s, k in enumerate(keys):
        part = np.zeros(k.size, dtype=np.int64)
        for j in range(P - 1):
            part += k.astype(np.uint64) >= res["thresholds"][s, j]
        assert np.array_equal(np.bincount(part, minlength=P), res["send_counts"][s])
        for d in range(P):
            sel = np.nonzero(part == d)[0]
            got[d].append(np.stack([k[sel].astype(np.int64), np.full(sel.size, s), sel], axis=1))
    allk = np.concatenate(keys).astype(np.int64)
    src = np.concatenate([np.full(k.size, s) for s, k in enumerate(keys)])
    pos = np.concatenate([np.arange(k.size) for k in keys])
    order = np.lexsort((pos, src, allk))
    want = np.stack([allk[order], src[order], pos[order]], axis=1)
    parts = []
    for d in range(P):
        a = np.concatenate(got[d])
        parts.append(a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))])  # a stable local sort keeps (source, position) order
    assert np.array_equal(np.concatenate(parts), want)  # rank-order concatenation == the stable sort of the whole input
    # equal shards: balanced to within ONE source's count of the key a boundary falls on (exact when no key is hot)
    spread = int(res["recv_total"].max() - res["recv_total"].min())
    assert spread <= (2000 if kind in ("two", "hotkey") else 1), (spread, res["recv_total"])


def test_halves_plan_two_stable_sorts_equal_one_stable_sort():
    """The pipelined partitioned sort on the CPU, P simulated ranks: every source partitions its slice by plan_exchange_halves()
    into the receive buffers' [A: source 0, 1, ... | B: source 0, 1, ...] layout at the plan's offsets, every destination sorts
    its two halves separately (stable); the concatenation must be the oracle's stable sort of the whole input."""
    from oclradixsort_b200.dist import plan_exchange_halves
    rng = np.random.default_rng(21)
    for P, n, kind, a in ((2, 5000, "uniform", 500), (3, 4001, "skewtop", 500), (8, 3000, "uniform", 350), (5, 2000, "onedigit", 500), (4, 1, "uniform", 500),
                          (16, 700, "lowentropy", 650)):
        slices = []
        for r in range(P):
            m = n + 3 * r
            keys = rng.integers(0, 2**32, size=m, dtype=np.uint64).astype(np.uint32)
            if kind == "skewtop":
                keys >>= np.uint32(2 * r)
            elif kind == "onedigit":
                keys = (keys & np.uint32(0x00FFFFFF)) | np.uint32(0x33000000)
            elif kind == "lowentropy":
                keys = (keys & np.uint32(0x0F00000F)) * np.uint32(0x11)
            kv = np.empty((m, 2), dtype=np.uint32)
            kv[:, 0], kv[:, 1] = keys, np.arange(m, dtype=np.uint32) + np.uint32(r << 20)
            slices.append(kv)
        hist = np.stack([np.bincount(kv[:, 0] >> 24, minlength=256) for kv in slices])
        plan = plan_exchange_halves(hist, a)
        # plan invariants
        assert np.all(np.diff(plan["bin_to_part"].astype(int)) >= 0) and plan["part_counts"].sum() == sum(len(kv) for kv in slices)
        assert np.array_equal(plan["recv_a"] + plan["recv_b"], plan["recv_total"])
        recv = [np.full((int(plan["recv_total"][d]), 2), 0xFFFFFFFF, dtype=np.uint32) for d in range(P)]
        filled = [np.zeros(int(plan["recv_total"][d]), dtype=bool) for d in range(P)]
        for s_rank, kv in enumerate(slices):
            part = plan["bin_to_part"][kv[:, 0] >> 24]
            for v in range(2 * P):
                run = kv[part == v]  # boolean indexing keeps input order: the stable partition
                assert len(run) == plan["part_counts"][s_rank, v]
                off = int(plan["part_offset"][s_rank, v])
                assert not filled[v // 2][off:off + len(run)].any()
                recv[v // 2][off:off + len(run)] = run
                filled[v // 2][off:off + len(run)] = True
        assert all(f.all() for f in filled)
        outs = []
        for d in range(P):
            ra = int(plan["recv_a"][d])
            outs.append(po.sort_pairs(recv[d][:ra].copy()))
            outs.append(po.sort_pairs(recv[d][ra:].copy()))
        whole = np.concatenate(slices)
        assert np.array_equal(np.concatenate(outs), po.sort_pairs(whole.copy())), (P, n, kind, a)


def test_numa_binding_is_best_effort_without_a_gpu():
    """bind_to_gpu_numa_node never raises: without a CUDA device (this container) or without a NUMA entry it returns None and
    leaves the process affinity alone."""
    from oclradixsort_b200.dist import bind_to_gpu_numa_node
    before = os.sched_getaffinity(0)
    res = bind_to_gpu_numa_node(0)
    assert res is None or (isinstance(res, dict) and res["cpus"] >= 1)
    if res is None:
        assert os.sched_getaffinity(0) == before


def test_exchange_plans_hold_their_invariants_on_random_histograms():
    """Property check (seeded random histograms incl. sparse, constant and single-source ones) of the two host plans every rank
    derives: digit ranges are contiguous and cover all 256 digits, halves cut each range once, send / receive counts add up,
    offsets tile every receive buffer exactly once in [A: source 0.. | B: source 0..] order."""
    from oclradixsort_b200.dist import plan_exchange_halves
    rng = np.random.default_rng(99)
    for trial in range(120):
        P = int(rng.integers(1, 17))
        mode = trial % 4
        hist = rng.integers(0, 1000, size=(P, 256)).astype(np.int64)
        if mode == 1:
            hist *= (rng.random((P, 256)) < 0.03)          # sparse: most digits empty
        elif mode == 2:
            hist[:] = 7                                     # every digit equally full
        elif mode == 3:
            hist[1:] = 0                                    # one source holds everything
        a = int(rng.choice([0, 250, 500, 900, 1000]))
        base, plan = plan_exchange(hist), plan_exchange_halves(hist, a)
        edges, mids = plan["edges"], plan["mids"]
        assert edges == base["edges"] and edges[0] == 0 and edges[-1] == 256 and all(x <= y for x, y in zip(edges, edges[1:]))
        assert all(edges[d] <= mids[d] <= edges[d + 1] for d in range(P))
        assert np.array_equal(plan["bin_to_part"] // 2, base["bin_to_rank"])
        assert np.array_equal(plan["part_counts"][:, 0::2] + plan["part_counts"][:, 1::2], base["send_counts"])
        assert np.array_equal(plan["recv_total"], base["recv_total"]) and int(plan["recv_total"].sum()) == int(hist.sum())
        for d in range(P):  # the (source, half) runs tile destination d's buffer without gaps or overlap, A before B, sources in order
            runs = sorted((int(plan["part_offset"][s_, 2 * d + h]), int(plan["part_counts"][s_, 2 * d + h]), h, s_) for h in (0, 1) for s_ in range(P))
            at = 0
            for off, cnt, h, s_ in runs:
                if cnt:
                    assert off == at, (trial, d, h, s_)
                    at += cnt
            assert at == int(plan["recv_total"][d])
            order = [(h, s_) for off, cnt, h, s_ in runs if cnt]
            assert order == sorted(order)
