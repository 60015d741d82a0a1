"""GPU tests of the multi-GPU sort: its single-GPU building blocks against numpy, and (when the box has >= 2 GPUs)
the whole partitioned sort under torchrun + NCCL against the oracle."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_histogram_and_stable_partition_match_numpy():
    import torch

    import oclradixsort_b200 as ob
    from oclradixsort_b200.dist import CudaLocalOps
    d = ob.DeviceUtils.allocate(ob.TYPE_CL, 0, cuda_stream=torch.cuda.current_stream().cuda_stream)
    p = ob.Pprims()
    ops = CudaLocalOps(d, p)
    rng = np.random.default_rng(3)
    for n in (1, 6143, 6144, 6145, 1_000_003):
        kv = np.empty((n, 2), dtype=np.uint32)
        kv[:, 0] = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
        kv[:, 1] = np.arange(n, dtype=np.uint32)
        src = torch.from_numpy(kv.view(np.int64).reshape(-1).copy()).cuda()
        hist = ops.histogram(src, n).cpu().numpy()
        top = kv[:, 0] >> 24
        assert np.array_equal(hist, np.bincount(top, minlength=256))
        edges = np.sort(rng.integers(0, 257, size=7))
        lut = np.searchsorted(edges, np.arange(256), side="right").astype(np.uint8)  # 8 contiguous parts, some may be empty
        counts = np.bincount(lut[top], minlength=8)
        dst = ops.empty(n)
        ops.partition(src, dst, n, lut, counts)
        torch.cuda.synchronize()
        got = dst[:n].cpu().numpy().view(np.uint32).reshape(n, 2)
        want = kv[np.argsort(lut[top], kind="stable")]
        assert np.array_equal(got, want), n
    ops.release()
    p.release()
    ob.DeviceUtils.deallocate(d)


def test_device_plan_matches_host_plan():
    import torch

    import oclradixsort_b200 as ob
    from oclradixsort_b200._lib import check, lib
    from oclradixsort_b200.dist import plan_exchange
    d = ob.DeviceUtils.allocate(ob.TYPE_CL, 0, cuda_stream=torch.cuda.current_stream().cuda_stream)
    rng = np.random.default_rng(11)
    for P in (1, 2, 3, 8):
        for kind in ("uniform", "skew", "onebin", "empty"):
            hist = rng.integers(0, 5000, size=(P, 256)).astype(np.int64)
            if kind == "skew":
                hist[:, 10:20] *= 300
            elif kind == "onebin":
                hist[:] = 0
                hist[:, 77] = 1234
            elif kind == "empty":
                hist[:] = 0
            plan = plan_exchange(hist)
            for me in range(P):
                peers = np.arange(P, dtype=np.int64) * (1 << 40) + (1 << 30)
                g = torch.from_numpy(hist.reshape(-1).copy()).cuda()
                pd = torch.from_numpy(peers).cuda()
                lut = torch.zeros(256, dtype=torch.uint8, device="cuda")
                base = torch.zeros(256, dtype=torch.int64, device="cuda")
                cnts = torch.zeros(2, dtype=torch.int64, device="cuda")
                st = torch.ones(1, dtype=torch.int32, device="cuda")
                cap = int(plan["recv_total"].max())
                check(lib().b200rs_dist_plan(d.handle, ctypes.c_void_p(g.data_ptr()), P, me, ctypes.c_void_p(pd.data_ptr()), cap, 4242,
                                             ctypes.c_void_p(lut.data_ptr()), ctypes.c_void_p(base.data_ptr()), ctypes.c_void_p(cnts.data_ptr()),
                                             ctypes.c_void_p(st.data_ptr())), "b200rs_dist_plan")
                torch.cuda.synchronize()
                assert np.array_equal(lut.cpu().numpy(), plan["bin_to_rank"]), (P, kind)
                assert np.array_equal(base.cpu().numpy()[:P], peers + 8 * plan["recv_offset"][me]), (P, kind, me)
                assert cnts.cpu().tolist() == [4242, int(plan["recv_total"][me])] and int(st.item()) == 0
                if cap > 0:  # one pair less capacity than needed: aborted, nothing to scatter or sort
                    check(lib().b200rs_dist_plan(d.handle, ctypes.c_void_p(g.data_ptr()), P, me, ctypes.c_void_p(pd.data_ptr()), cap - 1, 4242,
                                                 ctypes.c_void_p(lut.data_ptr()), ctypes.c_void_p(base.data_ptr()), ctypes.c_void_p(cnts.data_ptr()),
                                                 ctypes.c_void_p(st.data_ptr())), "b200rs_dist_plan")
                    torch.cuda.synchronize()
                    assert cnts.cpu().tolist() == [0, 0] and int(st.item()) == 1
    ob.DeviceUtils.deallocate(d)


def test_device_halves_plan_matches_host_plan():
    """dist_plan_halves_kernel (the plan of the pipelined partitioned sort) against plan_exchange_halves() for every rank of
    P = 1 .. 16, balanced / skewed / one-digit / empty histograms and three shares of half A."""
    import torch

    import oclradixsort_b200 as ob
    from oclradixsort_b200._lib import check, lib
    from oclradixsort_b200.dist import plan_exchange_halves
    d = ob.DeviceUtils.allocate(ob.TYPE_CL, 0, cuda_stream=torch.cuda.current_stream().cuda_stream)
    rng = np.random.default_rng(12)
    stage_base = 1 << 45
    for P in (1, 2, 3, 8, 16):
        for kind in ("uniform", "skew", "onebin", "empty"):
            hist = rng.integers(0, 5000, size=(P, 256)).astype(np.int64)
            if kind == "skew":
                hist[:, 10:20] *= 300
            elif kind == "onebin":
                hist[:] = 0
                hist[:, 77] = 1234
            elif kind == "empty":
                hist[:] = 0
            for a in (500, 300, 1000):
                plan = plan_exchange_halves(hist, a)
                peers = np.arange(P, dtype=np.int64) * (1 << 40) + (1 << 30)
                cap = int(plan["recv_total"].max())
                for me in range(P):
                    g = torch.from_numpy(hist.reshape(-1).copy()).cuda()
                    pd = torch.from_numpy(peers).cuda()
                    lut = torch.zeros(256, dtype=torch.uint8, device="cuda")
                    base = torch.zeros(64, dtype=torch.int64, device="cuda")
                    cnts = torch.zeros(2, dtype=torch.int64, device="cuda")
                    st = torch.ones(1, dtype=torch.int32, device="cuda")
                    out = torch.zeros(52, dtype=torch.int64, device="cuda")
                    check(lib().b200rs_dist_plan_halves(d.handle, ctypes.c_void_p(g.data_ptr()), P, me, ctypes.c_void_p(pd.data_ptr()), cap, stage_base, 4242, a,
                                                        ctypes.c_void_p(lut.data_ptr()), ctypes.c_void_p(base.data_ptr()), ctypes.c_void_p(cnts.data_ptr()),
                                                        ctypes.c_void_p(st.data_ptr()), ctypes.c_void_p(out.data_ptr())), "b200rs_dist_plan_halves")
                    torch.cuda.synchronize()
                    o = out.cpu().numpy()
                    assert np.array_equal(lut.cpu().numpy(), plan["bin_to_part"]), (P, kind, a)
                    assert int(st.item()) == 0 and o[0] == 0
                    assert cnts.cpu().tolist() == [4242, int(plan["recv_total"][me])]
                    assert (o[1], o[2], o[3]) == (plan["recv_total"][me], plan["recv_a"][me], plan["recv_b"][me]), (P, kind, a, me)
                    got_base = base.cpu().numpy()[:2 * P]
                    running = 0
                    for dest in range(P):
                        final_b = peers[dest] + 8 * plan["part_offset"][me, 2 * dest + 1]
                        assert got_base[2 * dest] == peers[dest] + 8 * plan["part_offset"][me, 2 * dest]
                        assert o[36 + dest] == final_b
                        if dest == me:
                            assert got_base[2 * dest + 1] == final_b and o[20 + dest] == 0
                        else:  # staged: runs in destination order, each starting on a 128-byte boundary
                            assert got_base[2 * dest + 1] == stage_base + 8 * running and o[4 + dest] == running
                            assert o[20 + dest] == plan["part_counts"][me, 2 * dest + 1]
                            running += (int(plan["part_counts"][me, 2 * dest + 1]) + 15) // 16 * 16
                if cap > 0:  # one pair less capacity than needed: aborted
                    check(lib().b200rs_dist_plan_halves(d.handle, ctypes.c_void_p(g.data_ptr()), P, 0, ctypes.c_void_p(pd.data_ptr()), cap - 1, stage_base, 4242, a,
                                                        ctypes.c_void_p(lut.data_ptr()), ctypes.c_void_p(base.data_ptr()), ctypes.c_void_p(cnts.data_ptr()),
                                                        ctypes.c_void_p(st.data_ptr()), ctypes.c_void_p(out.data_ptr())), "b200rs_dist_plan_halves")
                    torch.cuda.synchronize()
                    assert cnts.cpu().tolist() == [0, 0] and int(st.item()) == 1 and int(out[0].item()) == 1
    ob.DeviceUtils.deallocate(d)


def test_partitioned_sort_two_gpus_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    world = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    assert r.stdout.count("bit-exact=True") == 19, r.stdout[-3000:]  # 3 inputs x {nccl, p2p/dest, p2p/bins} + 3 skewed inputs x {nccl, p2p/dest} + 4 pipelined


def test_cpp_caller_of_the_partitioned_sort(tmp_path):
    """tests/cpp/dist_dropin.cpp: Tahoe::Pprims::radixSortDistributed (b200rs_dist_sort_pairs_u32) driven from C++, one host
    thread per GPU, collectives = pthread barrier + peer copies; checked against std::stable_sort of the whole input."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    exe = os.path.join(ROOT, "tools", "_build", "dist_dropin")
    if not os.path.exists(exe):
        subprocess.run(["make", "-s", "-C", ROOT, "dist_test"], check=True)
    for world, n in ((2, 300007), (min(torch.cuda.device_count(), 4), 1 << 20), (2, 3_400_007)):  # (the last one: pipelined form)
        r = subprocess.run([exe, str(world), str(n)], cwd=tmp_path, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "DIST DROPIN OK" in r.stdout, (r.stdout + r.stderr)[-2000:]


def _device_ops():
    import torch

    import oclradixsort_b200 as ob
    from oclradixsort_b200.dist import CudaLocalOps
    d = ob.DeviceUtils.allocate(ob.TYPE_CL, 0, cuda_stream=torch.cuda.current_stream().cuda_stream)
    p = ob.Pprims()
    return ob, d, p, CudaLocalOps(d, p)


def test_exchange_kernel_matches_numpy_stable_partition():
    """b200rs_exchange_pairs (the few-parts kernel with bulk copies) on one GPU: the parts' base addresses point into one
    local buffer, so the result must equal numpy's stable partition; sizes around the tile shapes tried (2048 / 4096 pairs; the 3584-pair tile's edges are in
    tools/sanitize_small.py), 1..32 parts (4 and 5 part bits), runs
    that start on odd element boundaries (ragged 16-byte ends), empty parts."""
    import torch

    from oclradixsort_b200._lib import check, lib
    ob, d, p, ops = _device_ops()
    rng = np.random.default_rng(5)
    for n in (1, 2, 2047, 2048, 2049, 4095, 4096, 4097, 123457, 1_000_003):
        for parts in (1, 2, 3, 8, 16, 17, 32):
            kv = np.empty((n, 2), dtype=np.uint32)
            kv[:, 0] = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
            kv[:, 1] = np.arange(n, dtype=np.uint32)
            top = kv[:, 0] >> 24
            edges = np.sort(rng.integers(0, 257, size=parts - 1))
            lut = np.searchsorted(edges, np.arange(256), side="right").astype(np.uint8)
            counts = np.bincount(lut[top], minlength=parts)
            gap = 3  # every part starts 3 pairs after the previous one ends: odd and even destination phases
            starts = np.cumsum(counts + gap) - (counts + gap)
            src = torch.from_numpy(kv.view(np.int64).reshape(-1).copy()).cuda()
            dst = torch.full((n + gap * parts + 8,), -1, dtype=torch.int64, device="cuda")
            lut_d = torch.from_numpy(lut).cuda()
            base_d = torch.from_numpy((np.uint64(dst.data_ptr()) + 8 * starts.astype(np.uint64)).view(np.int64)).cuda()
            need = ctypes.c_size_t(0)
            fn = lib().b200rs_exchange_pairs
            check(fn(d.handle, None, n, 24, 8, None, None, parts, None, None, ctypes.byref(need)), "size")
            temp = torch.empty(need.value + 256, dtype=torch.uint8, device="cuda")
            tp = (temp.data_ptr() + 255) // 256 * 256
            check(fn(d.handle, ctypes.c_void_p(src.data_ptr()), n, 24, 8, ctypes.c_void_p(lut_d.data_ptr()), ctypes.c_void_p(base_d.data_ptr()), parts, None,
                     ctypes.c_void_p(tp), ctypes.byref(need)), "b200rs_exchange_pairs")
            torch.cuda.synchronize()
            got = dst.cpu().numpy()
            order = np.argsort(lut[top], kind="stable")
            want = np.full(got.shape, -1, dtype=np.int64)
            sorted_pairs = kv[order].view(np.int64).reshape(-1)
            at = 0
            for q in range(parts):
                want[starts[q]:starts[q] + counts[q]] = sorted_pairs[at:at + counts[q]]
                at += counts[q]
            assert np.array_equal(got, want), (n, parts)  # (also: nothing written outside the runs)
    ops.release(); p.release(); ob.DeviceUtils.deallocate(d)


def test_filtered_histograms_and_splitter_exchange_match_numpy():
    import torch
    ob, d, p, ops = _device_ops()
    rng = np.random.default_rng(6)
    for n in (1, 5000, 777_777):
        kv = np.empty((n, 2), dtype=np.uint32)
        kv[:, 0] = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32) & rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
        kv[:, 1] = np.arange(n, dtype=np.uint32)
        keys = kv[:, 0]
        src = torch.from_numpy(kv.view(np.int64).reshape(-1).copy()).cuda()
        for shift in (16, 8, 0):
            prefixes = np.unique(keys >> np.uint32(shift + 8))[:5].astype(np.uint32) if shift + 8 < 32 else np.zeros(1, dtype=np.uint32)
            prefixes = np.concatenate([prefixes, [np.uint32(0x00ABCDEF >> shift)]]).astype(np.uint32)
            h = ops.filtered_histograms(src, n, shift, prefixes).cpu().numpy().reshape(len(prefixes), 256)
            high = (keys.astype(np.uint64) >> np.uint64(shift + 8)).astype(np.uint32)
            digit = (keys >> np.uint32(shift)) & np.uint32(255)
            for j, pre in enumerate(prefixes):
                assert np.array_equal(h[j], np.bincount(digit[high == pre], minlength=256)), (n, shift, j)
        for parts in (1, 2, 5, 8):
            thr = np.sort(rng.integers(0, 2**32 + 1, size=parts - 1, dtype=np.uint64))
            if parts > 2:
                thr[1] = thr[0]  # two boundaries on the same key: an empty part
            part = np.zeros(n, dtype=np.int64)
            for t in thr:
                part += keys.astype(np.uint64) >= t
            counts = np.bincount(part, minlength=parts)
            starts = np.cumsum(counts + 1) - (counts + 1)
            dst = torch.full((n + parts + 8,), -1, dtype=torch.int64, device="cuda")
            ops.exchange_by_splitters(src, n, thr, np.uint64(dst.data_ptr()) + 8 * starts.astype(np.uint64))
            torch.cuda.synchronize()
            got = dst.cpu().numpy()
            sorted_pairs = kv[np.argsort(part, kind="stable")].view(np.int64).reshape(-1)
            want = np.full(got.shape, -1, dtype=np.int64)
            at = 0
            for q in range(parts):
                want[starts[q]:starts[q] + counts[q]] = sorted_pairs[at:at + counts[q]]
                at += counts[q]
            assert np.array_equal(got, want), (n, parts)
    ops.release(); p.release(); ob.DeviceUtils.deallocate(d)
