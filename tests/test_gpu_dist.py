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


def test_partitioned_sort_two_gpus_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    world = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    assert r.stdout.count("bit-exact=True") == 9, r.stdout[-2000:]  # 3 inputs x {nccl, p2p/dest, p2p/bins}
