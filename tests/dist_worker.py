"""Worker for tests/test_gpu_dist.py: run under torchrun, one rank per GPU (NCCL).  Every rank generates its slice
from (seed, rank), the partitioned sort runs, rank 0 gathers everything on the CPU and compares the rank-order
concatenation with the oracle's stable sort of the concatenated input (bit-exact)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oclradixsort_b200 as ob  # noqa: E402
from oclradixsort_b200.dist import DistributedPairSorter  # noqa: E402


def make_input(kind, rank, n):
    rng = np.random.default_rng(7 + rank)
    keys = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    if kind == "lowentropy":
        keys = (keys & np.uint32(0x0F00000F)) * np.uint32(0x11)
    elif kind == "skewtop":
        keys = keys >> np.uint32(2 * rank)
    elif kind == "allequal":
        keys[:] = 0xDEADBEEF
    elif kind == "hotdigit":     # one top digit holds everything: digit ranges cannot balance it
        keys = (keys & np.uint32(0x00FFFFFF)) | np.uint32(0x5A000000)
    elif kind == "and3":
        keys = keys & rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32) & rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    kv = np.empty((n, 2), dtype=np.uint32)
    kv[:, 0], kv[:, 1] = keys, np.arange(n, dtype=np.uint32) + np.uint32(rank << 26)
    return kv


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = ob.DeviceUtils.allocate(ob.TYPE_CL, local, cuda_stream=torch.cuda.current_stream().cuda_stream)
    pp = ob.Pprims()
    ok = True
    modes = [("nccl", "dest"), ("p2p", "dest"), ("p2p", "bins")]
    cases = [("uniform", (1 << 20) + 17 * rank), ("lowentropy", 300_000), ("skewtop", 500_001)]
    # skewed inputs with the DEFAULT slack: the digit-range plan overflows and the sort re-plans with exact splitters
    skewed = [("allequal", 400_000), ("hotdigit", 300_001), ("and3", (1 << 19) + 5)]
    runs = [(c, m, float(world) + 0.5) for c in cases for m in modes] + [(c, m, 1.25) for c in skewed for m in modes[:2]]
    # receive capacity >= 2^22 pairs: the PIPELINED form of b200rs_dist_sort_pairs_u32 (two halves per destination, the second
    # one moved by copy engines while the first is sorted); ragged sizes, an input whose plan overflows (splitter fallback)
    # and one whose top-digit ranges cannot be halved evenly
    runs += [(("uniform", 3_000_001 + 4099 * rank), modes[1], 1.5), (("skewtop", 2_600_003), modes[1], float(world) + 0.5), (("and3", 3_500_001), modes[1], 1.25),
             (("lowentropy", 4_300_000 + rank), modes[1], float(world) + 0.5)]
    for (kind, n), (exchange, layout), slack in runs:
        kv = make_input(kind, rank, n)
        src = torch.from_numpy(kv.view(np.int64).reshape(-1).copy()).cuda()
        sorter = DistributedPairSorter(dev, pp, n + 64, dist, slack=slack, exchange=exchange, layout=layout)
        out, m = sorter.sort(src, n)
        out2, m2 = sorter.sort(src, n)  # a second call reuses the receive buffers
        assert m2 == m
        torch.cuda.synchronize()
        sizes = [None] * world
        dist.all_gather_object(sizes, (n, m))
        outs = [None] * world
        dist.gather_object(out.cpu().numpy().view(np.uint32).reshape(m, 2), outs if rank == 0 else None, dst=0)
        if rank == 0:
            from oracle import pyoracle as po
            whole = np.concatenate([make_input(kind, r, sizes[r][0]) for r in range(world)])
            same = np.array_equal(np.concatenate(outs), po.sort_pairs(whole))
            print(f"dist {kind} {exchange}/{layout} slack {slack}: per-rank in/out {sizes} bit-exact={same}", flush=True)
            ok = ok and same
        sorter.release()
    pp.release()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
