"""tools/dist_perf.py -- torchrun worker: times the partitioned sort per exchange mode and per phase (dev probe)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import oclradixsort_b200 as ob
from oclradixsort_b200.dist import DistributedPairSorter

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
modes = [tuple(m.split("/")) for m in sys.argv[2].split(",")] if len(sys.argv) > 2 else [("nccl", "dest"), ("p2p", "dest"), ("p2p", "bins")]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 6
n = 1 << log2n
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    dev = ob.DeviceUtils.allocate(ob.TYPE_CL, local, cuda_stream=st.cuda_stream)
    pp = ob.Pprims()
    g = torch.Generator(device="cuda").manual_seed(5 + rank)
    kv = torch.empty((n, 2), device="cuda", dtype=torch.int32)
    chunk = 1 << 28  # generated in pieces: the temporaries of a 2^31-element randint / arange would double the footprint
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        kv[c0:c1, 0] = torch.randint(-2**31, 2**31, (c1 - c0,), device="cuda", dtype=torch.int32, generator=g)
        kv[c0:c1, 1] = torch.arange(c0, c1, device="cuda", dtype=torch.int64).to(torch.int32)
    src = kv.view(torch.int64).reshape(-1)
    for exchange, layout in modes:
        sorter = DistributedPairSorter(dev, pp, n, dist, exchange=exchange, layout=layout)
        ts = []
        for it in range(iters):
            dist.barrier(); st.synchronize()
            if it == iters - 1: dev.toggleProfiling(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); out, m = sorter.sort(src, n); e1.record(st); st.synchronize()
            ts.append(e0.elapsed_time(e1))
        prof = dev.readProfile(64); dev.toggleProfiling(False)
        t = torch.tensor([min(ts[1:-1])], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            ks = ", ".join(f"{e['kernel'].replace('onesweep_pairs_','p')}={e['ms']:.2f}" for e in prof)
            print(f"[{world} GPUs, 2^{log2n}/GPU] {exchange}/{layout}: {t.item():.2f} ms/step -> {world*n/t.item()/1e6:.1f} Gkeys/s total | kernels(ms): {ks} | sum={sum(e['ms'] for e in prof):.2f}", flush=True)
        sorter.release()
    # single-GPU local sort for reference
    del sorter
    w = src.clone(); ts = []
    for it in range(3):
        w.copy_(src); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); pp.radixSort(dev, ob.Buffer(dev, n, ob.PAIR_DTYPE, ptr=w.data_ptr()), n, 32); e1.record(st); st.synchronize(); ts.append(e0.elapsed_time(e1))
    if rank == 0: print(f"local sort alone: {min(ts[1:]):.2f} ms -> {n/min(ts[1:])/1e6:.1f} Gkeys/s", flush=True)
    pp.release()
dist.barrier(); dist.destroy_process_group()
