"""tools/prof_once.py -- run each hot-path entry point once (after one warm-up) for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oclradixsort_b200 as ob

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
which = sys.argv[2] if len(sys.argv) > 2 else "keys,pairs,scan"
n = 1 << log2n
d = ob.DeviceUtils.allocate(ob.TYPE_CL, 0)
p = ob.Pprims()
g = torch.Generator(device="cuda").manual_seed(1)
for rep in range(2):
    if "keys" in which:
        k = torch.randint(-2**31, 2**31, (n,), device="cuda", dtype=torch.int32, generator=g)
        torch.cuda.synchronize()
        p.radixSort(d, ob.Buffer(d, n, np.uint32, ptr=k.data_ptr()), n); d.waitForCompletion()
    if "pairs" in which:
        kv = torch.randint(-2**31, 2**31, (n, 2), device="cuda", dtype=torch.int32, generator=g)
        torch.cuda.synchronize()
        p.radixSort(d, ob.Buffer(d, n, ob.PAIR_DTYPE, ptr=kv.data_ptr()), n); d.waitForCompletion()
    if "scan" in which:
        s = torch.randint(-2**31, 2**31, (n,), device="cuda", dtype=torch.int32, generator=g)
        torch.cuda.synchronize()
        b = ob.Buffer(d, n, np.uint32, ptr=s.data_ptr())
        p.scan(d, b, b, n); d.waitForCompletion()
p.release()
