"""tools/prof_dist.py <log2n> <kind> -- one warm-up and one profiled key sort of a given distribution (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oclradixsort_b200 as ob
from key_distributions import make

log2n = int(sys.argv[1]); kind = sys.argv[2]
n = 1 << log2n
d = ob.DeviceUtils.allocate(ob.TYPE_CL, 0)
p = ob.Pprims()
g = torch.Generator(device="cuda").manual_seed(7)
src = make(kind, n, g)
work = torch.empty_like(src)
buf = ob.Buffer(d, n, np.uint32, ptr=work.data_ptr())
for rep in range(2):
    work.copy_(src); torch.cuda.synchronize()
    p.radixSort(d, buf, n, 32); d.waitForCompletion()
p.release()
