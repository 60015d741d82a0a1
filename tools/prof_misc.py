"""tools/prof_misc.py -- one scan, one fill and one copy at 2^28 u32 (for ncu captures of the non-sort kernels)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oclradixsort_b200 as ob
n = 1 << 28
d = ob.DeviceUtils.allocate(ob.TYPE_CL, 0)
p = ob.Pprims()
s = torch.randint(-2**31, 2**31, (n,), device="cuda", dtype=torch.int32)
o = torch.empty_like(s)
torch.cuda.synchronize()
a, b = ob.Buffer(d, n, np.uint32, ptr=s.data_ptr()), ob.Buffer(d, n, np.uint32, ptr=o.data_ptr())
for rep in range(2):
    p.scan(d, b, a, n); p.fill(d, b, 7, n); p.copy(d, b, a, n); d.waitForCompletion()
p.release()
