"""tools/size_scaling.py -- per-kernel times of one pair sort and one key sort at 2^24 .. 2^30 (developer probe)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oclradixsort_b200 as ob
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    d = ob.DeviceUtils.allocate(ob.TYPE_CL, 0, cuda_stream=st.cuda_stream)
    p = ob.Pprims()
    g = torch.Generator(device="cuda").manual_seed(1)
    for what, width in (("pairs", 2), ("keys", 1)):
        for log2n in range(24, 31):
            n = 1 << log2n
            src = torch.randint(-2**31, 2**31, (n, width), device="cuda", dtype=torch.int32, generator=g)
            work = torch.empty_like(src)
            buf = ob.Buffer(d, n, np.uint32 if width == 1 else ob.PAIR_DTYPE, ptr=work.data_ptr())
            best = None
            for it in range(4):
                work.copy_(src); d.toggleProfiling(True); p.radixSort(d, buf, n, 32)
                prof = d.readProfile(); d.toggleProfiling(False)
                ps = [e["ms"] for e in prof if e["kernel"].startswith("onesweep")]
                best = min(ps) if best is None else min(best, min(ps))
            print(f"{what} 2^{log2n}: best scatter pass {best:.4f} ms = {2*n*4*width/best/1e6:.0f} GB/s; hist {prof[0]['ms']:.4f}", flush=True)
            del src, work
    p.release()
