"""tools/variant_sweep.py <log2n> <keys|pairs> <n_variants> -- time each compiled onesweep variant (dev knob env vars)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oclradixsort_b200 as ob

log2n = int(sys.argv[1]); what = sys.argv[2]; nvar = int(sys.argv[3])
n = 1 << log2n
peak = 6555.2
width = 1 if what == "keys" else 2
env = "B200RS_KEYS_VARIANT" if what == "keys" else "B200RS_PAIRS_VARIANT"
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    d = ob.DeviceUtils.allocate(ob.TYPE_CL, 0, cuda_stream=st.cuda_stream)
    p = ob.Pprims()
    g = torch.Generator(device="cuda").manual_seed(1)
    src = torch.randint(-2**31, 2**31, (n, width), device="cuda", dtype=torch.int32, generator=g)
    work = torch.empty_like(src)
    ref = None
    buf = ob.Buffer(d, n, np.uint32 if width == 1 else ob.PAIR_DTYPE, ptr=work.data_ptr())
    for v in range(nvar):
        os.environ[env] = str(v)
        times = []
        for it in range(5):
            work.copy_(src)
            if it == 4: d.toggleProfiling(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); p.radixSort(d, buf, n, 32); e1.record(st); st.synchronize()
            times.append(e0.elapsed_time(e1))
        prof = d.readProfile(); d.toggleProfiling(False)
        if ref is None: ref = work.clone()
        same = bool((work == ref).all())
        t = min(times[1:4]); bpk = 36 * width
        ps = [e["ms"] for e in prof if e["kernel"].startswith("onesweep")]
        print(f"{what} variant {v}: total {t:.3f} ms {n/t/1e6:.1f} G/s ({n*bpk/t/1e6/peak:.1%} of peak); onesweep pass avg {sum(ps)/len(ps):.3f} ms = {2*n*4*width/(sum(ps)/len(ps))/1e6/peak:.1%}; hist {prof[0]['ms']:.3f} ms; same_as_v0={same}")
    p.release()
