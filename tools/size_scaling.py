"""tools/size_scaling.py [scan] -- per-kernel times of one pair sort and one key sort at 2^24 .. 2^30, then the scan from 1K to 2^30
(with `scan`: only the scan).  Developer probe."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oclradixsort_b200 as ob
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    d = ob.DeviceUtils.allocate(ob.TYPE_CL, 0, cuda_stream=st.cuda_stream)
    p = ob.Pprims()
    g = torch.Generator(device="cuda").manual_seed(1)
    for what, width in (() if "scan" in sys.argv[1:] else (("pairs", 2), ("keys", 1))):
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
                if not ps:  # 32-bit key sorts from 201 M keys take the MSD pipeline (no scatter pass): tools/msd_probe.py times that
                    break
                best = min(ps) if best is None else min(best, min(ps))
            if best is None:
                print(f"{what} 2^{log2n}: MSD pipeline, " + ", ".join(f"{e['kernel']} {e['ms']:.3f}" for e in prof), flush=True)
                continue
            print(f"{what} 2^{log2n}: best scatter pass {best:.4f} ms = {2*n*4*width/best/1e6:.0f} GB/s; hist {prof[0]['ms']:.4f}", flush=True)
            del src, work
    p.release()

# ---- BASELINE config 3: exclusive scan, 1K .. 2^30, full-range inputs (wrap-around), checked against torch.cumsum ----
with torch.cuda.stream(st):
    d = ob.DeviceUtils.allocate(ob.TYPE_CL, 0, cuda_stream=st.cuda_stream)
    p = ob.Pprims()
    g = torch.Generator(device="cuda").manual_seed(2)
    print("| scan n | ms | Gelem/s | GB/s (8 B/elem) | total ok | result ok |\n|---|---:|---:|---:|---|---|")
    for log2n in list(range(10, 31, 2)) + [20.5]:
        n = (1 << int(log2n)) + (1 if log2n != int(log2n) else 0)  # 20.5 stands for 2^20 + 1 (the reference refuses >= 2^20)
        s = torch.randint(-2**31, 2**31, (n,), device="cuda", dtype=torch.int32, generator=g)
        o = torch.empty_like(s)
        sb, db = ob.Buffer(d, n, np.uint32, ptr=s.data_ptr()), ob.Buffer(d, n, np.uint32, ptr=o.data_ptr())
        best = 1e9
        for it in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); p.scan(d, db, sb, n); e1.record(st); st.synchronize()
            best = min(best, e0.elapsed_time(e1))
        total = p.scan(d, db, sb, n, sumOut=True)
        c = torch.cumsum(s.to(torch.int64), 0)
        want = (c - s).to(torch.int32)
        ok = bool(torch.equal(o, want)); tok = total == (int(c[-1].item()) & 0xFFFFFFFF)
        print(f"| {n} | {best:.4f} | {n/best/1e6:.1f} | {n*8/best/1e6:.0f} | {'yes' if tok else 'NO'} | {'yes' if ok else 'NO'} |", flush=True)
        del s, o, c, want
    p.release()
