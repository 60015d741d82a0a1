"""tools/quick_perf.py -- developer probe: device-resident timings of the three entry points.
Not the bench (bench.py is); prints per-kernel times from the library's own event log."""
import sys, os, ctypes, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oclradixsort_b200 as ob

def main():
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
    n = 1 << log2n
    peak = 6555.2
    torch.cuda.set_device(0)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        d = ob.DeviceUtils.allocate(ob.TYPE_CL, 0, cuda_stream=st.cuda_stream)
        p = ob.Pprims()
        g = torch.Generator(device="cuda").manual_seed(1)
        for what, width in (("keys", 1), ("pairs", 2)):
            src = torch.randint(-2**31, 2**31, (n, width), device="cuda", dtype=torch.int32, generator=g)
            work = torch.empty_like(src)
            buf = ob.Buffer(d, n, np.uint32 if width == 1 else ob.PAIR_DTYPE, ptr=work.data_ptr())
            for bits in (32, 16):
                times = []
                for it in range(6):
                    work.copy_(src)
                    if it == 5: d.toggleProfiling(True)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(st); p.radixSort(d, buf, n, bits); e1.record(st)
                    st.synchronize()
                    times.append(e0.elapsed_time(e1))
                prof = d.readProfile(); d.toggleProfiling(False)
                t = min(times[1:5])
                passes = (bits + 7) // 8
                bpk = (4 + 8 * passes) * width
                print(f"{what} 2^{log2n} bits={bits}: {t:.3f} ms  {n/t/1e6:.1f} Gelem/s  {n*bpk/t/1e6:.0f} GB/s = {n*bpk/t/1e6/peak:.1%} of measured peak")
                for e in prof:
                    print(f"     {e['kernel']:28s} {e['ms']:.3f} ms  {e['bytes']/e['ms']/1e6:.0f} GB/s ({e['bytes']/e['ms']/1e6/peak:.1%})")
            del src, work
        s = torch.randint(-2**31, 2**31, (n,), device="cuda", dtype=torch.int32, generator=g)
        o = torch.empty_like(s)
        sb, db = ob.Buffer(d, n, np.uint32, ptr=s.data_ptr()), ob.Buffer(d, n, np.uint32, ptr=o.data_ptr())
        times = []
        for it in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); p.scan(d, db, sb, n); e1.record(st); st.synchronize()
            times.append(e0.elapsed_time(e1))
        t = min(times[1:])
        print(f"scan 2^{log2n}: {t:.3f} ms {n/t/1e6:.1f} Gelem/s {n*8/t/1e6:.0f} GB/s = {n*8/t/1e6/peak:.1%}")
        # torch copy as the yardstick
        for it in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); o.copy_(s); e1.record(st); st.synchronize()
        print(f"torch copy: {n*8/e0.elapsed_time(e1)/1e6:.0f} GB/s")
        p.release()

if __name__ == "__main__":
    main()
