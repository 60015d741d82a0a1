"""tools/sweep.py -- developer probe: time the variants selected through the B200RS_*_VARIANT knobs.
usage: sweep.py [log2n] [keys=0,1,..] [pairs=0,1,..] [scan=0,1,..]   (not the bench; bench.py is)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oclradixsort_b200 as ob

PEAK = 6555.2


def main():
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
    sel = {"keys": [0], "pairs": [0], "scan": [0]}
    dist = "uniform"
    for a in sys.argv[2:]:
        k, v = a.split("=")
        if k == "dist":
            dist = v  # uniform | sorted (key = i * 2^32/n, pairs: value = i)
            continue
        sel[k] = [int(x) for x in v.split(",")] if v else []
    n = 1 << log2n
    torch.cuda.set_device(0)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        d = ob.DeviceUtils.allocate(ob.TYPE_CL, 0, cuda_stream=st.cuda_stream)
        p = ob.Pprims()
        g = torch.Generator(device="cuda").manual_seed(1)
        for what, width, env in (("keys", 1, "B200RS_KEYS_VARIANT"), ("pairs", 2, "B200RS_PAIRS_VARIANT")):
            if not sel[what]:
                continue
            src = torch.randint(-2**31, 2**31, (n, width), device="cuda", dtype=torch.int32, generator=g)
            if dist == "sorted":
                src[:, 0] = (torch.arange(n, device="cuda", dtype=torch.int64) * (2**32 // n)).to(torch.int32)
                if width == 2: src[:, 1] = torch.arange(n, device="cuda", dtype=torch.int32)
            work = torch.empty_like(src)
            ref = None
            buf = ob.Buffer(d, n, np.uint32 if width == 1 else ob.PAIR_DTYPE, ptr=work.data_ptr())
            for v in sel[what]:
                os.environ[env] = str(v)
                times = []
                for it in range(6):
                    work.copy_(src)
                    if it == 5: d.toggleProfiling(True)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(st); p.radixSort(d, buf, n, 32); e1.record(st)
                    st.synchronize()
                    times.append(e0.elapsed_time(e1))
                prof = d.readProfile(); d.toggleProfiling(False)
                # all variants must agree bit for bit (variant 0 of this run is the yardstick)
                if ref is None: ref = work.clone(); ok = "ref"
                else: ok = "same" if torch.equal(ref, work) else "DIFFERENT"
                t = min(times[1:5]); bpk = 36 * width
                ks = ", ".join(f"{e['kernel'].split('_')[0]}{e['kernel'][-1]} {e['ms']:.3f}" for e in prof)
                print(f"{what} v{v} {dist} 2^{log2n}: {t:.3f} ms {n/t/1e6:.1f} Gelem/s {n*bpk/t/1e6/PEAK:.1%} [{ks}] {ok}", flush=True)
            os.environ.pop(env, None)
            del src, work, ref
        if sel["scan"]:
            s = torch.randint(-2**31, 2**31, (n,), device="cuda", dtype=torch.int32, generator=g)
            o = torch.empty_like(s)
            want = (torch.cumsum(s.to(torch.int64), 0) - s).to(torch.int32)
            sb, db = ob.Buffer(d, n, np.uint32, ptr=s.data_ptr()), ob.Buffer(d, n, np.uint32, ptr=o.data_ptr())
            for v in sel["scan"]:
                os.environ["B200RS_SCAN_VARIANT"] = str(v)
                times = []
                for it in range(8):
                    o.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(st); p.scan(d, db, sb, n); e1.record(st); st.synchronize()
                    times.append(e0.elapsed_time(e1))
                t = min(times[1:])
                print(f"scan v{v} 2^{log2n}: {t:.3f} ms {n/t/1e6:.1f} Gelem/s {n*8/t/1e6/PEAK:.1%} {'ok' if torch.equal(o, want) else 'WRONG'}", flush=True)
            os.environ.pop("B200RS_SCAN_VARIANT", None)
        p.release()


if __name__ == "__main__":
    main()
