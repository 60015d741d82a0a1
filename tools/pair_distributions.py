"""tools/pair_distributions.py [log2n] -- u32/u32 pair sort (value = input index) across key distributions on one B200:
ms, Gpairs/s, fraction of the HBM roofline (72 B/pair).  Verified on the device: keys sorted, values increasing inside
equal-key runs (= the stable order, i.e. the reference's result)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oclradixsort_b200 as ob
from key_distributions import make, PEAK


def main():
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
    n = 1 << log2n
    kinds = sys.argv[2:] or ["uniform", "and3", "distinct16", "allequal", "sorted", "reversed"]
    st = torch.cuda.Stream()
    print("| n | distribution | ms | Gpairs/s | of HBM roofline | stable-sorted |\n|---|---|---:|---:|---:|---|")
    with torch.cuda.stream(st):
        d = ob.DeviceUtils.allocate(ob.TYPE_CL, 0, cuda_stream=st.cuda_stream)
        p = ob.Pprims()
        g = torch.Generator(device="cuda").manual_seed(7)
        for kind in kinds:
            src = torch.empty((n, 2), device="cuda", dtype=torch.int32)
            src[:, 0] = make(kind, n, g)
            src[:, 1] = torch.arange(n, device="cuda", dtype=torch.int32)
            work = torch.empty_like(src)
            buf = ob.Buffer(d, n, ob.PAIR_DTYPE, ptr=work.data_ptr())
            ts = []
            for it in range(4):
                work.copy_(src)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st); p.radixSort(d, buf, n, 32); e1.record(st); st.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = min(ts[1:])
            k = work[:, 0].to(torch.int64) & 0xFFFFFFFF
            v = work[:, 1].to(torch.int64) & 0xFFFFFFFF
            dk, dv = k[1:] - k[:-1], v[1:] - v[:-1]
            ok = bool((dk >= 0).all()) and bool(((dk > 0) | (dv > 0)).all()) and int(v.sum()) == n * (n - 1) // 2
            print(f"| 2^{log2n} | {kind} | {t:.3f} | {n/t/1e6:.1f} | {72*n/t/1e6/PEAK:.1%} | {'yes' if ok else 'NO'} |", flush=True)
            del src, work, k, v, dk, dv
        p.release()


if __name__ == "__main__":
    main()
