"""tools/key_distributions.py [log2n ...] -- BASELINE.json configs[3]: u32 key sort across key distributions on one B200.
Prints one markdown table row per (size, distribution): ms, Gkeys/s, fraction of the HBM roofline (36 B/key, or 20 B/key
for sortBits=16).  Results are verified on the device (sortedness on the sorted bits + multiset sum)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oclradixsort_b200 as ob

try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0


def make(kind, n, g):
    u = lambda: torch.randint(-2**31, 2**31, (n,), device="cuda", dtype=torch.int32, generator=g)
    if kind in ("uniform", "sortbits16"): return u()
    if kind == "and3": return u() & u() & u()
    if kind == "and4": return u() & u() & u() & u()
    if kind == "distinct16": return (torch.randint(0, 16, (n,), device="cuda", dtype=torch.int32, generator=g) * 0x01010101 * 7)
    if kind == "distinct2": return (torch.randint(0, 2, (n,), device="cuda", dtype=torch.int32, generator=g) * 0x7f3f1f0f)
    if kind == "allequal": return torch.full((n,), 0x5A5A5A5A, device="cuda", dtype=torch.int32)
    if kind == "sorted": return (torch.arange(n, device="cuda", dtype=torch.int64) * (2**32 // n) - 2**31).to(torch.int32) ^ torch.tensor(-2**31, device="cuda", dtype=torch.int32)
    if kind == "reversed": return make("sorted", n, g).flip(0).contiguous()
    raise ValueError(kind)


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [28]
    kinds = ["uniform", "and3", "and4", "distinct16", "distinct2", "allequal", "sorted", "reversed", "sortbits16"]
    st = torch.cuda.Stream()
    print("| n | distribution | sortBits | ms | Gkeys/s | of HBM roofline |\n|---|---|---:|---:|---:|---:|")
    with torch.cuda.stream(st):
        d = ob.DeviceUtils.allocate(ob.TYPE_CL, 0, cuda_stream=st.cuda_stream)
        p = ob.Pprims()
        g = torch.Generator(device="cuda").manual_seed(7)
        for log2n in sizes:
            n = 1 << log2n
            for kind in kinds:
                bits = 16 if kind == "sortbits16" else 32
                src = make(kind, n, g)
                work = torch.empty_like(src)
                buf = ob.Buffer(d, n, np.uint32, ptr=work.data_ptr())
                ts = []
                for it in range(4):
                    work.copy_(src)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(st); p.radixSort(d, buf, n, bits); e1.record(st); st.synchronize()
                    ts.append(e0.elapsed_time(e1))
                t = min(ts[1:])
                per_kernel = ""
                if os.environ.get("B200RS_TOOL_PROFILE"):  # per-kernel times of one more (untimed) sort
                    work.copy_(src); d.toggleProfiling(True); p.radixSort(d, buf, n, bits)
                    per_kernel = " " + ", ".join(f"{e['kernel'].replace('onesweep_keys_', '').replace('digit_', '')} {e['ms']:.3f}" for e in d.readProfile())
                    d.toggleProfiling(False)
                mask = (1 << bits) - 1
                k = work.to(torch.int64) & mask
                ok = bool((k[1:] >= k[:-1]).all()) and int(work.to(torch.int64).sum().item()) == int(src.to(torch.int64).sum().item())
                del k
                bpk = 4 + 8 * ((bits + 7) // 8)
                print(f"| 2^{log2n} | {kind} | {bits} | {t:.3f} | {n/t/1e6:.1f} | {n*bpk/t/1e6/PEAK:.1%} |{'' if ok else ' WRONG'}{per_kernel}", flush=True)
                del src, work
        p.release()


if __name__ == "__main__":
    main()
