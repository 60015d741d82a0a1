"""tools/msd_probe.py [check] [perf log2n ...] -- developer probe of the key-only MSD path (csrc/b200rs_msd.cuh).
check: b200rs_sort_keys_u32_msd against torch.sort over sizes x distributions (bit-exact), reports whether the path applied.
perf:  per-kernel times (the library's event log) and the share of the HBM roofline (36 B/key) at the given sizes.
Not the bench (bench.py is)."""
import sys, os, ctypes, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oclradixsort_b200 as ob
from oclradixsort_b200._lib import check, lib

try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0


def make(kind, n, g):
    u = lambda: torch.randint(-2**31, 2**31, (n,), device="cuda", dtype=torch.int32, generator=g)
    i64 = lambda: torch.arange(n, device="cuda", dtype=torch.int64)
    wrap = lambda x: (((x & 0xFFFFFFFF) ^ 0x80000000) - 0x80000000).to(torch.int32)  # u32 value (in an int64) -> the same bits as int32
    if kind == "uniform": return u()
    if kind == "sorted": return wrap(i64() * max(1, 2**32 // n))
    if kind == "reversed": return make("sorted", n, g).flip(0).contiguous()
    if kind == "and3": return u() & u() & u()
    if kind == "allequal": return torch.full((n,), 0x5A5A5A5A, device="cuda", dtype=torch.int32)
    if kind == "distinct16": return torch.randint(0, 16, (n,), device="cuda", dtype=torch.int32, generator=g) * 0x01010101 * 7
    if kind == "dups_low": return (u() & -65536) | 0x1234          # every bucket holds copies of ONE value: counting overflows -> robust route
    if kind == "dups8": return (u() & ~0x0700)                     # pairs of bins merged: more equal keys per bucket
    if kind == "low16": return u() & 0xFFFF                         # one bucket holds everything: not eligible
    if kind == "staircase": return wrap((i64() // 37) * 37 * max(1, 2**32 // n))   # runs of 37 equal keys, presorted
    if kind == "blocks": return wrap(((i64() * 2654435761) & 0xFFFF0000) | (i64() & 0xFFFF))  # scrambled buckets, ordered inside
    raise ValueError(kind)


def sort_ref(x):
    return (torch.sort(x.to(torch.int64) & 0xFFFFFFFF).values).to(torch.int64)


class Runner:
    def __init__(self):
        self.st = torch.cuda.Stream()
        with torch.cuda.stream(self.st):
            self.d = ob.DeviceUtils.allocate(ob.TYPE_CL, 0, cuda_stream=self.st.cuda_stream)
        self.temp = None

    def sort_msd(self, t, n):
        need = ctypes.c_size_t(0)
        used = ctypes.c_int(-1)
        fn = lib().b200rs_sort_keys_u32_msd
        check(fn(self.d.handle, None, n, None, ctypes.byref(need), ctypes.byref(used)), "size query")
        if self.temp is None or self.temp.numel() < need.value:
            self.temp = torch.empty(need.value + 256, device="cuda", dtype=torch.uint8)
        tp = (self.temp.data_ptr() + 255) // 256 * 256
        have = ctypes.c_size_t(need.value)
        check(fn(self.d.handle, ctypes.c_void_p(t.data_ptr()), n, ctypes.c_void_p(tp), ctypes.byref(have), ctypes.byref(used)), "sort_keys_u32_msd")
        return used.value


def do_check(r):
    g = torch.Generator(device="cuda").manual_seed(5)
    sizes = [2, 3, 5, 31, 257, 1000, 4096, 4097, 8191, 8192, 8193, 65537, (1 << 18) + 3, (1 << 20) - 1, (1 << 22) + 5, 1 << 24, (1 << 25) + 7, 1 << 26]
    kinds = ["uniform", "sorted", "reversed", "and3", "allequal", "distinct16", "dups_low", "dups8", "low16", "staircase", "blocks"]
    bad = 0
    with torch.cuda.stream(r.st):
        for n in sizes:
            row = []
            for kind in kinds:
                src = make(kind, n, g)
                want = sort_ref(src)
                work = src.clone()
                used = r.sort_msd(work, n)
                r.st.synchronize()
                ok = torch.equal(work.to(torch.int64) & 0xFFFFFFFF, want)
                bad += 0 if ok else 1
                row.append(f"{kind}:{'msd' if used else 'lsd'}:{'ok' if ok else 'WRONG'}")
                del src, want, work
            print(f"n={n}: " + " ".join(row), flush=True)
    print("CHECK", "FAILED" if bad else "passed", bad, flush=True)
    return bad


def do_perf(r, sizes, kinds):
    g = torch.Generator(device="cuda").manual_seed(7)
    with torch.cuda.stream(r.st):
        for log2n in sizes:
            n = 1 << log2n
            for kind in kinds:
                src = make(kind, n, g)
                work = torch.empty_like(src)
                ts, used = [], 0
                for it in range(5):
                    work.copy_(src)
                    if it == 4: r.d.toggleProfiling(True)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(r.st); used = r.sort_msd(work, n); e1.record(r.st); r.st.synchronize()
                    ts.append(e0.elapsed_time(e1))
                prof = r.d.readProfile(); r.d.toggleProfiling(False)
                t = min(ts[1:4])
                ok = torch.equal(work.to(torch.int64) & 0xFFFFFFFF, sort_ref(src))
                ks = ", ".join(f"{e['kernel'].replace('msd_', '').replace('_keys', '')} {e['ms']:.3f}" for e in prof)
                print(f"2^{log2n} {kind:10s} {'msd' if used else 'lsd'} {t:.3f} ms {n/t/1e6:.1f} Gkeys/s {36*n/t/1e6/PEAK:.1%} [{ks}] {'ok' if ok else 'WRONG'}", flush=True)
                del src, work


if __name__ == "__main__":
    r = Runner()
    args = sys.argv[1:] or ["check"]
    rc = 0
    if "check" in args:
        rc = do_check(r)
    if "perf" in args:
        i = args.index("perf")
        sizes = [int(a) for a in args[i + 1:] if a.isdigit()] or [28]
        kinds = [a for a in args[i + 1:] if not a.isdigit()] or ["uniform", "sorted", "reversed", "and3", "blocks"]
        do_perf(r, sizes, kinds)
    sys.exit(1 if rc else 0)
