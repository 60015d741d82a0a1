"""tools/host_sweep.py -- BASELINE.json configs[0]: the reference unit test's size sweep
(Demo.Sort32 / Demo.SortKeyValue inputs, n = 1K .. 1024K, srand(123); UnitTest/main.cpp:105-171) timed on the reference's
own Host-backend path (one host core; oracle/_ref when built, else the restated oracle) and on the B200 path for the same
inputs: device-resident (CUDA events, best of 20) and through the host-buffer C-ABI call (pinned memory, copies included).
Every GPU result is compared with the CPU result and with the reference's hash table (tests/golden/reference_hashes.json).
Prints a markdown table.  Checker use of oracle/ only (this is a measurement tool, not the product path)."""
import ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oclradixsort_b200 as ob
from oclradixsort_b200._lib import check, lib
from oracle import pyoracle as po

golden = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "reference_hashes.json")))
want_hash = {("keys", e["n"]): e["out"] for e in golden["sort32"]}
want_hash.update({("pairs", e["n"]): e["out"] for e in golden["sortkeyvalue"]})
kind = "reference (oracle/_ref)" if po.have_ref() else "restated oracle"
st = torch.cuda.Stream()
print(f"CPU column: {kind}, 1 thread, best of 3.  GPU: {torch.cuda.get_device_name(0)}\n")
print("| case | n | CPU ms | CPU Mkeys/s | GPU device-resident ms | GPU Mkeys/s | GPU incl. host copies ms | speed-up incl. copies | bit-exact, hash |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---|")
with torch.cuda.stream(st):
    d = ob.DeviceUtils.allocate(ob.TYPE_CL, 0, cuda_stream=st.cuda_stream)
    p = ob.Pprims()
    for what, sizes, gen in (("keys", [e["n"] for e in golden["sort32"]], po.gen_sort32), ("pairs", [e["n"] for e in golden["sortkeyvalue"]], po.gen_keyvalue)):
        for n in sizes:
            data = gen(n)
            cpu_t, cpu_out = 1e9, None
            for _ in range(3):
                w = data.copy()
                if po.have_ref():
                    t = po.ref_time_hostbackend(w, what == "pairs")
                else:
                    t0 = time.perf_counter(); w = po.sort_pairs(w) if what == "pairs" else po.sort_u32(w); t = time.perf_counter() - t0
                cpu_t, cpu_out = min(cpu_t, t), w
            dtype = ob.PAIR_DTYPE if what == "pairs" else np.uint32
            src = torch.from_numpy(data.view(np.uint8).copy()).cuda()
            work = torch.empty_like(src)
            buf = ob.Buffer(d, n, dtype, ptr=work.data_ptr())
            best = 1e9
            for _ in range(21):
                work.copy_(src)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st); p.radixSort(d, buf, n, 32); e1.record(st); st.synchronize()
                best = min(best, e0.elapsed_time(e1))
            got = work.cpu().numpy().view(dtype)
            pinned = ctypes.c_void_p()
            check(lib().b200rs_host_alloc(d.handle, data.nbytes, ctypes.byref(pinned)), "b200rs_host_alloc")
            host = np.ctypeslib.as_array(ctypes.cast(pinned, ctypes.POINTER(ctypes.c_uint8)), shape=(data.nbytes,))
            fn = lib().b200rs_sort_pairs_u32_host if what == "pairs" else lib().b200rs_sort_keys_u32_host
            e2e = 1e9
            for _ in range(6):
                host[:] = data.view(np.uint8)
                t0 = time.perf_counter(); check(fn(d.handle, pinned, n, 32), "host sort"); e2e = min(e2e, time.perf_counter() - t0)
            ok = np.array_equal(got, cpu_out.view(dtype)) and np.array_equal(host.view(dtype), cpu_out.view(dtype)) and f"{po.fnv1a64(got):016x}" == want_hash[(what, n)]
            check(lib().b200rs_host_free(d.handle, pinned), "b200rs_host_free")
            print(f"| {what} | {n} | {cpu_t*1e3:.3f} | {n/cpu_t/1e6:.1f} | {best:.4f} | {n/best/1e3:.0f} | {e2e*1e3:.4f} | {cpu_t/e2e:.1f} x | {'yes' if ok else 'NO'} |", flush=True)
    p.release()
