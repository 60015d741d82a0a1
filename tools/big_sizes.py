"""tools/big_sizes.py -- one-off check at the sizes of BASELINE.json configs[4] shards and beyond 32-bit element counts:
2^31 pairs (16 GiB), 2^32 + 12345 keys (16 GiB), 2^32 + 7 scan elements.  Verified on the device by properties
(sortedness, stability through value = index order inside equal-key runs, multiset sums; scan against chunked torch.cumsum)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oclradixsort_b200 as ob

st = torch.cuda.Stream()
with torch.cuda.stream(st):
    d = ob.DeviceUtils.allocate(ob.TYPE_CL, 0, cuda_stream=st.cuda_stream)
    p = ob.Pprims()
    g = torch.Generator(device="cuda").manual_seed(11)

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); fn(); e1.record(st); st.synchronize()
        return e0.elapsed_time(e1)

    # ---- 2^31 pairs ----
    n = 1 << 31
    kv = torch.empty((n, 2), device="cuda", dtype=torch.int32)
    CH = 1 << 28
    for c in range(0, n, CH):
        kv[c:c + CH, 0] = torch.randint(0, 1 << 20, (CH,), device="cuda", dtype=torch.int32, generator=g) * 4099  # many duplicates
        kv[c:c + CH, 1] = torch.arange(c, c + CH, device="cuda", dtype=torch.int64).to(torch.int32)
    t_first = timed(lambda: p.radixSort(d, ob.Buffer(d, n, ob.PAIR_DTYPE, ptr=kv.data_ptr()), n, 32))  # includes the 16 GiB scratch allocation
    for c in range(0, n, CH):  # scramble again (odd multiplier: a bijection on u32 that keeps the duplicates), value = position
        kv[c:c + CH, 0] *= -1640531535
        kv[c:c + CH, 1] = torch.arange(c, c + CH, device="cuda", dtype=torch.int64).to(torch.int32)
    ksum = sum(int(kv[c:c + CH, 0].to(torch.int64).sum().item()) for c in range(0, n, CH))
    t = timed(lambda: p.radixSort(d, ob.Buffer(d, n, ob.PAIR_DTYPE, ptr=kv.data_ptr()), n, 32))
    ok = True
    prev_k = prev_v = None
    for c in range(0, n, CH):
        k = kv[c:c + CH, 0].to(torch.int64) & 0xFFFFFFFF
        v = kv[c:c + CH, 1].to(torch.int64) & 0xFFFFFFFF
        ok &= bool((k[1:] >= k[:-1]).all())
        same = k[1:] == k[:-1]
        ok &= bool((v[1:][same] > v[:-1][same]).all())
        if prev_k is not None:
            ok &= (int(k[0]) > prev_k) or (int(k[0]) == prev_k and int(v[0]) > prev_v)
        prev_k, prev_v = int(k[-1]), int(v[-1])
        ksum -= int(kv[c:c + CH, 0].to(torch.int64).sum().item())
        del k, v, same
    print(f"pairs 2^31: {t:.2f} ms, {n/t/1e6:.1f} Gpairs/s ({72*n/t/1e6/6549.1:.1%} of measured HBM), first call incl. scratch allocation {t_first:.1f} ms, {'OK' if ok and ksum == 0 else 'WRONG'}", flush=True)
    del kv
    torch.cuda.empty_cache()

    # ---- 2^32 + 12345 keys ----
    n = (1 << 32) + 12345
    keys = torch.empty(n, device="cuda", dtype=torch.int32)
    for c in range(0, n, CH):
        m = min(CH, n - c)
        keys[c:c + m] = torch.randint(-2**31, 2**31, (m,), device="cuda", dtype=torch.int32, generator=g)
    t_first = timed(lambda: p.radixSort(d, ob.Buffer(d, n, np.uint32, ptr=keys.data_ptr()), n, 32))
    for c in range(0, n, CH):
        keys[c:c + CH] *= -1640531535
    ksum = sum(int(keys[c:c + CH].to(torch.int64).sum().item()) for c in range(0, n, CH))
    t = timed(lambda: p.radixSort(d, ob.Buffer(d, n, np.uint32, ptr=keys.data_ptr()), n, 32))
    ok = True
    prev = -1
    for c in range(0, n, CH):
        k = keys[c:c + CH].to(torch.int64) & 0xFFFFFFFF
        ok &= bool((k[1:] >= k[:-1]).all()) and int(k[0]) >= prev
        prev = int(k[-1])
        ksum -= int(keys[c:c + CH].to(torch.int64).sum().item())
        del k
    print(f"keys 2^32+12345: {t:.2f} ms, {n/t/1e6:.1f} Gkeys/s ({36*n/t/1e6/6549.1:.1%}), first call {t_first:.1f} ms, {'OK' if ok and ksum == 0 else 'WRONG'}", flush=True)

    # ---- scan of 2^32 + 7 elements (reuse the buffer: values masked to 4 bits so chunk sums stay exact) ----
    n = (1 << 32) + 7
    keys &= 0xF
    src = keys[:n]
    dst = torch.empty(n, device="cuda", dtype=torch.int32)
    t_first = timed(lambda: p.scan(d, ob.Buffer(d, n, np.uint32, ptr=dst.data_ptr()), ob.Buffer(d, n, np.uint32, ptr=src.data_ptr()), n))
    t = timed(lambda: p.scan(d, ob.Buffer(d, n, np.uint32, ptr=dst.data_ptr()), ob.Buffer(d, n, np.uint32, ptr=src.data_ptr()), n))
    ok = True
    carry = 0
    for c in range(0, n, CH):
        s = src[c:c + CH].to(torch.int64)
        want = (torch.cumsum(s, 0) - s + carry) & 0xFFFFFFFF
        ok &= bool(((dst[c:c + CH].to(torch.int64) & 0xFFFFFFFF) == want).all())
        carry += int(s.sum().item())
        del s, want
    print(f"scan 2^32+7: {t:.2f} ms, {n/t/1e6:.1f} Gelem/s ({8*n/t/1e6/6549.1:.1%}), first call {t_first:.2f} ms, {'OK' if ok else 'WRONG'}", flush=True)
    p.release()
