"""tools/xp_probe.py [log2n] [parts ...] -- the exchange kernel (b200rs_exchange_pairs) on ONE GPU, every part's base inside one
local buffer: what the SM side of the kernel can do when no link is involved (dev probe; with `once` as last argument: one
warm-up and one launch, for ncu)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oclradixsort_b200 as ob
from oclradixsort_b200._lib import check, lib

args = sys.argv[1:]
once = "once" in args
args = [a for a in args if a != "once"]
log2n = int(args[0]) if args else 28
parts_list = [int(a) for a in args[1:]] or [2, 8, 16, 32]
n = 1 << log2n
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    d = ob.DeviceUtils.allocate(ob.TYPE_CL, 0, cuda_stream=st.cuda_stream)
    g = torch.Generator(device="cuda").manual_seed(5)
    kv = torch.empty((n, 2), device="cuda", dtype=torch.int32)
    kv[:, 0] = torch.randint(-2**31, 2**31, (n,), device="cuda", dtype=torch.int32, generator=g)
    kv[:, 1] = torch.arange(n, device="cuda", dtype=torch.int32)
    src = kv.view(torch.int64).reshape(-1)
    dst = torch.empty(n + 64, dtype=torch.int64, device="cuda")
    fn = lib().b200rs_exchange_pairs
    for parts in parts_list:
        lut = (np.arange(256) * parts // 256).astype(np.uint8)
        top = torch.bincount(((kv[:, 0].to(torch.int64) & 0xFFFFFFFF) >> 24), minlength=256).cpu().numpy()
        counts = np.bincount(lut, weights=top, minlength=parts).astype(np.int64)
        starts = np.cumsum(counts) - counts
        lut_d = torch.from_numpy(lut).cuda()
        base_d = torch.from_numpy((np.uint64(dst.data_ptr()) + 8 * starts.astype(np.uint64)).view(np.int64)).cuda()
        need = ctypes.c_size_t(0)
        check(fn(d.handle, None, n, 24, 8, None, None, parts, None, None, ctypes.byref(need)), "size")
        temp = torch.empty(need.value + 256, dtype=torch.uint8, device="cuda")
        tp = (temp.data_ptr() + 255) // 256 * 256
        ts = []
        for it in range(2 if once else 5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            check(fn(d.handle, ctypes.c_void_p(src.data_ptr()), n, 24, 8, ctypes.c_void_p(lut_d.data_ptr()), ctypes.c_void_p(base_d.data_ptr()), parts, None,
                     ctypes.c_void_p(tp), ctypes.byref(need)), "b200rs_exchange_pairs")
            e1.record(st); st.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = min(ts[1:])
        ok = bool((((dst[:n] & 0xFFFFFFFF) >> 24)[1:] >= ((dst[:n] & 0xFFFFFFFF) >> 24)[:-1]).all()) if parts == 256 else True
        print(f"2^{log2n} pairs, {parts} local parts: {t:.3f} ms (memset of the look-back table included) = {16*n/t/1e9:.2f} TB/s read+write", flush=True)
