mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_prims.py -x -q) 2>&1 | tail -15 > gpurun_out/s22_pytest_prims.log
cd gpurun_out && timeout 120 ../tools/_build/prims_dropin > s22_prims_dropin.txt 2>&1; cd ..
timeout 300 python - > gpurun_out/s22_prims_perf.txt 2>&1 <<'PY'
import ctypes, numpy as np, torch
import oclradixsort_b200 as ob
from oclradixsort_b200._lib import lib, check
d = ob.DeviceUtils.allocate(ob.TYPE_CL); p = ob.Pprims()
n = 1 << 28
a, b = ob.Buffer(d, n, np.uint32), ob.Buffer(d, n, np.uint32)
d.toggleProfiling(True)
for _ in range(5):
    p.fill(d, a, 7, n); p.copy(d, b, a, n)
for e in d.readProfile():
    print(e["kernel"], f'{e["ms"]:.4f} ms', f'{e["bytes"]/e["ms"]/1e6:.0f} GB/s')
a.release(); b.release(); p.release(); ob.DeviceUtils.deallocate(d)
PY
