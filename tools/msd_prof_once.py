"""tools/msd_prof_once.py [log2n] [kind] -- two key sorts through the MSD entry point (the second one is what ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msd_probe import Runner, make
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
kind = sys.argv[2] if len(sys.argv) > 2 else "uniform"
n = 1 << log2n
r = Runner()
g = torch.Generator(device="cuda").manual_seed(3)
with torch.cuda.stream(r.st):
    src = make(kind, n, g)
    for rep in range(2):
        work = src.clone()
        r.st.synchronize()
        used = r.sort_msd(work, n)
        r.st.synchronize()
print("used", used)
