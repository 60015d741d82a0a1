#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: Gkeys/s of the u32/u32 key-value radix sort.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log2-pairs-per-gpu L]

A "step" is one sort of one batch of synthetic pairs (uniform random u32 keys, value = index), 2^L pairs
per GPU (default 2^28 = the largest size of BASELINE.json configs[1]).  Every step sorts a fresh, unsorted
buffer that is already resident in HBM; buffers are 2 GiB each, far larger than the 126 MB L2, so nothing a
step reads is cached from the previous one.

  value      whole-job Gkeys/s with inputs resident in HBM, device time (CUDA events), max over ranks
  e2e        same metric through the C-ABI HOST-buffer call a drop-in caller makes (b200rs_sort_pairs_u32_host: ONE pinned
             host array copied in, sorted, copied back, all inside the timed region); batch_* = the pipelined batch
             entry (b200rs_sort_pairs_u32_host_batch, copies of neighbouring arrays overlap the sort); buffer_api_* =
             the reference caller's own sequence through adl::Buffer (getHostPtr / fill / returnHostPtr / radixSort /
             getHostPtr, UnitTest/main.cpp:118-139) via the Python mirror
  roofline   dominant kernel (one scatter pass): algorithmic bytes (2 x 8 B per pair) / its average launch
             time from the library's own CUDA events, against the measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline  the UNMODIFIED reference's Adl Host-backend sort (oracle/_ref, built from /root/reference) on
             this box's host CPU, 1 thread (the reference is serial), on a bounded sample

  keys       BASELINE.json configs[3] / north_star target #1: u32 KEY sort of 2^L uniform keys, device-resident, with its
             share of the HBM roofline at 36 B/key (first-class field; other distributions under extra)
  parity     every run (any N) checks its own output on the device, outside the timed region: keys sorted, values
             strictly increasing inside equal-key runs (value = global input index, so the stable result is unique)
             and two 64-bit multiset hashes equal to the input's <=> bit-exact with the oracle; N > 1
             also sorts a small input first and compares it bit for bit with the oracle on rank 0
--impl reference times that same reference CPU path for the same metric/config (rank 0 only under torchrun): the
same 2^L pairs per step, at most 5 timed steps so the run ends within minutes.
N > 1 (torchrun, one rank per GPU, NCCL): the partitioned sort of oclradixsort_b200.dist (b200rs_dist_sort_pairs_u32) --
top-digit histogram, all-gather of the histograms, on-device plan with two halves per destination, ONE exchange kernel
(half A: peer stores over NVLink into CUDA-IPC-mapped receive buffers; half B: staged locally, moved by copy engines
while half A is already being sorted on a second stream), local sorts; weak scaling (2^L pairs per GPU); config5 = the
same at BASELINE config 5's shard size (2^31 pairs per GPU).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Gkeys/s sort (u32/u32 key-value pairs, uniform keys)"
UNIT = "Gkeys/s"


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.004):
        super().__init__(daemon=True)
        self.index, self.period_s = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(self.period_s)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def reference_arm(args, rank: int, json_out) -> None:
    """The reference's own CPU implementation of the path, timed on this box's host cores, on the SAME config as our arm:
    2^L uniform pairs per step (about 6.8 s per step at L = 28 on one core -- the reference is serial).  The number of
    timed steps is capped so the run ends within minutes; the cap is reported in `steps`."""
    if rank != 0:
        return
    import numpy as np
    from oracle import pyoracle as po
    log2n = args.ref_log2_sample if args.ref_log2_sample is not None else args.log2_pairs_per_gpu
    n = 1 << log2n
    steps = min(args.steps, args.ref_max_steps) if log2n >= 26 else args.steps
    warmup = min(args.warmup, 1) if log2n >= 26 else args.warmup
    kind = "reference" if po.have_ref() else "port"
    rng = np.random.default_rng(1234)
    src = np.empty(n, dtype=po.PAIR_DTYPE)
    src["key"] = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    src["value"] = np.arange(n, dtype=np.uint32)

    def one_step(i) -> float:
        kv = src.copy()
        kv["key"] ^= np.uint32((0x9E3779B1 * (i + 1)) & 0xFFFFFFFF)  # a different key set per step (XOR keeps it a uniform permutation)
        if kind == "reference":
            return po.ref_time_hostbackend(kv, pairs=True)
        t0 = time.perf_counter()
        po.lib().oracle_sort_pairs(ctypes.c_void_p(kv.ctypes.data), n, 32)
        return time.perf_counter() - t0

    for i in range(warmup):
        one_step(i)
    times = [one_step(warmup + i) for i in range(steps)]
    total = sum(times)
    value = n * steps / total / 1e9
    sample = f"2^{log2n} uniform pairs per step, {steps} timed steps" + ("" if log2n == args.log2_pairs_per_gpu else f" (sample of the 2^{args.log2_pairs_per_gpu}-pair workload)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": 1e3 * total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic",
        "config": {"workload": f"kv_sort_u32u32_uniform_2^{args.log2_pairs_per_gpu}_pairs_per_gpu", "sort_bits": 32, "pairs_per_gpu": 1 << args.log2_pairs_per_gpu,
                   "pairs_per_step_measured": n, "steps_requested": args.steps,
                   "reference_path": "Adl Host backend: Pprims::radixSort -> RadixSort::sort (serial, 1 thread)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=json_out, flush=True)


def cpu_baseline(log2_sample: int):
    import numpy as np
    from oracle import pyoracle as po
    n = 1 << log2_sample
    rng = np.random.default_rng(99)
    kind = "reference" if po.have_ref() else "port"
    best = None
    for _ in range(2):
        kv = np.empty(n, dtype=po.PAIR_DTYPE)
        kv["key"] = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
        kv["value"] = np.arange(n, dtype=np.uint32)
        if kind == "reference":
            t = po.ref_time_hostbackend(kv, pairs=True)
        else:
            t0 = time.perf_counter()
            po.lib().oracle_sort_pairs(ctypes.c_void_p(kv.ctypes.data), n, 32)
            t = time.perf_counter() - t0
        best = t if best is None else min(best, t)
    return {"value": n / best / 1e9, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"2^{log2_sample} uniform pairs, best of 2, reference Host-backend Pprims::radixSort on 1 host thread"}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2-pairs-per-gpu", type=int, default=28)
    ap.add_argument("--ref-log2-sample", type=int, default=None, help="pairs per step of the CPU reference arm (default: the workload's own size)")
    ap.add_argument("--ref-max-steps", type=int, default=5, help="cap on the timed steps of the CPU reference arm at full size (6.8 s per step at 2^28)")
    ap.add_argument("--no-config5", action="store_true", help="N > 1: skip the extra block at BASELINE config 5's shard size (2^31 pairs per GPU)")
    ap.add_argument("--cpu-log2-sample", type=int, default=26, help="pairs of the cpu_baseline sample")
    ap.add_argument("--skip-extras", action="store_true", help="only the headline numbers (no key-only / scan extras)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    # Contract: rank 0 prints ONE JSON line on stdout.  Libraries write there too (NCCL prints "NCCL version ..." on stdout
    # when NCCL_DEBUG=VERSION, as on the GPU boxes), so file descriptor 1 is pointed at stderr for the whole run and the
    # JSON line goes to the saved descriptor.
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    if args.impl == "reference":
        reference_arm(args, rank, json_out)
        return

    import numpy as np
    import torch

    import oclradixsort_b200 as ob
    from oclradixsort_b200._lib import check, lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    numa = None
    if world > 1:
        import torch.distributed as dist
        from oclradixsort_b200.dist import bind_to_gpu_numa_node
        numa = bind_to_gpu_numa_node(local_rank)  # before any pinned allocation: every rank's host buffers next to its GPU
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = 1 << args.log2_pairs_per_gpu
    stream = torch.cuda.Stream()
    peak, peak_src = measured_hbm_peak()
    with torch.cuda.stream(stream):
        dev = ob.DeviceUtils.allocate(ob.TYPE_CL, local_rank, cuda_stream=stream.cuda_stream)
        pp = ob.Pprims()
        gen = torch.Generator(device="cuda").manual_seed(1000 + rank)
        nbuf = args.steps + args.warmup

        def fresh_pairs(count=None):
            """count uniform pairs; value = GLOBAL input index (rank * count + i, mod 2^32), so the stable result is unique."""
            count = n if count is None else count
            kv = torch.empty((count, 2), device="cuda", dtype=torch.int32)
            kv[:, 0] = torch.randint(-2**31, 2**31, (count,), device="cuda", dtype=torch.int32, generator=gen)  # uniform over all 32 bits
            kv[:, 1] = (torch.arange(count, device="cuda", dtype=torch.int64) + rank * count).to(torch.int32)
            return kv

        def multiset_hash(x64):
            """Two order-independent 64-bit sums (mod 2^64) of non-linear mixes of every 8-byte element."""
            a = x64 * -7046029254386353131  # 0x9E3779B97F4A7C15; int64 products wrap
            a = a ^ (a >> 29)
            b = (x64 ^ 0x632BE59BD9B4E019) * -4417276706812531889  # 0xC2B2AE3D27D4EB4F
            b = b ^ (b >> 31)
            return torch.stack([a.sum(), (b * (b | 1)).sum()])

        CHUNK = 1 << 26  # the checks walk the data in slices so that their int64 temporaries stay small

        def hash_chunks(x64):
            h = torch.zeros(2, device="cuda", dtype=torch.int64)
            for o in range(0, x64.numel(), CHUNK):
                h += multiset_hash(x64[o:o + CHUNK])
            return h

        def check_sorted_output(out64, hash_in, total_pairs, unique_values):
            """SURVEY.md 8c, O(n) on the device: sorted by key, values strictly increasing inside equal-key runs, same multiset
            as the input.  out64: this rank's sorted pairs (int64, key = low half).  Collective when dist is active."""
            m = out64.numel()
            ok = True
            for o in range(0, m, CHUNK):
                part = out64[o:min(m, o + CHUNK + 1)]  # one element of overlap: the edge between slices is checked too
                k = part & 0xFFFFFFFF
                ok = ok and (bool((k[1:] >= k[:-1]).all()) if part.numel() > 1 else True)
                if unique_values and part.numel() > 1:
                    v = (part >> 32) & 0xFFFFFFFF
                    same = k[1:] == k[:-1]
                    ok = ok and bool((v[1:][same] > v[:-1][same]).all())
                    del v, same
                del k
            h = hash_chunks(out64)
            first, last = (int(out64[0].item()), int(out64[-1].item())) if m else (0, 0)
            lo, hi = (first & 0xFFFFFFFF, last & 0xFFFFFFFF) if m else (2**32, -1)
            first_v, last_v = (first >> 32) & 0xFFFFFFFF, (last >> 32) & 0xFFFFFFFF
            if dist is not None:
                dist.all_reduce(h)  # sums mod 2^64
                stats = torch.tensor([m, lo, hi, int(ok), first_v, last_v], device="cuda", dtype=torch.int64)
                allstats = [torch.empty_like(stats) for _ in range(world)]
                dist.all_gather(allstats, stats)
                rows = [t.tolist() for t in allstats]
                nonempty = [r for r in rows if r[0] > 0]
                ok = all(r[3] == 1 for r in rows) and sum(r[0] for r in rows) == total_pairs
                for a_, b_ in zip(nonempty, nonempty[1:]):  # rank r's largest key <= rank r+1's smallest; equal keys keep value order across the edge
                    ok = ok and (a_[2] < b_[1] or (a_[2] == b_[1] and (not unique_values or a_[5] < b_[4])))
            else:
                ok = ok and m == total_pairs
            return ok and bool((h == hash_in).all())

        def input_hash(buf):
            h = hash_chunks(buf.view(torch.int64).reshape(-1))
            if dist is not None:
                dist.all_reduce(h)
            return h

        sorter = None
        if world > 1:
            from oclradixsort_b200.dist import DistributedPairSorter
            sorter = DistributedPairSorter(dev, pp, n, dist)

        # ---- N > 1: a small input first, gathered on rank 0 and compared bit for bit with the oracle (the checker) ----
        small_parity = None
        if sorter is not None:
            ns = 1 << 18
            small = fresh_pairs(ns)
            small[:, 0] &= 0x00FFFFFF if rank % 2 else -1  # a low-entropy half: equal keys across ranks exercise stability
            small64 = small.view(torch.int64).reshape(-1)
            out_s, m_s = sorter.sort(small64, ns)
            cnt = torch.tensor([m_s], device="cuda", dtype=torch.int64)
            cnts = [torch.empty_like(cnt) for _ in range(world)]
            dist.all_gather(cnts, cnt)
            cap = max(int(c.item()) for c in cnts)
            padded = torch.zeros(cap, device="cuda", dtype=torch.int64)
            padded[:m_s] = out_s[:m_s]
            outs = [torch.empty_like(padded) for _ in range(world)]
            ins = [torch.empty_like(small64) for _ in range(world)]
            dist.all_gather(outs, padded)
            dist.all_gather(ins, small64)
            if rank == 0:
                from oracle import pyoracle as po
                got = np.concatenate([o[: int(c.item())].cpu().numpy() for o, c in zip(outs, cnts)]).view(po.PAIR_DTYPE)
                want = po.sort_pairs(np.concatenate([i.cpu().numpy() for i in ins]).view(po.PAIR_DTYPE))
                small_parity = bool(np.array_equal(got, want))
                if not small_parity:
                    raise SystemExit("bench.py: the distributed sort of the small input differs from the oracle")
            del small, small64, padded, outs, ins

        bufs = [fresh_pairs() for _ in range(nbuf)]
        handles = [ob.Buffer(dev, n, ob.PAIR_DTYPE, ptr=b.data_ptr()) for b in bufs]
        inputs64 = [b.view(torch.int64).reshape(-1) for b in bufs]
        unique_values = world * n <= 2**32
        hash_last_in = input_hash(bufs[-1])

        last = {}

        def step(i):
            if sorter is None:
                pp.radixSort(dev, handles[i], n, 32)
            else:
                last["out"] = sorter.sort_async(inputs64[i], n)  # stream-ordered: no host round trip inside a step

        sampler = ClockSampler(local_rank)  # NVML is initialised here, well before the timed region: with 8 ranks doing it at once
        #                                     right before the first timed step, that step took 12.9 ms instead of 8.1
        for i in range(args.warmup):
            step(i)
        stream.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        launches0 = dev.launch_count()
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        if dist is not None:
            # ranks leave the host-side barrier milliseconds apart; a stream-ordered all-reduce lines the GPUs up so that
            # the first timed step does not include the other ranks' launch skew (the host barrier + synchronize stay)
            align = torch.zeros(1, device="cuda")
            dist.all_reduce(align)
        e0.record(stream)
        for i in range(args.warmup, nbuf):
            step(i)
            marks[i - args.warmup].record(stream)  # per-step split times (reported, not used for `value`)
        e1.record(stream)
        stream.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        clocks = sampler.stop()
        launches = dev.launch_count() - launches0
        ms = e0.elapsed_time(e1)
        step_ms = [(e0 if k == 0 else marks[k - 1]).elapsed_time(marks[k]) for k in range(args.steps)]
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
            dist.all_reduce(lt)
            launches = int(lt.item())
        ms_per_step = ms / args.steps
        value = world * n / (ms_per_step * 1e-3) / 1e9

        # ---- parity of the last timed step's output, on the device, outside the timed region ----
        if sorter is None:
            ok_sorted = check_sorted_output(inputs64[-1], hash_last_in, n, unique_values)
        else:
            m = sorter.finish()  # element count of the last step (raises if a rank's share overflowed)
            ok_sorted = check_sorted_output(last["out"][:m], hash_last_in, world * n, unique_values)
        if not ok_sorted:
            raise SystemExit("bench.py: output of the last timed step is not the stable sort of its input")
        parity = {"checked": "last timed step, on device: sorted by key, values increasing inside equal-key runs, two 64-bit multiset hashes equal to the input's",
                  "ok": True, "value_is_global_index": bool(unique_values)}
        if small_parity is not None:
            parity["small_input_vs_oracle"] = {"pairs": world << 18, "bit_exact": small_parity}

        # ---- roofline of the dominant kernel, from the library's per-launch CUDA events ----
        roofline = None
        extras = {}
        if sorter is None:
            dev.toggleProfiling(True)
            probe = [fresh_pairs() for _ in range(3)]
            for b in probe:
                pp.radixSort(dev, ob.Buffer(dev, n, ob.PAIR_DTYPE, ptr=b.data_ptr()), n, 32)
            prof = dev.readProfile(64)
            dev.toggleProfiling(False)
            del probe
            scatter = [e for e in prof if e["kernel"].startswith("onesweep")]
            hist = [e for e in prof if e["kernel"].startswith("digit_histogram")]
            avg_ms = sum(e["ms"] for e in scatter) / len(scatter)
            achieved = scatter[0]["bytes"] / (avg_ms * 1e-3) / 1e9
            traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
            try:
                with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                    traffic = json.load(f).get(f"onesweep_pairs_dram_bytes_per_launch_2^{args.log2_pairs_per_gpu}")
            except Exception:
                pass
            roofline = {"bound": "hbm", "kernel": "onesweep2_kernel<uint2, 320 threads x 20 pairs> (one scatter pass, 4 per sort)", "achieved": achieved, "peak": peak,
                        "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                        "avg_launch_ms": avg_ms, "algorithmic_bytes_per_launch": scatter[0]["bytes"],
                        "whole_sort": {"algorithmic_bytes": 72 * n, "achieved": 72 * n / (ms_per_step * 1e-3) / 1e9,
                                       "frac": 72 * n / (ms_per_step * 1e-3) / 1e9 / peak},
                        "histogram_kernel_ms": sum(e["ms"] for e in hist) / len(hist),
                        "scatter_share_of_step": sum(e["ms"] for e in scatter) / 3 / ms_per_step}
        del bufs, handles, inputs64

        # ---- end to end through the C-ABI host-buffer entry points (pinned host memory) ----
        # `value`: b200rs_sort_pairs_u32_host -- the call a drop-in caller makes for one array: pinned host -> device copy,
        # sort, device -> host copy, all inside the timed region (nothing to overlap with).  batch_*: the pipelined batch
        # entry (E2E_BATCH arrays per call, copies of neighbouring arrays overlap the sort).  buffer_api_*: the reference
        # caller's own sequence through adl::Buffer (map / fill / unmap / radixSort / map, UnitTest/main.cpp:118-139).
        e2e = None
        if sorter is None:
            E2E_BATCH = 6
            src = fresh_pairs().cpu()
            hosts = [torch.empty((n, 2), dtype=torch.int32).pin_memory() for _ in range(E2E_BATCH)]
            ptrs = (ctypes.c_void_p * E2E_BATCH)(*[h.data_ptr() for h in hosts])

            def refill():
                for i, h in enumerate(hosts):
                    h.copy_(src)
                    h[:, 0] ^= (0x9E3779B1 * (i + 1)) & 0x7FFFFFFF  # a different key set per array (XOR keeps it a uniform permutation)

            times = []
            for r in range(3):
                refill()
                t0 = time.perf_counter()
                check(lib().b200rs_sort_pairs_u32_host_batch(dev.handle, ptrs, E2E_BATCH, n, 32), "b200rs_sort_pairs_u32_host_batch")
                times.append(time.perf_counter() - t0)
            for h in hosts:
                hk = h[:, 0].to(torch.int64) & 0xFFFFFFFF
                assert bool((hk[1:] >= hk[:-1]).all()), "e2e result not sorted"
            t_batch = min(times[1:]) / E2E_BATCH
            single = []
            for r in range(4):
                hosts[0].copy_(src)
                t0 = time.perf_counter()
                check(lib().b200rs_sort_pairs_u32_host(dev.handle, ctypes.c_void_p(hosts[0].data_ptr()), n, 32), "b200rs_sort_pairs_u32_host")
                single.append(time.perf_counter() - t0)
            t_single = min(single[1:])
            # the reference caller's sequence through the Buffer API: getHostPtr, fill, returnHostPtr, radixSort, getHostPtr
            bapi = []
            hb = ob.Buffer(dev, n, ob.PAIR_DTYPE)
            src_np = src.numpy().view(np.uint32).reshape(-1)
            for r in range(3):
                t0 = time.perf_counter()
                m_ = hb.getHostPtr(n)
                m_.view(np.uint32).reshape(-1)[:] = src_np
                hb.returnHostPtr(m_)
                pp.radixSort(dev, hb, n, 32)
                m_ = hb.getHostPtr(n)
                dev.waitForCompletion()
                first_key = int(m_["key"][0])
                hb.returnHostPtr(m_)
                bapi.append(time.perf_counter() - t0)
            hb.release()
            e2e = {"value": n / t_single / 1e9, "unit": UNIT, "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * n, "ms_per_step": 1e3 * t_single,
                   "api": "b200rs_sort_pairs_u32_host (one pinned host array: copy in, sort, copy back; what a drop-in caller sees)",
                   "batch_value": n / t_batch / 1e9, "batch_ms_per_array": 1e3 * t_batch,
                   "batch_api": f"b200rs_sort_pairs_u32_host_batch ({E2E_BATCH} pinned host arrays per call, each copied in, sorted, copied back; pipelined)",
                   "buffer_api_value": n / min(bapi[1:]) / 1e9, "buffer_api_ms": 1e3 * min(bapi[1:]),
                   "buffer_api": "adl.Buffer getHostPtr / fill (host memcpy of 2 GiB included) / returnHostPtr / Pprims.radixSort / getHostPtr (UnitTest/main.cpp:118-139)"}
            check(lib().b200rs_device_release_scratch(dev.handle), "release_scratch")
            del hosts, src
        else:
            e2e = sorter.e2e(fresh_pairs, n, world)

        # ---- keys (north_star target #1, BASELINE configs[3]) and the other single-GPU configs, device-resident ----
        keys_line = None
        if sorter is None and not args.skip_extras:
            def timeit(fn, make, reps=6):
                ts = []
                for _ in range(reps):
                    x = make()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(stream); fn(x); b.record(stream); stream.synchronize()
                    ts.append(a.elapsed_time(b))
                return ts[1:], x
            u = lambda: torch.randint(-2**31, 2**31, (n,), device="cuda", dtype=torch.int32, generator=gen)
            to_i32 = lambda x: (((x & 0xFFFFFFFF) ^ 0x80000000) - 0x80000000).to(torch.int32)
            presorted = lambda: to_i32(torch.arange(n, device="cuda", dtype=torch.int64) * (2**32 // n))
            sort_keys = lambda bits: (lambda x: pp.radixSort(dev, ob.Buffer(dev, n, np.uint32, ptr=x.data_ptr()), n, bits))

            def keys_case(make, bits=32):
                ts, out = timeit(sort_keys(bits), make)
                k = out.to(torch.int64) & ((1 << bits) - 1)
                ok = bool((k[1:] >= k[:-1]).all())
                del k
                t = statistics.median(ts)
                bpk = 4 + 8 * ((bits + 7) // 8)
                return {"ms": t, "ms_min": min(ts), "gkeys_s": n / t / 1e6, "roofline_frac": bpk * n / t / 1e6 / peak, "sorted": ok}
            uni = keys_case(u)
            keys_line = dict(uni, workload=f"keys_u32_uniform_2^{args.log2_pairs_per_gpu}", algorithmic_bytes_per_key=36,
                             path="MSD pipeline (joint top-16 histogram, 2 unstable 8-bit partition passes, counting sort per bucket) when every top-16 bucket is small, else 4 LSD passes",
                             target="north_star: >= 0.70 of the HBM roofline at 2^28 keys")
            try:  # DRAM bytes of the MSD chain's kernels from the committed ncu --set full capture (the path moves 28 B/key, the roofline figure uses 36)
                with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                    msd = json.load(f).get(f"msd_keys_2^{args.log2_pairs_per_gpu}")
                if msd:
                    keys_line["traffic"] = sum(v["dram_bytes"] for v in msd.values())
                    keys_line["path_algorithmic_bytes"] = sum(v["algorithmic"] for v in msd.values())
            except Exception:
                pass
            ts_scan, _ = timeit(lambda x: pp.scan(dev, ob.Buffer(dev, n, np.uint32, ptr=x.data_ptr()), ob.Buffer(dev, n, np.uint32, ptr=x.data_ptr()), n), u)
            t_scan = min(ts_scan)
            extras = {
                f"keys_u32_2^{args.log2_pairs_per_gpu}_uniform": uni,
                f"keys_u32_2^{args.log2_pairs_per_gpu}_presorted": keys_case(presorted),
                f"keys_u32_2^{args.log2_pairs_per_gpu}_reversed": keys_case(lambda: presorted().flip(0).contiguous()),
                f"keys_u32_2^{args.log2_pairs_per_gpu}_and3": keys_case(lambda: u() & u() & u()),
                f"keys_u32_2^{args.log2_pairs_per_gpu}_sortbits16": keys_case(u, 16),
                f"scan_u32_2^{args.log2_pairs_per_gpu}": {"ms": t_scan, "gelem_s": n / t_scan / 1e6, "roofline_frac": 8 * n / t_scan / 1e6 / peak},
            }

        # ---- N > 1: a skewed input (keys = AND of three uniform words: a third of them share the top byte 0), same size ----
        # From 4 ranks on the digit-range plan overflows the receive buffers (default slack 1.25), so the sorter re-plans with exact
        # quantile splitters (four rounds of filtered histograms, host round trips included in the time); with 2 ranks the digit
        # ranges still fit.  Outside the headline; parity checked the same way.
        skewed = None
        if sorter is not None and not args.skip_extras:
            try:
                sk = fresh_pairs()
                for _ in range(2):
                    sk[:, 0] &= torch.randint(-2**31, 2**31, (n,), device="cuda", dtype=torch.int32, generator=gen)
                h_sk = input_hash(sk)
                sk64 = sk.view(torch.int64).reshape(-1)
                ts_sk = []
                for _ in range(3):
                    dist.barrier()
                    torch.cuda.synchronize()
                    e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0_.record(stream); out_sk, m_sk = sorter.sort(sk64, n); e1_.record(stream)
                    stream.synchronize()
                    ts_sk.append(e0_.elapsed_time(e1_))
                t_sk = torch.tensor([min(ts_sk[1:])], device="cuda", dtype=torch.float64)
                dist.all_reduce(t_sk, op=dist.ReduceOp.MAX)
                sizes_sk = torch.tensor([m_sk], device="cuda", dtype=torch.int64)
                all_sk = [torch.empty_like(sizes_sk) for _ in range(world)]
                dist.all_gather(all_sk, sizes_sk)
                ok_sk = check_sorted_output(out_sk[:m_sk], h_sk, world * n, world * n <= 2**32)
                skewed = {"workload": f"kv_sort_u32u32_and3_2^{args.log2_pairs_per_gpu}_pairs_per_gpu", "ms_per_step": float(t_sk.item()),
                          "value": world * n / float(t_sk.item()) / 1e6, "unit": UNIT, "pairs_per_rank_out": [int(x.item()) for x in all_sk], "parity_ok": bool(ok_sk),
                          "path": getattr(sorter, "last_path", "?") + " (digit-range plan when it fits the receive buffers, else exact quantile splitters: SplitterPlan + exchange by splitters)"}
                del sk, sk64, out_sk
            except Exception as exc:  # noqa: BLE001 -- an extra must never cost the headline line
                skewed = {"error": repr(exc)[:300]}

        # ---- N > 1: the same sort at BASELINE config 5's shard size (2^31 pairs per GPU: 2^32 / 2^33 / 2^34 pairs in all) ----
        config5 = None
        if sorter is not None and not args.no_config5:
            sorter.release()
            sorter = None
            torch.cuda.empty_cache()
            n5 = 1 << 31
            free_b, _ = torch.cuda.mem_get_info()
            if free_b > 5 * 8 * n5 + (8 << 30):
                from oclradixsort_b200.dist import DistributedPairSorter
                s5 = DistributedPairSorter(dev, pp, n5, dist)
                in5 = [fresh_pairs(n5).view(torch.int64).reshape(-1) for _ in range(2)]
                h5 = input_hash(in5[1])
                for i in range(2):
                    s5.sort_async(in5[i], n5)
                stream.synchronize()
                dist.barrier()
                torch.cuda.synchronize()
                align = torch.zeros(1, device="cuda")
                dist.all_reduce(align)
                a5, b5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps5 = 4
                a5.record(stream)
                for i in range(reps5):
                    out5 = s5.sort_async(in5[i & 1], n5)
                b5.record(stream)
                stream.synchronize()
                dist.barrier()
                t5 = torch.tensor([a5.elapsed_time(b5) / reps5], device="cuda", dtype=torch.float64)
                dist.all_reduce(t5, op=dist.ReduceOp.MAX)
                m5 = s5.finish()
                ok5 = check_sorted_output(out5[:m5], h5, world * n5, False)
                # one GPU on one shard of the same size, for the ratio north_star asks for
                one = in5[0]
                t1s = []
                for i in range(3):
                    one.copy_(in5[1])
                    e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0_.record(stream); pp.radixSort(dev, ob.Buffer(dev, n5, ob.PAIR_DTYPE, ptr=one.data_ptr()), n5, 32); e1_.record(stream)
                    stream.synchronize()
                    t1s.append(e0_.elapsed_time(e1_))
                t1 = torch.tensor([min(t1s[1:])], device="cuda", dtype=torch.float64)
                t1max = t1.clone()
                dist.all_reduce(t1, op=dist.ReduceOp.MIN)   # the ratio is taken against the FASTEST rank's one-GPU sort (all ranks run it at the same time)
                dist.all_reduce(t1max, op=dist.ReduceOp.MAX)
                ms5, ms1 = float(t5.item()), float(t1.item())
                config5 = {"pairs_per_gpu": n5, "total_pairs": world * n5, "ms_per_step": ms5, "value": world * n5 / ms5 / 1e6, "unit": UNIT, "steps": reps5,
                           "one_gpu_shard_ms": ms1, "one_gpu_shard_ms_slowest_rank": float(t1max.item()), "one_gpu_shard_value": n5 / ms1 / 1e6, "ratio_to_one_gpu": (world * n5 / ms5) / (n5 / ms1),
                           "parity_ok": bool(ok5), "note": "values wrap mod 2^32 at this size: parity = sortedness + multiset hashes"}
                if not ok5:
                    raise SystemExit("bench.py: config-5 output is not a sort of its input")
                del in5, out5, one
                s5.release()
            else:
                config5 = {"skipped": f"not enough free device memory for 2^31 pairs per GPU ({free_b >> 30} GiB free)"}

        pp.release()
        if sorter is not None:
            sorter.release()

    cpu = cpu_baseline(args.cpu_log2_sample) if (rank == 0 and world == 1) else None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"kv_sort_u32u32_uniform_2^{args.log2_pairs_per_gpu}_pairs_per_gpu", "sort_bits": 32,
                       "pairs_per_gpu": n, "l2": "every step sorts a different 2 GiB buffer (inputs larger than L2, no flush needed)",
                       "parallelism": "single GPU" if world == 1 else f"msd-partitioned over {world} GPUs (top-digit histogram, all-gather of the histograms, on-device plan with two halves per destination; one exchange kernel: half A by peer stores over NVLink, half B staged and moved by copy engines while half A is sorted on a second stream; local LSD sorts)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "numa_rank0": numa, "skewed": skewed,
            "step_ms_rank0": [round(x, 3) for x in step_ms],
        }
        if keys_line:
            line["keys"] = keys_line
        if config5:
            line["config5"] = config5
        if extras:
            line["extra"] = extras
        print(json.dumps(line), file=json_out, flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
