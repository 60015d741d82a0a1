#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: Gkeys/s of the u32/u32 key-value radix sort.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log2-pairs-per-gpu L]

A "step" is one sort of one batch of synthetic pairs (uniform random u32 keys, value = index), 2^L pairs
per GPU (default 2^28 = the largest size of BASELINE.json configs[1]).  Every step sorts a fresh, unsorted
buffer that is already resident in HBM; buffers are 2 GiB each, far larger than the 126 MB L2, so nothing a
step reads is cached from the previous one.

  value      whole-job Gkeys/s with inputs resident in HBM, device time (CUDA events), max over ranks
  e2e        same metric through the C-ABI HOST-buffer call (b200rs_sort_pairs_u32_host_batch): for every array pinned
             host -> device copy, sort, device -> host copy inside the timed region; the call overlaps the copies of
             neighbouring arrays (single_call_* = one array through b200rs_sort_pairs_u32_host, no overlap)
  roofline   dominant kernel (one scatter pass): algorithmic bytes (2 x 8 B per pair) / its average launch
             time from the library's own CUDA events, against the measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline  the UNMODIFIED reference's Adl Host-backend sort (oracle/_ref, built from /root/reference) on
             this box's host CPU, 1 thread (the reference is serial), on a bounded sample

--impl reference times that same reference CPU path for the same metric/config (rank 0 only under torchrun).
N > 1 (torchrun, one rank per GPU, NCCL): the partitioned sort of oclradixsort_b200.dist -- top-digit
histogram all-reduce, bucket exchange over NVLink, local sort; weak scaling (2^L pairs per GPU).
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
    """The reference's own CPU implementation of the path, timed on this box's host cores."""
    if rank != 0:
        return
    import numpy as np
    from oracle import pyoracle as po
    n = 1 << args.ref_log2_sample
    kind = "reference" if po.have_ref() else "port"
    rng = np.random.default_rng(1234)

    def one_step() -> float:
        kv = np.empty(n, dtype=po.PAIR_DTYPE)
        kv["key"] = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
        kv["value"] = np.arange(n, dtype=np.uint32)
        if kind == "reference":
            return po.ref_time_hostbackend(kv, pairs=True)
        t0 = time.perf_counter()
        po.lib().oracle_sort_pairs(ctypes.c_void_p(kv.ctypes.data), n, 32)
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        one_step()
    times = [one_step() for _ in range(args.steps)]
    total = sum(times)
    value = n * args.steps / total / 1e9
    sample = f"2^{args.ref_log2_sample} uniform pairs per step (bounded sample of the 2^{args.log2_pairs_per_gpu}-pair workload)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic",
        "config": {"workload": f"kv_sort_u32u32_uniform_2^{args.log2_pairs_per_gpu}_pairs_per_gpu", "sort_bits": 32,
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
    ap.add_argument("--ref-log2-sample", type=int, default=24, help="pairs per step of the CPU reference arm")
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
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = 1 << args.log2_pairs_per_gpu
    stream = torch.cuda.Stream()
    peak, peak_src = measured_hbm_peak()
    with torch.cuda.stream(stream):
        dev = ob.DeviceUtils.allocate(ob.TYPE_CL, local_rank, cuda_stream=stream.cuda_stream)
        pp = ob.Pprims()
        gen = torch.Generator(device="cuda").manual_seed(1000 + rank)
        nbuf = args.steps + args.warmup

        def fresh_pairs():
            kv = torch.empty((n, 2), device="cuda", dtype=torch.int32)
            kv[:, 0] = torch.randint(-2**31, 2**31, (n,), device="cuda", dtype=torch.int32, generator=gen)  # uniform over all 32 bits
            kv[:, 1] = torch.arange(n, device="cuda", dtype=torch.int32)
            return kv

        sorter = None
        if world > 1:
            from oclradixsort_b200.dist import DistributedPairSorter
            sorter = DistributedPairSorter(dev, pp, n, dist)

        bufs = [fresh_pairs() for _ in range(nbuf)]
        handles = [ob.Buffer(dev, n, ob.PAIR_DTYPE, ptr=b.data_ptr()) for b in bufs]
        inputs64 = [b.view(torch.int64).reshape(-1) for b in bufs]

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

        # sanity: the last timed step's output is sorted, stable, and a permutation of what was generated
        if sorter is None:
            k64 = bufs[-1][:, 0].to(torch.int64) & 0xFFFFFFFF
            same = k64[1:] == k64[:-1]
            v = bufs[-1][:, 1]
            ok_sorted = bool((k64[1:] >= k64[:-1]).all()) and bool((v[1:][same] > v[:-1][same]).all()) \
                and int(v.to(torch.int64).sum().item()) == n * (n - 1) // 2
            del k64, same
        else:
            m = sorter.finish()  # element count of the last step (raises if a rank's share overflowed)
            out = last["out"][:m]
            k64 = out & 0xFFFFFFFF  # key = low half of the 8-byte pair
            ok_local = bool((k64[1:] >= k64[:-1]).all()) if m > 1 else True
            lo = int(k64[0].item()) if m else 2**32
            hi = int(k64[-1].item()) if m else -1
            stats = torch.tensor([m, lo, hi, int(ok_local)], device="cuda", dtype=torch.int64)
            allstats = [torch.empty_like(stats) for _ in range(world)]
            dist.all_gather(allstats, stats)
            rows = [t.tolist() for t in allstats]
            nonempty = [r for r in rows if r[0] > 0]
            ok_sorted = all(r[3] == 1 for r in rows) and sum(r[0] for r in rows) == world * n \
                and all(a[2] <= b[1] for a, b in zip(nonempty, nonempty[1:]))  # rank r's largest key <= rank r+1's smallest
            del k64
        if not ok_sorted:
            raise SystemExit("bench.py: output of the last timed step is not a stable sort of its input")

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
        # `value`: b200rs_sort_pairs_u32_host_batch -- E2E_BATCH host arrays of 2^L pairs, each copied in, sorted and copied
        # back inside the timed region; the call pipelines them over two device buffers so both directions of the host
        # link are busy.  `single_call_ms`: one array through b200rs_sort_pairs_u32_host (nothing to overlap with).
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
            for r in range(3):
                hosts[0].copy_(src)
                t0 = time.perf_counter()
                check(lib().b200rs_sort_pairs_u32_host(dev.handle, ctypes.c_void_p(hosts[0].data_ptr()), n, 32), "b200rs_sort_pairs_u32_host")
                single.append(time.perf_counter() - t0)
            e2e = {"value": n / t_batch / 1e9, "unit": UNIT, "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * n, "ms_per_step": 1e3 * t_batch,
                   "api": f"b200rs_sort_pairs_u32_host_batch ({E2E_BATCH} pinned host arrays per call, each copied in, sorted, copied back; pipelined)",
                   "single_call_ms": 1e3 * min(single[1:]), "single_call_value": n / min(single[1:]) / 1e9,
                   "single_call_api": "b200rs_sort_pairs_u32_host (one pinned host array: copy in, sort, copy back, no overlap possible)"}
            check(lib().b200rs_device_release_scratch(dev.handle), "release_scratch")
            del hosts, src
        else:
            e2e = sorter.e2e(fresh_pairs, n, world)

        # ---- extras: the other single-GPU configs of BASELINE.json, device-resident ----
        if sorter is None and not args.skip_extras:
            def timeit(fn, make, reps=5):
                ts = []
                for _ in range(reps):
                    x = make()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(stream); fn(x); b.record(stream); stream.synchronize()
                    ts.append(a.elapsed_time(b))
                return min(ts[1:])
            mk = lambda: torch.randint(-2**31, 2**31, (n,), device="cuda", dtype=torch.int32, generator=gen)
            t_keys = timeit(lambda x: pp.radixSort(dev, ob.Buffer(dev, n, np.uint32, ptr=x.data_ptr()), n, 32), mk)
            t_keys16 = timeit(lambda x: pp.radixSort(dev, ob.Buffer(dev, n, np.uint32, ptr=x.data_ptr()), n, 16), mk)
            t_scan = timeit(lambda x: pp.scan(dev, ob.Buffer(dev, n, np.uint32, ptr=x.data_ptr()), ob.Buffer(dev, n, np.uint32, ptr=x.data_ptr()), n), mk)
            extras = {
                f"keys_u32_2^{args.log2_pairs_per_gpu}_uniform": {"ms": t_keys, "gkeys_s": n / t_keys / 1e6, "roofline_frac": 36 * n / t_keys / 1e6 / peak},
                f"keys_u32_2^{args.log2_pairs_per_gpu}_sortbits16": {"ms": t_keys16, "gkeys_s": n / t_keys16 / 1e6, "roofline_frac": 20 * n / t_keys16 / 1e6 / peak},
                f"scan_u32_2^{args.log2_pairs_per_gpu}": {"ms": t_scan, "gelem_s": n / t_scan / 1e6, "roofline_frac": 8 * n / t_scan / 1e6 / peak},
            }
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
                       "parallelism": "single GPU" if world == 1 else f"msd-partitioned over {world} GPUs (histogram all-reduce + NVLink exchange + local LSD)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "step_ms_rank0": [round(x, 3) for x in step_ms],
        }
        if extras:
            line["extra"] = extras
        print(json.dumps(line), file=json_out, flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
