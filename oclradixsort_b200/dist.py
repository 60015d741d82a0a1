"""Partitioned key-value sort across the GPUs of one box (new capability -- the reference is
single-device, one cl_context + one in-order queue, Adl/CL/AdlCL.inl:284-303).

One process per GPU, torch.distributed for the plumbing.  Rank r holds a slice of the input; slices in rank order
are the global input order.  For inputs beyond one GPU's HBM:

  1. local histogram of the TOP key digit (8 bits)                        b200rs_digit_histogram_pairs
  2. all-gather of the per-rank histograms (256 x P counts)               dist.all_gather_into_tensor  (NCCL)
     -> every rank derives the same plan ON THE DEVICE: contiguous digit ranges -> ranks with about N/P pairs each
  3. ONE kernel: local STABLE partition by destination rank, every run stored straight into the destination GPU's
     receive buffer over NVLink (peer memory mapped through CUDA IPC, bulk copies)   b200rs_exchange_pairs
  4. barrier (an all-reduce of one word): every rank's stores have landed
  5. local stable LSD sort of what was received                           b200rs_sort_pairs_u32_devn
Steps 1-5 are ONE C call, b200rs_dist_sort_pairs_u32, which takes the two collectives as callbacks (the library links no
communication library); this module supplies torch.distributed for them.  exchange="nccl" keeps the baseline: partition
into a local send buffer + dist.all_to_all_single.

From 2^22 pairs of receive capacity the C call runs steps 2-5 PIPELINED (plan_exchange_halves() below is the host mirror
of its plan): every destination's digit range is cut into halves A | B; the kernel of step 3 stores half A over NVLink
and stages half B locally; copy engines move half B to the peers while half A is already being sorted on a second
stream; half B is sorted when it has landed (two barriers instead of one; the call reads the plan back once, so it
returns after the plan kernel has run -- everything after that is stream-ordered).

Skew.  Digit ranges cannot split a hot top digit; when the plan of step 2 would overflow a receive buffer the sort is
re-planned with EXACT QUANTILE SPLITTERS (SplitterPlan): three more rounds of 256-bin histograms restricted to the bins
the boundaries fall in narrow every boundary down to one 32-bit key, and a key that is itself too frequent is split by
source rank (all copies of a key on lower ranks go first: stability is kept), so any distribution -- all keys equal
included -- is balanced to within one source's count of one key.

Bit-exactness: 3 and 5 are stable and 4 keeps (source rank, position) order = global input order inside every
destination, and destinations own disjoint, increasing key ranges, so the concatenation of the ranks' outputs is
exactly the single stable sort of the concatenated input.  Outputs differ in length under skew; a destination
whose share exceeds its receive capacity raises (no silent truncation, no CPU fallback).

The device work is injected through a small `ops` object so the protocol (plan, split sizes, ordering) can be
exercised on CPU with gloo in tests/ -- the product uses CudaLocalOps only.
"""
from __future__ import annotations

import ctypes

import numpy as np

from ._lib import B200RSError, check, lib

TOP_SHIFT, TOP_BITS, NUM_BINS = 24, 8, 256
ERR_CAPACITY = -7  # B200RS_ERR_CAPACITY


def plan_exchange(hist: np.ndarray) -> dict:
    """hist[s, b] = number of pairs on source rank s whose top digit is b  (shape P x 256, any integer dtype).

    Returns the plan every rank computes identically:
      bin_to_rank[256]   owner of each digit: contiguous, non-decreasing ranges with totals close to N/P
      send_counts[s, d]  pairs rank s sends to rank d
      recv_total[d]      pairs rank d ends up with
    """
    hist = np.asarray(hist, dtype=np.int64)
    P = hist.shape[0]
    assert hist.shape == (P, NUM_BINS)
    totals = hist.sum(axis=0)
    cum = np.concatenate([[0], np.cumsum(totals)])  # cum[b] = pairs with digit < b
    N = int(cum[-1])
    # boundary r (between rank r-1 and r) = the digit edge closest to r*N/P (exact integer compare, ties to the
    # lower digit -- the same rule as dist_plan_kernel); edges are kept non-decreasing
    edges = [0]
    for r in range(1, P):
        b = int(np.argmin(np.abs(cum * P - r * N)))
        edges.append(max(b, edges[-1]))
    edges.append(NUM_BINS)
    bin_to_rank = np.zeros(NUM_BINS, dtype=np.uint8)
    for r in range(P):
        bin_to_rank[edges[r]:edges[r + 1]] = r
    send_counts = np.zeros((P, P), dtype=np.int64)
    for d in range(P):
        send_counts[:, d] = hist[:, edges[d]:edges[d + 1]].sum(axis=1)
    # where source s writes inside destination d's receive buffer (pairs):
    #   dest layout  [source 0 | source 1 | ...]                         recv_offset[s, d]
    #   bins layout  [digit b0: source 0, source 1, ... | digit b0+1: ...]  bin_offset[s, b]
    recv_offset = np.cumsum(send_counts, axis=0) - send_counts
    bin_offset = np.zeros((P, NUM_BINS), dtype=np.int64)
    for d in range(P):
        lo, hi = edges[d], edges[d + 1]
        digit_start = np.cumsum(totals[lo:hi]) - totals[lo:hi]          # start of each owned digit's block
        within = np.cumsum(hist[:, lo:hi], axis=0) - hist[:, lo:hi]       # sources before s, same digit
        bin_offset[:, lo:hi] = digit_start[None, :] + within
    return {"bin_to_rank": bin_to_rank, "edges": edges, "send_counts": send_counts, "recv_total": send_counts.sum(axis=0), "total": N,
            "recv_offset": recv_offset, "bin_offset": bin_offset}


def plan_exchange_halves(hist: np.ndarray, a_permille: int = 500) -> dict:
    """The plan of the PIPELINED partitioned sort (dist_plan_halves_kernel in csrc/b200rs_sort.cu; this is its host mirror
    and the CPU tests' reference): the digit ranges of plan_exchange(), each cut once more at the digit boundary where half A
    holds closest to a_permille / 1000 of the destination's pairs (ties to the lower digit).

      bin_to_part[256]      2 * destination + half (0 = A: the lower digits of the destination's range)
      mids[d]               first digit of destination d's half B
      part_counts[s, v]     pairs source s sends to part v
      recv_a[d], recv_b[d]  pairs destination d receives per half
      part_offset[s, v]     where source s writes inside destination v // 2's receive buffer (pairs): the buffer reads
                            [A: source 0, source 1, ... | B: source 0, source 1, ...]
    """
    base = plan_exchange(hist)
    hist = np.asarray(hist, dtype=np.int64)
    P = hist.shape[0]
    edges = base["edges"]
    totals = hist.sum(axis=0)
    cum = np.concatenate([[0], np.cumsum(totals)])
    mids = []
    for d in range(P):
        lo, hi = edges[d], edges[d + 1]
        tot = int(cum[hi] - cum[lo])
        ks = np.arange(lo, hi + 1)
        dist_k = np.abs(1000 * (cum[ks] - cum[lo]) - a_permille * tot)
        mids.append(int(ks[int(np.argmin(dist_k))]))  # argmin returns the first (lowest digit) of equal distances
    bin_to_part = np.zeros(NUM_BINS, dtype=np.uint8)
    part_counts = np.zeros((P, 2 * P), dtype=np.int64)
    for d in range(P):
        bin_to_part[edges[d]:mids[d]] = 2 * d
        bin_to_part[mids[d]:edges[d + 1]] = 2 * d + 1
        part_counts[:, 2 * d] = hist[:, edges[d]:mids[d]].sum(axis=1)
        part_counts[:, 2 * d + 1] = hist[:, mids[d]:edges[d + 1]].sum(axis=1)
    recv_a, recv_b = part_counts[:, 0::2].sum(axis=0), part_counts[:, 1::2].sum(axis=0)
    before = np.cumsum(part_counts, axis=0) - part_counts  # pairs of lower ranks in the same part
    part_offset = before.copy()
    part_offset[:, 1::2] += recv_a[None, :]
    return {"bin_to_part": bin_to_part, "edges": edges, "mids": mids, "part_counts": part_counts, "recv_a": recv_a, "recv_b": recv_b,
            "recv_total": recv_a + recv_b, "part_offset": part_offset, "total": base["total"]}


class SplitterPlan:
    """Exact quantile splitters from per-source digit histograms, refined one digit per round (pure numpy; every rank
    computes the same plan from the same all-gathered histograms).

      plan = SplitterPlan(H0)                H0[s, b]: pairs on source s with top digit b
      for level in 1, 2, 3:
          plan.refine(H)                     H[s, j, d]: pairs on source s whose key bits above digit (3 - level) equal
                                             plan.prefixes()[j] and whose digit (3 - level) is d
      res = plan.finish()

    Boundary j (between destination j and j + 1) aims at t_j = j * N // P pairs below it.  After the last round it is an
    exact key K_j with, per source, the number of pairs below K_j and equal to K_j; the pairs equal to K_j go below the
    boundary for the first tie_j sources only -- source order is input order, so this keeps the sort stable."""

    def __init__(self, hist0):
        h = np.asarray(hist0, dtype=np.int64)
        self.P = h.shape[0]
        assert h.shape == (self.P, NUM_BINS)
        self.n_src = h.sum(axis=1)
        self.N = int(self.n_src.sum())
        self.J = self.P - 1
        self.targets = np.array([(j * self.N) // self.P for j in range(1, self.P)], dtype=np.int64)
        self.below = np.zeros((self.J, self.P), dtype=np.int64)   # pairs on source s below boundary j's current prefix
        self.prefix = np.zeros(self.J, dtype=np.int64)
        self.eq = np.zeros((self.J, self.P), dtype=np.int64)
        self.level = 0
        self._step(np.broadcast_to(h[:, None, :], (self.P, self.J, NUM_BINS)))

    def _step(self, H):
        """H[s, j, d] for the digit below the current prefixes."""
        tot = H.sum(axis=0)  # [J, 256]
        for j in range(self.J):
            r = int(self.targets[j] - self.below[j].sum())
            cum = np.concatenate([[0], np.cumsum(tot[j])])
            d = int(np.searchsorted(cum, r, side="right")) - 1
            d = min(max(d, 0), NUM_BINS - 1)
            self.below[j] += H[:, j, :d].sum(axis=1)
            self.prefix[j] = self.prefix[j] * NUM_BINS + d
            self.eq[j] = H[:, j, d]

    def prefixes(self):
        return self.prefix.astype(np.uint32)

    def refine(self, H):
        H = np.asarray(H, dtype=np.int64).reshape(self.P, self.J, NUM_BINS)
        self.level += 1
        self._step(H)

    def finish(self):
        assert self.level == 3 or self.J == 0
        P, J = self.P, self.J
        L = np.zeros((P + 1, P), dtype=np.int64)  # L[j, s]: pairs of source s that go below boundary j
        thr = np.zeros((P, max(J, 1)), dtype=np.uint64)
        ties = []
        for j in range(J):
            r = int(self.targets[j] - self.below[j].sum())
            ceq = np.concatenate([[0], np.cumsum(self.eq[j])])
            tie = int(np.argmin(np.abs(ceq - r)))  # whole sources whose copies of K_j go below: the count closest to the target
            ties.append(tie)
            L[j + 1] = self.below[j] + np.where(np.arange(P) < tie, self.eq[j], 0)
            thr[:, j] = np.uint64(self.prefix[j]) + (np.arange(P) < tie).astype(np.uint64)
        L[P] = self.n_src
        for j in range(1, P + 1):  # boundaries that coincide (a key hot enough to span several ranks) stay ordered
            L[j] = np.maximum(L[j], L[j - 1])
        send_counts = (L[1:] - L[:-1]).T.copy()  # [s, d]
        recv_offset = np.cumsum(send_counts, axis=0) - send_counts
        return {"thresholds": thr[:, :J], "keys": self.prefix.copy(), "ties": ties, "send_counts": send_counts, "recv_offset": recv_offset,
                "recv_total": send_counts.sum(axis=0), "total": self.N}


class CudaLocalOps:
    """The device side of the protocol on one B200, through the C ABI.  Buffers are torch int64 tensors (one
    element = one {key, value} pair, key in the low half -- the AoS layout of Tahoe::uint2)."""

    def __init__(self, device, pprims):
        import torch
        self.torch = torch
        self.device, self.pprims = device, pprims
        self.cuda = torch.device("cuda", device.device_idx)
        # the library's kernels run on the handle's stream, torch's collectives and copies on torch's current stream: they
        # must be the same stream, or nothing orders a histogram before its all-gather, a peer scatter before the barrier ...
        if device.stream() != torch.cuda.current_stream(self.cuda).cuda_stream:
            raise ValueError("DistributedPairSorter: create the Device on torch's current stream "
                             "(DeviceUtils.allocate(..., cuda_stream=torch.cuda.current_stream().cuda_stream)) and keep that stream current")
        self._hist = torch.zeros(NUM_BINS, dtype=torch.int64, device=self.cuda)
        self._lut = torch.zeros(NUM_BINS, dtype=torch.uint8, device=self.cuda)
        self._counts = torch.zeros(NUM_BINS, dtype=torch.int64, device=self.cuda)
        self._temp = None

    def empty(self, n):
        return self.torch.empty(max(int(n), 1), dtype=self.torch.int64, device=self.cuda)

    def histogram(self, pairs, n):
        check(lib().b200rs_digit_histogram_pairs(self.device.handle, ctypes.c_void_p(pairs.data_ptr()), n, TOP_SHIFT, TOP_BITS,
                                                 ctypes.c_void_p(self._hist.data_ptr())), "b200rs_digit_histogram_pairs")
        return self._hist

    def partition(self, src, dst, n, bin_to_part: np.ndarray, part_counts: np.ndarray):
        t = self.torch
        self._lut.copy_(t.from_numpy(np.ascontiguousarray(bin_to_part, dtype=np.uint8)), non_blocking=False)
        counts = np.zeros(NUM_BINS, dtype=np.int64)
        counts[: len(part_counts)] = part_counts
        self._counts.copy_(t.from_numpy(counts), non_blocking=False)
        need = ctypes.c_size_t(0)
        fn = lib().b200rs_partition_pairs
        check(fn(self.device.handle, None, None, n, TOP_SHIFT, TOP_BITS, None, None, None, ctypes.byref(need)), "b200rs_partition_pairs (size)")
        if self._temp is None or self._temp.numel() < need.value:
            self._temp = t.empty(need.value, dtype=t.uint8, device=self.cuda)
        have = ctypes.c_size_t(self._temp.numel())
        check(fn(self.device.handle, ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(dst.data_ptr()), n, TOP_SHIFT, TOP_BITS,
                 ctypes.c_void_p(self._lut.data_ptr()), ctypes.c_void_p(self._counts.data_ptr()), ctypes.c_void_p(self._temp.data_ptr()),
                 ctypes.byref(have)), "b200rs_partition_pairs")

    def local_sort(self, pairs, m):
        from .adl import PAIR_DTYPE, Buffer
        self.pprims.radixSort(self.device, Buffer(self.device, m, PAIR_DTYPE, ptr=pairs.data_ptr()), m, 32)

    # ---- peer-memory exchange (one process per GPU: CUDA IPC) ----
    supports_p2p = True

    def alloc_exported(self, n):
        """(int64 tensor view, device address, 64-byte IPC handle) of a fresh cudaMalloc'ed buffer of n pairs."""
        p = ctypes.c_void_p()
        check(lib().b200rs_malloc(self.device.handle, int(n) * 8, ctypes.byref(p)), "b200rs_malloc")
        handle = ctypes.create_string_buffer(64)
        check(lib().b200rs_ipc_export(self.device.handle, p, handle), "b200rs_ipc_export")
        return _tensor_from_ptr(self.torch, p.value, int(n), self.cuda), p.value, handle.raw

    def import_peer(self, handle: bytes) -> int:
        p = ctypes.c_void_p()
        check(lib().b200rs_ipc_import(self.device.handle, handle, ctypes.byref(p)), "b200rs_ipc_import")
        return p.value

    def release_peer(self, addr: int):
        check(lib().b200rs_ipc_release(self.device.handle, ctypes.c_void_p(addr)), "b200rs_ipc_release")

    def free(self, addr: int):
        check(lib().b200rs_free(self.device.handle, ctypes.c_void_p(addr)), "b200rs_free")

    # ---- stream-ordered pipeline (no host round trip): plan, scatter and local sort all read device-side tables ----
    def plan_async(self, gathered, world, rank, peers_dev, capacity, n):
        t = self.torch
        if not hasattr(self, "_cnts"):
            self._cnts = t.zeros(2, dtype=t.int64, device=self.cuda)    # [0] pairs to scatter, [1] pairs received
            self._status = t.zeros(1, dtype=t.int32, device=self.cuda)  # 1 = capacity exceeded, nothing exchanged
        check(lib().b200rs_dist_plan(self.device.handle, ctypes.c_void_p(gathered.data_ptr()), world, rank, ctypes.c_void_p(peers_dev.data_ptr()),
                                     capacity, n, ctypes.c_void_p(self._lut.data_ptr()), ctypes.c_void_p(self._counts.data_ptr()),
                                     ctypes.c_void_p(self._cnts.data_ptr()), ctypes.c_void_p(self._status.data_ptr())), "b200rs_dist_plan")

    def scatter_async(self, src, n):
        t = self.torch
        need = ctypes.c_size_t(0)
        fn = lib().b200rs_scatter_pairs_to_parts
        check(fn(self.device.handle, None, n, TOP_SHIFT, TOP_BITS, None, None, None, None, ctypes.byref(need)), "b200rs_scatter_pairs_to_parts (size)")
        if self._temp is None or self._temp.numel() < need.value:
            self._temp = t.empty(need.value, dtype=t.uint8, device=self.cuda)
        have = ctypes.c_size_t(self._temp.numel())
        check(fn(self.device.handle, ctypes.c_void_p(src.data_ptr()), n, TOP_SHIFT, TOP_BITS, ctypes.c_void_p(self._lut.data_ptr()),
                 ctypes.c_void_p(self._counts.data_ptr()), ctypes.c_void_p(self._cnts.data_ptr()), ctypes.c_void_p(self._temp.data_ptr()),
                 ctypes.byref(have)), "b200rs_scatter_pairs_to_parts")

    def local_sort_devn(self, pairs, capacity):
        need = ctypes.c_size_t(0)
        fn = lib().b200rs_sort_pairs_u32_devn
        check(fn(self.device.handle, None, capacity, None, 32, None, ctypes.byref(need)), "b200rs_sort_pairs_u32_devn (size)")
        temp = self.pprims._scratch(self.device, need.value)
        have = ctypes.c_size_t(temp.getSize())
        check(fn(self.device.handle, ctypes.c_void_p(pairs.data_ptr()), capacity, ctypes.c_void_p(self._cnts.data_ptr() + 8), 32,
                 ctypes.c_void_p(temp.m_ptr), ctypes.byref(have)), "b200rs_sort_pairs_u32_devn")

    def read_counts(self):
        """(pairs received, status) -- synchronises."""
        c = self._cnts.cpu()
        return int(c[1].item()), int(self._status.cpu().item())

    def scatter(self, src, n, digit_to_part: np.ndarray, part_base_addr: np.ndarray):
        """Stable partition of src straight into the parts' base addresses (local or peer memory)."""
        t = self.torch
        self._lut.copy_(t.from_numpy(np.ascontiguousarray(digit_to_part, dtype=np.uint8)))
        self._counts.copy_(t.from_numpy(np.ascontiguousarray(part_base_addr, dtype=np.uint64).view(np.int64)))
        need = ctypes.c_size_t(0)
        fn = lib().b200rs_scatter_pairs_to_parts
        check(fn(self.device.handle, None, n, TOP_SHIFT, TOP_BITS, None, None, None, None, ctypes.byref(need)), "b200rs_scatter_pairs_to_parts (size)")
        if self._temp is None or self._temp.numel() < need.value:
            self._temp = t.empty(need.value, dtype=t.uint8, device=self.cuda)
        have = ctypes.c_size_t(self._temp.numel())
        check(fn(self.device.handle, ctypes.c_void_p(src.data_ptr()), n, TOP_SHIFT, TOP_BITS, ctypes.c_void_p(self._lut.data_ptr()),
                 ctypes.c_void_p(self._counts.data_ptr()), None, ctypes.c_void_p(self._temp.data_ptr()), ctypes.byref(have)),
              "b200rs_scatter_pairs_to_parts")

    def to_host_matrix(self, t):
        return t.cpu().numpy()

    # ---- the whole fast path as one C call (b200rs_dist_sort_pairs_u32) with torch.distributed behind its callbacks ----
    def dist_sort_async(self, src, n, dist, world, rank, peers, capacity, flag):
        t = self.torch
        if not hasattr(self, "_cnts"):
            self._cnts = t.zeros(2, dtype=t.int64, device=self.cuda)    # [0] pairs to scatter, [1] pairs received
            self._status = t.zeros(1, dtype=t.int32, device=self.cuda)  # 1 = capacity exceeded, nothing exchanged
        if not hasattr(self, "_comm"):
            cuda = self.cuda

            def allgather(user, send, recv, nbytes):
                try:
                    dist.all_gather_into_tensor(_tensor_from_ptr(t, recv, world * nbytes // 8, cuda), _tensor_from_ptr(t, send, nbytes // 8, cuda))
                    return 0
                except Exception:  # noqa: BLE001 -- must not unwind through the C frame
                    return -1

            def barrier(user):
                try:
                    dist.all_reduce(flag)  # stream-ordered: every rank's stores have landed before anyone's next kernel
                    return 0
                except Exception:  # noqa: BLE001
                    return -1

            AG = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)
            BR = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p)

            class Comm(ctypes.Structure):
                _fields_ = [("rank", ctypes.c_int), ("world", ctypes.c_int), ("allgather", AG), ("barrier", BR), ("user", ctypes.c_void_p)]

            self._callbacks = (AG(allgather), BR(barrier))  # kept alive as long as the ops object
            self._comm = Comm(rank, world, self._callbacks[0], self._callbacks[1], None)
            self._recv_base = (ctypes.c_uint64 * world)(*[int(a) for a in peers])
        fn = lib().b200rs_dist_sort_pairs_u32
        need = ctypes.c_size_t(0)
        check(fn(self.device.handle, ctypes.byref(self._comm), self._recv_base, capacity, None, n, None, None, None, ctypes.byref(need)),
              "b200rs_dist_sort_pairs_u32 (size)")
        temp = self.pprims._scratch(self.device, need.value)
        have = ctypes.c_size_t(temp.getSize())
        check(fn(self.device.handle, ctypes.byref(self._comm), self._recv_base, capacity, ctypes.c_void_p(src.data_ptr()), n,
                 ctypes.c_void_p(self._cnts.data_ptr()), ctypes.c_void_p(self._status.data_ptr()), ctypes.c_void_p(temp.m_ptr), ctypes.byref(have)),
              "b200rs_dist_sort_pairs_u32")

    # ---- splitter plan (skewed inputs) ----
    def filtered_histograms(self, pairs, n, shift, prefixes: np.ndarray):
        t = self.torch
        J = len(prefixes)
        pre = t.from_numpy(np.ascontiguousarray(prefixes, dtype=np.uint32).view(np.int32)).to(self.cuda)
        out = t.zeros(J * NUM_BINS, dtype=t.int64, device=self.cuda)
        check(lib().b200rs_filtered_histograms_pairs(self.device.handle, ctypes.c_void_p(pairs.data_ptr()), n, shift, ctypes.c_void_p(pre.data_ptr()), J,
                                                     ctypes.c_void_p(out.data_ptr())), "b200rs_filtered_histograms_pairs")
        return out

    def exchange_by_splitters(self, src, n, thresholds: np.ndarray, part_base_addr: np.ndarray):
        """Stable partition by key thresholds straight into the parts' base addresses (local or peer memory)."""
        t = self.torch
        parts = len(part_base_addr)
        thr = t.from_numpy(np.ascontiguousarray(np.concatenate([thresholds, [0]]), dtype=np.uint64).view(np.int64)).to(self.cuda)
        base = t.from_numpy(np.ascontiguousarray(part_base_addr, dtype=np.uint64).view(np.int64)).to(self.cuda)
        fn = lib().b200rs_exchange_pairs_by_splitters
        need = ctypes.c_size_t(0)
        check(fn(self.device.handle, None, n, None, None, parts, None, ctypes.byref(need)), "b200rs_exchange_pairs_by_splitters (size)")
        if self._temp is None or self._temp.numel() < need.value + 256:
            self._temp = t.empty(need.value + 256, dtype=t.uint8, device=self.cuda)
        tp = (self._temp.data_ptr() + 255) // 256 * 256
        have = ctypes.c_size_t(need.value)
        check(fn(self.device.handle, ctypes.c_void_p(src.data_ptr()), n, ctypes.c_void_p(thr.data_ptr()), ctypes.c_void_p(base.data_ptr()), parts,
                 ctypes.c_void_p(tp), ctypes.byref(have)), "b200rs_exchange_pairs_by_splitters")
        self._keep = (thr, base)  # alive until the kernel has run (the next call replaces them after a stream-ordered collective)

    def partition_by_splitters(self, src, dst, n, thresholds: np.ndarray, part_counts: np.ndarray):
        starts = np.cumsum(part_counts) - part_counts
        self.exchange_by_splitters(src, n, thresholds, np.uint64(dst.data_ptr()) + 8 * starts.astype(np.uint64))

    def release(self):
        self._temp = None


class DistributedPairSorter:
    """sort(pairs, n) -> (sorted_pairs_of_this_rank, m).  `pairs`: int64 tensor (or adl.Buffer of PAIR_DTYPE)
    holding this rank's n input pairs; it is left unchanged.  The result is a view into an internal receive
    buffer, valid until the next call."""

    def __init__(self, device, pprims, capacity_pairs: int, dist, ops=None, slack: float = 1.25, exchange: str | None = None,
                 layout: str = "dest"):
        """exchange: "p2p"  -- ONE kernel partitions and stores straight into the destination GPUs' receive buffers
                               (peer memory over NVLink, mapped with CUDA IPC); default on GPUs
                     "nccl" -- partition into a local send buffer, then dist.all_to_all_single
           layout (p2p only): "dest" = receive buffer ordered [source 0 | source 1 | ...];
                              "bins" = ordered by top digit, then source (already grouped by the top digit)."""
        self.dist = dist
        self.ops = ops if ops is not None else CudaLocalOps(device, pprims)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        # receive capacity; a plan that exceeds it raises.  Every rank must use the SAME value (the on-device plan and the
        # overflow decision compare every destination's share with it): the minimum over the ranks' requests.
        caps = [None] * self.world
        dist.all_gather_object(caps, int(capacity_pairs * slack) + 1024)
        self.capacity = min(caps)
        self.send_capacity = int(capacity_pairs)
        self.exchange = exchange or ("p2p" if getattr(self.ops, "supports_p2p", False) else "nccl")
        self.layout = layout
        self.last_plan = None
        self.peers = None
        if self.exchange == "p2p":
            self.send = None
            self.recv, self.recv_addr, handle = self.ops.alloc_exported(self.capacity)
            handles = [None] * self.world
            dist.all_gather_object(handles, handle)
            self.peers = [self.recv_addr if r == self.rank else self.ops.import_peer(handles[r]) for r in range(self.world)]
            self._flag = self.ops.empty(1)
            import torch
            self._peers_dev = torch.tensor(self.peers, dtype=torch.int64, device=self.ops.cuda)
            dist.barrier()
        else:
            self.send = self.ops.empty(capacity_pairs)
            self.recv = self.ops.empty(self.capacity)

    def _as_tensor(self, pairs, n):
        if hasattr(pairs, "m_ptr"):  # adl.Buffer wrapping device memory: view it as int64 without copying
            import torch
            from .adl import PAIR_DTYPE
            assert pairs.dtype == PAIR_DTYPE
            return _tensor_from_ptr(torch, pairs.m_ptr, n, self.ops.cuda)
        return pairs

    def sort_async(self, pairs, n: int):
        """p2p/dest only: enqueue the whole partitioned sort -- one call of b200rs_dist_sort_pairs_u32 (histogram ->
        all-gather -> on-device plan -> fused partition + peer stores -> barrier -> local sort; pipelined in two halves from
        2^22 pairs of capacity, see the module docstring; below that no host round trip at all).  Returns the receive buffer;
        call finish() for the element count (it synchronises; a plan that overflowed a receive buffer exchanged nothing and
        reports status 1)."""
        assert self.exchange == "p2p" and self.layout == "dest"
        src = self._as_tensor(pairs, n)
        self.ops.dist_sort_async(src, n, self.dist, self.world, self.rank, self.peers, self.capacity, self._flag)
        return self.recv

    def finish(self) -> int:
        m, status = self.ops.read_counts()
        if status != 0:
            raise B200RSError(ERR_CAPACITY, f"distributed sort: a rank's share exceeds the receive capacity of {self.capacity} pairs "
                                            "(digit-range plan; DistributedPairSorter.sort re-plans with exact splitters)")
        return m

    def _gather_matrix(self, t, cols):
        g = t.new_empty(self.world * t.numel())
        self.dist.all_gather_into_tensor(g, t)
        return self.ops.to_host_matrix(g).reshape(self.world, *cols)

    def sort_with_splitters(self, pairs, n: int):
        """The skew-proof path: exact quantile splitters (SplitterPlan) from four rounds of 256-bin histograms, ties split by
        source rank; then the same fused exchange (thresholds instead of a digit table) and the local sort."""
        ops, dist = self.ops, self.dist
        src = self._as_tensor(pairs, n)
        plan = SplitterPlan(self._gather_matrix(ops.histogram(src, n), (NUM_BINS,)))  # (also orders this step after the previous local sorts)
        if self.world > 1:
            for level in (1, 2, 3):
                h = ops.filtered_histograms(src, n, TOP_SHIFT - 8 * level, plan.prefixes())
                plan.refine(self._gather_matrix(h, (self.world - 1, NUM_BINS)))
        res = plan.finish()
        self.last_plan = res
        m = int(res["recv_total"][self.rank])
        if int(res["recv_total"].max()) > self.capacity:  # same decision on every rank: nobody enters the exchange
            raise B200RSError(ERR_CAPACITY, f"distributed sort: rank {int(res['recv_total'].argmax())} would receive {int(res['recv_total'].max())} pairs "
                                            f"even with exact splitters (one key on one source exceeds the slack), capacity {self.capacity}")
        if self.exchange == "p2p":
            base = np.asarray(self.peers, dtype=np.uint64) + 8 * res["recv_offset"][self.rank].astype(np.uint64)
            ops.exchange_by_splitters(src, n, res["thresholds"][self.rank], base)
            dist.all_reduce(self._flag)  # every rank's stores have landed before anyone sorts
        else:
            assert n <= self.send_capacity
            send_counts = res["send_counts"][self.rank]
            ops.partition_by_splitters(src, self.send, n, res["thresholds"][self.rank], send_counts)
            recv_counts = res["send_counts"][:, self.rank]
            dist.all_to_all_single(self.recv[:m], self.send[:n], [int(c) for c in recv_counts], [int(c) for c in send_counts])
        if m:
            ops.local_sort(self.recv, m)
        return self.recv[:m], m

    def sort(self, pairs, n: int):
        if self.exchange == "p2p" and self.layout == "dest" and hasattr(self.ops, "dist_sort_async"):
            out = self.sort_async(pairs, n)
            m, status = self.ops.read_counts()
            self.last_path = "digit ranges" if status == 0 else "quantile splitters"
            if status == 0:
                return out[:m], m
            return self.sort_with_splitters(pairs, n)  # the digit-range plan overflowed (skew): nothing was exchanged
        ops, dist = self.ops, self.dist
        src = self._as_tensor(pairs, n)
        hist = ops.histogram(src, n)
        gathered = hist.new_empty(self.world * NUM_BINS)
        dist.all_gather_into_tensor(gathered, hist)
        plan = plan_exchange(ops.to_host_matrix(gathered).reshape(self.world, NUM_BINS))
        self.last_plan = plan
        m = int(plan["recv_total"][self.rank])
        if int(plan["recv_total"].max()) > self.capacity:  # same decision on every rank: nobody enters the collective
            if hasattr(ops, "filtered_histograms"):
                return self.sort_with_splitters(pairs, n)
            raise B200RSError(ERR_CAPACITY, f"distributed sort: rank {int(plan['recv_total'].argmax())} would receive "
                                            f"{int(plan['recv_total'].max())} pairs, capacity {self.capacity}")
        if self.exchange == "p2p":
            # The all-gather above doubles as the barrier that makes every rank's previous result dead before
            # anyone overwrites a receive buffer.  One kernel: partition + stores into peer memory.
            peers = np.asarray(self.peers, dtype=np.uint64)
            base = np.zeros(NUM_BINS, dtype=np.uint64)
            if self.layout == "bins":
                lut = np.arange(NUM_BINS, dtype=np.uint8)
                base[:] = peers[plan["bin_to_rank"]] + 8 * plan["bin_offset"][self.rank].astype(np.uint64)
            else:
                lut = plan["bin_to_rank"]
                base[: self.world] = peers + 8 * plan["recv_offset"][self.rank].astype(np.uint64)
            ops.scatter(src, n, lut, base)
            dist.all_reduce(self._flag)  # every rank's stores have landed (kernel completion + collective) before anyone sorts
        else:
            assert n <= self.send_capacity, "the send buffer holds capacity_pairs pairs"
            send_counts = plan["send_counts"][self.rank]
            ops.partition(src, self.send, n, plan["bin_to_rank"], send_counts)
            recv_counts = plan["send_counts"][:, self.rank]
            dist.all_to_all_single(self.recv[:m], self.send[:n], [int(c) for c in recv_counts], [int(c) for c in send_counts])
        if m:
            ops.local_sort(self.recv, m)
        return self.recv[:m], m

    def e2e(self, make_pairs, n: int, world: int):
        """Same sort with HOST buffers: pinned host -> device, distributed sort, device -> pinned host, timed end to end."""
        import time

        import torch
        host_in = make_pairs().view(torch.int64).reshape(-1).cpu().pin_memory()
        host_out = torch.empty(self.capacity, dtype=torch.int64).pin_memory()
        dev_in = self.ops.empty(n)
        times = []
        for _ in range(3):
            self.dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dev_in.copy_(host_in, non_blocking=True)
            out, m = self.sort(dev_in, n)  # finish() inside: the count is needed to size the copy back
            host_out[:m].copy_(out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            self.dist.barrier()
            times.append(time.perf_counter() - t0)
        t = torch.tensor([min(times[1:])], device="cuda", dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        t = float(t.item())
        return {"value": world * n / t / 1e9, "unit": "Gkeys/s", "h2d_bytes_per_step": 8 * n * world, "d2h_bytes_per_step": 8 * n * world,
                "ms_per_step": 1e3 * t, "api": "DistributedPairSorter.sort with pinned host input/output per rank"}

    def release(self):
        if self.peers is not None:
            import torch
            torch.cuda.synchronize()
            self.dist.barrier()  # nobody is still storing into a buffer that is about to be unmapped / freed
            for r, addr in enumerate(self.peers):
                if r != self.rank:
                    self.ops.release_peer(addr)
            self.recv = None
            self.ops.free(self.recv_addr)
            self.peers = None
        self.send = self.recv = None
        self.ops.release()


def bind_to_gpu_numa_node(device_index: int):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE it allocates pinned host buffers: with one
    process per GPU the staging memory of every rank then sits on the socket next to its GPU (first touch) instead of wherever
    the launcher started the process, and host <-> device copies do not cross the socket interconnect.  Best effort: returns
    {"node": N, "cpus": count} or None when the topology cannot be read (no sysfs entry, single node, no permission)."""
    import os
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return {"node": node, "cpus": len(allowed)}
    except Exception:  # noqa: BLE001 -- an optimisation, never a requirement
        return None


def _tensor_from_ptr(torch, ptr: int, n: int, device):
    """Zero-copy int64 view of `n` pairs at device address `ptr` (CUDA array interface)."""

    class _Raw:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3}

    return torch.as_tensor(_Raw(), device=device)
