"""Merging the per-GPU partial results of a sharded run.

One process per GPU; `torch.distributed` is the plumbing.  The path has exactly three exchange steps:
  after pass 1 and pass 2 -- bitwise OR of the Bloom bit arrays and SUM of the insert counters,
  after pass 3          -- SUM of the covariate count tables.
NCCL has no bitwise-OR reduction, so the OR is this repo's own.  Where the filters are shareable virtual-memory-API
memory (KBBQ_FILTER_VMM=1) and the devices support NVSwitch multicast, the switch does it (`kbbq_bloom_or_allreduce_mc`:
`multimem.ld_reduce.or` + `multimem.st`, 1/N of the array per GPU on the wire).  Otherwise, on one NVLink/NVSwitch box, ONE kernel per rank over
peer-mapped memory (`kbbq_bloom_or_allreduce_p2p`: rank r ORs slice r of every rank's bit array and writes the result
into every rank's array -- reduce-scatter and all-gather fused, no staging, 2(N-1)/N of the array per GPU on the
wire); where peer mapping is not available (KBBQ_MERGE=allgather, or the mapping fails) the fallback all-gathers
slices with NCCL and ORs them locally (`kbbq_bloom_or_from`).  The sums are plain all-reduces on int64 views.

`Comm` only needs a tensor view of the state to merge, so the same code runs over NCCL on device memory (product)
and over gloo on host memory (the world_size-2 CPU tests, where the oracle port plays the compute engine).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import api


class _DevMem:
    """Zero-copy view of library-owned device memory for torch (CUDA array interface v2)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def device_tensor(ptr: int, nbytes: int, device: int) -> torch.Tensor:
    return torch.as_tensor(_DevMem(ptr, nbytes), device=f"cuda:{device}")


class Comm:
    def __init__(self, rank: int, world: int):
        self.rank, self.world = rank, world
        self.last_merge = None  # which transport merged the most recent filter: "nvls", "p2p" or "allgather"

    @property
    def merge_mode(self) -> str:
        """KBBQ_MERGE = auto (default) | nvls | p2p | allgather.  auto: through the switch from four GPUs on (1 + 1/N of the
        array per GPU and direction against the peer-memory kernel's 2(N-1)/N: fewer bytes from N = 4), the peer-memory
        kernel below.  Whatever is asked for falls back when it is not to be had (nvls -> p2p -> allgather)."""
        import os

        m = os.environ.get("KBBQ_MERGE", "auto")
        if m == "auto":
            m = "nvls" if self.world >= 4 else "p2p"
        return m

    def wants_nvls(self) -> bool:
        return self.world > 1 and self.merge_mode == "nvls" and isinstance(self, GpuComm)

    # ---- scalars -------------------------------------------------------------------------------------------------
    def exclusive_prefix(self, x: int, device=None) -> int:
        """sum of x over the ranks below this one: the ordinal of this rank's first k-mer slot."""
        t = torch.tensor([x], dtype=torch.int64, device=device if device is not None else "cpu")
        out = [torch.zeros_like(t) for _ in range(self.world)]
        dist.all_gather(out, t)
        return int(sum(int(o.item()) for o in out[: self.rank]))

    # ---- generic tensor merges (either backend) ---------------------------------------------------------------------
    # elements per all-gather slice: bounds the staging buffer to world x 256 MiB whatever the filter size
    # (the human-scale filters are 26 and 43 GB)
    SLICE_ELEMS = 32 << 20

    def or_reduce_(self, words: torch.Tensor, or_into=None) -> None:
        """words |= every other rank's words, slice by slice.
        `or_into(first_elem, n_elems, src_tensor)` performs words[first:first+n] |= src (this repo's kernel on GPUs)."""
        if self.world == 1:
            return
        flat = words.view(-1)
        n = flat.numel()
        step = min(n, self.SLICE_ELEMS)
        gathered = torch.empty(self.world * step, dtype=words.dtype, device=words.device)
        for lo in range(0, n, step):
            m = min(step, n - lo)
            g = gathered[: self.world * m]
            dist.all_gather_into_tensor(g, flat[lo : lo + m])
            g = g.view(self.world, m)
            for r in range(self.world):
                if r == self.rank:
                    continue
                if or_into is None:
                    flat[lo : lo + m].bitwise_or_(g[r])
                else:
                    or_into(lo, m, g[r])

    def sum_reduce_(self, t: torch.Tensor) -> None:
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)

    # ---- the three exchange steps on a GPU context --------------------------------------------------------------------
    def _p2p_handles(self, ctx: api.Context, which: int):
        """every rank's IPC handle of the filter's bit array, or None when this run merges by all-gather"""
        import os

        if self.merge_mode not in ("p2p", "nvls") or not isinstance(self, GpuComm):
            return None
        key = (id(ctx), which, ctx.bloom_device_words(which))
        cache = self.__dict__.setdefault("_handles", {})
        if key not in cache:
            try:
                h = ctx.bloom_ipc_handle(which)
                ok = 1
            except api.KbbqError:  # virtual-memory-API filters (KBBQ_FILTER_VMM) have no IPC handle
                h, ok = np.zeros(64, dtype=np.uint8), 0
            if self.sum_int(ok) != self.world:
                cache[key] = None
                return None
            mine = torch.from_numpy(h).to(f"cuda:{ctx.device}")
            allh = torch.empty(self.world * 64, dtype=torch.uint8, device=mine.device)
            dist.all_gather_into_tensor(allh, mine)
            cache[key] = allh.cpu().numpy().reshape(self.world, 64)
        return cache[key]

    # ---- NVSwitch multicast (NVLS): the switch ORs the copies -------------------------------------------------------------
    def _agree(self, ok: bool) -> bool:
        """every rank succeeded (one all-reduce; also a barrier)"""
        return self.sum_int(1 if ok else 0) == self.world

    def _mc_ready(self, ctx: api.Context, which: int) -> bool:
        """Set the multicast object of this filter up once: rank 0 creates it and hands its file descriptor to the other
        ranks over a Unix socket, every rank adds its device, then binds its copy and maps the object.  Every stage is
        agreed on by all ranks before the next one starts (binding blocks until every device has been added), and any
        failure anywhere makes every rank fall back to another merge for good."""
        import os
        import socket

        if not self.wants_nvls():
            return False
        state = self.__dict__.setdefault("_mc", {})
        key = (id(ctx), which, ctx.bloom_device_words(which))
        if key in state:
            return state[key]
        state[key] = False
        if not self._agree(ctx.bloom_mc_supported(which) and not os.environ.get("KBBQ_NVLS_FAIL")):  # (the variable: tests of the fallback)
            return False
        path = "/tmp/kbbq_mc_%s_%d_%d.sock" % (os.environ.get("MASTER_PORT", "0"), which, len(state))
        ok, srv, fd = True, None, -1
        try:  # stage 1: the object and its descriptor
            if self.rank == 0:
                fd = ctx.bloom_mc_create(which, self.world)
                if os.path.exists(path):
                    os.unlink(path)
                srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
                srv.bind(path)
                srv.listen(self.world)
                srv.settimeout(60)
        except Exception as e:
            print(f"[rank {self.rank}] multicast object: {e}", flush=True)
            ok = False
        if not self._agree(ok):
            if srv is not None:
                srv.close()
            return False
        try:  # stage 2: the descriptor travels (SCM_RIGHTS), the other ranks import it
            if self.rank == 0:
                for _ in range(self.world - 1):
                    conn, _ = srv.accept()
                    socket.send_fds(conn, [b"mc"], [fd])
                    conn.close()
            else:
                c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
                c.settimeout(60)
                c.connect(path)
                _, fds, _, _ = socket.recv_fds(c, 16, 1)
                c.close()
                ctx.bloom_mc_import(which, self.world, fds[0])
                os.close(fds[0])
        except Exception as e:
            print(f"[rank {self.rank}] multicast descriptor: {e}", flush=True)
            ok = False
        finally:
            if srv is not None:
                srv.close()
                try:
                    os.unlink(path)
                except OSError:
                    pass
        if not self._agree(ok):
            return False
        try:  # stage 3: every device joins
            ctx.bloom_mc_add_device(which)
        except api.KbbqError as e:
            print(f"[rank {self.rank}] multicast add device: {e}", flush=True)
            ok = False
        if not self._agree(ok):
            return False
        try:  # stage 4: bind this rank's copy, map the object
            ctx.bloom_mc_bind(which)
        except api.KbbqError as e:
            print(f"[rank {self.rank}] multicast bind: {e}", flush=True)
            ok = False
        state[key] = self._agree(ok)
        return state[key]

    def mc_ready(self, ctx: api.Context, which: int) -> bool:
        """set the multicast merge of this filter up (collective); False when it is not to be had"""
        return self._mc_ready(ctx, which)

    def merge_filter(self, ctx: api.Context, which: int) -> None:
        if self.world == 1:
            return
        g = ctx.bloom_info(which)
        nbytes = g.table_bits // 8
        counter = device_tensor(ctx.bloom_device_counter(which), 8, ctx.device).view(torch.int64)
        if self._mc_ready(ctx, which):
            ctx.sync()          # this rank's inserts are in memory ...
            self.barrier()      # ... and so are everybody else's
            ctx.bloom_or_allreduce_mc(which, self.rank, self.world)
            ctx.sync()
            self.sum_reduce_(counter)  # doubles as the barrier after: every slice is written everywhere
            torch.cuda.current_stream().synchronize()
            self.last_merge = "nvls"
            return
        handles = self._p2p_handles(ctx, which)
        if handles is not None and not self.__dict__.get("_p2p_broken"):
            ctx.sync()          # this rank's inserts are in memory ...
            self.barrier()      # ... and so are everybody else's
            try:
                ctx.bloom_or_allreduce_p2p(which, self.rank, self.world, handles)
                ok = 1
            except api.KbbqError as e:  # no peer access between these devices: every rank must agree to fall back
                print(f"[rank {self.rank}] peer-memory merge unavailable ({e}); using the all-gather merge", flush=True)
                ok = 0
            ctx.sync()
            if self.min_int(ok) == 1:   # doubles as the barrier after: every slice is written everywhere
                self.sum_reduce_(counter)
                torch.cuda.current_stream().synchronize()
                self.last_merge = "p2p"
                return
            self._p2p_broken = True
        ctx.sync()  # the library works on its own stream
        words = device_tensor(ctx.bloom_device_words(which), nbytes, ctx.device).view(torch.int64)

        def or_kernel(first, n, src):
            # src holds words [first, first+n) of a peer; kbbq_bloom_or_from indexes both sides by absolute word
            torch.cuda.current_stream().synchronize()
            ctx.bloom_or_from(which, src.data_ptr() - 8 * first, first, n)
            ctx.sync()

        self.or_reduce_(words, or_kernel)
        self.sum_reduce_(counter)
        torch.cuda.current_stream().synchronize()
        self.last_merge = "allgather"

    def merge_counts(self, ctx: api.Context) -> None:
        if self.world == 1:
            return
        ptr, n = ctx.covariates_device()
        ctx.sync()
        t = device_tensor(ptr, n * 8, ctx.device).view(torch.int64)
        self.sum_reduce_(t)
        torch.cuda.current_stream().synchronize()

    def exclusive_prefix_gpu(self, x: int, device: int) -> int:
        return self.exclusive_prefix(x, device=f"cuda:{device}")


class GpuComm(Comm):
    """Comm whose scalar exchanges also go through NCCL (tensors must live on the rank's GPU)."""

    def __init__(self, rank: int, world: int, device: int):
        super().__init__(rank, world)
        self.device = device

    def exclusive_prefix(self, x: int, device=None) -> int:
        return super().exclusive_prefix(x, device=f"cuda:{self.device}")

    def max_float(self, v: float) -> float:
        t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{self.device}")
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def min_int(self, v: int) -> int:
        t = torch.tensor([v], dtype=torch.int64, device=f"cuda:{self.device}")
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return int(t.item())

    def sum_int(self, v: int) -> int:
        t = torch.tensor([v], dtype=torch.int64, device=f"cuda:{self.device}")
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())

    def barrier(self):
        if self.world > 1:
            dist.barrier(device_ids=[self.device])


def shard_bounds(n_reads: int, world: int) -> np.ndarray:
    """contiguous shards of reads; any assignment gives identical results, contiguous keeps file order per rank"""
    return np.linspace(0, n_reads, world + 1).astype(np.int64)
