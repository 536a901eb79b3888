"""One process per GPU: the exchange step of the distributed 2-D NTT as a single all-to-all.

The reference moves every (rows_p x cols_q) block with a hand-rolled all-to-all over fresh TCP
connections (PlonkSlave.fft2Prepare -> PlonkPeer.fftExchange, src/worker.rs:280-345, 412-438).
Here the row-phase kernel already stores its output in exchange layout (W contiguous blocks), so
the whole step is ONE `all_to_all_single` over NVLink (NCCL) on the library's device buffers, and
the column-phase kernel reads the received [r][cols] matrix with strided loads: no pack, unpack or
transpose kernels.  This is the only collective on the path; MSM shards need none.

The same code runs on CPU under the `gloo` backend against the kernel-logic emulator (tests only).
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist


class _CudaBuffer:
    """Expose a raw device pointer through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {
            "shape": (nbytes // 8,), "typestr": "<i8", "data": (ptr, False), "version": 3, "strides": None,
        }


def as_tensor(ptr: int, nbytes: int, cuda: bool) -> torch.Tensor:
    """int64 view of `nbytes` at `ptr` (device memory when cuda, host memory otherwise)."""
    if cuda:
        return torch.as_tensor(_CudaBuffer(ptr, nbytes), device="cuda")
    arr = np.ctypeslib.as_array((C.c_int64 * (nbytes // 8)).from_address(ptr))
    return torch.from_numpy(arr)


def make_exchange(group=None, cuda: bool | None = None):
    """Returns exchange(send_ptr, recv_ptr, block_elems) for PlonkSlave.fft2_prepare: block q of my
    send buffer goes to rank q, landing as block `rank` of its recv buffer."""
    world = dist.get_world_size(group)
    if cuda is None:
        cuda = dist.get_backend(group) == "nccl"

    def exchange(send_ptr: int, recv_ptr: int, block_elems: int):
        nbytes = block_elems * 32 * world
        send = as_tensor(send_ptr, nbytes, cuda)
        recv = as_tensor(recv_ptr, nbytes, cuda)
        dist.all_to_all_single(recv, send, group=group)
        if cuda:
            torch.cuda.current_stream().synchronize()

    return exchange


def make_stream_ordered_exchange(ctx, group=None):
    """The exchange enqueued on the context's compute stream (NCCL only): no host synchronisation, so a
    multi-GPU transform is as asynchronous as a single-GPU one.  Pair it with
    Context.fft_exchange_begin_async; block q of send goes to rank q as for make_exchange."""
    world = dist.get_world_size(group)
    stream = torch.cuda.ExternalStream(ctx.compute_stream())

    def exchange(send_ptr: int, recv_ptr: int, block_elems: int):
        nbytes = block_elems * 32 * world
        with torch.cuda.stream(stream):   # the collective starts after the row kernels and the column kernels wait for it
            dist.all_to_all_single(as_tensor(recv_ptr, nbytes, True), as_tensor(send_ptr, nbytes, True), group=group)

    exchange.stream_ordered = True
    return exchange


def msm_shard(n_bases: int, rank: int, world: int):
    """MsmWorkload of `rank`: the global index range of dispatcher2.rs:870-881."""
    return rank * n_bases // world, (rank + 1) * n_bases // world


def attach_peers(ctx, arena_bytes: int, group=None):
    """Fused exchange set-up (one process per GPU of one box): every rank creates its receive arena,
    the CUDA IPC handles are swapped with all_gather, every rank maps the others' arenas.  After
    this PlonkSlave.fft2_prepare / Context.fft_dev_rows_p2p store straight into peer memory and
    the only synchronisation left is a barrier between the row and the column phase."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    handle = ctx.peer_arena_create(arena_bytes)
    handles = [None] * world
    dist.all_gather_object(handles, handle, group=group)
    for q, h in enumerate(handles):
        if q != rank:
            ctx.peer_attach(q, h)
    dist.barrier(group=group)
    return ctx.peer_ready()


def parse_cpulist(text: str):
    """'0-3,8,10-11' (sysfs cpulist format) -> {0, 1, 2, 3, 8, 10, 11}"""
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_local_cpus(device_index: int, sysfs: str = "/sys/bus/pci/devices"):
    """CPUs of the NUMA node the GPU hangs off (its PCIe root complex), from sysfs; None when unknown
    (no such file, a single-node box, or a VM that hides the topology)."""
    try:
        p = torch.cuda.get_device_properties(device_index)
        path = f"{sysfs}/{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0/local_cpulist"
        with open(path) as f:
            cpus = parse_cpulist(f.read())
    except (OSError, ValueError, AttributeError, RuntimeError, AssertionError):
        return None
    return cpus or None


@contextlib.contextmanager
def near_gpu(device_index: int, cpus=None):
    """Run the calling thread on the CPUs next to its GPU while it allocates, fills and hands over pinned
    host buffers.  One process per GPU on a two-socket box otherwise leaves about half of the ranks with their
    staging buffers on the far socket (first touch decides the node), and every host<->device copy of those ranks
    crosses the inter-socket link, which all of them share.  Yields a description of what was done; the previous
    affinity is restored on exit (threads that already exist, e.g. an OpenMP pool, keep theirs)."""
    try:
        before = os.sched_getaffinity(0)
    except (AttributeError, OSError):
        yield "unchanged (no sched_getaffinity)"
        return
    local = cpus if cpus is not None else gpu_local_cpus(device_index)
    want = (local & before) if local else set()
    if not want or want == before:
        yield "unchanged (" + ("GPU-local CPU list unknown" if not local else "GPU-local CPUs outside this process's set" if not want
                               else "already local / single node") + ")"
        return
    try:
        os.sched_setaffinity(0, want)
    except OSError as exc:
        yield f"unchanged (sched_setaffinity: {exc})"
        return
    try:
        yield f"{len(want)} of {len(before)} CPUs: the GPU's NUMA node"
    finally:
        try:
            os.sched_setaffinity(0, before)
        except OSError:
            pass
