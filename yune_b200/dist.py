"""Multi-GPU rule of the hot path: the path shards by SAMPLE INDEX.

Every rank renders all pixels for a disjoint range of sample indices into its own fp32 sum buffer; there is no traffic
while rendering.  One SUM reduce of the buffers to rank 0 follows, then rank 0 divides by the total sample count and
tonemaps.  Because random numbers are addressed by (seed, pixel, sample), the reduced image does not depend on the number
of ranks (up to fp32 summation order).  torch.distributed is plumbing only (NCCL over NVLink on GPUs, gloo in CPU tests).
"""
import torch
import torch.distributed as dist


def shard_samples(total_spp, rank, world_size):
    """Contiguous, balanced split of [0, total_spp): returns (spp_begin, spp_count) of `rank`."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(int(total_spp), int(world_size))
    begin = rank * base + min(rank, extra)
    return begin, base + (1 if rank < extra else 0)


def reduce_sum_to_root(sum_tensor, root=0):
    """In-place SUM reduce of the accumulation buffer (W*H*4 fp32) to `root`; the single exchange step of the path."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(sum_tensor, dst=root, op=dist.ReduceOp.SUM)
    return sum_tensor


class _DevicePointer:
    """Expose a raw device pointer (from yune_sum_device_ptr) to torch through __cuda_array_interface__."""

    def __init__(self, ptr, n, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 3, "strides": None}


def sum_buffer_as_tensor(renderer):
    """Zero-copy torch view of a RendererCore's device-resident accumulation buffer (for NCCL): the int64 fixed-point buffer when
    the context accumulates deterministically (default; after the reduce call yune_sum_refresh on the root), else the fp32 sums."""
    import ctypes as C
    p, n = C.c_void_p(), C.c_size_t()
    if renderer.cl_manager.getOption("deterministic"):
        renderer.cl_manager.check(renderer._lib.yune_sum_fixed_device_ptr(renderer._ctx, C.byref(p), C.byref(n)))
        return torch.as_tensor(_DevicePointer(p.value, n.value // 8, "<i8"), device="cuda")
    renderer.cl_manager.check(renderer._lib.yune_sum_device_ptr(renderer._ctx, C.byref(p), C.byref(n)))
    return torch.as_tensor(_DevicePointer(p.value, n.value // 4), device="cuda")
