"""Row-sharded FLAT index across the GPUs of one box: one process per GPU (torch.distributed), each rank
holds a contiguous block of rows in its own HBM and answers every query locally; one all-gather of the
per-rank [B,k] results (NCCL over NVLink on GPUs, gloo in CPU tests) and a k-way merge with the same
(distance,label) order give a result identical to the single-GPU one.

Reference analog: the cluster fan-out + merge of src/query/fanout.cc:159-220 over gRPC
(src/coordinator/coordinator.proto:127-157) — here without the network.
"""
import ctypes as C

import numpy as np

from . import _lib as L


def shard_bounds(n_rows, world_size, rank):
    """Contiguous row block of `rank`: sizes differ by at most one row."""
    base, rem = divmod(int(n_rows), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def merge_topk_host(dist, labels, counts, k):
    """k-way merge of per-shard ascending results on the host (used by the gloo/CPU tests of the
    multi-rank plumbing and as the specification of the GPU merge kernel).
    dist/labels: [G,B,k]; counts: [G,B].  Returns ([B,k] dist, [B,k] labels, [B] n)."""
    G, B, _ = dist.shape
    out_d = np.full((B, k), np.inf, np.float32)
    out_l = np.full((B, k), np.iinfo(np.uint64).max, np.uint64)
    out_n = np.zeros(B, np.uint32)
    for b in range(B):
        items = []
        for g in range(G):
            c = int(counts[g, b])
            items.extend(zip(dist[g, b, :c].tolist(), labels[g, b, :c].tolist()))
        items.sort()
        items = items[:k]
        out_n[b] = len(items)
        for j, (d, l) in enumerate(items):
            out_d[b, j] = d
            out_l[b, j] = l
    return out_d, out_l, out_n


def packed_result_bytes(B, k):
    """Size of one rank's packed result block: labels u64 [B][k] | dist f32 [B][k] | n u32 [B], padded to 256 bytes
    (the host-side statement of vkgpu_packed_result_bytes, include/vkgpu.h)."""
    return (B * k * 12 + B * 4 + 255) & ~255


def pack_result_host(dist, labels, counts):
    """numpy [B,k] f32, [B,k] u64, [B] u32 -> one uint8 block in the packed layout."""
    B, k = dist.shape
    blk = np.zeros(packed_result_bytes(B, k), np.uint8)
    blk[: B * k * 8] = np.ascontiguousarray(labels, np.uint64).view(np.uint8).reshape(-1)
    blk[B * k * 8: B * k * 12] = np.ascontiguousarray(dist, np.float32).view(np.uint8).reshape(-1)
    blk[B * k * 12: B * k * 12 + B * 4] = np.ascontiguousarray(counts, np.uint32).view(np.uint8).reshape(-1)
    return blk


def unpack_results_host(all_blocks, G, B, k):
    """The G gathered blocks (uint8, rank order) -> ([G,B,k] dist, [G,B,k] labels, [G,B] counts)."""
    n = packed_result_bytes(B, k)
    blocks = np.asarray(all_blocks, np.uint8).reshape(G, n)
    labels = np.stack([blocks[g, : B * k * 8].copy().view(np.uint64).reshape(B, k) for g in range(G)])
    dist = np.stack([blocks[g, B * k * 8: B * k * 12].copy().view(np.float32).reshape(B, k) for g in range(G)])
    counts = np.stack([blocks[g, B * k * 12: B * k * 12 + B * 4].copy().view(np.uint32) for g in range(G)])
    return dist, labels, counts


class ShardedFlat:
    """One rank's view of a row-sharded FLAT index.  `local` is a VectorFlat holding this rank's rows with
    GLOBAL labels.  search_device() runs local search -> all_gather -> merge entirely on the device."""

    def __init__(self, local_index, dist_module=None, device=None):
        self.local = local_index
        self.dist = dist_module
        self.world = dist_module.get_world_size() if dist_module is not None and dist_module.is_initialized() else 1
        self.rank = dist_module.get_rank() if self.world > 1 else 0
        self.device = device
        self._lib = L.lib()

    def search_device(self, d_Q, k, stream_ptr, out=None):
        """d_Q: torch CUDA tensor [B,dim] fp32.  Returns torch tensors (dist [B,k], labels [B,k] int64 view of
        u64, n [B] int32) on the device, merged over all ranks (every rank gets the full result)."""
        import torch

        B = d_Q.shape[0]
        dev = d_Q.device
        if out is None:
            out = self.alloc_out(B, k, dev)
        loc_d, loc_l, loc_n, all_d, all_l, all_n, mer_d, mer_l, mer_n = out[:9]
        L.check(self._lib.vkgpu_search_batch_device(self.local.handle(), d_Q.data_ptr(), B, k, 0, loc_d.data_ptr(),
                                                    loc_l.data_ptr(), loc_n.data_ptr(), stream_ptr))
        if self.world == 1:
            return loc_d, loc_l, loc_n
        # the single exchange step of the path: ONE all-gather of the rank's packed block
        # (labels | distances | counts = B*k*12 + B*4 bytes, vkgpu_packed_result_bytes), then the k-way merge
        loc_packed, all_packed = out[9], out[10]
        self.dist.all_gather_into_tensor(all_packed, loc_packed)
        L.check(self._lib.vkgpu_merge_topk_packed_device(dev.index, all_packed.data_ptr(), self.world, B, k,
                                                         mer_d.data_ptr(), mer_l.data_ptr(), mer_n.data_ptr(),
                                                         stream_ptr))
        return mer_d, mer_l, mer_n

    def search_host(self, h_Q, k, stream_ptr, out, pinned):
        """End-to-end form of search_device for HOST buffers: `h_Q` is a pinned [B,dim] fp32 torch tensor,
        `pinned` = (dist [B,k] f32, labels [B,k] i64, n [B] i32) pinned host tensors that receive the merged
        result.  One H2D copy of the queries, the sharded search + exchange + merge on the device, one D2H copy
        of the final [B,k] — the per-shard partial results never visit the host."""
        import torch

        d_Q = out[11]
        d_Q.copy_(h_Q, non_blocking=True)
        res = self.search_device(d_Q, k, stream_ptr, out)
        for dst, src in zip(pinned, res):
            dst.copy_(src, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return pinned

    def alloc_out(self, B, k, dev):
        import torch

        G = self.world
        mk = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)
        # the local result lives inside one packed block (labels | distances | counts) so that the exchange is a
        # single collective; loc_* are views into it
        nbytes = int(self._lib.vkgpu_packed_result_bytes(B, k))
        loc_packed = torch.zeros((nbytes,), dtype=torch.uint8, device=dev)
        all_packed = torch.empty((G * nbytes,), dtype=torch.uint8, device=dev)
        loc_l = loc_packed[: B * k * 8].view(torch.int64).view(B, k)
        loc_d = loc_packed[B * k * 8: B * k * 12].view(torch.float32).view(B, k)
        loc_n = loc_packed[B * k * 12: B * k * 12 + B * 4].view(torch.int32)
        return (loc_d, loc_l, loc_n,
                mk((G, B, k), torch.float32), mk((G, B, k), torch.int64), mk((G, B), torch.int32),
                mk((B, k), torch.float32), mk((B, k), torch.int64), mk((B,), torch.int32), loc_packed, all_packed,
                mk((B, self.local.dimensions_), torch.float32))
