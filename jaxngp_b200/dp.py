"""Ray-sharded data parallelism: what each rank owns and the one exchange per step.

Training: every rank draws its own ray batch (disjoint index streams), runs the whole step locally and
exchanges ONE flat gradient buffer [hash-table grad | MLP grads] per step.  The all-reduce is split into its two
halves around the optimizer (ZeRO-1 style): reduce-scatter the gradient, run Adam on this rank's 1/N of the
parameters and moments (the optimizer is HBM-bound at 28 B/parameter, so its time shrinks with N), all-gather
the updated parameters.  Same bytes on NVLink as one all-reduce; the optimizer divides by the world size
(NgpAdamDescriptor.grad_scale) and replicas stay bit-identical.
Inference: image rows are dealt to ranks in interleaved tiles; the only collective is the final gather.
"""
import torch
import torch.distributed as dist


def allreduce_flat_gradients(flat_grads: torch.Tensor, group=None) -> torch.Tensor:
    """SUM all-reduce of the flat gradient buffer, in place (NCCL on GPUs, gloo in the CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    return flat_grads


def shard_bounds(n_padded: int, rank: int, world_size: int):
    """[begin, end) of the flat parameter buffer owned by `rank` (n_padded is a multiple of 4 * world_size)."""
    if n_padded % (4 * world_size) != 0:
        raise ValueError(f"flat buffer length {n_padded} is not a multiple of 4 * world_size = {4 * world_size}: shards would be "
                         "unaligned or leave a tail nobody owns (Trainer pads to lcm(32, 4 * world_size))")
    per = n_padded // world_size
    return rank * per, (rank + 1) * per


def reduce_scatter_flat_gradients(flat_grads: torch.Tensor, rank: int, world_size: int, group=None) -> torch.Tensor:
    """SUM reduce-scatter: returns this rank's shard of the summed gradient (a view into ``flat_grads``).
    Half of an all-reduce; the other half is ``all_gather_flat_parameters`` after the sharded optimizer step."""
    lo, hi = shard_bounds(flat_grads.numel(), rank, world_size)
    shard = flat_grads[lo:hi]
    # world_size == 1: a replica trainer inside a multi-rank job (bench.py's render model) exchanges nothing
    if world_size == 1 or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return shard
    if dist.get_backend(group) == "nccl":
        dist.reduce_scatter_tensor(shard, flat_grads, op=dist.ReduceOp.SUM, group=group)  # in place on the rank's slice
    else:  # gloo (CPU tests) has no reduce-scatter: same result through an all-reduce
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    return shard


def all_gather_flat_parameters(flat_params: torch.Tensor, rank: int, world_size: int, group=None) -> torch.Tensor:
    """Every rank updated its own shard of ``flat_params``: gather the shards in place."""
    if world_size > 1 and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        lo, hi = shard_bounds(flat_params.numel(), rank, world_size)
        if dist.get_backend(group) == "nccl":
            dist.all_gather_into_tensor(flat_params, flat_params[lo:hi], group=group)
        else:
            parts = [torch.empty(hi - lo, dtype=flat_params.dtype, device=flat_params.device) for _ in range(world_size)]
            dist.all_gather(parts, flat_params[lo:hi].clone(), group=group)
            flat_params.copy_(torch.cat(parts))
    return flat_params


def allreduce_density_grid(grid: torch.Tensor, group=None) -> torch.Tensor:
    """MAX all-reduce of the density grid after each rank updated its own share of the cells."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(grid, op=dist.ReduceOp.MAX, group=group)
    return grid


def tile_rows(height: int, rank: int, world_size: int, tile: int = 32):
    """Row indices of the image owned by `rank`: interleaved `tile`-row bands (empty-space rays are cheap,
    so contiguous slabs would be unbalanced)."""
    rows = torch.arange(height)
    return rows[((rows // tile) % world_size) == rank]


class ImageGather:
    """Final gather of a tile-sharded frame ("no collective beyond the final gather").  ``tile_rows`` is a
    pure function of (height, rank, world), so every rank knows every rank's rows up front: ONE
    ``all_gather_into_tensor`` of the padded u8 row bands per frame and one precomputed index scatter -- no size
    exchange, no host synchronisation."""

    def __init__(self, height: int, width: int, channels: int, rank: int, world_size: int, device, dtype=torch.uint8,
                 tile: int = 32, group=None):
        self.group, self.world, self.rank = group, world_size, rank
        rows = [tile_rows(height, r, world_size, tile) for r in range(world_size)]
        self.max_rows = max(int(r.numel()) for r in rows)
        self.n_local = int(rows[rank].numel())
        self.local_rows = rows[rank].to(device)
        self.send = torch.zeros(self.max_rows, width, channels, dtype=dtype, device=device)
        self.recv = torch.empty(world_size, self.max_rows, width, channels, dtype=dtype, device=device)
        # where row k of rank r's band goes in the full image, and which (r, k) pairs are real rows
        src = torch.cat([torch.arange(r.numel()) + i * self.max_rows for i, r in enumerate(rows)])
        self.src = src.to(device)
        self.dst = torch.cat(rows).to(device)
        self.out = torch.empty(height, width, channels, dtype=dtype, device=device)

    def __call__(self, local_pixels: torch.Tensor) -> torch.Tensor:
        """local_pixels [n_local_rows, W, C] -> the full image [H, W, C] on every rank."""
        if self.world == 1 or not (dist.is_available() and dist.is_initialized()):
            self.out[self.local_rows] = local_pixels
            return self.out
        self.send[: self.n_local].copy_(local_pixels)
        if dist.get_backend(self.group) == "nccl":
            dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
        else:
            parts = [torch.empty_like(self.send) for _ in range(self.world)]
            dist.all_gather(parts, self.send, group=self.group)
            self.recv.copy_(torch.stack(parts))
        self.out[self.dst] = self.recv.reshape(self.world * self.max_rows, *self.recv.shape[2:])[self.src]
        return self.out


def gather_image(local_rows: torch.Tensor, local_pixels: torch.Tensor, height: int, group=None) -> torch.Tensor:
    """All-gather the per-rank row bands into the full image [height, W, C] on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        out = torch.empty((height,) + tuple(local_pixels.shape[1:]), dtype=local_pixels.dtype, device=local_pixels.device)
        out[local_rows] = local_pixels
        return out
    counts = [torch.zeros(1, dtype=torch.int64, device=local_pixels.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([local_rows.numel()], dtype=torch.int64, device=local_pixels.device), group=group)
    max_rows = int(max(int(c) for c in counts))
    pad_rows = torch.full((max_rows,), -1, dtype=torch.int64, device=local_pixels.device)
    pad_rows[: local_rows.numel()] = local_rows.to(local_pixels.device)
    pad_pix = torch.zeros((max_rows,) + tuple(local_pixels.shape[1:]), dtype=local_pixels.dtype, device=local_pixels.device)
    pad_pix[: local_rows.numel()] = local_pixels
    all_rows = [torch.empty_like(pad_rows) for _ in range(world)]
    all_pix = [torch.empty_like(pad_pix) for _ in range(world)]
    dist.all_gather(all_rows, pad_rows, group=group)
    dist.all_gather(all_pix, pad_pix, group=group)
    out = torch.empty((height,) + tuple(local_pixels.shape[1:]), dtype=local_pixels.dtype, device=local_pixels.device)
    for r, p in zip(all_rows, all_pix):
        ok = r >= 0
        out[r[ok]] = p[ok]
    return out
