"""Gradient exchange fused with the optimizer over NVLink (``csrc/exchange.cu``).

``dp.py`` exchanges the flat gradient as reduce-scatter -> Adam on the rank's shard -> all-gather (two NCCL launches
around our optimizer kernel).  ``PeerExchange`` does the same step as ONE kernel per rank: the flat parameter and
gradient buffers live in symmetric memory (every rank maps every peer's copy, plus one multicast mapping when the
NVSwitch offers it), and ``ngp_adam_step_exchange`` reads the summed shard straight from the peers, applies Adam and
writes the new parameters into every replica.

The default at world_size > 1 (``Trainer(exchange="auto")``; ``NGP_B200_EXCHANGE=nccl`` selects ``dp.py``).  Measured on
2 B200s (tools/exchange_check.py, C2 buffer of 12.2 M floats): replicas bit-identical, parameters and moments EQUAL to
the NCCL arm's, 0.158 ms (multimem) / 0.240 ms (per-peer loads) against 0.198 ms for the three NCCL-arm launches.
There is no CPU path; ``tests/test_dp_gloo.py`` covers the host-side layout only.
"""
import os

import torch
import torch.distributed as dist

from . import _lib, descriptors

#: first u32 word of the signal pads used by ngp_adam_step_exchange (torch's own barrier()/put_signal use the low words)
SIGNAL_BASE = 1024
MAX_WORLD = 8


MODES = ("auto", "nccl", "peer", "peer-p2p")


def requested_mode(default="auto"):
    """``auto`` (the fused kernel, NCCL if peers cannot be mapped), ``nccl`` (dp.py), ``peer`` (multicast if available,
    else per-peer loads) or ``peer-p2p`` (never multicast)."""
    mode = os.environ.get("NGP_B200_EXCHANGE", default)
    if mode not in MODES:
        raise ValueError(f"NGP_B200_EXCHANGE must be one of {MODES}, got {mode!r}")
    return mode


def blocks_for(pad_bytes: int, world: int, want: int = 64) -> int:
    """CTAs per launch: each needs ``world`` signal words above SIGNAL_BASE; at most one per SM (they spin on peers).

    64 by default: the kernel is bound by requests in flight over the link, not by SMs, and the SMs it leaves alone run the
    next batch's march beside it.  Training step on C2, ms, by CTA count: N = 8: 32 0.576, 64 0.578, 96 0.584, 128 0.596,
    148 0.599; N = 2: 64 0.567, 128 0.571 (profiles/exchange_knobs_r02.txt)."""
    room = (pad_bytes // 4 - SIGNAL_BASE) // world
    if room < 1:
        raise _lib.NgpError(f"signal pad of {pad_bytes} bytes has no room above word {SIGNAL_BASE} for {world} ranks")
    return max(1, min(want, room, 148))


class PeerExchange:
    """Symmetric flat buffers of one trainer and the fused exchange launch over them.

    ``numel`` is the padded length of the flat buffers (a multiple of 4 * world); rank r owns
    ``dp.shard_bounds(numel, r, world)``.  Construction is collective: every rank of ``group`` must build it with the
    same ``numel``."""

    def __init__(self, numel: int, rank: int, world_size: int, device, group=None, mode="peer", n_blocks=None):
        import torch.distributed._symmetric_memory as symm_mem  # imported here: CPU-only hosts never reach this class

        if not (dist.is_available() and dist.is_initialized()):
            raise _lib.NgpError("the peer exchange needs an initialised NCCL process group (one process per GPU)")
        if world_size < 2 or world_size > MAX_WORLD:
            raise _lib.NgpError(f"the peer exchange supports 2..{MAX_WORLD} ranks of one NVLink domain, got {world_size}")
        if numel % (4 * world_size) != 0:
            raise _lib.NgpError(f"flat buffer length {numel} is not a multiple of 4 * world_size")
        group = group if group is not None else dist.group.WORLD
        self.rank, self.world, self.device = rank, world_size, torch.device(device)
        self.params = symm_mem.empty(numel, dtype=torch.float32, device=self.device)
        self.grads = symm_mem.empty(numel, dtype=torch.float32, device=self.device)
        self.params.zero_()
        self.grads.zero_()
        self._h_params = symm_mem.rendezvous(self.params, group=group)
        self._h_grads = symm_mem.rendezvous(self.grads, group=group)
        if self._h_grads.rank != rank or self._h_grads.world_size != world_size:
            raise _lib.NgpError("symmetric-memory rank/world differ from the trainer's")
        # The handles describe the ALLOCATION a tensor lives in (torch may carve several tensors out of one pooled
        # block): peer r's copy of a tensor sits at buffer_ptrs[r] + (this tensor's offset inside the local block), and
        # symmetric allocation makes that offset the same on every rank.  Same for the multicast mapping.
        def mapped(tensor, handle):
            bases = [int(p) for p in handle.buffer_ptrs]
            offset = tensor.data_ptr() - bases[rank]
            if offset < 0:
                raise _lib.NgpError("symmetric tensor lies below the block its handle describes")
            if offset % 16 != 0:
                raise _lib.NgpError("symmetric tensor is not 16-byte aligned inside its block")
            mc = int(handle.multicast_ptr)
            return [b + offset for b in bases], (mc + offset if mc else 0)

        g_ptrs, mc_g = mapped(self.grads, self._h_grads)
        p_ptrs, mc_p = mapped(self.params, self._h_params)
        self.use_multimem = mode == "peer" and mc_p != 0 and mc_g != 0
        self._mc_params, self._mc_grads = (mc_p, mc_g) if self.use_multimem else (0, 0)
        # device arrays of u64 addresses (torch has no uint64 arithmetic; int64 carries the same bits)
        as_dev = lambda ptrs: torch.tensor(ptrs, dtype=torch.int64, device=self.device)  # noqa: E731
        self._grads_ptrs, self._params_ptrs = as_dev(g_ptrs), as_dev(p_ptrs)
        self._signal_ptrs = as_dev([int(p) for p in self._h_grads.signal_pad_ptrs])
        # how long a block waits for a peer before the kernel traps (a rank that never arrives fails the launch with an
        # error instead of spinning on the GPU until somebody kills the job); 0 = the kernel's default, 20 s
        self.timeout_ms = int(os.environ.get("NGP_B200_EXCHANGE_TIMEOUT_MS", "0"))
        if n_blocks is None and os.environ.get("NGP_B200_EXCHANGE_BLOCKS"):
            n_blocks = int(os.environ["NGP_B200_EXCHANGE_BLOCKS"])
        self.n_blocks = blocks_for(int(symm_mem.get_signal_pad_size()), world_size) if n_blocks is None else int(n_blocks)
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)  # every replica zeroed and mapped before the first launch touches a peer

    def step(self, step_dev, adam_m, adam_v, adam_desc: bytes, shard_begin: int):
        """Enqueue the fused exchange + Adam on the current stream.  ``adam_desc`` is the shard's NgpAdamDescriptor
        (grad_scale = 1 / world)."""
        desc = descriptors.make_adam_exchange_descriptor(adam_desc, shard_begin, self.rank, self.world,
                                                         self.use_multimem, self.n_blocks, SIGNAL_BASE, self.timeout_ms)
        _lib.call("ngp_adam_step_exchange", [step_dev, adam_m, adam_v, self._grads_ptrs, self._params_ptrs,
                                             self._signal_ptrs, self._mc_grads, self._mc_params], desc)
