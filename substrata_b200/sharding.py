"""Multi-GPU mix: sources shard across ranks (one process per GPU), each rank mixes its slice into its partial bus, one
sum over ranks combines the buses, then the (linear, O(1)) SH-HRIR decode + stereo mix runs on the summed bus.

The sum of the [18][frames] f32 buses plus 4 "mixer used" flag floats -- 37 KB at 512 frames -- is the only exchange
step of the path.  On GPUs it runs inside the engine's own kernels over NVLink peer memory
(sbm_peer_export / sbm_peer_connect, see ShardedMixer); an all-reduce on the engine's stream between sbm_mix_begin and
sbm_mix_finish is kept as the comparison arm and for the CPU tests (gloo, host engine on the device-emulation build).
Per-source state never moves.  `torch.distributed` is plumbing only."""
import ctypes as C

import numpy as np


def shard_slice(n_sources, rank, world):
    """Contiguous block partition of source indices: [lo, hi) for `rank`."""
    lo = (n_sources * rank) // world
    hi = (n_sources * (rank + 1)) // world
    return lo, hi


class _DevPtr:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class ShardedMixer:
    """Wraps one rank's MixEngine.

    exchange="peer" (GPUs, the product path): the ranks swap the IPC handles of their gather areas once
    (torch.distributed is only the transport for those 64 bytes), sbm_peer_connect maps them, and from then on
    sbm_mix itself is the collective -- K3 stores the bus into every rank's area over NVLink, the gather kernel in front
    of K4 sums them; no NCCL call and no host synchronisation per buffer.
    exchange="allreduce": sbm_mix_begin -> all_reduce(bus) on the engine's stream -> sbm_mix_finish (NCCL on GPUs for
    comparison; gloo on the host-emulation build in the CPU tests)."""

    def __init__(self, engine, group=None, device=None, exchange=None):
        import torch
        import torch.distributed as dist
        self.engine, self.dist, self.group, self.torch = engine, dist, group, torch
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.exchange = exchange or ("peer" if device is not None else "allreduce")
        ptr, n = engine.bus_device()
        if device is None:      # host memory (device-emulation build): view it in place
            buf = (C.c_float * n).from_address(ptr)
            self.bus = torch.from_numpy(np.frombuffer(buf, dtype=np.float32))
            self.stream = None
            self.exchange = "allreduce"
        else:
            self.bus = torch.as_tensor(_DevPtr(ptr, n), device=device)
            self.stream = torch.cuda.ExternalStream(engine.stream(), device=device)
        if self.exchange == "peer" and self.world > 1:
            mine = engine.peer_export()
            handles = [None] * self.world
            dist.all_gather_object(handles, mine, group=group)
            engine.peer_connect(dist.get_rank(group), self.world, handles)
            dist.barrier(group)

    def mix(self, out=None):
        if self.exchange == "peer" or self.world == 1:
            return self.engine.mix(out)
        self.engine.mix_begin()
        if self.stream is not None:
            with self.torch.cuda.stream(self.stream):
                self.dist.all_reduce(self.bus, group=self.group)
        else:
            self.dist.all_reduce(self.bus, group=self.group)
        return self.engine.mix_finish(out)

    def submit(self):
        """Pipelined form (peer exchange or single GPU only): enqueue one buffer, collect it later."""
        assert self.exchange == "peer" or self.world == 1, "submit/collect needs the in-kernel exchange"
        self.engine.mix_submit()

    def collect(self, out=None):
        return self.engine.mix_collect(out)
