"""Multi-GPU mix: sources shard across ranks (one process per GPU), each rank mixes its slice into its partial bus, one
sum over ranks combines the buses, then the (linear, O(1)) SH-HRIR decode + stereo mix runs on the summed bus.

The collective is the only exchange step of the path: an all-reduce of the [18][frames] f32 bus plus 4
"mixer used" flag floats -- 37 KB at 512 frames -- issued on the engine's own CUDA stream between sbm_mix_begin and
sbm_mix_finish, so it is ordered after K3 and before K4 without host synchronisation.  Per-source state never moves.
`torch.distributed` is plumbing only (NCCL on GPUs; gloo in the CPU tests, which run the host engine on the
device-emulation build)."""
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
    """Wraps one rank's MixEngine.  mix() == sbm_mix_begin -> all_reduce(bus) -> sbm_mix_finish."""

    def __init__(self, engine, group=None, device=None):
        import torch
        import torch.distributed as dist
        self.engine, self.dist, self.group, self.torch = engine, dist, group, torch
        ptr, n = engine.bus_device()
        if device is None:      # host memory (device-emulation build): view it in place
            buf = (C.c_float * n).from_address(ptr)
            self.bus = torch.from_numpy(np.frombuffer(buf, dtype=np.float32))
            self.stream = None
        else:
            self.bus = torch.as_tensor(_DevPtr(ptr, n), device=device)
            self.stream = torch.cuda.ExternalStream(engine.stream(), device=device)

    def mix(self, out=None):
        self.engine.mix_begin()
        if self.dist.is_initialized() and self.dist.get_world_size(self.group) > 1:
            if self.stream is not None:
                with self.torch.cuda.stream(self.stream):
                    self.dist.all_reduce(self.bus, group=self.group)
            else:
                self.dist.all_reduce(self.bus, group=self.group)
        return self.engine.mix_finish(out)
