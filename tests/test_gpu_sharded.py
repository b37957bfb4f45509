"""The N > 1 path on real GPUs (needs >= 2 devices; skipped otherwise): one process per GPU, each rank runs the CUDA
engine on its shard of the sources and the buses are exchanged over NVLink peer memory inside the kernels
(sbm_peer_export / sbm_peer_connect -- K3 peer stores + arrival counters, gather kernel in front of K4).  Every rank
must hold the output the oracle computes for the whole source set, buffer after buffer, including the SH-HRIR decode
of the summed ambisonic bus and the cross-rank "mixer used" flags.  torch.distributed (gloo) only carries the 64-byte
IPC handles."""
import os
import subprocess
import sys

import pytest

from harness import ROOT

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch, torch.distributed as dist
import substrata_b200 as sb
from substrata_b200.sharding import ShardedMixer, shard_slice
from harness import KIND_HRTF, KIND_PAN, OracleBackend, load_sh_hrir, load_sh_hrir_order1, positions, quat_yaw, rel_rms, synth_pcm
rank, world = int(sys.argv[1]), int(sys.argv[2])
torch.cuda.set_device(rank)
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[3], rank=rank, world_size=world)
F, N, NB = 512, 600, 12
rng = np.random.default_rng(78)
pcm = [synth_pcm(rng, 3001) for _ in range(N)]
pos = positions(rng, N); vol = rng.uniform(0.25, 1, N).astype(np.float32)
start = rng.integers(0, 3001, N); rates = rng.choice([48000, 44100], N)
# the last rank's shard is ambisonic-only and silent for the first buffers: exercises the cross-rank flags and the decoder flush
kinds = np.array([KIND_HRTF if s >= N - N // world else (KIND_PAN if s % 2 == 0 else KIND_HRTF) for s in range(N)])
hr = load_sh_hrir()
def build(be, sel):
    ids = []
    for s in sel:
        sid = be.source_create(int(kinds[s])); ids.append(sid)
        be.source_attach_buffer(sid, be.buffer_create(pcm[s]), int(start[s]), int(rates[s]), True, False)
    ids = np.array(ids, np.int32)
    be.set_source_positions(ids, pos[sel]); be.set_source_volumes(ids, vol[sel])
    return ids
def script(be, ids_of, b):
    t = b * F / 48000.0
    be.set_head([t, 0, 0], quat_yaw(0.5 * t))
    # room effects from buffer 3 on: the mono reflections send is one more row of the exchanged bus, its "fed" flag travels with
    # the others, and every rank runs the (O(1)) room kernel on the summed send
    if b == 0: be.set_room_hrir(load_sh_hrir_order1())
    if b == 3:
        be.enable_room_effects(True)
        be.set_reflection_properties([0.5, 0.0, 1.0], [0, 0, 0, 1], [30.0, 8.0, 24.0], [0.8, 0.8, 0.3, 0.8, 0.8, 0.2], 0.0, 0.7)
    if b == 7: be.enable_room_effects(False)
    last = np.arange(N - N // world, N)
    sel = [ids_of[s] for s in last if s in ids_of]
    if sel and b == 0: be.set_source_volumes(np.array(sel, np.int32), np.zeros(len(sel), np.float32))
    if sel and b == 4: be.set_source_volumes(np.array(sel, np.int32), vol[[s for s in last if s in ids_of]])
lo, hi = shard_slice(N, rank, world)
eng = sb.MixEngine(F, sh_hrir=hr, max_sources=N, device=rank)
mine = np.arange(lo, hi)
ids = build(eng, mine)
ids_of = dict(zip(mine.tolist(), ids.tolist()))
mixer = ShardedMixer(eng, device="cuda:%d" % rank, exchange="peer")
outs, valids = [], []
for b in range(NB):
    script(eng, ids_of, b)
    v, o = mixer.mix()
    outs.append(o.copy()); valids.append(v)
got = np.stack(outs)
if rank == 0:
    orc = OracleBackend(F, sh_hrir=hr)
    oids = build(orc, np.arange(N))
    o_of = dict(zip(range(N), oids.tolist()))
    ref, rv = [], []
    for b in range(NB):
        script(orc, o_of, b)
        v, o = orc.mix()
        ref.append(o); rv.append(v)
    err = rel_rms(got, np.stack(ref))
    print("REL_RMS", err, "VALID", valids, rv)
    assert valids == rv
    assert err <= 1e-5, err
t = torch.from_numpy(got.copy()); t0 = t.clone()
dist.broadcast(t0, 0)
assert torch.equal(t, t0), "ranks disagree"
dist.barrier(); dist.destroy_process_group()
'''


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_memory_sharded_mix_matches_oracle(tmp_path, world):
    if _gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    port = str(31500 + os.getpid() % 2000 + world)
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(world), port], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(world)]
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=300)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    assert "REL_RMS" in outs[0]
