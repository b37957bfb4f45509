"""The N > 1 path on CPU: world_size 2 over gloo.  Each rank runs the REAL host engine (device-emulation build) on its
shard of the sources; the partial buses are summed with one all-reduce between sbm_mix_begin and sbm_mix_finish
(substrata_b200/sharding.py); every rank must then hold the output the oracle computes for the whole source set --
including the SH-HRIR decode of the summed ambisonic bus and the cross-rank "mixer used" flags."""
import os
import subprocess
import sys

import numpy as np
import pytest

from harness import ROOT

WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch, torch.distributed as dist
import substrata_b200 as sb
from substrata_b200.sharding import ShardedMixer, shard_slice
from harness import KIND_HRTF, KIND_PAN, OracleBackend, load_sh_hrir, positions, quat_yaw, rel_rms, synth_pcm
rank, world = int(sys.argv[1]), int(sys.argv[2])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[3], rank=rank, world_size=world)
EMU = os.path.join({root!r}, "tests", "native", "libsbm_emu.so")
F, N, NB = 256, 40, 8
rng = np.random.default_rng(77)
pcm = [synth_pcm(rng, 2048) for _ in range(N)]
pos = positions(rng, N); vol = rng.uniform(0.25, 1, N).astype(np.float32)
start = rng.integers(0, 2048, N); rates = rng.choice([48000, 44100], N)
# rank 1's shard contains ambisonic sources only from buffer 3 on (they start silent): exercises the cross-rank flags
kinds = np.array([KIND_PAN if s % 2 == 0 else KIND_HRTF for s in range(N)])
hr = load_sh_hrir()
def build(be, sel):
    ids = []
    for s in sel:
        sid = be.source_create(int(kinds[s])); ids.append(sid)
        be.source_attach_buffer(sid, be.buffer_create(pcm[s]), int(start[s]), int(rates[s]), True, False)
    ids = np.array(ids, np.int32)
    be.set_source_positions(ids, pos[sel]); be.set_source_volumes(ids, vol[sel])
    return ids
lo, hi = shard_slice(N, rank, world)
eng = sb.MixEngine(F, sh_hrir=hr, max_sources=N, _lib_path=EMU)
build(eng, np.arange(lo, hi))
mixer = ShardedMixer(eng)
outs, valids = [], []
for b in range(NB):
    t = b * F / 48000.0
    eng.set_head([t, 0, 0], quat_yaw(0.5 * t))
    v, o = mixer.mix()
    outs.append(o.copy()); valids.append(v)
got = np.stack(outs)
if rank == 0:
    orc = OracleBackend(F, sh_hrir=hr)
    build(orc, np.arange(N))
    ref = []
    for b in range(NB):
        t = b * F / 48000.0
        orc.set_head([t, 0, 0], quat_yaw(0.5 * t))
        ref.append(orc.mix()[1])
    err = rel_rms(got, np.stack(ref))
    print("REL_RMS", err, "VALID", valids)
    assert err <= 1e-5, err
# every rank holds the same output
t = torch.from_numpy(got.copy()); t0 = t.clone()
dist.broadcast(t0, 0)
assert torch.equal(t, t0)
dist.barrier(); dist.destroy_process_group()
'''


def test_two_rank_sharded_mix_matches_oracle(tmp_path):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "native")])
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2", port], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    assert "REL_RMS" in outs[0]
