"""GPU parity tests proper: the CUDA path, through the C ABI (include/sbm_audio.h), against
  (a) the committed reference outputs (tests/golden/scenarios.npz, made by the compiled reference),
  (b) the oracle (oracle/libmix_oracle.so) on the same seeded inputs at sizes it finishes in seconds,
  (c) size-independent properties at BASELINE.json's sizes (linearity of the mix in the source set).
Bar (BASELINE.json north_star): <= 1e-5 relative RMS over the interleaved stereo output of consecutive buffers."""
import os

import numpy as np
import pytest

import substrata_b200 as sb
from harness import (GOLDEN_DIR, KIND_HRTF, KIND_PAN, KIND_STEREO, OracleBackend, Scenario, load_sh_hrir, rel_rms, churn_golden, run_churn, run_pipelined)
from golden.make_golden import GOLDEN_SCENARIOS

pytestmark = pytest.mark.gpu
TOL = 1e-5


def engine(F, hrir=None, max_sources=4096, **kw):
    return sb.MixEngine(F, sh_hrir=hrir, max_sources=max_sources, **kw)


@pytest.mark.parametrize("name", sorted(GOLDEN_SCENARIOS))
def test_cuda_matches_reference_golden(name):
    kw, F, nb, taps = GOLDEN_SCENARIOS[name]
    g = np.load(os.path.join(GOLDEN_DIR, "scenarios.npz"))
    hr = load_sh_hrir()
    out, valid = Scenario(frames=F, **kw).run(engine(F, hr[:, :taps] if taps else hr), nb)
    assert valid == g[name + "__valid"].tolist()
    assert rel_rms(out, g[name]) <= TOL


@pytest.mark.parametrize("kw,F,nb,taps", [
    # C1 / C2 / C4 / C3 / C5 shapes at oracle-friendly sizes, moving listener (gain ramps every buffer)
    (dict(n_sources=16, kind=KIND_PAN, loop_len=48000), 512, 8, 0),
    (dict(n_sources=4096, kind=KIND_PAN, feeder="resample", loop_len=4096, seed=21), 512, 8, 0),
    (dict(n_sources=4096, kind=KIND_PAN, feeder="doppler", loop_len=4096, seed=22), 512, 8, 0),
    (dict(n_sources=2048, kind=KIND_HRTF, loop_len=4096, seed=23), 1024, 6, 128),
    (dict(n_sources=1024, kind=KIND_HRTF, feeder="resample", loop_len=4096, seed=24, volume_toggle=True), 512, 12, 0),
    (dict(n_sources=300, mixed_kinds=True, occlusion_toggle=True, volume_toggle=True, seed=25, n_shared_buffers=7, feeder="resample"), 256, 12, 0),
    (dict(n_sources=64, kind=KIND_PAN, feeder="ring", seed=26, rates=(48000, 44100), rate_probs=(0.5, 0.5)), 256, 10, 0),
    (dict(n_sources=256, kind=KIND_HRTF, feeder="ring", seed=37, rates=(48000, 24000), rate_probs=(0.7, 0.3), batched_ring_push=True), 256, 12, 0),
    (dict(n_sources=200, kind=KIND_PAN, loop_len=100, feeder="resample", seed=27), 1024, 6, 0),     # loops shorter than a buffer
    (dict(n_sources=100, kind=KIND_HRTF, seed=28), 32, 10, 0),                                      # smallest buffer Resonance accepts
    (dict(n_sources=100, mixed_kinds=True, seed=29), 1000, 5, 0),                                   # frames not a multiple of the tile
    # loop lengths that are not a multiple of 4 samples: the wrapped window's second TMA segment is re-aligned (kMdAdjust)
    (dict(n_sources=512, kind=KIND_PAN, feeder="doppler", loop_len=4099, seed=32), 512, 8, 0),
    (dict(n_sources=512, kind=KIND_HRTF, feeder="doppler", loop_len=4098, seed=33), 512, 8, 0),
    (dict(n_sources=512, kind=KIND_HRTF, feeder="doppler", loop_len=4096, seed=34, volume_toggle=True), 256, 8, 0),
    (dict(n_sources=300, mixed_kinds=True, occlusion_toggle=True, volume_toggle=True, seed=35, feeder="resample", loop_len=3001), 1024, 8, 0),
    (dict(n_sources=300, mixed_kinds=True, occlusion_toggle=True, seed=36, feeder="doppler", loop_len=2049), 480, 8, 0),
    # three mix sources per AudioSource (the bike), retuned every buffer: the generic multi-read-head path
    (dict(n_sources=600, mixed_kinds=True, feeder="mix3", loop_len=4099, seed=38, n_shared_buffers=31, occlusion_toggle=True), 512, 8, 0),
    (dict(n_sources=200, kind=KIND_HRTF, feeder="mix3", loop_len=2048, seed=39, n_shared_buffers=17), 1024, 6, 0),
    # room effects (reflections sends in the ROOM instantiations of both mix kernels, room kernel, order-1 decode) at sizes where
    # every CTA works through several batches
    (dict(n_sources=4096, kind=KIND_HRTF, loop_len=4096, seed=41, volume_toggle=True, room=dict(on=1, off=9, moves={5: ((2.0, 0.0, 1.0), (24.0, 7.0, 18.0))})), 512, 12, 0),
    (dict(n_sources=8192, kind=KIND_PAN, feeder="doppler", loop_len=4096, seed=42, room=dict(on=0)), 512, 6, 0),
    (dict(n_sources=900, mixed_kinds=True, occlusion_toggle=True, volume_toggle=True, feeder="resample", seed=43, loop_len=3001,
          room=dict(on=2, cutoff=1200.0, rot=(0.0, 0.25881905, 0.0, 0.96592583))), 256, 12, 0),
    (dict(n_sources=2048, kind=KIND_HRTF, feeder="doppler", loop_len=4098, seed=44, room=dict(on=1)), 1024, 5, 0),
    (dict(n_sources=300, mixed_kinds=True, occlusion_toggle=True, seed=45, room=dict(on=0, off=4)), 480, 14, 0),
])
def test_cuda_matches_oracle(kw, F, nb, taps):
    hr = load_sh_hrir()
    h = hr[:, :taps] if taps else hr
    sc = Scenario(frames=F, **kw)
    oo, ov = sc.run(OracleBackend(F, sh_hrir=h), nb)
    go, gv = sc.run(engine(F, h, max_sources=max(4096, kw["n_sources"])), nb)
    assert ov == gv
    assert rel_rms(go, oo) <= TOL


@pytest.mark.parametrize("kw,F", [
    (dict(n_sources=600, mixed_kinds=True, feeder="resample", volume_toggle=True, occlusion_toggle=True, seed=61), 512),
    (dict(n_sources=128, kind=KIND_HRTF, feeder="ring", rates=(48000, 24000), rate_probs=(0.5, 0.5), seed=62, batched_ring_push=True), 256),
    (dict(n_sources=2048, kind=KIND_PAN, feeder="doppler", loop_len=4096, seed=63), 512),
])
def test_cuda_submit_collect_equals_synchronous_mix(kw, F):
    """Two buffers in flight (sbm_mix_submit / sbm_mix_collect) give bit-identical buffers to sbm_mix, and match the oracle."""
    hr = load_sh_hrir()
    a, va = Scenario(frames=F, **kw).run(engine(F, hr), 10)
    b, vb = run_pipelined(Scenario(frames=F, **kw), engine(F, hr), 10)
    o, vo = Scenario(frames=F, **kw).run(OracleBackend(F, sh_hrir=hr), 10)
    assert va == vb == vo
    assert np.array_equal(a, b)
    assert rel_rms(b, o) <= TOL


def test_ring_push_from_pinned_memory_equals_staged_push():
    """A packet payload in page-locked memory (sbm_host_alloc_pinned) skips the staging copy; the rings must end up the same."""
    F, nv = 256, 96
    rng = np.random.default_rng(11)
    pos = rng.uniform([-20, -5, -20], [20, 5, 20], (nv, 3)).astype(np.float32) + np.float32(1.5)
    outs = []
    for pinned in (False, True):
        e = engine(F, load_sh_hrir(), max_sources=nv)
        first = e.source_create(KIND_HRTF, nv)
        ids = np.arange(first, first + nv, dtype=np.int32)
        for sid in ids:
            e.source_attach_ring(int(sid), 48000)
        e.set_source_positions(ids, pos)
        offs = np.arange(nv + 1, dtype=np.uint64) * np.uint64(480)
        prng = np.random.default_rng(12)
        res = []
        for b in range(8):
            payload = (0.3 * prng.uniform(-1, 1, nv * 480)).astype(np.float32)
            if b % 2 == 0:          # 480-sample packets every other 256-frame buffer
                if pinned:
                    buf = sb.pinned_array(nv * 480)
                    buf[:] = payload
                    e.sources_ring_push_flat(ids, buf, offs)
                else:
                    e.sources_ring_push_flat(ids, payload, offs)
            res.append(e.mix()[1].copy())
        outs.append(np.stack(res))
    assert outs[0].any() and np.array_equal(outs[0], outs[1])


def test_cuda_source_churn_matches_oracle():
    """Sources created and destroyed between every pair of buffers, one-shots finishing on their own (AudioEngine.cpp:900-943,
    1022-1052): outputs, valid flags and finished-id lists follow the oracle while slots and list positions are recycled."""
    hr = load_sh_hrir()
    oo, ov, of = run_churn(OracleBackend(512, sh_hrir=hr), 512, n_buffers=12, n_start=1500, per_buffer=200)
    go, gv, gf = run_churn(engine(512, hr, max_sources=4096), 512, n_buffers=12, n_start=1500, per_buffer=200)
    assert ov == gv and of == gf
    assert rel_rms(go, oo) <= TOL


def test_cuda_source_churn_matches_reference_golden():
    ro, rv, rf = churn_golden()
    go, gv, gf = run_churn(engine(256, load_sh_hrir(), max_sources=1024), 256)
    assert gv == rv and gf == rf
    assert rel_rms(go, ro) <= TOL


def test_cuda_feeder_bookkeeping():
    F = 256
    rng = np.random.default_rng(3)
    pcm = [(0.4 * rng.uniform(-1, 1, n)).astype(np.float32) for n in (700, 1024, 5000)]
    res = {}
    for be in (OracleBackend(F), engine(F)):
        bufs = [be.buffer_create(p) for p in pcm]
        a = be.source_create(KIND_PAN); be.source_attach_buffer(a, bufs[0], 0, 48000, False, True)
        b = be.source_create(KIND_PAN); be.source_attach_buffer(b, bufs[1], 100, 44100, True, False)
        c = be.source_create(KIND_PAN); be.source_attach_buffer(c, bufs[2], 0, 48000, False, False); be.source_set_eof_marker(c, 1000)
        be.set_source_positions(np.array([a, b, c], np.int32), np.array([[2, 0, -3], [-4, 1, 2], [0, 0, -5]], np.float32))
        outs, fin, q = [], [], []
        for _ in range(8):
            v, o = be.mix()
            outs.append(o); fin.append(be.poll_finished()); q.append(be.source_query(b)["cur_read_i"])
        res[be.name] = (np.stack(outs), fin, q)
    assert res["oracle"][1] == res["b200"][1] and res["oracle"][2] == res["b200"][2]
    assert rel_rms(res["b200"][0], res["oracle"][0]) <= 1e-6


def test_cuda_vs_fp64_bus_oracle_16k():
    """SURVEY finding 9: at larger N the f32 summation order matters; compare against the fp64-bus oracle too."""
    F, nb = 512, 4
    sc = Scenario(n_sources=16384, frames=F, kind=KIND_PAN, feeder="doppler", loop_len=2048, seed=31)
    oo, _ = sc.run(OracleBackend(F, double_bus=True), nb)
    go, _ = sc.run(engine(F, max_sources=16384), nb)
    assert rel_rms(go, oo) <= TOL


def test_linearity_at_full_size():
    """Size-independent property at C4's per-rank size (131 072 doppler sources, F = 512): the mix is linear in the source
    set, so mix(A u B) == mix(A) + mix(B) for a split of the same sources over two engines (what sharding relies on)."""
    import torch
    F, N, L = 512, 131072, 4096
    rng = np.random.default_rng(77)
    n_buf = 256
    pcm = (0.3 * rng.uniform(-1, 1, (n_buf, L))).astype(np.float32)
    pos = rng.uniform([-40, -10, -40], [40, 10, 40], (N, 3)).astype(np.float32)
    pos[np.linalg.norm(pos, axis=1) < 1.05] += 2.0
    vol = rng.uniform(0.25, 1.0, N).astype(np.float32)
    start = rng.uniform(0, L, N)
    delta = rng.uniform(0.92, 1.10, N)
    mixf = rng.uniform(0.5, 1.0, N).astype(np.float32)

    def run(sel):
        e = engine(F, max_sources=len(sel))
        bufs = [e.buffer_create(p) for p in pcm]
        first = e.source_create(KIND_PAN, len(sel))
        ids = np.arange(first, first + len(sel), dtype=np.int32)
        e.sources_attach_mix1(ids, np.array(bufs, np.int32)[sel % n_buf], start[sel], delta[sel], mixf[sel])
        e.set_source_positions(ids, pos[sel]); e.set_source_volumes(ids, vol[sel])
        outs = []
        for b in range(3):
            e.set_head([0.01 * b, 0, 0], [0, np.sin(0.005 * b), 0, np.cos(0.005 * b)])
            outs.append(e.mix()[1].astype(np.float64))
        return np.stack(outs)

    al = np.arange(N)
    whole, a, b = run(al), run(al[: N // 2]), run(al[N // 2:])
    assert rel_rms(a + b, whole) <= TOL


def test_linearity_at_c3_size():
    """C3's full size (65 536 ambisonic sources, 1024-frame buffers, 128-tap SH-HRIRs): mix(A u B) == mix(A) + mix(B) --
    encode, bus sum and the SH-domain decode are linear, and the decode's tail carry must add up across buffers too."""
    F, N, L, taps = 1024, 65536, 4096, 128
    rng = np.random.default_rng(79)
    n_buf = 128
    pcm = (0.3 * rng.uniform(-1, 1, (n_buf, L))).astype(np.float32)
    pos = rng.uniform([-40, -10, -40], [40, 10, 40], (N, 3)).astype(np.float32)
    pos[np.linalg.norm(pos, axis=1) < 1.05] += 2.0
    vol = rng.uniform(0.25, 1.0, N).astype(np.float32)
    start = rng.integers(0, L, N)
    hr = load_sh_hrir()[:, :taps]

    def run(sel):
        e = engine(F, hr, max_sources=len(sel))
        bufs = np.array([e.buffer_create(p) for p in pcm], np.int32)
        first = e.source_create(KIND_HRTF, len(sel))
        ids = np.arange(first, first + len(sel), dtype=np.int32)
        e.sources_attach_buffer(ids, bufs[sel % n_buf], start[sel], np.full(len(sel), 48000, np.int32))
        e.set_source_positions(ids, pos[sel]); e.set_source_volumes(ids, vol[sel])
        outs = []
        for b in range(3):
            e.set_head([0.02 * b, 0, 0], [0, np.sin(0.004 * b), 0, np.cos(0.004 * b)])
            outs.append(e.mix()[1].astype(np.float64))
        return np.stack(outs)

    al = np.arange(N)
    whole, a, b = run(al), run(al[0::2]), run(al[1::2])
    assert rel_rms(a + b, whole) <= TOL


def test_still_scene_is_time_invariant_at_c5_size():
    """C5's source count (32 768 ambient loops, 256-frame buffers): with listener and sources at rest and loops whose length
    is a multiple of the buffer, every loop period reproduces the same buffers bit for bit (no drift in read indices,
    no stray gain ramps, deterministic reduction order)."""
    F, N, L = 256, 32768, 1024
    rng = np.random.default_rng(80)
    n_buf = 64
    pcm = (0.3 * rng.uniform(-1, 1, (n_buf, L))).astype(np.float32)
    pos = rng.uniform([-40, -10, -40], [40, 10, 40], (N, 3)).astype(np.float32)
    pos[np.linalg.norm(pos, axis=1) < 1.05] += 2.0
    e = engine(F, load_sh_hrir(), max_sources=N)
    bufs = np.array([e.buffer_create(p) for p in pcm], np.int32)
    first = e.source_create(KIND_HRTF, N)
    ids = np.arange(first, first + N, dtype=np.int32)
    e.sources_attach_buffer(ids, bufs[np.arange(N) % n_buf], (rng.integers(0, L // F, N) * F), np.full(N, 48000, np.int32))
    e.set_source_positions(ids, pos)
    outs = [e.mix()[1].copy() for _ in range(3 * L // F)]
    period = L // F
    for k in range(period, 2 * period):          # second period == third period (the first one contains the decoder's start-up)
        assert np.array_equal(outs[k], outs[k + period])


def test_resampler_dropin_matches_reference_golden():
    """glare::AudioResampler drop-in against outputs of the reference class (tests/golden/resampler.npz: the reference's
    own test pattern, AudioResampler.cpp:213-287 -- ramp src[i] = i fed in 4-sample destination chunks, 13 rate pairs),
    including the number of source samples it asks for per call."""
    from golden.make_golden import RESAMPLER_PAIRS
    g = np.load(os.path.join(GOLDEN_DIR, "resampler.npz"))
    rs = sb.AudioResampler([s for s, _ in RESAMPLER_PAIRS if _ == 48000], 48000)
    keys = ["%d_%d" % (s, d) for s, d in RESAMPLER_PAIRS if d == 48000]
    fed = np.zeros(rs.count, np.int64)
    outs = [[] for _ in keys]
    for step in range(64):
        need = rs.num_src_samples_needed(4)
        for k, key in enumerate(keys):
            assert int(need[k]) == int(g[key + "__needs"][step])
        chunks = [np.arange(fed[k], fed[k] + int(need[k]), dtype=np.float32) for k in range(rs.count)]
        fed += need.astype(np.int64)
        o = rs.resample(4, chunks)
        for k in range(rs.count):
            outs[k].append(o[k])
    for k, key in enumerate(keys):
        assert np.max(np.abs(np.concatenate(outs[k]) - g[key]) / np.maximum(1.0, np.abs(g[key]))) <= 1e-6
    for s, d in RESAMPLER_PAIRS:
        if d == 48000:
            continue
        r = sb.AudioResampler([s], d)
        key = "%d_%d" % (s, d)
        fed1, o1 = 0, []
        for step in range(64):
            need = int(r.num_src_samples_needed(4)[0])
            assert need == int(g[key + "__needs"][step])
            o1.append(r.resample(4, [np.arange(fed1, fed1 + need, dtype=np.float32)])[0])
            fed1 += need
        assert np.max(np.abs(np.concatenate(o1) - g[key]) / np.maximum(1.0, np.abs(g[key]))) <= 1e-6
