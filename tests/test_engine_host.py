"""No-GPU tests of the product's HOST side and of the arithmetic the kernels share with it.

The real host engine (substrata_b200/csrc/engine.cc: id/slot tables, patch merging, dirty uploads, rings, finished
sources) is linked against tests/native/device_emu.cc -- a host emulation of the device seam that calls the same
per-source (k1_core.h) and per-sample (k2_core.h) functions the CUDA kernels call -- and driven through the C ABI with
the scenarios whose reference outputs are committed under tests/golden/.  The kernels' parallel decomposition itself is
only checked on the GPU (tests/test_gpu_parity.py)."""
import os
import subprocess

import numpy as np
import pytest

import substrata_b200 as sb
from harness import (GOLDEN_DIR, KIND_HRTF, KIND_PAN, ROOT, OracleBackend, Scenario, load_sh_hrir, rel_rms, churn_golden, run_churn, run_pipelined)
from golden.make_golden import GOLDEN_SCENARIOS

EMU_DIR = os.path.join(ROOT, "tests", "native")
EMU = os.path.join(EMU_DIR, "libsbm_emu.so")
TOL = 1e-5          # the product bar (BASELINE.json north_star)


@pytest.fixture(scope="module", autouse=True)
def build_emu():
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])


def emu_engine(F, hrir=None, max_sources=1024):
    return sb.MixEngine(F, sh_hrir=hrir, max_sources=max_sources, _lib_path=EMU)


@pytest.mark.parametrize("name", sorted(GOLDEN_SCENARIOS))
def test_host_engine_matches_reference_golden(name):
    kw, F, nb, taps = GOLDEN_SCENARIOS[name]
    g = np.load(os.path.join(GOLDEN_DIR, "scenarios.npz"))
    hr = load_sh_hrir()
    out, valid = Scenario(frames=F, **kw).run(emu_engine(F, hr[:, :taps] if taps else hr), nb)
    assert valid == g[name + "__valid"].tolist()
    assert rel_rms(out, g[name]) <= TOL


def test_feeder_bookkeeping_matches_oracle():
    """Rows a2/a17 through the C ABI: one-shot removal, EOF marker, looping resampled read index (AudioEngine.cpp:393-425,496-517)."""
    F = 256
    rng = np.random.default_rng(3)
    pcm = [(0.4 * rng.uniform(-1, 1, n)).astype(np.float32) for n in (700, 1024, 5000)]
    res = {}
    for be in (OracleBackend(F), emu_engine(F)):
        bufs = [be.buffer_create(p) for p in pcm]
        a = be.source_create(KIND_PAN); be.source_attach_buffer(a, bufs[0], 0, 48000, False, True)
        b = be.source_create(KIND_PAN); be.source_attach_buffer(b, bufs[1], 100, 44100, True, False)
        c = be.source_create(KIND_PAN); be.source_attach_buffer(c, bufs[2], 0, 48000, False, False); be.source_set_eof_marker(c, 1000)
        ids = np.array([a, b, c], np.int32)
        be.set_source_positions(ids, np.array([[2, 0, -3], [-4, 1, 2], [0, 0, -5]], np.float32))
        outs, fin, q = [], [], []
        for _ in range(8):
            v, o = be.mix()
            outs.append(o); fin.append(be.poll_finished()); q.append(be.source_query(b)["cur_read_i"])
        res[be.name] = (np.stack(outs), fin, q)
    oo, of, oq = res["oracle"]
    eo, ef, eq = res["b200"]
    assert of == ef and oq == eq
    assert rel_rms(eo, oo) <= 1e-6


def test_slot_reuse_and_destroy():
    """DestroySource + CreateSoundObjectSource churn (AudioEngine.cpp:900-943, playOneShotSound :1022-1052): ids keep
    increasing, slots are recycled with clean DSP state."""
    F = 256
    rng = np.random.default_rng(5)
    pcm = (0.3 * rng.uniform(-1, 1, 4096)).astype(np.float32)
    outs = {}
    for be in (OracleBackend(F), emu_engine(F, max_sources=4)):
        buf = be.buffer_create(pcm)
        o = []
        ids = []
        for round_ in range(6):
            sid = be.source_create(KIND_PAN)
            ids.append(sid)
            be.source_attach_buffer(sid, buf, 17 * round_, 48000, True, False)
            be.set_source_positions([sid], np.array([[1.0 + round_, 0.5, -2.0]], np.float32))
            for _ in range(2):
                o.append(be.mix()[1])
            if round_ % 2 == 1:
                be.source_destroy(ids[-2]); be.source_destroy(ids[-1])
        outs[be.name] = (np.stack(o), ids)
    assert outs["oracle"][1] == outs["b200"][1] == list(range(6))
    assert rel_rms(outs["b200"][0], outs["oracle"][0]) <= 1e-6


def test_capacity_and_argument_errors():
    e = emu_engine(256, max_sources=2)
    e.source_create(KIND_PAN, 2)
    with pytest.raises(sb.SbmError) as ei:
        e.source_create(KIND_PAN)
    assert ei.value.code == -5
    with pytest.raises(sb.SbmError):
        e.source_create(KIND_HRTF)            # engine created without SH-HRIRs
    # queued setters ignore unknown ids like Resonance (resonance_audio_api_impl.cc:551-554) and count them
    e.set_source_positions([99], np.zeros((1, 3), np.float32))
    e.mix()
    assert e.stats()["ignored_setter_items"] == 1
    buf = e.buffer_create(np.zeros(100, np.float32))
    with pytest.raises(sb.SbmError):          # AudioEngine::addSource rejects rates outside 8000..48000 (AudioEngine.cpp:884-885)
        e.source_attach_buffer(0, buf, 0, 96000)
    with pytest.raises(sb.SbmError):
        sb.MixEngine(16, _lib_path=EMU)       # Resonance accepts 32..16384 frames (resonance_audio_api_impl.cc:78-94)


def test_empty_engine_is_invalid_and_silent():
    e = emu_engine(128)
    v, o = e.mix()
    assert v == 0 and not o.any()


def test_batched_ring_push_equals_per_source_pushes():
    """sbm_source_ring_push_batch (one staged H2D + scatter) must leave every ring exactly as the
    per-source sbm_source_ring_push calls do -- including ring wrap, dropped packets (underflow) and mixed rates."""
    F, nb = 256, 12
    hr = load_sh_hrir()
    kw = dict(n_sources=24, frames=F, kind=KIND_HRTF, feeder="ring", rates=(48000, 24000), rate_probs=(0.5, 0.5), seed=91)
    a, va = Scenario(**kw).run(emu_engine(F, hr), nb)
    b, vb = Scenario(batched_ring_push=True, **kw).run(emu_engine(F, hr), nb)
    o, vo = Scenario(**kw).run(OracleBackend(F, sh_hrir=hr), nb)
    assert va == vb == vo
    assert np.array_equal(a, b)
    assert rel_rms(b, o) <= TOL


def test_batched_ring_push_is_all_or_nothing():
    e = emu_engine(256)
    ids = [e.source_create(KIND_PAN) for _ in range(2)]
    for sid in ids:
        e.source_attach_ring(sid, 48000, capacity=1000)
    e.sources_ring_push(ids, [np.ones(600, np.float32), np.ones(300, np.float32)])
    with pytest.raises(sb.SbmError) as ei:      # the first ring would overflow: neither packet is accepted
        e.sources_ring_push([ids[1], ids[0]], [np.ones(100, np.float32), np.ones(500, np.float32)])
    assert ei.value.code == -5
    assert e.source_ring_size(ids[0]) == 600 and e.source_ring_size(ids[1]) == 300
    e.sources_ring_push([ids[1], ids[1]], [np.ones(100, np.float32), np.ones(50, np.float32)])      # the same id twice appends in order
    assert e.source_ring_size(ids[1]) == 450
    with pytest.raises(sb.SbmError):
        e.sources_ring_push([12345], [np.ones(4, np.float32)])


@pytest.mark.parametrize("kw,F", [
    (dict(n_sources=40, mixed_kinds=True, feeder="resample", volume_toggle=True, occlusion_toggle=True, seed=61), 256),
    (dict(n_sources=24, kind=KIND_HRTF, feeder="ring", rates=(48000, 24000), rate_probs=(0.5, 0.5), seed=62, batched_ring_push=True), 256),
])
def test_submit_collect_equals_synchronous_mix(kw, F):
    """sbm_mix_submit / sbm_mix_collect (two buffers in flight) must produce exactly the buffers sbm_mix produces."""
    hr = load_sh_hrir()
    a, va = Scenario(frames=F, **kw).run(emu_engine(F, hr), 10)
    b, vb = run_pipelined(Scenario(frames=F, **kw), emu_engine(F, hr), 10)
    assert va == vb and np.array_equal(a, b)


def test_submit_collect_reports_finished_sources_and_limits_depth():
    F = 128
    e = emu_engine(F)
    rng = np.random.default_rng(4)
    buf = e.buffer_create((0.3 * rng.uniform(-1, 1, 300)).astype(np.float32))
    sid = e.source_create(KIND_PAN)
    e.source_attach_buffer(sid, buf, 0, 48000, False, True)       # one-shot: stops after 300 samples
    e.set_source_positions([sid], np.array([[1.0, 0.0, -2.0]], np.float32))
    e.mix_submit(); e.mix_submit()
    with pytest.raises(sb.SbmError):
        e.mix_submit()                                               # at most two in flight
    with pytest.raises(sb.SbmError):
        e.mix_begin()                                                # the synchronous form needs an empty pipeline
    fin = []
    for _ in range(4):
        e.mix_collect(); fin += e.poll_finished()
        e.mix_submit()
    e.mix_collect(); e.mix_collect()
    fin += e.poll_finished()
    assert fin == [sid]
    with pytest.raises(sb.SbmError):
        e.mix_collect()                                              # nothing in flight
    v, o = e.mix()                                                   # and the synchronous call works again
    assert v == 0


def test_source_churn_matches_oracle():
    """Sources created and destroyed between every pair of buffers, one-shots finishing on their own: outputs, valid flags
    and the finished-id lists must follow the oracle (ids keep increasing; slots, list positions and DSP state are recycled)."""
    hr = load_sh_hrir()
    oo, ov, of = run_churn(OracleBackend(256, sh_hrir=hr), 256)
    eo, ev, ef = run_churn(emu_engine(256, hr, max_sources=1024), 256)
    assert ov == ev and of == ef
    assert rel_rms(eo, oo) <= TOL


def test_source_churn_matches_reference_golden():
    ro, rv, rf = churn_golden()
    eo, ev, ef = run_churn(emu_engine(256, load_sh_hrir(), max_sources=1024), 256)
    assert ev == rv and ef == rf
    assert rel_rms(eo, ro) <= TOL


@pytest.mark.parametrize("name,override", [
    ("c2", dict(n=300, pool=8)), ("c3", dict(n=200, pool=8, frames=256)), ("c4_rank", dict(n=400, pool=8)),
    ("c5", dict(n=260, voices=40, pool=8)),
])
def test_full_scene_harness_on_host_engine(name, override):
    """The full-size scene harness (tests/test_gpu_fullsize.py runs it at BASELINE's sizes on the GPU) at toy sizes: the host
    engine on the emulated device against the partitioned oracle, f32 and fp64 bus."""
    from harness import full_scene_parity
    for dbl in (False, True):
        r = full_scene_parity(name, lambda sc: emu_engine(sc.F, sc.hrir, max_sources=sc.n), parts=3, double_bus=dbl, **override)
        assert r["valid_equal"] and r["rel_rms"] <= TOL, r


def test_setters_never_block_and_are_race_free_under_tsan():
    """Four threads hammer every queued setter while the mixer thread runs sbm_mix_submit / sbm_mix_collect and churns sources:
    ThreadSanitizer build of the real host engine on the emulated device (tests/native/test_setters_tsan.cc).  No TSAN report,
    and the setters' latency stays in microseconds -- they never wait for the mixer, which here spends milliseconds per
    buffer inside the engine lock (Resonance: lock-free task queue, RA/utils/lockless_task_queue.cc:43-51)."""
    subprocess.check_call(["make", "-s", "-C", EMU_DIR, "test_setters_tsan"])
    r = subprocess.run([os.path.join(EMU_DIR, "test_setters_tsan")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ThreadSanitizer" not in r.stderr, r.stderr[-4000:]
    f = dict(zip(r.stdout.split()[0::2], r.stdout.split()[1::2]))
    assert int(f["buffers"]) >= 4 and int(f["setter_calls"]) > 1000
    assert float(f["median_us"]) < 100.0 and float(f["p999_us"]) < 2000.0, r.stdout      # (worst_us is OS scheduling noise under TSAN)


def test_mix_source_entries_and_buffers_are_recycled():
    """MixZ ranges return to a free list when their source dies (1365 bike creations used to exhaust the table), and
    sbm_buffer_destroy frees a PCM buffer once its last source is gone while rejecting new attachments."""
    F = 256
    e = sb.MixEngine(F, max_sources=8, max_mix_sources=6, _lib_path=EMU)
    rng = np.random.default_rng(5)
    bufs = [e.buffer_create((0.3 * rng.uniform(-1, 1, 1000)).astype(np.float32)) for _ in range(3)]
    for k in range(40):      # 40 x 3 mix sources through a 6-entry table
        sid = e.source_create(KIND_PAN)
        e.source_attach_mix(sid, 48000, bufs, [0.0, 10.0, 20.0], [1.0, 0.5, 1.87], [0.6, 0.3, 0.45])
        e.set_source_positions([sid], np.array([[2.0, 0.0, -3.0]], np.float32))
        v, o = e.mix()
        assert v == 1 and o.any()
        if k % 2:
            e.source_attach_mix(sid, 48000, bufs[:1], [0.0], [1.0], [1.0])      # re-attach with another length: old range recycled too
            e.mix()
        e.source_destroy(sid)
    # buffer life time
    sid = e.source_create(KIND_PAN)
    e.source_attach_buffer(sid, bufs[0], 0, 48000)
    assert e._l.sbm_buffer_destroy(e.h, bufs[0]) == 0           # still referenced: stays alive for its source ...
    v, o = e.mix()
    assert v == 1
    s2 = e.source_create(KIND_PAN)
    with pytest.raises(sb.SbmError):                              # ... but takes no new attachments
        e.source_attach_buffer(s2, bufs[0], 0, 48000)
    e.source_destroy(sid)
    e.mix()
    assert e._l.sbm_buffer_destroy(e.h, bufs[0]) == -4          # gone
    assert e._l.sbm_buffer_destroy(e.h, bufs[1]) == 0 and e._l.sbm_buffer_destroy(e.h, 99) == -4


def test_scattered_and_contiguous_setters_reach_the_device():
    """The per-buffer upload block: a few scattered ids (index list + scatter) and a whole batch of consecutive ids (dense
    range fast path) must leave the same state as the oracle sees."""
    F, N = 256, 3000
    rng = np.random.default_rng(8)
    pcm = (0.3 * rng.uniform(-1, 1, 2048)).astype(np.float32)
    outs = {}
    for be in (OracleBackend(F), sb.MixEngine(F, max_sources=N, _lib_path=EMU)):
        r = np.random.default_rng(9)
        b = be.buffer_create(pcm)
        first = be.sources_create(KIND_PAN, N) if hasattr(be, "sources_create") else be.source_create(KIND_PAN, N)
        ids = np.arange(first, first + N, dtype=np.int32)
        be.sources_attach_buffer(ids, np.full(N, b, np.int32), r.integers(0, 2048, N), np.full(N, 48000, np.int32))
        res = []
        for k in range(6):
            if k % 2 == 0:
                be.set_source_positions(ids, r.uniform(-20, 20, (N, 3)).astype(np.float32) + np.float32(1.5))      # all: dense
                be.set_source_volumes(ids[100:2900], r.uniform(0.2, 1, 2800).astype(np.float32))
            else:
                sel = np.sort(r.choice(N, 37, replace=False))
                be.set_source_positions(ids[sel], r.uniform(-20, 20, (37, 3)).astype(np.float32) + np.float32(1.5))  # few: scattered
                be.set_source_occlusions(ids[sel[:5]], np.ones(5, np.float32))
            res.append(be.mix()[1].copy())
        outs[be.name] = np.stack(res)
    assert rel_rms(outs["b200"], outs["oracle"]) <= TOL


def test_partitioned_oracle_equals_single_oracle():
    """The threaded oracle of the full-size runs (block partitions summed in f64) renders what ONE oracle engine renders,
    including the buffers in which every 8th source is switched off and on again (the GainMixer sweep is shared state)."""
    from harness import BigScene, FULL_SCENES
    kw = dict(FULL_SCENES["c3"]); kw.update(n=2048, pool=16)
    sc = BigScene(**kw)
    a, va = sc.run(sc.oracle())
    b, vb = sc.run_oracle_partitioned(16)
    assert va == vb
    for k in range(sc.nb):
        assert rel_rms(b[k], a[k]) <= 3e-6


def test_streaming_boundary_of_the_audio_engine_mirror():
    """SURVEY 8f row 1 through the glare::AudioEngine mirror on the emulated device (tests/native/test_streaming_host.cc):
    addSourceFromStreamingSoundFile / needNewStreamerForPath / the stream maps, two sources sharing one decoder start in sync
    (AudioEngine.cpp:1124-1131), a decoder-reported sampling rate (StreamerThread.cpp:98-105), a paused source with its own
    decoder unpaused by seekToStartAndUnpauseAudio, removeSource freeing the decoder -- output against the oracle fed with
    exactly the samples the synthetic decoder produced."""
    subprocess.check_call(["make", "-s", "-C", EMU_DIR, "test_streaming_host"])
    r = subprocess.run([os.path.join(EMU_DIR, "test_streaming_host")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rel_rms" in r.stdout


def test_out_device_names_the_block_of_the_last_buffer():
    """sbm_out_device: the engine keeps one result block per buffer in flight; the pointer it returns must name the block of the
    buffer finished (or submitted) LAST -- on the emulated device that block is plain host memory and can be read back."""
    import ctypes
    F = 128
    e = emu_engine(F)
    rng = np.random.default_rng(5)
    buf = e.buffer_create((0.5 * rng.uniform(-1, 1, 4000)).astype(np.float32))
    s = e.source_create(KIND_PAN); e.source_attach_buffer(s, buf, 0, 48000, True, False)
    e.set_source_positions(np.array([s], np.int32), np.array([[1.0, 0.0, -2.0]], np.float32))
    seen = set()
    for k in range(5):
        _, o = e.mix()
        ptr = ctypes.c_void_p()
        assert sb.lib(EMU).sbm_out_device(e.h, ctypes.byref(ptr)) == 0
        seen.add(ptr.value)
        dev = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_float)), shape=(2 * F,))
        assert np.array_equal(dev, np.asarray(o).reshape(-1))
    assert len(seen) == 2          # two blocks, used alternately
