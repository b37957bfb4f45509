"""Pins oracle/mix_oracle.c against the reference itself: (a) the committed golden outputs of the compiled reference
(tests/golden/scenarios.npz, made by tests/golden/make_golden.py) and (b) when oracle/_ref/libsubstrata_ref.so is
present, the live reference on further scenarios.  Bar: <= 1e-6 relative RMS at these sizes (the reference's own
hash-order-vs-id-order f32 summation noise is 2e-7 at N = 64, SURVEY finding 9), an order below the 1e-5 product bar."""
import os

import numpy as np
import pytest

from harness import (GOLDEN_DIR, KIND_HRTF, KIND_PAN, OracleBackend, RefBackend, Scenario, load_sh_hrir, ref_library_path, run_churn,
                     rel_rms)
from golden.make_golden import GOLDEN_SCENARIOS

TOL = 1e-6


@pytest.mark.parametrize("name", sorted(GOLDEN_SCENARIOS))
def test_oracle_matches_golden(name):
    kw, F, nb, taps = GOLDEN_SCENARIOS[name]
    g = np.load(os.path.join(GOLDEN_DIR, "scenarios.npz"))
    hr = load_sh_hrir()
    be = OracleBackend(F, sh_hrir=hr[:, :taps] if taps else hr)
    out, valid = Scenario(frames=F, **kw).run(be, nb)
    assert valid == g[name + "__valid"].tolist()
    assert rel_rms(out, g[name]) <= TOL


needs_ref = pytest.mark.skipif(ref_library_path() is None, reason="compiled reference (oracle/_ref) not present")


@needs_ref
@pytest.mark.parametrize("kw,F,nb,taps", [
    (dict(n_sources=200, kind=KIND_PAN, feeder="doppler", seed=7), 512, 6, 0),
    (dict(n_sources=200, kind=KIND_HRTF, feeder="resample", seed=8, volume_toggle=True), 512, 12, 0),
    (dict(n_sources=100, mixed_kinds=True, occlusion_toggle=True, seed=9, n_shared_buffers=5), 256, 12, 0),
    (dict(n_sources=50, kind=KIND_HRTF, seed=10, moving=False), 1024, 4, 64),
])
def test_oracle_matches_live_reference(kw, F, nb, taps):
    hr = load_sh_hrir()
    sc = Scenario(frames=F, **kw)
    ro, rv = sc.run(RefBackend(F, hrir_taps=taps), nb)
    oo, ov = sc.run(OracleBackend(F, sh_hrir=hr[:, :taps] if taps else hr), nb)
    assert rv == ov
    assert rel_rms(oo, ro) <= TOL


@needs_ref
def test_feeder_state_and_eof_bookkeeping():
    """Rows a2/a17: non-looping sources stop at the end of their buffer (remove_on_finish) or at the EOF marker, and
    cur_read_i advances identically (AudioEngine.cpp:393-425,496-517)."""
    F = 256
    rng = np.random.default_rng(3)
    pcm = [(0.4 * rng.uniform(-1, 1, n)).astype(np.float32) for n in (700, 1024, 5000)]
    res = {}
    for make in (lambda: RefBackend(F), lambda: OracleBackend(F)):
        be = make()
        bufs = [be.buffer_create(p) for p in pcm]
        a = be.source_create(KIND_PAN); be.source_attach_buffer(a, bufs[0], 0, 48000, False, True)   # one-shot
        b = be.source_create(KIND_PAN); be.source_attach_buffer(b, bufs[1], 100, 44100, True, False)  # looping, resampled
        c = be.source_create(KIND_PAN); be.source_attach_buffer(c, bufs[2], 0, 48000, False, False); be.source_set_eof_marker(c, 1000)
        ids = np.array([a, b, c], np.int32)
        be.set_source_positions(ids, np.array([[2, 0, -3], [-4, 1, 2], [0, 0, -5]], np.float32))
        outs, fin, q = [], [], []
        for _ in range(8):
            v, o = be.mix()
            outs.append(o); fin.append(be.poll_finished()); q.append(be.source_query(b))
        res[be.name] = (np.stack(outs), fin, q)
    ro, rf, rq = res["reference"]
    oo, of, oq = res["oracle"]
    assert rf == of
    assert [x["cur_read_i"] for x in rq] == [x["cur_read_i"] for x in oq]
    assert rel_rms(oo, ro) <= TOL


@needs_ref
def test_source_churn_oracle_vs_reference():
    """Create / destroy churn between buffers and one-shots finishing on their own: the oracle follows the compiled reference
    (outputs, valid flags, finished ids) -- the reference's GainMixer processor bookkeeping and DestroySource ordering."""
    hr = load_sh_hrir()
    oo, ov, of = run_churn(OracleBackend(256, sh_hrir=hr), 256)
    ro, rv, rf = run_churn(RefBackend(256), 256)
    assert ov == rv and of == rf
    assert rel_rms(oo, ro) <= TOL
