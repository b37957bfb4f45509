"""Oracle parity at BASELINE.json's OWN sizes (harness.FULL_SCENES), through the C ABI on the B200:

  C2  4 096 resampled + panned sources, F = 512                        vs the f32-bus oracle
  C3  65 536 ambisonic sources, F = 1024, 128-tap SH-HRIRs             vs the f32-bus oracle (partitioned)
  C4  one rank slice, 131 072 doppler + panned sources, F = 512        vs the single f32-bus oracle summing in id order
      AND vs the fp64-bus oracle (SURVEY.md finding 9: both numbers)
  C4  all 1 048 576 doppler + panned sources on one GPU                vs the fp64-bus oracle (partitioned over the host cores)
  C5  1 024 voice rings + 32 768 ambient loops in ONE engine, F = 256  vs the f32-bus oracle (partitioned)

Every scene moves the listener every buffer (gain ramps on all sources: the device's closed-form ramp, sbm_core.h plan_gain,
is therefore exercised at 10^6 sources), switches every 8th source off and on again (-60 dB rules, mixer-processor sweep)
and, for C5, drops voice packets (ring underflow).  Bar: <= 1e-5 relative RMS over all buffers (BASELINE.json)."""
import numpy as np
import pytest

import substrata_b200 as sb
from harness import FULL_SCENES, BigScene, full_scene_parity, rel_rms

pytestmark = pytest.mark.gpu
TOL = 1e-5


def make_engine(sc):
    return sb.MixEngine(sc.F, sh_hrir=sc.hrir, max_sources=sc.n)


@pytest.mark.parametrize("name", ["c2", "c3", "c5", "c4_1m"])
def test_full_size_scene_matches_oracle(name):
    r = full_scene_parity(name, make_engine)
    print(r)
    assert r["valid_equal"]
    assert r["rel_rms"] <= TOL, r


def test_c4_rank_slice_vs_f32_and_fp64_oracles():
    sc = BigScene(**FULL_SCENES["c4_rank"])
    go, gv = sc.run(make_engine(sc))
    o32, v32 = sc.run(sc.oracle())                              # one engine, f32 bus, id order: the reference's arithmetic
    o64, v64 = sc.run_oracle_partitioned(8, double_bus=True)
    e32, e64 = rel_rms(go, o32), rel_rms(go, o64)
    print({"scene": "c4_rank", "rel_rms_vs_f32_bus": e32, "rel_rms_vs_fp64_bus": e64, "f32_vs_fp64_oracle": rel_rms(o32, o64)})
    assert gv == v32 == v64
    assert e32 <= TOL and e64 <= TOL


def test_pipelined_full_size_is_bit_identical():
    """sbm_mix_submit / sbm_mix_collect at C4's rank-slice size reproduce sbm_mix bit for bit (two buffers in flight)."""
    sc = BigScene(**FULL_SCENES["c4_rank"])
    a, _ = sc.run(make_engine(sc))
    e = make_engine(sc)
    sel = np.arange(sc.n)
    ids = sc.build(e, sel)
    outs = []
    for b in range(sc.nb):
        sc.step(e, ids, sel, b)
        e.mix_submit()
        if b >= 1:
            outs.append(e.mix_collect()[1].copy())
    outs.append(e.mix_collect()[1].copy())
    assert np.array_equal(a, np.stack(outs))


def test_overlapped_params_and_mix_kernels_match_the_serial_order(monkeypatch):
    """SBM_OVERLAP=1: large stereo-bus scenes run K1 as a persistent grid BESIDE the mix kernel, which polls K1's per-tile stamps
    before it fetches records (k1_params.cu / k2_mix.cu wait_records_ready / engine.cc mix_begin_locked).  Opt-in -- it measured
    slower than the serial order -- but kept correct: same records, another grid shape (partial sums in another
    order), so equal to rounding, deterministic, and within the parity bar against the oracle."""
    sc = BigScene(**FULL_SCENES["c4_rank"])
    b, vb = sc.run(make_engine(sc))
    monkeypatch.setenv("SBM_OVERLAP", "1")
    a, va = sc.run(make_engine(sc))
    a2, _ = sc.run(make_engine(sc))
    monkeypatch.delenv("SBM_OVERLAP")
    assert va == vb
    assert np.array_equal(a, a2)
    assert not np.array_equal(a, b)          # (the overlapped shape really ran: 3 CTAs per SM, not 4)
    assert rel_rms(a, b) <= 2e-6
    o64, v64 = sc.run_oracle_partitioned(8, double_bus=True)
    assert va == v64 and rel_rms(a, o64) <= TOL
