"""Regenerates the committed golden fixtures from the COMPILED REFERENCE (oracle/_ref/libsubstrata_ref.so).

Run here (where /root/reference exists):   make -C oracle ref && python tests/golden/make_golden.py

Outputs
  substrata_b200/data/sh_hrir_order1_48k.f32  order-1 SADIE SH-HRIRs [4][256] (the reflections' decoder, room effects), same source
  substrata_b200/data/sh_hrir_order3_48k.f32  order-3 SADIE SH-HRIRs [16][256] f32 exactly as the reference decoder
                                              receives them (CreateShHrirsFromAssets, RA/dsp/sh_hrir_creator.cc:60-70).
                                              This is input DATA of the path (the host passes the taps through the C ABI).
  tests/golden/encoder_cells.npz              GetEncodingCoeffs at the centre of one-degree cells (full stored quadrant + every 5th cell of the sphere).
  tests/golden/scenarios.npz                  interleaved stereo output of the reference for the scenarios in
                                              GOLDEN_SCENARIOS (tests/harness.py Scenario), 8-14 buffers each.
  tests/golden/resampler.npz                  glare::AudioResampler outputs for the 13 rate pairs of AudioResampler::test().
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from harness import ROOT, KIND_HRTF, KIND_PAN, RefBackend, Scenario, ref_lib, run_churn  # noqa: E402

SPECIAL_POSITIONS = [(0, 0, 0), (0, 5, 0), (0, -5, 0), (3, 0, 0), (-3, 0, 0), (0, 0, 3), (0, 0, -3), (600, 0, 0), (0.5, 0, 0),
                     (2, 2, -2), (-1, 0, -1), (0, 1e-3, -1), (499.5, 0, 0), (1, 0, 0), (0, 0, -1)]

GOLDEN_SCENARIOS = {
    # name: (Scenario kwargs, frames, n_buffers, hrir_taps)
    "c1_pan16": (dict(n_sources=16, kind=KIND_PAN, loop_len=48000), 512, 8, 0),
    "hq16": (dict(n_sources=16, kind=KIND_HRTF, loop_len=48000), 512, 8, 0),
    "c2_resample_pan": (dict(n_sources=96, kind=KIND_PAN, feeder="resample"), 512, 8, 0),
    "c4_doppler_pan": (dict(n_sources=96, kind=KIND_PAN, feeder="doppler"), 512, 8, 0),
    "c3_hq_128tap": (dict(n_sources=64, kind=KIND_HRTF), 1024, 8, 128),
    "c5_ring_hq": (dict(n_sources=48, kind=KIND_HRTF, feeder="ring", rates=(48000, 24000), rate_probs=(0.5, 0.5)), 256, 10, 0),
    "mixed_toggles": (dict(n_sources=64, mixed_kinds=True, volume_toggle=True, occlusion_toggle=True), 512, 14, 0),
    "f480_hq": (dict(n_sources=32, kind=KIND_HRTF), 480, 8, 0),
    "short_loops": (dict(n_sources=40, kind=KIND_PAN, loop_len=300, feeder="resample"), 256, 8, 0),
    # loop lengths that are not a multiple of 4 samples with a pitch / doppler read head (re-aligned second TMA segment)
    "doppler_odd_loop_pan": (dict(n_sources=48, kind=KIND_PAN, feeder="doppler", loop_len=4099, seed=101), 512, 8, 0),
    "doppler_odd_loop_hq": (dict(n_sources=48, kind=KIND_HRTF, feeder="doppler", loop_len=4098, seed=102, volume_toggle=True), 256, 10, 0),
    "hq_resample_1024": (dict(n_sources=40, kind=KIND_HRTF, feeder="resample", loop_len=3001, seed=103, occlusion_toggle=True), 1024, 8, 0),
    # the bike: three pitch-shifted mix sources per AudioSource, retuned every buffer (AudioEngine.cpp:435-458, BikePhysics.cpp:712-740)
    "bike_mix3_pan": (dict(n_sources=24, kind=KIND_PAN, feeder="mix3", loop_len=4099, seed=111, n_shared_buffers=9), 512, 8, 0),
    # master volume moves, including to silence and back (every source crosses the -60 dB rules at once; `valid` goes false
    # once the decoder tail is flushed)
    "master_moves": (dict(n_sources=48, mixed_kinds=True, feeder="resample", seed=113, master_script={1: 0.5, 3: 0.0, 15: 0.8}), 256, 20, 0),
    # sources on the axes, at the poles, at the listener's own position, on one-degree cell edges, at the near-field threshold and
    # out of range; the listener at rest for three buffers (exact cell edges), then moving
    "special_positions_hq": (dict(n_sources=30, kind=KIND_HRTF, seed=114, positions_override=SPECIAL_POSITIONS, moving=False), 256, 6, 0),
    "special_positions_mixed": (dict(n_sources=45, mixed_kinds=True, seed=115, positions_override=SPECIAL_POSITIONS, feeder="resample"), 512, 8, 0),
    "bike_mix3_hq": (dict(n_sources=24, kind=KIND_HRTF, feeder="mix3", loop_len=4096, seed=112, n_shared_buffers=9, volume_toggle=True), 256, 10, 0),
    # room effects as Substrata drives them inside a parcel (AudioEngine.cpp:738-768, GUIClient.cpp:6886-6896): reflections sends,
    # six wall reflections cross-faded while the listener moves and when the room changes, head rotation, order-1 decode; enabled
    # mid-stream, the room replaced, disabled again (the delay line and the decoder play out)
    "room_hq": (dict(n_sources=40, kind=KIND_HRTF, loop_len=3000, seed=121, volume_toggle=True,
                     room=dict(on=0, off=20, moves={9: ((0.0, 0.0, 0.0), (20.0, 6.0, 20.0))})), 256, 34, 0),
    "room_pan_doppler": (dict(n_sources=48, kind=KIND_PAN, feeder="doppler", loop_len=4099, seed=122, room=dict(on=2)), 512, 10, 0),
    # ... with occluded (low-passed) sources, a low-pass in the reflections (cut-off frequency) and a rotated room
    "room_mixed_lpf": (dict(n_sources=60, mixed_kinds=True, occlusion_toggle=True, volume_toggle=True, feeder="resample", seed=123,
                            room=dict(on=1, off=14, cutoff=800.0, rot=(0.0, 0.38268343, 0.0, 0.92387953), pos=(0.5, 0.0, -1.0), dims=(16.0, 5.0, 12.0))), 256, 20, 0),
    "room_hq_1024": (dict(n_sources=48, kind=KIND_HRTF, feeder="doppler", loop_len=4096, seed=124, room=dict(on=1)), 1024, 6, 0),
}

RESAMPLER_PAIRS = [(8000, 48000), (12000, 48000), (16000, 48000), (24000, 48000), (44100, 48000), (48000, 48000),
                   (6000, 48000), (8000, 44100), (12000, 44100), (16000, 44100), (24000, 44100), (44100, 44100), (48000, 44100)]


def main():
    L = ref_lib()
    if L is None:
        raise SystemExit("build oracle/_ref first: make -C oracle ref")
    # 1. SH-HRIR taps
    L.ref_dump_sh_hrirs.restype = C.c_int
    L.ref_dump_sh_hrirs.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    ch = C.c_int(0)
    T = L.ref_dump_sh_hrirs(48000, None, 0, C.byref(ch))
    hr = np.zeros((ch.value, T), np.float32)
    L.ref_dump_sh_hrirs(48000, hr.ctypes.data, hr.size, C.byref(ch))
    os.makedirs(os.path.join(ROOT, "substrata_b200", "data"), exist_ok=True)
    hr.tofile(os.path.join(ROOT, "substrata_b200", "data", "sh_hrir_order3_48k.f32"))
    # ... and the order-1 set the reflections are decoded with (room effects)
    L.ref_dump_sh_hrirs_order1.restype = C.c_int
    L.ref_dump_sh_hrirs_order1.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    T1 = L.ref_dump_sh_hrirs_order1(48000, None, 0, C.byref(ch))
    hr1 = np.zeros((ch.value, T1), np.float32)
    L.ref_dump_sh_hrirs_order1(48000, hr1.ctypes.data, hr1.size, C.byref(ch))
    hr1.tofile(os.path.join(ROOT, "substrata_b200", "data", "sh_hrir_order1_48k.f32"))
    # 2. encoder cells
    L.ref_encoding_coeffs.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p]
    # every cell of the stored quadrant (0..90 x 0..90) plus every 5th cell of the full sphere (symmetry flips)
    pairs = [(a, e) for a in range(0, 91) for e in range(0, 91)]
    pairs += [(a, e) for a in range(-180, 180, 5) for e in range(-90, 91, 5)]
    pairs = np.array(pairs, np.int32)
    cells = np.zeros((len(pairs), 16), np.float32)
    tmp = np.zeros(16, np.float32)
    for i, (a, e) in enumerate(pairs):
        # centre of the one-degree cell (truncation toward zero: cell e covers [e, e+1) for e >= 0, (e-1, e] for e < 0)
        ee = min(max(e + (0.5 if e >= 0 else -0.5), -89.9), 89.9)
        aa = a + (0.5 if a >= 0 else -0.5)
        L.ref_encoding_coeffs(np.float32(np.deg2rad(aa)), np.float32(np.deg2rad(ee)), 0.0, tmp.ctypes.data)
        cells[i] = tmp
    np.savez_compressed(os.path.join(HERE, "encoder_cells.npz"), pairs=pairs, cells=cells)
    # 3. scenarios
    out = {}
    for name, (kw, F, nb, taps) in GOLDEN_SCENARIOS.items():
        sc = Scenario(frames=F, **kw)
        o, v = sc.run(RefBackend(F, hrir_taps=taps), nb)
        out[name] = o
        out[name + "__valid"] = np.array(v, np.int32)
    # create / destroy churn with one-shots finishing on their own (tests/harness.py run_churn): outputs, valid flags, finished ids
    co, cv, cf = run_churn(RefBackend(256), 256)
    out["churn"] = co
    out["churn__valid"] = np.array(cv, np.int32)
    out["churn__fin_flat"] = np.array([i for f in cf for i in f], np.int32)
    out["churn__fin_off"] = np.cumsum([0] + [len(f) for f in cf]).astype(np.int32)
    # The reference sums its sources in unordered_map order (RA/node/node.h:163), i.e. in an order that depends on heap addresses:
    # a regenerated scenario differs from the committed one by that f32 order noise (1e-7 .. 3e-7 relative).  Entries that already
    # exist are therefore kept as committed unless --force is given; new scenarios are added.
    path = os.path.join(HERE, "scenarios.npz")
    if os.path.exists(path) and "--force" not in sys.argv:
        old = dict(np.load(path))
        for k, v in old.items():
            if k in out and out[k].shape == v.shape:
                out[k] = v
    np.savez_compressed(path, **out)
    # 4. resampler (the reference's own test pattern: ramp src[i] = i, 4-sample chunks; AudioResampler.cpp:213-287)
    L.ref_resampler_create.restype = C.c_void_p
    L.ref_resampler_create.argtypes = [C.c_int, C.c_int]
    L.ref_resampler_num_src_needed.restype = C.c_size_t
    L.ref_resampler_num_src_needed.argtypes = [C.c_void_p, C.c_size_t]
    L.ref_resampler_resample.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    L.ref_resampler_destroy.argtypes = [C.c_void_p]
    rs = {}
    for (s, d) in RESAMPLER_PAIRS:
        r = L.ref_resampler_create(s, d)
        src_i, outs, needs = 0, [], []
        for _ in range(64):
            n = L.ref_resampler_num_src_needed(r, 4)
            src = np.arange(src_i, src_i + n, dtype=np.float32)
            src_i += n
            dst = np.zeros(4, np.float32)
            L.ref_resampler_resample(r, dst.ctypes.data, 4, src.ctypes.data, n)
            outs.append(dst)
            needs.append(n)
        L.ref_resampler_destroy(r)
        rs["%d_%d" % (s, d)] = np.concatenate(outs)
        rs["%d_%d__needs" % (s, d)] = np.array(needs, np.int64)
    np.savez_compressed(os.path.join(HERE, "resampler.npz"), **rs)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
