"""The C++ host mirror of glare::AudioEngine / glare::AudioSource / glare::AudioResampler (substrata_b200/host/) driven
the way Substrata's client drives the reference (addSource(AudioSourceRef), setHeadTransform per buffer, the mixing-loop
body, the device callback): its callback output must match the oracle on the same source set and listener path."""
import os
import struct
import subprocess
import tempfile

import numpy as np
import pytest

from harness import (KIND_HRTF, KIND_PAN, KIND_STEREO, ROOT, OracleBackend, load_sh_hrir, positions, quat_yaw, rel_rms, synth_pcm)
from substrata_b200.build import build_library

NATIVE = os.path.join(ROOT, "tests", "native")
PKG = os.path.join(ROOT, "substrata_b200")


def build_driver():
    build_library()
    exe = os.path.join(NATIVE, "test_audio_engine")
    src = os.path.join(NATIVE, "test_audio_engine.cc")
    deps = [src, os.path.join(PKG, "libsubstrata_b200_host.so"), os.path.join(PKG, "host", "AudioEngine.h")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + os.path.join(PKG, "host"), "-o", exe, src, "-L" + PKG,
                               "-l:libsubstrata_b200_host.so", "-l:libsubstrata_b200.so", "-Wl,-rpath," + PKG, "-lpthread"])
    return exe


def test_host_mirror_compiles_and_links():
    """No GPU needed: the mirror and its driver build against the C ABI."""
    assert os.path.exists(build_driver())


def scenario(n, F, nb, binaural, seed):
    rng = np.random.default_rng(seed)
    L = 4096
    pcm = np.stack([synth_pcm(rng, L) for _ in range(n)])
    pos = positions(rng, n)
    vol = rng.uniform(0.25, 1.0, n).astype(np.float32)
    rates = rng.choice([48000, 44100, 24000], n).astype(np.int32)
    start = rng.integers(0, L, n)
    nonspatial = (rng.uniform(size=n) < 0.1).astype(np.int32)
    streaming = ((rng.uniform(size=n) < 0.25) & (nonspatial == 0)).astype(np.int32)     # voice-like sources fed packet by packet
    heads = [(np.array([b * F / 48000.0, 0, 0], np.float32), quat_yaw(np.deg2rad(30.0) * b * F / 48000.0)) for b in range(nb)]
    return dict(n=n, F=F, nb=nb, L=L, pcm=pcm, pos=pos, vol=vol, rates=rates, start=start, nonspatial=nonspatial, streaming=streaming, heads=heads, binaural=binaural)


def write_scenario(path, sc, hr):
    with open(path, "wb") as f:
        f.write(struct.pack("<6i", sc["n"], sc["F"], sc["nb"], sc["L"], int(sc["binaural"]), hr.shape[1] if sc["binaural"] else 0))
        if sc["binaural"]:
            f.write(np.ascontiguousarray(hr, np.float32).tobytes())
        for s in range(sc["n"]):
            f.write(struct.pack("<4iQ4f", int(sc["nonspatial"][s]), int(sc["streaming"][s]), int(sc["rates"][s]), 1, int(sc["start"][s]), float(sc["vol"][s]), *[float(v) for v in sc["pos"][s]]))
            f.write(sc["pcm"][s].tobytes())
        for hp, hq in sc["heads"]:
            f.write(np.concatenate([hp, hq]).astype(np.float32).tobytes())


def oracle_run(sc, hr):
    be = OracleBackend(sc["F"], sh_hrir=hr if sc["binaural"] else None)
    ids = []
    for s in range(sc["n"]):
        kind = KIND_STEREO if sc["nonspatial"][s] else (KIND_HRTF if sc["binaural"] else KIND_PAN)
        sid = be.source_create(kind)
        if sc["streaming"][s]:
            be.source_attach_ring(sid, int(sc["rates"][s]))
        else:
            be.source_attach_buffer(sid, be.buffer_create(sc["pcm"][s]), int(sc["start"][s]), int(sc["rates"][s]), True, False)
        ids.append(sid)
    ids = np.array(ids, np.int32)
    be.set_source_positions(ids, sc["pos"]); be.set_source_volumes(ids, sc["vol"])
    outs = []
    F = sc["F"]
    for b, (hp, hq) in enumerate(sc["heads"]):
        be.set_head(hp, hq)
        if b % 5 != 3:      # the driver's packet schedule (tests/native/test_audio_engine.cc)
            for s in range(sc["n"]):
                if sc["streaming"][s]:
                    m = int(480 * ((F + 479) // 480) * float(sc["rates"][s]) / 48000.0)
                    if b * m + m <= sc["L"]:
                        be.source_ring_push(int(ids[s]), sc["pcm"][s][b * m:(b + 1) * m])
        outs.append(np.clip(be.mix()[1], -1.0, 1.0))          # the device callback clamps (AudioEngine.cpp:215-216)
    return np.stack(outs)


@pytest.mark.gpu
@pytest.mark.parametrize("binaural", [True, False])
def test_audio_engine_mirror_matches_oracle(binaural):
    exe = build_driver()
    hr = load_sh_hrir()
    sc = scenario(48, 256, 10, binaural, seed=5 + int(binaural))
    with tempfile.TemporaryDirectory() as d:
        sp, op = os.path.join(d, "sc.bin"), os.path.join(d, "out.bin")
        write_scenario(sp, sc, hr)
        subprocess.check_call([exe, sp, op])
        got = np.fromfile(op, np.float32).reshape(sc["nb"], 2 * sc["F"])
    assert rel_rms(got, oracle_run(sc, hr)) <= 1e-5


@pytest.mark.gpu
def test_audio_engine_mirror_threaded_mix_loop():
    """The ResonanceThread-style loop (4 buffers of look-ahead) keeps the callback fed: output is non-silent and clamped."""
    exe = build_driver()
    hr = load_sh_hrir()
    sc = scenario(16, 256, 12, True, seed=9)
    with tempfile.TemporaryDirectory() as d:
        sp, op = os.path.join(d, "sc.bin"), os.path.join(d, "out.bin")
        write_scenario(sp, sc, hr)
        subprocess.check_call([exe, sp, op, "threaded"])
        got = np.fromfile(op, np.float32).reshape(sc["nb"], 2 * sc["F"])
    assert np.abs(got[2:]).max() > 1e-3 and np.abs(got).max() <= 1.0
