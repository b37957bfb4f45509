"""Known-answer tests of the CUDA path through the C ABI, with the golden values of Resonance Audio's own unit tests
(the reference has no end-to-end golden; these are the per-node values its tests pin, reached here by driving one source
with a constant signal so that the output IS the gain under test):
  * RA/dsp/stereo_panner_test.cc:33-86            pan law at 9 directions,
  * RA/dsp/gain_processor_test.cc:49-179          linear gain ramp, 1/2048 per sample (0.1 -> 0.6: 0.1, 0.1004883, ...),
  * RA/graph/ambisonic_mixing_encoder_node_test.cc:115-143   the 16 encoding coefficients at azimuth 36 / elevation 18 degrees,
  * RA/dsp/distance_attenuation.cc:26-52           the logarithmic roll-off at Substrata's defaults (closed form),
plus the device callback's clamp (AudioEngine.cpp:215-216) is NOT applied by sbm_mix (it is the host mirror's job)."""
import numpy as np
import pytest

import substrata_b200 as sb
from harness import KIND_HRTF, KIND_PAN, KIND_STEREO, load_sh_hrir

pytestmark = pytest.mark.gpu
EPS = 1e-6          # kEpsilonFloat, the tolerance of the reference's tests (RA/base/constants_and_types.h:84)
F = 64


def position(az_deg, el_deg, dist=1.0):
    """World position of a direction in Resonance's convention (+x right, +y up, -z forward; azimuth counter-clockwise
    from the front: az = atan2(-x, -z), el = atan2(y, hypot(x, z)); RA/base/spherical_angle.cc:42-49)."""
    az, el = np.deg2rad(az_deg), np.deg2rad(el_deg)
    return np.array([-dist * np.cos(el) * np.sin(az), dist * np.sin(el), -dist * np.cos(el) * np.cos(az)], np.float32)


def one_source(kind, pos, volume=1.0, hrir=None, pcm_value=1.0):
    e = sb.MixEngine(F, sh_hrir=hrir, max_sources=4)
    buf = e.buffer_create(np.full(4 * F, pcm_value, np.float32))
    sid = e.source_create(kind)
    e.source_attach_buffer(sid, buf, 0, 48000, True, False)
    e.set_source_positions([sid], np.asarray(pos, np.float32).reshape(1, 3))
    e.set_source_volumes([sid], np.array([volume], np.float32))
    return e, sid


@pytest.mark.parametrize("az,el,check", [
    (-90, 0, lambda l, r: abs(l - 0.0) < EPS and abs(r - 1.0) < EPS),
    (-45, 0, lambda l, r: abs((r - l) - 1.0 / np.sqrt(2.0)) < EPS),
    (0, 0, lambda l, r: abs(l - r) < EPS),
    (45, 0, lambda l, r: abs((l - r) - 1.0 / np.sqrt(2.0)) < EPS),
    (90, 0, lambda l, r: abs(l - 1.0) < EPS and abs(r) < EPS),
    (0, 45, lambda l, r: abs(l - 0.5) < EPS and abs(r - 0.5) < EPS),
    (0, -60, lambda l, r: abs(l - 0.5) < EPS and abs(r - 0.5) < EPS),
    (0, 90, lambda l, r: abs(l - 0.5) < EPS and abs(r - 0.5) < EPS),
    (45, 45, lambda l, r: abs(l - 0.75) < 2 * EPS and abs(r - 0.25) < 2 * EPS),
])
def test_stereo_pan_law_through_the_mixer(az, el, check):
    # distance 0.9 <= near-field threshold: distance attenuation is exactly 1; first buffer: gains are set, not ramped
    e, _ = one_source(KIND_PAN, position(az, el, 0.9))
    valid, out = e.mix()
    assert valid == 1
    l, r = out[0::2], out[1::2]
    assert np.ptp(l) == 0 and np.ptp(r) == 0          # constant gain over the buffer
    assert check(float(l[0]), float(r[0])), (az, el, l[0], r[0])


def test_gain_ramp_sequence():
    """GainProcessor: 0.1 -> 0.6 ramps over |0.5| * 2048 = 1024 samples in steps of 0.5 / 1024 (gain_processor_test.cc: 0.1,
    0.1004883, 0.1009766, ...), continuing across buffers; a non-spatial source with a constant signal shows the gain itself."""
    e, sid = one_source(KIND_STEREO, [0, 0, 0], volume=0.1)
    _, first = e.mix()
    np.testing.assert_allclose(first, np.float32(0.1), atol=EPS)            # first call: no ramp
    e.set_source_volumes([sid], np.array([0.6], np.float32))
    got = np.concatenate([e.mix()[1][0::2] for _ in range(1024 // F + 1)])
    step = np.float32(0.5) / np.float32(1024.0)
    exp = np.float32(0.1) + step * np.arange(1024, dtype=np.float32)
    np.testing.assert_allclose(got[:3], [0.1, 0.1004883, 0.1009766], atol=EPS)
    np.testing.assert_allclose(got[:1024], exp, atol=2e-6)
    np.testing.assert_allclose(got[1024:], np.float32(0.6), atol=2e-6)


def test_encoder_node_vector():
    """ambisonic_mixing_encoder_node_test.cc:115-143: a source at azimuth 36 / elevation 18 degrees encodes a unit signal
    into these 16 ACN/SN3D coefficients (read back from the rank's ambisonic bus)."""
    hr = load_sh_hrir()
    pos = 0.9 * np.array([-0.55901699, 0.30901699, -0.76942088], np.float32)
    e, _ = one_source(KIND_HRTF, pos, hrir=hr)
    e.mix()
    bus = e.ambisonic_bus()
    exp = [1.0, 0.55901699, 0.30901699, 0.76942088, 0.74498856, 0.29920441, -0.35676274, 0.41181955, 0.24206145, 0.64679299,
           0.51477443, -0.17888019, -0.38975424, -0.24620746, 0.16726035, -0.21015578]
    assert np.ptp(bus, axis=1).max() == 0
    np.testing.assert_allclose(bus[:, 0], exp, atol=EPS)


@pytest.mark.parametrize("d", [0.5, 1.0, 2.0, 5.0, 37.5, 250.0, 499.0, 501.0])
def test_log_distance_attenuation_at_substrata_defaults(d):
    """ComputeLogarithmicDistanceAttenuation with min 0 (-> near-field threshold 1) and max 500 (distance_attenuation.cc:26-52)."""
    e, _ = one_source(KIND_PAN, position(0, 0, d))
    valid, out = e.mix()
    if d > 500.0:
        assert valid == 0 or not out.any()
        return
    if d <= 1.0:
        att = 1.0
    else:
        a_max = 1.0 / (1.0 + 499.0)
        att = (1.0 / (d - 1.0 + 1.0) - a_max) / (1.0 - a_max)
    if att < 1e-3:          # below -60 dB the GainNode drops the source (gain_node.cc:52-60)
        assert not out.any()
        return
    np.testing.assert_allclose(out[0::2], 0.5 * att, rtol=2e-6, atol=1e-7)      # front: both pan gains are 0.5
    np.testing.assert_allclose(out[1::2], 0.5 * att, rtol=2e-6, atol=1e-7)


def test_occlusion_mono_pole_impulse_response():
    """MonoPoleFilter (mono_pole_filter_test.cc:33-57): impulse response (1 - a) a^n; the coefficient comes from the smoothed
    occlusion, occ <- o + 0.75 (occ - o) per buffer and a = 1 - 1 / (occ + 1)^4 (occlusion_node.cc:53-99,
    occlusion_calculator.cc:43-50): intensity 4 on the first buffer gives occ = 1, a = 0.9375."""
    e = sb.MixEngine(F, max_sources=4)
    pcm = np.zeros(4 * F, np.float32)
    pcm[0] = 1.0
    buf = e.buffer_create(pcm)
    sid = e.source_create(KIND_PAN)
    e.source_attach_buffer(sid, buf, 0, 48000, True, False)
    e.set_source_positions([sid], position(0, 0, 0.9).reshape(1, 3))
    e.set_source_occlusions([sid], np.array([4.0], np.float32))
    _, out = e.mix()
    a = 0.9375
    exp = 0.5 * (1.0 - a) * a ** np.arange(F)          # front: both pan gains are 0.5
    np.testing.assert_allclose(out[0::2], exp, atol=EPS)
    np.testing.assert_allclose(out[1::2], exp, atol=EPS)
    _, out2 = e.mix()                                   # the filter state carries into the next buffer (and the coefficient moves on)
    assert abs(out2[0]) > 0 and np.all(np.diff(np.abs(out2[0::2])) < 0)


def test_mix_is_not_clamped_but_the_callback_is():
    e, _ = one_source(KIND_STEREO, [0, 0, 0], volume=1.0, pcm_value=3.0)
    _, out = e.mix()
    np.testing.assert_allclose(out, 3.0, atol=EPS)     # sbm_mix hands out the mix; AudioEngine::fillOutputBuffer clamps (host mirror)
