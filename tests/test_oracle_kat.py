"""Known-answer tests for oracle/mix_oracle.c, restated from the reference's own unit tests (gtest is not
vendored, so the values are re-checked here): resonance-audio/resonance_audio/dsp/stereo_panner_test.cc:33-86,
distance_attenuation_test.cc:29-128, gain_processor_test.cc:49-179, mono_pole_filter_test.cc:33-57,
occlusion_calculator_test.cc:107-137, ambisonics/ambisonic_lookup_table_test.cc:87-148 (spread 0),
graph/ambisonic_mixing_encoder_node_test.cc:115-143, ambisonics/ambisonic_binaural_decoder_test.cc:53-160 and
audio/AudioResampler.cpp:213-287.  Tolerance is the reference's kEpsilonFloat = 1e-6 unless stated."""
import ctypes as C
import os

import numpy as np
import pytest

from harness import GOLDEN_DIR, oracle_lib

EPS = 1e-6
_f, _i, _vp, _sz = C.c_float, C.c_int, C.c_void_p, C.c_size_t


@pytest.fixture(scope="module")
def L():
    lib = oracle_lib()
    lib.orc_log_distance_attenuation.restype = _f
    lib.orc_log_distance_attenuation.argtypes = [_vp, _vp, _f, _f]
    lib.orc_linear_distance_attenuation.restype = _f
    lib.orc_linear_distance_attenuation.argtypes = [_vp, _vp, _f, _f]
    lib.orc_update_attenuations.argtypes = [_f, _f, _f, _vp, _vp, _f, _f, _i, _f, _f, _f, _vp]
    lib.orc_stereo_pan_gains.argtypes = [_f, _f, _vp]
    lib.orc_encoding_coeffs.argtypes = [_f, _f, _vp]
    lib.orc_relative_direction.argtypes = [_vp, _vp, _vp, _vp]
    lib.orc_spherical_from_world.argtypes = [_vp, _vp]
    lib.orc_occlusion_filter_coefficient.restype = _f
    lib.orc_occlusion_filter_coefficient.argtypes = [_f, _f]
    lib.orc_directivity.restype = _f
    lib.orc_directivity.argtypes = [_f, _f, _f, _f]
    lib.orc_gain_apply.argtypes = [_vp, _f, _vp, _vp, _sz, _i]
    lib.orc_mono_pole.restype = _i
    lib.orc_mono_pole.argtypes = [_f, _vp, _vp, _vp, _sz]
    lib.orc_sh_decode.argtypes = [_vp, _i, _i, _vp, _vp, _i, _vp, _vp]
    lib.orc_resampler_init.argtypes = [_vp, _i, _i]
    lib.orc_resampler_num_src_needed.restype = _sz
    lib.orc_resampler_num_src_needed.argtypes = [_vp, _sz]
    lib.orc_resampler_resample.argtypes = [_vp, _vp, _sz, _vp, _sz]
    return lib


def v3(*a):
    return np.array(a, dtype=np.float32)


def pan(L, az_deg, el_deg):
    out = np.zeros(2, np.float32)
    rad = np.float32(np.pi / 180.0)
    L.orc_stereo_pan_gains(np.float32(az_deg) * rad, np.float32(el_deg) * rad, out.ctypes.data)
    return out


def test_stereo_pan_law(L):
    inv_sqrt2 = 1.0 / np.sqrt(2.0)
    g = pan(L, -90, 0); assert abs(g[0] - 0.0) < EPS and abs(g[1] - 1.0) < EPS
    g = pan(L, -45, 0); assert abs((g[1] - g[0]) - inv_sqrt2) < EPS
    g = pan(L, 0, 0); assert abs(g[0] - g[1]) < EPS
    g = pan(L, 45, 0); assert abs((g[0] - g[1]) - inv_sqrt2) < EPS
    g = pan(L, 90, 0); assert abs(g[0] - 1.0) < EPS and abs(g[1]) < EPS
    for el in (45, -60, 90, -90):
        g = pan(L, 0, el); assert abs(g[0] - 0.5) < EPS and abs(g[1] - 0.5) < EPS
    g = pan(L, 45, 45); assert abs(g[0] - 0.75) < EPS and abs(g[1] - 0.25) < EPS


def test_distance_attenuation(L):
    lis, src = v3(0, 3, 0), v3(0, 0, -4)
    assert abs(L.orc_log_distance_attenuation(lis.ctypes.data, src.ctypes.data, 1.0, 500.0) - 0.1983967) < EPS
    assert L.orc_log_distance_attenuation(lis.ctypes.data, src.ctypes.data, 1.0, 3.0) == 0.0
    assert L.orc_log_distance_attenuation(lis.ctypes.data, src.ctypes.data, 10.0, 500.0) == 1.0
    assert abs(L.orc_linear_distance_attenuation(lis.ctypes.data, src.ctypes.data, 1.0, 8.0) - 0.4285714) < EPS
    assert L.orc_linear_distance_attenuation(lis.ctypes.data, src.ctypes.data, 1.0, 3.0) == 0.0
    assert L.orc_linear_distance_attenuation(lis.ctypes.data, src.ctypes.data, 10.0, 8.0) == 1.0


def test_update_attenuation_parameters(L):
    out = np.zeros(4, np.float32)
    lis, src = v3(0, 0, 0), v3(0, 0, 1.5)
    L.orc_update_attenuations(0.5, 0.5, 2.0, lis.ctypes.data, src.ctypes.data, 1.0, 0.25, 2, 0.0, 500.0, 0.2, out.ctypes.data)
    np.testing.assert_allclose(out, [0.5, 0.1, 0.0125, 0.25], atol=EPS)


@pytest.mark.parametrize("accumulate", [0, 1])
def test_gain_processor_ramp(L, accumulate):
    expected = np.array([0.1, 0.1004883, 0.1009766, 0.1014648, 0.1019531, 0.1024414, 0.1029297, 0.1034180, 0.1039063, 0.1043945])
    x = np.ones(10, np.float32)
    y = np.full(10, 2.0, np.float32)
    st = np.array([0.1, 1.0], np.float32)
    L.orc_gain_apply(st.ctypes.data, 0.6, x.ctypes.data, y.ctypes.data, 10, accumulate)
    np.testing.assert_allclose(y, expected + (2.0 if accumulate else 0.0), atol=EPS)
    # longer than the ramp: ramp halts at ramp_length, then constant
    n = 2048
    x = np.ones(n, np.float32)
    y = np.full(n, 2.0, np.float32)
    st = np.array([0.1, 1.0], np.float32)
    L.orc_gain_apply(st.ctypes.data, 0.6, x.ctypes.data, y.ctypes.data, n, accumulate)
    ramp = int(np.float32(abs(np.float32(0.6) - np.float32(0.1))) * np.float32(2048))
    inc = np.float32((np.float32(0.6) - np.float32(0.1)) / np.float32(ramp))
    g = np.float32(0.1)
    exp = np.zeros(n, np.float32)
    for i in range(n):
        exp[i] = g
        if i < ramp:
            g = np.float32(g + inc)
    np.testing.assert_allclose(y, exp + (2.0 if accumulate else 0.0), atol=EPS)


@pytest.mark.parametrize("accumulate", [0, 1])
def test_gain_processor_first_call_does_not_ramp(L, accumulate):
    x = np.full(64, 0.1, np.float32)
    y = np.zeros(64, np.float32)
    st = np.zeros(2, np.float32)  # default-constructed: uninitialised
    L.orc_gain_apply(st.ctypes.data, 0.6, x.ctypes.data, y.ctypes.data, 64, accumulate)
    np.testing.assert_allclose(y, 0.1 * 0.6, atol=EPS)
    assert st[0] == np.float32(0.6) and st[1] == 1.0


def test_mono_pole_filter(L):
    x = np.zeros(10, np.float32); x[0] = 1.0
    y = np.zeros(10, np.float32)
    prev = np.zeros(1, np.float32)
    assert L.orc_mono_pole(0.5, prev.ctypes.data, x.ctypes.data, y.ctypes.data, 10) == 1
    np.testing.assert_allclose(y, 0.5 * 0.5 ** np.arange(10), atol=EPS)
    assert L.orc_mono_pole(0.0, prev.ctypes.data, x.ctypes.data, y.ctypes.data, 10) == 0


def test_occlusion_filter_coefficient(L):
    for d in (0.0, 0.3, 0.5, 0.7, 0.9):
        assert L.orc_occlusion_filter_coefficient(d, 0.0) == np.float32(1.0) - np.float32(d)
    assert L.orc_occlusion_filter_coefficient(1.5, 0.0) == 0.0
    for o in (0.01, 0.1, 1.0, 10.0, 100.0):
        b = np.float32(o) + np.float32(1.0)
        b2 = np.float32(b * b)
        exp = np.float32(1.0) - np.float32(1.0) / np.float32(b2 * b2)
        assert L.orc_occlusion_filter_coefficient(1.0, o) == exp
    # directivity: omnidirectional for alpha 0 (what Substrata leaves at default)
    assert L.orc_directivity(0.0, 1.0, 0.3, 0.2) == 1.0
    assert abs(L.orc_directivity(1.0, 2.0, 0.0, np.pi / 4) - 0.5) < EPS


def closed_form_sh(a, e):
    s3 = np.sqrt(3.0)
    return np.array([
        1.0, np.cos(e) * np.sin(a), np.sin(e), np.cos(e) * np.cos(a),
        s3 / 2 * np.cos(e) ** 2 * np.sin(2 * a), s3 / 2 * np.sin(2 * e) * np.sin(a), 0.5 * (3 * np.sin(e) ** 2 - 1),
        s3 / 2 * np.sin(2 * e) * np.cos(a), s3 / 2 * np.cos(e) ** 2 * np.cos(2 * a),
        np.sqrt(5 / 8) * np.cos(e) ** 3 * np.sin(3 * a), np.sqrt(15) / 2 * np.sin(e) * np.cos(e) ** 2 * np.sin(2 * a),
        np.sqrt(3 / 8) * np.cos(e) * (5 * np.sin(e) ** 2 - 1) * np.sin(a), 0.5 * np.sin(e) * (5 * np.sin(e) ** 2 - 3),
        np.sqrt(3 / 8) * np.cos(e) * (5 * np.sin(e) ** 2 - 1) * np.cos(a), np.sqrt(15) / 2 * np.sin(e) * np.cos(e) ** 2 * np.cos(2 * a),
        np.sqrt(5 / 8) * np.cos(e) ** 3 * np.cos(3 * a)])


def test_encoding_coefficients_closed_form(L):
    dirs = [(0, 0), (90, 0), (180, 0), (-90, 0), (0, 90), (0, -90), (10, 20), (-30, 40), (50, -60), (290, -80), (110, 5),
            (-120, 15), (130, -25), (220, -35)]
    out = np.zeros(16, np.float32)
    rad = np.float32(np.pi / 180.0)
    for az, el in dirs:
        a, e = np.float32(az) * rad, np.float32(el) * rad
        L.orc_encoding_coeffs(a, e, out.ctypes.data)
        np.testing.assert_allclose(out, closed_form_sh(float(a), float(e)), atol=2e-6, err_msg=str((az, el)))


def test_encoder_node_vector(L):
    # ambisonic_mixing_encoder_node_test.cc:115-143: position at azimuth 36 / elevation 18 degrees.
    pos = v3(-0.55901699, 0.30901699, -0.76942088)
    rel, ae, out = np.zeros(3, np.float32), np.zeros(2, np.float32), np.zeros(16, np.float32)
    L.orc_relative_direction(v3(0, 0, 0).ctypes.data, np.array([0, 0, 0, 1], np.float32).ctypes.data, pos.ctypes.data, rel.ctypes.data)
    L.orc_spherical_from_world(rel.ctypes.data, ae.ctypes.data)
    L.orc_encoding_coeffs(ae[0], ae[1], out.ctypes.data)
    exp = [1.0, 0.55901699, 0.30901699, 0.76942088, 0.74498856, 0.29920441, -0.35676274, 0.41181955, 0.24206145, 0.64679299,
           0.51477443, -0.17888019, -0.38975424, -0.24620746, 0.16726035, -0.21015578]
    np.testing.assert_allclose(out, exp, atol=EPS)


def test_encoder_cells_match_reference_bit_exact(L):
    g = np.load(os.path.join(GOLDEN_DIR, "encoder_cells.npz"))
    out = np.zeros(16, np.float32)
    bad = 0
    for (a, e), ref in zip(g["pairs"], g["cells"]):
        ee = min(max(e + (0.5 if e >= 0 else -0.5), -89.9), 89.9)
        aa = a + (0.5 if a >= 0 else -0.5)
        L.orc_encoding_coeffs(np.float32(np.deg2rad(aa)), np.float32(np.deg2rad(ee)), out.ctypes.data)
        bad += int(not np.array_equal(out, ref))
    assert bad == 0


def _ramp_data(channels, n):
    return (np.arange(channels * n, dtype=np.float64) * 0.001).astype(np.float32).reshape(channels, n)


@pytest.mark.parametrize("taps,exp_l,exp_r", [
    (17, [0.0042840000, 0.0087780003, 0.013486000, 0.018412000, 0.023560001, 0.028934000, 0.034538001, 0.040376000, 0.046452001,
          0.052770000, 0.059333999, 0.066147998, 0.073215999, 0.080541998, 0.088129997, 0.095983997, 0.10410800, 0.10638600],
     [0.0036720000, 0.0074840002, 0.011438000, 0.015536000, 0.019780001, 0.024172001, 0.028713999, 0.033408001, 0.038256001,
      0.043260001, 0.048422001, 0.053743999, 0.059227999, 0.064875998, 0.070689999, 0.076672003, 0.082823999, 0.084252000]),
    (18, [0.0045360001, 0.0092879999, 0.014260001, 0.019455999, 0.024879999, 0.030536000, 0.036428001, 0.042560000, 0.048935998,
          0.055560000, 0.062436000, 0.069568001, 0.076959997, 0.084615998, 0.092540003, 0.10073600, 0.10920800, 0.11796000],
     [0.0038880000, 0.0079199998, 0.012098000, 0.016424000, 0.020900000, 0.025528001, 0.030309999, 0.035248000, 0.040344000,
      0.045600001, 0.051018000, 0.056600001, 0.062348001, 0.068264000, 0.074349999, 0.080608003, 0.087040000, 0.093648002]),
    (19, [0.0047880001, 0.0097980006, 0.015034000, 0.020500001, 0.026200000, 0.032138001, 0.038318001, 0.044744000, 0.051419999,
          0.058350001, 0.065537997, 0.072987996, 0.080704004, 0.088689998, 0.096950002, 0.10548800, 0.11430800, 0.12341400],
     [0.0041040001, 0.0083560003, 0.012758000, 0.017312000, 0.022020001, 0.026883999, 0.031906001, 0.037087999, 0.042431999,
      0.047940001, 0.053613998, 0.059455998, 0.065467998, 0.071652003, 0.078010000, 0.084543996, 0.091256000, 0.098148003]),
])
def test_binaural_decoder_golden(L, taps, exp_l, exp_r):
    F = 18
    # GenerateAudioData increments a float accumulator by 0.001f (ambisonic_binaural_decoder_test.cc:31-45)
    def gen(ch, n):
        out = np.zeros((ch, n), np.float32)
        v = np.float32(0.0)
        for c in range(ch):
            for i in range(n):
                out[c, i] = v
                v = np.float32(v + np.float32(0.001))
        return out
    x, h = gen(4, F), gen(4, taps)
    hist = np.zeros((4, taps - 1), np.float32)
    l, r = np.zeros(F, np.float32), np.zeros(F, np.float32)
    L.orc_sh_decode(h.ctypes.data, 4, taps, hist.ctypes.data, x.ctypes.data, F, l.ctypes.data, r.ctypes.data)
    np.testing.assert_allclose(l, exp_l, atol=EPS)
    np.testing.assert_allclose(r, exp_r, atol=EPS)


class Resampler(C.Structure):
    _fields_ = [("src_rate", _i), ("dst_rate", _i), ("prev_dest", C.c_int64), ("prev0", C.c_int64), ("prev", _f * 2)]


PAIRS = [(8000, 48000), (12000, 48000), (16000, 48000), (24000, 48000), (44100, 48000), (48000, 48000), (6000, 48000),
         (8000, 44100), (12000, 44100), (16000, 44100), (24000, 44100), (44100, 44100), (48000, 44100)]


@pytest.mark.parametrize("src,dst", PAIRS)
def test_resampler_ramp_and_golden(L, src, dst):
    """AudioResampler::test() (audio/AudioResampler.cpp:213-287): a ramp src[i] = i resampled in 4-sample chunks must
    come out as out[i] ~= i * src/dst; plus bit-exact agreement with the compiled reference class (golden)."""
    g = np.load(os.path.join(GOLDEN_DIR, "resampler.npz"))
    r = Resampler()
    L.orc_resampler_init(C.byref(r), src, dst)
    src_i, outs, needs = 0, [], []
    for _ in range(64):
        n = L.orc_resampler_num_src_needed(C.byref(r), 4)
        s = np.arange(src_i, src_i + n, dtype=np.float32)
        src_i += n
        d = np.zeros(4, np.float32)
        L.orc_resampler_resample(C.byref(r), d.ctypes.data, 4, s.ctypes.data, n)
        outs.append(d)
        needs.append(n)
    out = np.concatenate(outs)
    np.testing.assert_allclose(out, np.arange(out.size) * src / dst, atol=1e-4)
    assert np.array_equal(out, g["%d_%d" % (src, dst)])
    assert np.array_equal(np.array(needs), g["%d_%d__needs" % (src, dst)])
