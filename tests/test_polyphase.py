"""vraudio::Resampler drop-in (include/sbm_audio.h sbm_polyphase_*, csrc/polyphase.cc + k6_polyphase.cu).

The bar is bit-exactness (the taps are summed in the reference's order with the reference's coefficient bits):
  * the oracle's restatement (oracle/mix_oracle.c orc_polyphase_*) against the committed outputs of the COMPILED REFERENCE
    (tests/golden/polyphase.npz, made by tests/golden/make_polyphase_golden.py) and, where oracle/_ref is built, live;
  * the known-answer tests of RA/dsp/resampler_test.cc restated (lengths, zero crossings, Dirac delay, supported rates);
  * the product's host logic (rate bookkeeping, filter generation, carried state) through the host emulation of the device seam;
  * -m gpu: the CUDA path through the C ABI against the golden vectors and the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import substrata_b200 as sb
from harness import GOLDEN_DIR, oracle_lib, ref_lib

EMU = os.path.join(os.path.dirname(os.path.abspath(__file__)), "native", "libsbm_emu.so")
_vp, _sz, _i = C.c_void_p, C.c_size_t, C.c_int


def _ptr(a):
    return a.ctypes.data_as(_vp)


class CResampler:
    """orc_polyphase_* / ref_polyphase_* behind the PolyphaseResampler method names."""

    def __init__(self, lib, prefix):
        self.l, self.px = lib, prefix
        f = self._fn("create", _vp, [])
        self.h = _vp(f())
        self.channels = 0

    def _fn(self, name, res, args):
        f = getattr(self.l, self.px + "polyphase_" + name)
        f.restype, f.argtypes = res, args
        return f

    def set_rate_and_num_channels(self, s, d, nch):
        self._fn("set_rate_and_channels", None, [_vp, _i, _i, _i])(self.h, s, d, nch)
        self.channels = nch

    def reset_state(self):
        self._fn("reset_state", None, [_vp])(self.h)

    def get_max_output_length(self, n):
        return self._fn("max_output_length", _sz, [_vp, _sz])(self.h, n)

    def get_next_output_length(self, n):
        return self._fn("next_output_length", _sz, [_vp, _sz])(self.h, n)

    def process(self, x):
        x = np.ascontiguousarray(np.atleast_2d(x), np.float32)
        n_out = self.get_next_output_length(x.shape[1])
        out = np.zeros((self.channels, max(n_out, 1)), np.float32)
        if self.px == "ref_":
            self._fn("process", None, [_vp, _vp, _sz, _i, _vp, _sz])(self.h, _ptr(x), x.shape[1], self.channels, _ptr(out), n_out)
        else:
            self._fn("process", None, [_vp, _vp, _sz, _vp, _sz])(self.h, _ptr(x), x.shape[1], _ptr(out), n_out)
        return out[:, :n_out]

    def close(self):
        if self.h:
            self._fn("destroy", None, [_vp])(self.h)
            self.h = None


def make(kind):
    if kind == "oracle":
        return CResampler(oracle_lib(), "orc_")
    if kind == "ref":
        if ref_lib() is None:
            pytest.skip("oracle/_ref not built")
        return CResampler(ref_lib(), "ref_")
    if kind == "emu":
        return sb.PolyphaseResampler(_lib_path=EMU)
    return sb.PolyphaseResampler()


def supported(kind, s, d):
    if kind == "oracle":
        return bool(oracle_lib().orc_polyphase_rates_supported(s, d))
    return sb.PolyphaseResampler.are_sample_rates_supported(s, d, _lib_path=EMU if kind == "emu" else None)


def golden_cases():
    g = np.load(os.path.join(GOLDEN_DIR, "polyphase.npz"))
    for name in sorted({k.split("/")[0] for k in g.files}):
        cfg = g[name + "/cfg"].tolist()
        yield name, cfg[0], cfg[1], cfg[2], cfg[3:], g[name + "/in"], g[name + "/out"]


def run_calls(r, src, dst, nch, lengths, x):
    r.set_rate_and_num_channels(src, dst, nch)
    outs, pos = [], 0
    for n in lengths:
        outs.append(r.process(x[:, pos:pos + n]))
        pos += n
    return np.concatenate(outs, axis=1)


def check_golden(kind):
    for name, src, dst, nch, lengths, x, want in golden_cases():
        r = make(kind)
        got = run_calls(r, src, dst, nch, lengths, x)
        r.close()
        assert got.shape == want.shape, name
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), "%s: not bit-exact (max diff %g)" % (name, np.abs(got - want).max())


RANDOM_CASES = [(44100, 48000, 2, [512, 1, 2, 3, 700, 13, 14, 15, 1024]), (48000, 22050, 16, [2048, 5, 2048]), (8000, 48000, 1, [100, 1, 1, 1, 64]),
                (96000, 48000, 4, [16384]), (48000, 32000, 3, [10, 10, 10, 3000]), (22050, 22050, 2, [100, 50])]


def check_against(kind, ref_kind):
    for ci, (src, dst, nch, lengths) in enumerate(RANDOM_CASES):
        x = np.random.default_rng(100 + ci).uniform(-1, 1, (nch, sum(lengths))).astype(np.float32)
        a, b = make(kind), make(ref_kind)
        ya, yb = run_calls(a, src, dst, nch, lengths, x), run_calls(b, src, dst, nch, lengths, x)
        # a rate change mid-stream keeps the state (SetRateAndNumChannels only regenerates the filter)
        a.set_rate_and_num_channels(dst, src, nch), b.set_rate_and_num_channels(dst, src, nch)
        za, zb = a.process(x[:, :257]), b.process(x[:, :257])
        a.reset_state(), b.reset_state()
        wa, wb = a.process(x[:, :100]), b.process(x[:, :100])
        a.close(), b.close()
        for u, v in ((ya, yb), (za, zb), (wa, wb)):
            assert u.shape == v.shape and np.array_equal(u.view(np.uint32), v.view(np.uint32)), (kind, ref_kind, src, dst)


def sine(freq, rate, n):
    return np.sin(2.0 * np.pi * freq * np.arange(n) / rate).astype(np.float32)[None, :]


def zero_crossings(y):
    sq = (np.abs(y) >= 0.5).astype(np.int32)
    return (1 + int(np.count_nonzero(sq[1:] != sq[:-1]))) // 2


def check_kats(kind):
    """resampler_test.cc, restated."""
    x = sine(1000, 44100, 441)
    # UpSampleTest / DownSampleTest (:39-115): ten periods stay ten periods
    for dst in (48000, 24000):
        r = make(kind)
        r.set_rate_and_num_channels(44100, dst, 1)
        y = r.process(x)
        assert y.shape[1] == dst // (44100 // 441)
        assert zero_crossings(y[0]) == 2 * 1000 // (44100 // 441)
        # ResetStateTest (:117-141)
        r.reset_state()
        assert np.array_equal(r.process(x), y)
        r.close()
    # TwoChannelTest (:143-176)
    r = make(kind)
    r.set_rate_and_num_channels(44100, 48000, 2)
    y = r.process(np.concatenate([sine(1000, 44100, 441), sine(2000, 44100, 441)]))
    assert zero_crossings(y[0]) == 20 and zero_crossings(y[1]) == 40
    r.close()
    # DiracImpulseUpsampleTest (:178-201): the impulse moves to 2 * 64 plus the filter's group delay, peak ~1
    r = make(kind)
    r.set_rate_and_num_channels(100, 200, 1)
    d = np.zeros((1, 128), np.float32)
    d[0, 64] = 1.0
    assert r.get_next_output_length(128) == 256
    y = r.process(d)
    assert y.shape[1] == 256
    cpp = 14      # max_rate 2 -> filter length 26 -> 13 taps per phase padded to 14; the centre tap sits 13 outputs late
    assert int(np.argmax(np.abs(y[0]))) == 128 + 13 and abs(y[0, 128 + 13] - 1.0) < 1e-3, (int(np.argmax(np.abs(y[0]))), cpp)
    r.close()
    # GetNextOutputLengthTest (:203-216)
    r = make(kind)
    r.set_rate_and_num_channels(44100, 22050, 1)
    assert r.get_next_output_length(10) == 5
    r.set_rate_and_num_channels(44100, 88200, 1)
    assert r.get_next_output_length(10) == 20
    r.close()
    # AreSampleRatesSupportedTest (:218-231)
    pairs = [(4000, 96000), (16000, 44100), (44100, 48000), (96000, 32000), (41999, 44100), (48000, 43210)]
    assert [supported(kind, s, d) for s, d in pairs] == [True, True, True, True, False, False]
    # OutputLengthTest (:238-259)
    for s, d in ((44100, 48000), (48000, 44100)):
        r = make(kind)
        r.set_rate_and_num_channels(s, d, 1)
        mx = r.get_max_output_length(512)
        z = np.zeros((1, 512), np.float32)
        for _ in range(100):
            n = r.get_next_output_length(512)
            assert n <= mx and mx - n <= 1
            assert r.process(z).shape[1] == n
        r.close()


# ------------------------------------------------------------------------------------------------ no GPU
def test_oracle_matches_reference_golden():
    check_golden("oracle")


def test_oracle_matches_compiled_reference_live():
    check_against("oracle", "ref")


def test_oracle_known_answers():
    check_kats("oracle")


def test_host_logic_matches_reference_golden_through_emulation():
    check_golden("emu")


def test_host_logic_matches_oracle_through_emulation():
    check_against("emu", "oracle")
    check_kats("emu")


def test_argument_errors_through_emulation():
    r = make("emu")
    with pytest.raises(sb.SbmError):
        r.set_rate_and_num_channels(48000, 43210, 1)      # unsupported pair (filter longer than kMaxSupportedNumFrames)
    with pytest.raises(sb.SbmError):
        r.set_rate_and_num_channels(48000, 44100, 17)
    r.set_rate_and_num_channels(48000, 44100, 2)
    x = np.zeros((2, 100), np.float32)
    out = np.zeros((2, 100), np.float32)
    assert r._l.sbm_polyphase_process(r.h, _ptr(x), 100, _ptr(out), 100) == -1      # SBM_ERR_INVALID_ARG: n_out is not GetNextOutputLength
    r.close()


def test_resample_sh_hrir_through_emulation():
    hr = sb.load_sh_hrir()
    o = make("oracle")
    o.set_rate_and_num_channels(48000, 44100, hr.shape[0])
    want = o.process(hr)
    o.close()
    got = sb.resample_sh_hrir(hr, 48000, 44100, _lib_path=EMU)
    assert got.shape == want.shape and np.array_equal(got, want)
    assert np.array_equal(sb.resample_sh_hrir(hr, 48000, 48000, _lib_path=EMU), hr)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_cuda_matches_reference_golden():
    check_golden("cuda")


@pytest.mark.gpu
def test_cuda_matches_oracle_bit_exact():
    check_against("cuda", "oracle")
    check_kats("cuda")


@pytest.mark.gpu
def test_cuda_resample_sh_hrir_and_max_sizes():
    hr = sb.load_sh_hrir()
    o = make("oracle")
    o.set_rate_and_num_channels(48000, 44100, hr.shape[0])
    want = o.process(hr)
    o.close()
    got = sb.resample_sh_hrir(hr, 48000, 44100)
    assert got.shape == want.shape and np.array_equal(got, want)
    # kMaxSupportedNumFrames x kMaxChannels, the longest supported filter (4000 -> 96000: up 24, 13 * 24 taps)
    x = np.random.default_rng(5).uniform(-1, 1, (16, 16384 // 24)).astype(np.float32)
    a, b = make("cuda"), make("oracle")
    ya, yb = run_calls(a, 4000, 96000, 16, [x.shape[1]], x), run_calls(b, 4000, 96000, 16, [x.shape[1]], x)
    assert np.array_equal(ya, yb)
    a.close(), b.close()
