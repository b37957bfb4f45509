"""Generates tests/golden/polyphase.npz from the COMPILED REFERENCE (oracle/_ref/libsubstrata_ref.so: the real vraudio::Resampler,
RA/dsp/resampler.cc) -- run in the development container, where /root/reference exists and oracle/Makefile has built _ref.

Each case: seeded planar input [channels][total], a list of Process call lengths, the concatenated output of those calls."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import harness  # noqa: E402

CASES = {
    # name: (source_hz, destination_hz, channels, call lengths)
    "up_44100_48000_mono": (44100, 48000, 1, [441, 441, 100, 7, 300]),
    "down_48000_44100_stereo": (48000, 44100, 2, [480, 123, 5, 9, 512]),
    "down_48000_24000_sh16": (48000, 24000, 16, [512, 256, 33]),
    "up_100_200_dirac": (100, 200, 1, [128]),
    "up_32000_48000_3ch": (32000, 48000, 3, [64, 3, 2, 640]),
    "hrir_48000_44100_sh16": (48000, 44100, 16, [512]),
}


def run_ref(src, dst, nch, lengths, x):
    ref = harness.ref_lib()
    ref.ref_polyphase_create.restype = C.c_void_p
    ref.ref_polyphase_next_output_length.restype = C.c_size_t
    ref.ref_polyphase_next_output_length.argtypes = [C.c_void_p, C.c_size_t]
    ref.ref_polyphase_process.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t]
    ref.ref_polyphase_set_rate_and_channels.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    ref.ref_polyphase_destroy.argtypes = [C.c_void_p]
    p = C.c_void_p(ref.ref_polyphase_create())
    ref.ref_polyphase_set_rate_and_channels(p, src, dst, nch)
    outs, pos = [], 0
    for n in lengths:
        seg = np.ascontiguousarray(x[:, pos:pos + n])
        pos += n
        n_out = ref.ref_polyphase_next_output_length(p, n)
        o = np.zeros((nch, max(n_out, 1)), np.float32)
        ref.ref_polyphase_process(p, seg.ctypes.data_as(C.c_void_p), n, nch, o.ctypes.data_as(C.c_void_p), n_out)
        outs.append(o[:, :n_out])
    ref.ref_polyphase_destroy(p)
    return np.concatenate(outs, axis=1)


def case_input(name, nch, total):
    if "dirac" in name:
        x = np.zeros((nch, total), np.float32)
        x[:, total // 2] = 1.0
        return x
    rng = np.random.default_rng(sum(map(ord, name)))
    return rng.uniform(-1.0, 1.0, (nch, total)).astype(np.float32)


def main():
    if harness.ref_lib() is None:
        raise SystemExit("oracle/_ref/libsubstrata_ref.so is not built (make -C oracle ref)")
    out = {}
    for name, (src, dst, nch, lengths) in CASES.items():
        x = case_input(name, nch, sum(lengths))
        y = run_ref(src, dst, nch, lengths, x)
        out[name + "/cfg"] = np.array([src, dst, nch] + lengths, np.int64)
        out[name + "/in"] = x
        out[name + "/out"] = y
        print(name, x.shape, "->", y.shape)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "polyphase.npz"), **out)


if __name__ == "__main__":
    main()
