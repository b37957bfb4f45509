"""substrata_b200 -- B200-native spatial-audio mixer behind Substrata's glare::AudioEngine / AudioResampler interface.

This package is the thin Python face of the product library ``libsubstrata_b200.so`` (C ABI: include/sbm_audio.h;
host engine in csrc/engine.cc; hand-written sm_100a kernels in csrc/*.cu).  There is no CPU path: loading fails loudly
if the library has not been built, and every call that needs the device fails if CUDA is unavailable.

  * ``lib()``          -- the ctypes handle (builds nothing; see ``substrata_b200.build``).
  * ``MixEngine``      -- one engine = one ``sbm_engine`` = "ResonanceThread's per-buffer body + Resonance Audio".
  * ``AudioResampler`` -- batched drop-in for glare::AudioResampler (audio/AudioResampler.h:23-42).
  * ``host.AudioEngine`` / ``host.AudioSource`` -- mirror of glare::AudioEngine's public surface (audio/AudioEngine.h:176-233).
"""
import ctypes as C
import os

import numpy as np

from .abi import bind, SbmConfig, SbmSourceState, SbmStats, ERROR_NAMES

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsubstrata_b200.so")
HRIR_PATH = os.path.join(HERE, "data", "sh_hrir_order3_48k.f32")
HRIR1_PATH = os.path.join(HERE, "data", "sh_hrir_order1_48k.f32")

KIND_PAN = 0        # vraudio::RenderingMode::kStereoPanning
KIND_HRTF = 3       # vraudio::RenderingMode::kBinauralHighQuality (AudioEngine.cpp:857)
KIND_STEREO = 100   # CreateStereoSource(2) (AudioEngine.cpp:867)

_lib = None


class SbmError(RuntimeError):
    def __init__(self, code, where, detail=""):
        self.code = code
        super().__init__("%s failed: %s%s" % (where, ERROR_NAMES.get(code, str(code)), (" -- " + detail) if detail else ""))


def lib(path=None):
    """The loaded product library.  `path` overrides the location (the no-GPU tests load the host-emulation build)."""
    global _lib
    if path is not None:
        return bind(C.CDLL(path))
    if _lib is None:
        override = os.environ.get("SBM_LIB_PATH")      # development: A/B-testing kernel variants on the GPU box
        if override:
            _lib = bind(C.CDLL(override))
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libsubstrata_b200.so is not built (run `python -m substrata_b200.build`); there is no fallback path")
        _lib = bind(C.CDLL(LIB_PATH))
    return _lib


def load_sh_hrir(taps=None):
    """Order-3 SADIE SH-HRIRs [16][256] f32 at 48 kHz, exactly what the reference's CreateShHrirsFromAssets returns
    (RA/dsp/sh_hrir_creator.cc:60-70; dumped by tests/golden/make_golden.py).  `taps` truncates (config C3: 128)."""
    a = np.fromfile(HRIR_PATH, dtype=np.float32).reshape(16, -1)
    return np.ascontiguousarray(a[:, :taps]) if taps else a


def load_sh_hrir_order1():
    """Order-1 SADIE SH-HRIRs [4][256] f32 at 48 kHz: what the reference's order-1 decoder (the reflections') receives."""
    return np.fromfile(HRIR1_PATH, dtype=np.float32).reshape(4, -1)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def pinned_array(n, dtype=np.float32, _lib_path=None):
    """A numpy array over page-locked host memory (sbm_host_alloc_pinned); kept alive for the life of the process."""
    l = lib(_lib_path) if _lib_path else lib()
    nbytes = int(n) * np.dtype(dtype).itemsize
    ptr = l.sbm_host_alloc_pinned(nbytes)
    if not ptr:
        raise MemoryError("sbm_host_alloc_pinned(%d)" % nbytes)
    buf = (C.c_ubyte * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=int(n))


class MixEngine:
    """ctypes wrapper of sbm_engine with the method names of the test harness backends (tests/harness.py)."""

    name = "b200"

    def __init__(self, frames, rate=48000, sh_hrir=None, max_sources=4096, max_mix_sources=None, device=0, _lib_path=None):
        self._l = lib(_lib_path)
        self.frames = int(frames)
        self.rate = int(rate)
        cfg = SbmConfig()
        cfg.frames_per_buffer = self.frames
        cfg.sample_rate_hz = self.rate
        cfg.device = int(device)
        cfg.max_sources = int(max_sources)
        cfg.max_mix_sources = int(max_mix_sources if max_mix_sources is not None else max_sources)
        self._hrir = None
        if sh_hrir is not None:
            self._hrir = _f32(sh_hrir)
            cfg.sh_hrir = self._hrir.ctypes.data_as(C.POINTER(C.c_float))
            cfg.sh_hrir_taps = int(self._hrir.shape[1])
        h = C.c_void_p()
        rc = self._l.sbm_create(C.byref(cfg), C.byref(h))
        if rc:
            raise SbmError(rc, "sbm_create", "is a B200 visible? (no CPU fallback exists)")
        self.h = h

    # -- plumbing -------------------------------------------------------------------------------------------------
    def _ck(self, rc, where):
        if rc:
            raise SbmError(rc, where, self._l.sbm_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self._l.sbm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- buffers / sources ----------------------------------------------------------------------------------------
    def buffer_create(self, samples):
        s = _f32(samples)
        out = C.c_int()
        self._ck(self._l.sbm_buffer_create(self.h, _p(s), s.size, C.byref(out)), "sbm_buffer_create")
        return out.value

    def buffer_adopt_device(self, device_ptr, n, stride_samples, count):
        out = C.c_int()
        self._ck(self._l.sbm_buffer_adopt_device(self.h, C.c_void_p(int(device_ptr)), int(n), int(stride_samples), int(count), C.byref(out)),
                 "sbm_buffer_adopt_device")
        return out.value

    def source_create(self, kind, count=1):
        out = C.c_int()
        self._ck(self._l.sbm_source_create(self.h, int(kind), int(count), C.byref(out)), "sbm_source_create")
        return out.value

    def source_destroy(self, ids):
        i = _i32(np.atleast_1d(ids))
        self._ck(self._l.sbm_source_destroy(self.h, i.size, _p(i)), "sbm_source_destroy")

    def source_attach_buffer(self, sid, buf, cur_read_i=0, sampling_rate=48000, looping=True, remove_on_finish=False):
        self.sources_attach_buffer([sid], [buf], [cur_read_i], [sampling_rate], [looping], [remove_on_finish])

    def sources_attach_buffer(self, ids, bufs, cur_read_i=None, sampling_rate=None, looping=None, remove_on_finish=None):
        i, b = _i32(ids), _i32(bufs)
        n = i.size
        cur = np.ascontiguousarray(cur_read_i if cur_read_i is not None else np.zeros(n), dtype=np.uint64)
        sr = _i32(sampling_rate if sampling_rate is not None else np.full(n, self.rate))
        lo = np.ascontiguousarray(looping if looping is not None else np.ones(n), dtype=np.uint8)
        rm = np.ascontiguousarray(remove_on_finish if remove_on_finish is not None else np.zeros(n), dtype=np.uint8)
        self._ck(self._l.sbm_source_attach_buffer(self.h, n, _p(i), _p(b), _p(cur), _p(sr), _p(lo), _p(rm)), "sbm_source_attach_buffer")

    def source_attach_ring(self, sid, sampling_rate=48000, capacity=48000):
        self._ck(self._l.sbm_source_attach_ring(self.h, int(sid), int(sampling_rate), int(capacity)), "sbm_source_attach_ring")

    def source_ring_push(self, sid, samples):
        s = _f32(samples)
        self._ck(self._l.sbm_source_ring_push(self.h, int(sid), _p(s), s.size), "sbm_source_ring_push")

    def sources_ring_push(self, ids, packets):
        """Batched push: packets[k] (1-D float arrays) is appended to the ring of source ids[k]; one H2D + one scatter kernel."""
        i = _i32(ids)
        offs = np.zeros(i.size + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(p) for p in packets])
        payload = _f32(np.concatenate([np.asarray(p, np.float32).ravel() for p in packets])) if i.size else _f32(np.zeros(1))
        self._ck(self._l.sbm_source_ring_push_batch(self.h, i.size, _p(i), _p(payload), _p(offs)), "sbm_source_ring_push_batch")

    def sources_ring_push_flat(self, ids, payload, offsets):
        """Same, for a pre-concatenated payload (float32) and offsets (uint64, len(ids) + 1): no Python-side copies."""
        self._ck(self._l.sbm_source_ring_push_batch(self.h, len(ids), _p(ids), _p(payload), _p(offsets)), "sbm_source_ring_push_batch")

    def source_ring_size(self, sid):
        out = C.c_uint64()
        self._ck(self._l.sbm_source_ring_size(self.h, int(sid), C.byref(out)), "sbm_source_ring_size")
        return out.value

    def source_attach_mix(self, sid, sampling_rate, bufs, sound_file_i, source_delta, mix_factor):
        b = _i32(bufs)
        p = np.ascontiguousarray(sound_file_i, dtype=np.float64)
        d = np.ascontiguousarray(source_delta, dtype=np.float64)
        m = _f32(mix_factor)
        self._ck(self._l.sbm_source_attach_mix(self.h, int(sid), int(sampling_rate), b.size, _p(b), _p(p), _p(d), _p(m)), "sbm_source_attach_mix")

    def sources_attach_mix1(self, ids, bufs, sound_file_i, source_delta, mix_factor):
        i, b = _i32(ids), _i32(bufs)
        p = np.ascontiguousarray(sound_file_i, dtype=np.float64)
        d = np.ascontiguousarray(source_delta, dtype=np.float64)
        m = _f32(mix_factor)
        self._ck(self._l.sbm_source_attach_mix1(self.h, i.size, _p(i), _p(b), _p(p), _p(d), _p(m)), "sbm_source_attach_mix1")

    def source_set_mix_params(self, sid, z, source_delta, mix_factor):
        i, zz = _i32([sid]), _i32([z])
        d = np.array([source_delta], dtype=np.float64)
        m = _f32([mix_factor])
        self._ck(self._l.sbm_source_set_mix_params(self.h, 1, _p(i), _p(zz), _p(d), _p(m)), "sbm_source_set_mix_params")

    def source_set_eof_marker(self, sid, n):
        self._ck(self._l.sbm_source_set_eof_marker(self.h, int(sid), int(n)), "sbm_source_set_eof_marker")

    def source_set_looping(self, sid, looping):
        self._ck(self._l.sbm_source_set_looping(self.h, int(sid), int(looping)), "sbm_source_set_looping")

    # -- setters --------------------------------------------------------------------------------------------------
    def set_source_positions(self, ids, xyz):
        i, x = _i32(ids), _f32(xyz)
        self._ck(self._l.sbm_set_source_positions(self.h, i.size, _p(i), _p(x)), "sbm_set_source_positions")

    def set_source_volumes(self, ids, vol):
        i, v = _i32(ids), _f32(vol)
        self._ck(self._l.sbm_set_source_volumes(self.h, i.size, _p(i), _p(v)), "sbm_set_source_volumes")

    def set_source_occlusions(self, ids, occ):
        i, v = _i32(ids), _f32(occ)
        self._ck(self._l.sbm_set_source_occlusions(self.h, i.size, _p(i), _p(v)), "sbm_set_source_occlusions")

    def set_head(self, pos, quat_xyzw):
        p, q = _f32(pos), _f32(quat_xyzw)
        self._ck(self._l.sbm_set_head(self.h, _p(p), _p(q)), "sbm_set_head")

    def set_master_volume(self, v):
        self._ck(self._l.sbm_set_master_volume(self.h, float(v)), "sbm_set_master_volume")

    # -- room effects (AudioEngine.cpp:738-768) ------------------------------------------------------------------
    def set_room_hrir(self, sh_hrir_order1):
        h = _f32(sh_hrir_order1)
        self._ck(self._l.sbm_set_room_sh_hrir(self.h, _p(h), int(h.shape[1])), "sbm_set_room_sh_hrir")

    def enable_room_effects(self, on):
        self._ck(self._l.sbm_enable_room_effects(self.h, int(bool(on))), "sbm_enable_room_effects")

    def set_reflection_properties(self, room_position, room_rotation_xyzw, room_dimensions, coefficients6, cutoff_frequency=0.0, gain=0.0):
        a, b, c, d = _f32(room_position), _f32(room_rotation_xyzw), _f32(room_dimensions), _f32(coefficients6)
        self._ck(self._l.sbm_set_reflection_properties(self.h, _p(a), _p(b), _p(c), _p(d), float(cutoff_frequency), float(gain)), "sbm_set_reflection_properties")

    # -- the per-buffer call --------------------------------------------------------------------------------------
    def mix(self, out=None):
        """sbm_mix: returns (valid, interleaved stereo [2*frames])."""
        if out is None:
            out = np.zeros(2 * self.frames, dtype=np.float32)
        valid = C.c_int()
        self._ck(self._l.sbm_mix(self.h, _p(out), C.byref(valid)), "sbm_mix")
        return valid.value, out

    def mix_submit(self):
        """Enqueue one whole buffer without waiting (at most two in flight); pair with mix_collect()."""
        self._ck(self._l.sbm_mix_submit(self.h), "sbm_mix_submit")

    def mix_collect(self, out=None):
        if out is None:
            out = np.zeros(2 * self.frames, dtype=np.float32)
        valid = C.c_int()
        self._ck(self._l.sbm_mix_collect(self.h, _p(out), C.byref(valid)), "sbm_mix_collect")
        return valid.value, out

    def mix_begin(self):
        self._ck(self._l.sbm_mix_begin(self.h), "sbm_mix_begin")

    def mix_finish(self, out=None, copy_out=True):
        if out is None and copy_out:
            out = np.zeros(2 * self.frames, dtype=np.float32)
        valid = C.c_int()
        self._ck(self._l.sbm_mix_finish(self.h, _p(out) if copy_out else None, C.byref(valid)), "sbm_mix_finish")
        return valid.value, out

    def peer_export(self):
        """64-byte CUDA IPC handle of this engine's gather area (multi-GPU bus exchange over peer memory)."""
        h = (C.c_ubyte * 64)()
        self._ck(self._l.sbm_peer_export(self.h, h, 64), "sbm_peer_export")
        return bytes(h)

    def peer_connect(self, rank, world, handles):
        """handles: the ranks' peer_export() results in rank order (exchanged by any host transport)."""
        blob = b"".join(handles)
        assert len(blob) == 64 * world
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        self._ck(self._l.sbm_peer_connect(self.h, int(rank), int(world), buf, 64), "sbm_peer_connect")

    def bus_device(self):
        ptr, n = C.c_void_p(), C.c_size_t()
        self._ck(self._l.sbm_bus_device(self.h, C.byref(ptr), C.byref(n)), "sbm_bus_device")
        return ptr.value, n.value

    def stream(self):
        return self._l.sbm_stream(self.h)

    def ambisonic_bus(self):
        out = np.zeros((16, self.frames), dtype=np.float32)
        self._ck(self._l.sbm_debug_get_ambisonic_bus(self.h, _p(out)), "sbm_debug_get_ambisonic_bus")
        return out

    def poll_finished(self, cap=1024):
        ids = np.zeros(cap, dtype=np.int32)
        n = C.c_int()
        self._ck(self._l.sbm_poll_finished(self.h, _p(ids), cap, C.byref(n)), "sbm_poll_finished")
        return ids[:n.value].tolist()

    def source_query(self, sid):
        st = SbmSourceState()
        rc = self._l.sbm_source_query(self.h, int(sid), C.byref(st))
        if rc == -4:
            return None
        self._ck(rc, "sbm_source_query")
        return {"cur_read_i": st.cur_read_i, "eof_marker": st.eof_marker, "ring_size": st.ring_size, "mix_pos0": st.mix_pos0,
                "direct_gain": st.direct_gain, "occlusion": st.occlusion, "alive": st.alive}

    def stats(self):
        st = SbmStats()
        self._ck(self._l.sbm_get_stats(self.h, C.byref(st)), "sbm_get_stats")
        return {k: getattr(st, k) for k, _ in SbmStats._fields_}

    def set_profiling(self, on=True):
        self._ck(self._l.sbm_set_profiling(self.h, int(on)), "sbm_set_profiling")


class AudioResampler:
    """Batched glare::AudioResampler (audio/AudioResampler.h:23-42): init(src, dst) / numSrcSamplesNeeded / resample."""

    def __init__(self, src_rates, dst_rate, device=0):
        self._l = lib()
        self.src_rates = _i32(np.atleast_1d(src_rates))
        self.count = self.src_rates.size
        h = C.c_void_p()
        rc = self._l.sbm_resampler_create(int(device), self.count, _p(self.src_rates), int(dst_rate), C.byref(h))
        if rc:
            raise SbmError(rc, "sbm_resampler_create")
        self.h = h

    def num_src_samples_needed(self, n_dst):
        out = np.zeros(self.count, dtype=np.uint64)
        rc = self._l.sbm_resampler_num_src_needed(self.h, int(n_dst), _p(out))
        if rc:
            raise SbmError(rc, "sbm_resampler_num_src_needed")
        return out

    def resample(self, n_dst, src_chunks):
        """src_chunks: one f32 array per stream holding exactly num_src_samples_needed(n_dst)[k] samples."""
        offs = np.zeros(self.count, dtype=np.uint64)
        tot = 0
        for k, c in enumerate(src_chunks):
            offs[k] = tot
            tot += len(c)
        src = _f32(np.concatenate([_f32(c) for c in src_chunks])) if tot else np.zeros(1, np.float32)
        dst = np.zeros((self.count, int(n_dst)), dtype=np.float32)
        rc = self._l.sbm_resampler_resample(self.h, _p(dst), int(n_dst), _p(src), _p(offs), tot)
        if rc:
            raise SbmError(rc, "sbm_resampler_resample")
        return dst

    def close(self):
        if getattr(self, "h", None):
            self._l.sbm_resampler_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PolyphaseResampler:
    """vraudio::Resampler (RA/dsp/resampler.h:27-130): Resonance's polyphase sample-rate converter, bit-exact on the GPU.

    Audio is planar [channels][frames] float32.  Method names follow the reference's."""

    def __init__(self, device=0, _lib_path=None):
        self._l = lib(_lib_path)
        h = C.c_void_p()
        rc = self._l.sbm_polyphase_create(int(device), C.byref(h))
        if rc:
            raise SbmError(rc, "sbm_polyphase_create")
        self.h = h
        self.channels = 0

    @staticmethod
    def are_sample_rates_supported(source_hz, destination_hz, _lib_path=None):
        return bool(lib(_lib_path).sbm_polyphase_rates_supported(int(source_hz), int(destination_hz)))

    def set_rate_and_num_channels(self, source_hz, destination_hz, num_channels):
        rc = self._l.sbm_polyphase_set_rate_and_channels(self.h, int(source_hz), int(destination_hz), int(num_channels))
        if rc:
            raise SbmError(rc, "sbm_polyphase_set_rate_and_channels")
        self.channels = int(num_channels)

    def reset_state(self):
        self._l.sbm_polyphase_reset_state(self.h)

    def get_max_output_length(self, input_length):
        return int(self._l.sbm_polyphase_max_output_length(self.h, int(input_length)))

    def get_next_output_length(self, input_length):
        return int(self._l.sbm_polyphase_next_output_length(self.h, int(input_length)))

    def process(self, x):
        x = _f32(np.atleast_2d(x))
        if x.shape[0] != self.channels:
            raise ValueError("expected [%d][frames]" % self.channels)
        n_out = self.get_next_output_length(x.shape[1])
        out = np.zeros((self.channels, max(n_out, 1)), dtype=np.float32)
        rc = self._l.sbm_polyphase_process(self.h, _p(x), x.shape[1], _p(out), n_out)
        if rc:
            raise SbmError(rc, "sbm_polyphase_process")
        return out[:, :n_out]

    def close(self):
        if getattr(self, "h", None):
            self._l.sbm_polyphase_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def resample_sh_hrir(sh_hrir, source_hz, destination_hz, device=0, _lib_path=None):
    """The resampling step of CreateShHrirsFromWav (RA/dsp/sh_hrir_creator.cc:42-55): [channels][taps] -> [channels][taps']."""
    l = lib(_lib_path)
    h = _f32(np.atleast_2d(sh_hrir))
    n = C.c_int(0)
    rc = l.sbm_resample_sh_hrir(int(device), _p(h), h.shape[0], h.shape[1], int(source_hz), int(destination_hz), None, 0, C.byref(n))
    if rc:
        raise SbmError(rc, "sbm_resample_sh_hrir")
    out = np.zeros((h.shape[0], max(n.value, 1)), dtype=np.float32)
    rc = l.sbm_resample_sh_hrir(int(device), _p(h), h.shape[0], h.shape[1], int(source_hz), int(destination_hz), _p(out), out.shape[1], C.byref(n))
    if rc:
        raise SbmError(rc, "sbm_resample_sh_hrir")
    return out[:, :n.value]
