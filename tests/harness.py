"""Shared test harness: uniform Python wrappers over the three implementations of the mix path.

  * ``RefBackend``     -- the compiled reference (oracle/_ref/libsubstrata_ref.so: real Resonance Audio +
                           real AudioResampler + restated feeder).  Optional: present where it was built.
  * ``OracleBackend``  -- the plain-C restatement (oracle/libmix_oracle.so).
  * ``substrata_b200.MixEngine`` -- the product (CUDA, through the C ABI of include/sbm_audio.h).

All three expose the same methods, so one scenario (``Scenario``) can be replayed on each and the
interleaved stereo outputs compared.  Everything under oracle/ is test infrastructure only.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

KIND_PAN = 0       # vraudio::RenderingMode::kStereoPanning
KIND_HRTF = 3      # vraudio::RenderingMode::kBinauralHighQuality (what AudioEngine.cpp:857 uses)
KIND_STEREO = 100  # non-spatial stereo source (AudioEngine.cpp:867)

_vp, _i, _f, _sz, _u64, _i64, _d = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_uint64, C.c_int64, C.c_double


def _ptr(a):
    return a.ctypes.data_as(_vp)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def build_oracle():
    so = os.path.join(ORACLE_DIR, "libmix_oracle.so")
    src = os.path.join(ORACLE_DIR, "mix_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])
    return so


def ref_library_path():
    p = os.path.join(ORACLE_DIR, "_ref", "libsubstrata_ref.so")
    return p if os.path.exists(p) else None


def load_sh_hrir_order1():
    """Order-1 SADIE SH-HRIRs [4][256]: the reflections' decoder (room effects); dumped from the compiled reference."""
    return np.fromfile(os.path.join(ROOT, "substrata_b200", "data", "sh_hrir_order1_48k.f32"), dtype=np.float32).reshape(4, -1)


def load_sh_hrir():
    """Order-3 SADIE SH-HRIRs [16][256] as the reference decoder receives them (dumped by make_golden.py)."""
    a = np.fromfile(os.path.join(ROOT, "substrata_b200", "data", "sh_hrir_order3_48k.f32"), dtype=np.float32)
    return a.reshape(16, -1)


class _CBackend:
    """Common ctypes plumbing for the ref_/orc_ entry points (identical signatures after the handle)."""

    prefix = None

    def _bind(self, lib):
        p = self.prefix
        sig = {
            "buffer_create": (_i, [_vp, _vp, _sz]),
            "source_create": (_i, [_vp, _i]),
            "source_attach_buffer": (None, [_vp, _i, _i, _u64, _i, _i, _i]),
            "source_attach_ring": (None, [_vp, _i, _i]),
            "source_ring_push": (None, [_vp, _i, _vp, _sz]),
            "source_attach_mix": (None, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
            "source_set_mix_params": (None, [_vp, _i, _i, _d, _f]),
            "source_set_eof_marker": (None, [_vp, _i, _i64]),
            "source_set_looping": (None, [_vp, _i, _i]),
            "set_source_positions": (None, [_vp, _i, _vp, _vp]),
            "set_source_volumes": (None, [_vp, _i, _vp, _vp]),
            "set_source_occlusions": (None, [_vp, _i, _vp, _vp]),
            "set_head": (None, [_vp, _vp, _vp]),
            "set_master_volume": (None, [_vp, _f]),
            "source_destroy": (None, [_vp, _i]),
            "mix": (_i, [_vp, _vp]),
            "get_ambisonic_bus": (_i, [_vp, _vp]),
            "poll_finished": (_i, [_vp, _vp, _i]),
            "source_query": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
            "destroy": (None, [_vp]),
        }
        self._fn = {}
        for name, (res, args) in sig.items():
            f = getattr(lib, p + name)
            f.restype, f.argtypes = res, args
            self._fn[name] = f

    # -- uniform API -------------------------------------------------------------------------------
    def buffer_create(self, samples):
        s = _f32(samples)
        return self._fn["buffer_create"](self.h, _ptr(s), s.size)

    def source_create(self, kind):
        return self._fn["source_create"](self.h, kind)

    def source_attach_buffer(self, sid, buf, cur_read_i=0, sampling_rate=48000, looping=True, remove_on_finish=False):
        self._fn["source_attach_buffer"](self.h, sid, buf, int(cur_read_i), int(sampling_rate), int(looping), int(remove_on_finish))

    def source_attach_ring(self, sid, sampling_rate=48000):
        self._fn["source_attach_ring"](self.h, sid, int(sampling_rate))

    def source_ring_push(self, sid, samples):
        s = _f32(samples)
        self._fn["source_ring_push"](self.h, sid, _ptr(s), s.size)

    def source_attach_mix(self, sid, sampling_rate, bufs, sound_file_i, source_delta, mix_factor):
        b = np.ascontiguousarray(bufs, dtype=np.int32)
        p = np.ascontiguousarray(sound_file_i, dtype=np.float64)
        d = np.ascontiguousarray(source_delta, dtype=np.float64)
        m = _f32(mix_factor)
        self._fn["source_attach_mix"](self.h, sid, int(sampling_rate), b.size, _ptr(b), _ptr(p), _ptr(d), _ptr(m))

    def source_set_mix_params(self, sid, z, source_delta, mix_factor):
        self._fn["source_set_mix_params"](self.h, sid, z, float(source_delta), float(mix_factor))

    def source_set_eof_marker(self, sid, n):
        self._fn["source_set_eof_marker"](self.h, sid, int(n))

    def source_set_looping(self, sid, looping):
        self._fn["source_set_looping"](self.h, sid, int(looping))

    def set_source_positions(self, ids, xyz):
        i = np.ascontiguousarray(ids, dtype=np.int32)
        x = _f32(xyz)
        self._fn["set_source_positions"](self.h, i.size, _ptr(i), _ptr(x))

    def set_source_volumes(self, ids, vol):
        i = np.ascontiguousarray(ids, dtype=np.int32)
        v = _f32(vol)
        self._fn["set_source_volumes"](self.h, i.size, _ptr(i), _ptr(v))

    def set_source_occlusions(self, ids, occ):
        i = np.ascontiguousarray(ids, dtype=np.int32)
        v = _f32(occ)
        self._fn["set_source_occlusions"](self.h, i.size, _ptr(i), _ptr(v))

    def set_head(self, pos, quat_xyzw):
        p, q = _f32(pos), _f32(quat_xyzw)
        self._fn["set_head"](self.h, _ptr(p), _ptr(q))

    def set_master_volume(self, v):
        self._fn["set_master_volume"](self.h, float(v))

    # room effects (AudioEngine.cpp:738-768): bound lazily, the same three calls on every backend
    def _room_fn(self, name, argtypes):
        f = getattr(self._room_lib(), self.prefix + name)
        f.restype, f.argtypes = None, argtypes
        return f

    def set_room_hrir(self, sh_hrir_order1):
        if self.prefix == "orc_":           # (the compiled reference loads its own order-1 asset)
            h = _f32(sh_hrir_order1)
            self._room_fn("set_room_hrir", [_vp, _vp, _i])(self.h, _ptr(h), h.shape[1])

    def enable_room_effects(self, on):
        self._room_fn("enable_room_effects", [_vp, _i])(self.h, int(bool(on)))

    def set_reflection_properties(self, room_position, room_rotation_xyzw, room_dimensions, coefficients6, cutoff_frequency=0.0, gain=0.0):
        a, b, c, d = _f32(room_position), _f32(room_rotation_xyzw), _f32(room_dimensions), _f32(coefficients6)
        self._room_fn("set_reflection_properties", [_vp, _vp, _vp, _vp, _vp, _f, _f])(self.h, _ptr(a), _ptr(b), _ptr(c), _ptr(d), float(cutoff_frequency), float(gain))

    def source_destroy(self, sid):
        self._fn["source_destroy"](self.h, sid)

    def mix(self):
        out = np.zeros(2 * self.frames, dtype=np.float32)
        valid = self._fn["mix"](self.h, _ptr(out))
        return valid, out

    def ambisonic_bus(self):
        out = np.zeros((16, self.frames), dtype=np.float32)
        ok = self._fn["get_ambisonic_bus"](self.h, _ptr(out))
        return out if ok else None

    def poll_finished(self, cap=1024):
        ids = np.zeros(cap, dtype=np.int32)
        n = self._fn["poll_finished"](self.h, _ptr(ids), cap)
        return ids[:n].tolist()

    def source_query(self, sid):
        cur, eof, ring, pos = _u64(0), _i64(0), _u64(0), _d(0)
        ok = self._fn["source_query"](self.h, sid, C.byref(cur), C.byref(eof), C.byref(ring), C.byref(pos))
        if not ok:
            return None
        return {"cur_read_i": cur.value, "eof_marker": eof.value, "ring_size": ring.value, "mix_pos0": pos.value}

    def close(self):
        if self.h:
            self._fn["destroy"](self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_oracle_lib = None
_ref_lib = None


def oracle_lib():
    global _oracle_lib
    if _oracle_lib is None:
        _oracle_lib = C.CDLL(build_oracle())
    return _oracle_lib


def ref_lib():
    global _ref_lib
    if _ref_lib is None:
        p = ref_library_path()
        if p is None:
            return None
        _ref_lib = C.CDLL(p)
    return _ref_lib


class OracleBackend(_CBackend):
    prefix = "orc_"
    name = "oracle"

    def _room_lib(self):
        return oracle_lib()

    def __init__(self, frames, rate=48000, sh_hrir=None, double_bus=False):
        lib = oracle_lib()
        self._bind(lib)
        lib.orc_create.restype, lib.orc_create.argtypes = _vp, [_i, _i, _vp, _i]
        self.frames = frames
        if sh_hrir is None:
            self.h = lib.orc_create(frames, rate, None, 0)
        else:
            hr = _f32(sh_hrir)
            self.h = lib.orc_create(frames, rate, _ptr(hr), hr.shape[1])
        if double_bus:
            lib.orc_set_bus_precision.argtypes = [_vp, _i]
            lib.orc_set_bus_precision(self.h, 1)
        lib.orc_sources_create.restype, lib.orc_sources_create.argtypes = _i, [_vp, _i, _i]
        lib.orc_sources_attach_buffer.restype, lib.orc_sources_attach_buffer.argtypes = None, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]
        lib.orc_sources_attach_mix1.restype, lib.orc_sources_attach_mix1.argtypes = None, [_vp, _i, _vp, _vp, _vp, _vp, _vp]
        self._lib = lib

    # batched forms (same names and argument meaning as substrata_b200.MixEngine's): the full-size parity runs
    def sources_create(self, kind, count):
        return self._lib.orc_sources_create(self.h, int(kind), int(count))

    def sources_attach_buffer(self, ids, bufs, cur_read_i=None, sampling_rate=None, looping=None, remove_on_finish=None):
        i, b = np.ascontiguousarray(ids, np.int32), np.ascontiguousarray(bufs, np.int32)
        n = i.size
        cur = np.ascontiguousarray(cur_read_i if cur_read_i is not None else np.zeros(n), dtype=np.uint64)
        sr = np.ascontiguousarray(sampling_rate if sampling_rate is not None else np.full(n, 48000), dtype=np.int32)
        lo = np.ascontiguousarray(looping if looping is not None else np.ones(n), dtype=np.uint8)
        rm = np.ascontiguousarray(remove_on_finish if remove_on_finish is not None else np.zeros(n), dtype=np.uint8)
        self._lib.orc_sources_attach_buffer(self.h, n, _ptr(i), _ptr(b), _ptr(cur), _ptr(sr), _ptr(lo), _ptr(rm))

    def sources_attach_mix1(self, ids, bufs, sound_file_i, source_delta, mix_factor):
        i, b = np.ascontiguousarray(ids, np.int32), np.ascontiguousarray(bufs, np.int32)
        p = np.ascontiguousarray(sound_file_i, dtype=np.float64)
        d = np.ascontiguousarray(source_delta, dtype=np.float64)
        m = _f32(mix_factor)
        self._lib.orc_sources_attach_mix1(self.h, i.size, _ptr(i), _ptr(b), _ptr(p), _ptr(d), _ptr(m))


class RefBackend(_CBackend):
    prefix = "ref_"
    name = "reference"

    def _room_lib(self):
        return ref_lib()

    def __init__(self, frames, rate=48000, hrir_taps=0):
        lib = ref_lib()
        if lib is None:
            raise RuntimeError("oracle/_ref/libsubstrata_ref.so not built")
        self._bind(lib)
        lib.ref_create.restype, lib.ref_create.argtypes = _vp, [_i, _i, _i]
        self.frames = frames
        self.h = lib.ref_create(frames, rate, hrir_taps)


# ---------------------------------------------------------------------------------------------------
# Scenario: one deterministic workload description that can be replayed on any backend.
# ---------------------------------------------------------------------------------------------------
def quat_yaw(yaw_rad):
    """Rotation about +y (up in Resonance's world convention, RA/base/misc_math.h:82-87), xyzw order."""
    return np.array([0.0, np.sin(yaw_rad / 2), 0.0, np.cos(yaw_rad / 2)], dtype=np.float32)


def synth_pcm(rng, n, f_lo=100.0, f_hi=4000.0, rate=48000):
    """0.5 sin(2 pi f t) + 0.1 noise."""
    f = rng.uniform(f_lo, f_hi)
    t = np.arange(n, dtype=np.float64) / rate
    return (0.5 * np.sin(2 * np.pi * f * t) + 0.1 * rng.uniform(-1, 1, n)).astype(np.float32)


def positions(rng, n, boundary_safe=False):
    """Uniform in [-40,40]x[-10,10]x[-40,40] m rejecting d < 1.05 m; the boundary-safe variant draws
    (azimuth, elevation) at k+0.5 degrees so that no direction sits near a lookup-table cell edge at t = 0."""
    if boundary_safe:
        az = np.deg2rad(rng.integers(-180, 180, n) + 0.5)
        el = np.deg2rad(rng.integers(-80, 80, n) + 0.5)
        d = rng.uniform(1.5, 40.0, n)
        x = -d * np.cos(el) * np.sin(az)
        y = d * np.sin(el)
        z = -d * np.cos(el) * np.cos(az)
        return np.stack([x, y, z], 1).astype(np.float32)
    out = np.zeros((n, 3), dtype=np.float32)
    k = 0
    while k < n:
        p = rng.uniform([-40, -10, -40], [40, 10, 40], (n, 3))
        p = p[np.linalg.norm(p, axis=1) >= 1.05]
        m = min(n - k, len(p))
        out[k:k + m] = p[:m]
        k += m
    return out


class Scenario:
    """A seeded source set + per-buffer event script.

    kinds:      per-source KIND_*.
    feeder:     per-source "native" | "resample" (rate != 48000) | "doppler" (one mix_source) | "ring".
    The listener starts at the origin and, when `moving`, translates 1 m/s along +x and yaws 30 deg/s,
    updated once per buffer -- which exercises the 1/2048-per-sample gain ramps.
    """

    def __init__(self, n_sources, frames, seed=12345, kind=KIND_PAN, feeder="native", loop_len=16384, moving=True,
                 rates=(44100, 24000, 16000, 48000), rate_probs=(0.5, 0.2, 0.15, 0.15), boundary_safe=False,
                 volume_toggle=False, occlusion_toggle=False, mixed_kinds=False, n_shared_buffers=0, batched_ring_push=False,
                 master_script=None, positions_override=None, room=None):
        self.n, self.F, self.seed = n_sources, frames, seed
        rng = np.random.default_rng(seed)
        self.moving = moving
        self.volume_toggle, self.occlusion_toggle = volume_toggle, occlusion_toggle
        self.pos = positions(rng, n_sources, boundary_safe)
        if positions_override is not None:      # hand-picked positions (cycled over the sources): axes, poles, the listener's own position, out of range
            po = np.asarray(positions_override, np.float32)
            self.pos = po[np.arange(n_sources) % len(po)].copy()
        self.vol = rng.uniform(0.25, 1.0, n_sources).astype(np.float32)
        if mixed_kinds:
            self.kinds = rng.choice([KIND_PAN, KIND_HRTF, KIND_STEREO], n_sources, p=[0.45, 0.45, 0.1])
        else:
            self.kinds = np.full(n_sources, kind)
        self.feeder = feeder
        nb = n_shared_buffers if n_shared_buffers else n_sources
        self.pcm = [synth_pcm(rng, loop_len) for _ in range(nb)]
        self.buf_of = np.arange(n_sources) % nb
        self.start = rng.integers(0, loop_len, n_sources)
        self.rates = rng.choice(rates, n_sources, p=rate_probs) if feeder == "resample" else np.full(n_sources, 48000)
        v = rng.uniform(-30, 30, (n_sources, 2))
        self.delta = (343.0 - v[:, 0]) / (343.0 - v[:, 1])  # doppler_factor, AudioEngine.cpp:140-145
        self.mixf = rng.uniform(0.5, 1.0, n_sources).astype(np.float32)
        self.ring_rng_seed = seed + 1
        self.batched_ring_push = batched_ring_push
        self.master_script = master_script or {}      # buffer index -> master volume (AudioEngine::setMasterVolume, AudioEngine.cpp:805-811)
        # room effects script (AudioEngine.cpp:738-768, GUIClient.cpp:6886-6896): {"on": buffer, "off": buffer or None, "moves": {buffer: (pos, dims)},
        # "pos", "dims", "rot", "cutoff"} -- Substrata's fixed coefficients and gain unless given
        self.room = room

    def build(self, be):
        bufs = [be.buffer_create(p) for p in self.pcm]
        ids = []
        for s in range(self.n):
            sid = be.source_create(int(self.kinds[s]))
            ids.append(sid)
            b = bufs[self.buf_of[s]]
            if self.feeder in ("native", "resample"):
                be.source_attach_buffer(sid, b, int(self.start[s]), int(self.rates[s]), True, False)
            elif self.feeder == "doppler":
                be.source_attach_mix(sid, 48000, [b], [float(self.start[s])], [float(self.delta[s])], [float(self.mixf[s])])
            elif self.feeder == "mix3":
                # the bike: three mix sources per AudioSource, each with its own sound file, read head, pitch and level
                # (BikePhysics.cpp:712-740; AudioEngine.cpp:435-458)
                nb_ = len(bufs)
                zb = [bufs[(self.buf_of[s] + z) % nb_] for z in range(3)]
                be.source_attach_mix(sid, 48000, zb, [float(self.start[s]) * (0.5 + 0.25 * z) for z in range(3)],
                                     [float(self.delta[s]) * self.MIX3_PITCH[z] for z in range(3)],
                                     [float(self.mixf[s]) * self.MIX3_LEVEL[z] for z in range(3)])
            elif self.feeder == "ring":
                be.source_attach_ring(sid, int(self.rates[s]))
            else:
                raise ValueError(self.feeder)
        self.ids = np.array(ids, dtype=np.int32)
        be.set_source_positions(self.ids, self.pos)
        be.set_source_volumes(self.ids, self.vol)
        return self.ids

    MIX3_PITCH = (1.0, 0.5, 1.87)
    MIX3_LEVEL = (0.6, 0.3, 0.45)

    ROOM_COEFFS = (0.8, 0.8, 0.3, 0.8, 0.8, 0.2)      # AudioEngine.cpp:762-765
    ROOM_GAIN = 0.7                                     # AudioEngine.cpp:766

    def _room_events(self, be, b):
        r = self.room
        if b == 0:
            be.set_room_hrir(load_sh_hrir_order1())
        props = None
        if b == r.get("on", 0):
            be.enable_room_effects(True)
            props = (r.get("pos", (1.0, 0.5, 2.0)), r.get("dims", (30.0, 8.0, 24.0)))
        if b in r.get("moves", {}):
            props = r["moves"][b]
        if props is not None:
            be.set_reflection_properties(props[0], r.get("rot", (0.0, 0.0, 0.0, 1.0)), props[1], r.get("coeffs", self.ROOM_COEFFS),
                                         r.get("cutoff", 0.0), r.get("gain", self.ROOM_GAIN))
        if r.get("off") is not None and b == r["off"]:
            be.enable_room_effects(False)

    def step(self, be, b):
        """Events before buffer `b`."""
        if self.room is not None:
            self._room_events(be, b)
        if self.feeder == "mix3" and b > 0:
            # the game thread retunes the engine sound every frame: new source_delta / mix_factor for one of the three
            # mix sources of every other AudioSource
            z = b % 3
            for s in range(0, self.n, 2):
                k = 1.0 + 0.03 * np.sin(0.7 * b + s)
                be.source_set_mix_params(int(self.ids[s]), z, float(self.delta[s]) * self.MIX3_PITCH[z] * k,
                                         float(self.mixf[s]) * self.MIX3_LEVEL[z] * (2.0 - k))
        if self.moving:
            t = b * self.F / 48000.0
            be.set_head([1.0 * t, 0.0, 0.0], quat_yaw(np.deg2rad(30.0) * t))
        if b in self.master_script:
            be.set_master_volume(self.master_script[b])
        if self.volume_toggle:
            sel = self.ids[::4]
            if b % 6 == 2:
                be.set_source_volumes(sel, np.zeros(len(sel), np.float32))
            elif b % 6 == 4:
                be.set_source_volumes(sel, self.vol[::4])
        if self.occlusion_toggle:
            sel = self.ids[1::3]
            if b % 5 == 1:
                be.set_source_occlusions(sel, np.ones(len(sel), np.float32))
            elif b % 5 == 3:
                be.set_source_occlusions(sel, np.zeros(len(sel), np.float32))
        if self.feeder == "ring":
            # 480-sample packets every 10 ms as opus delivers them (ClientUDPHandlerThread.cpp:196); an occasional
            # dropped packet exercises the underflow zero-pad (AudioEngine.cpp:466-473).
            rng = np.random.default_rng(self.ring_rng_seed + b)
            n_pk = int(np.ceil(self.F / 480.0)) + (1 if b == 0 else 0)
            ids, packets = [], []
            for s in range(self.n):
                if b % 7 == 5 and s % 3 == 0:
                    continue
                scale = self.rates[s] / 48000.0
                m = int(n_pk * 480 * scale)
                ids.append(int(self.ids[s])); packets.append((0.3 * rng.uniform(-1, 1, m)).astype(np.float32))
            if self.batched_ring_push and hasattr(be, "sources_ring_push"):
                be.sources_ring_push(ids, packets)       # sbm_source_ring_push_batch: one H2D + scatter kernel
            else:
                for sid, pk in zip(ids, packets):
                    be.source_ring_push(sid, pk)

    def run(self, be, n_buffers):
        self.build(be)
        outs, valids = [], []
        for b in range(n_buffers):
            self.step(be, b)
            v, o = be.mix()
            outs.append(o)
            valids.append(v)
        return np.stack(outs), valids


def run_churn(be, frames, n_buffers=10, n_start=400, per_buffer=60, seed=5, kinds=(KIND_PAN, KIND_HRTF)):
    """playOneShotSound / removeSource churn (AudioEngine.cpp:900-943,1022-1052): every buffer a batch of sources is destroyed
    and a batch of new ones (one-shots with remove_on_finish among them) is created; ids keep increasing, slots are recycled.
    Returns (outputs, valid flags, finished ids per buffer)."""
    rng = np.random.default_rng(seed)
    pcm = [synth_pcm(rng, L) for L in (700, 1500, 4096, 3001)]
    bufs = [be.buffer_create(p) for p in pcm]
    alive = []

    def add(n):
        for _ in range(n):
            kind = int(rng.choice(kinds))
            sid = be.source_create(kind)
            b = int(rng.integers(0, len(bufs)))
            one_shot = rng.uniform() < 0.3
            rate = int(rng.choice([48000, 44100, 24000]))
            be.source_attach_buffer(sid, bufs[b], int(rng.integers(0, len(pcm[b]))), rate, not one_shot, bool(one_shot))
            pos = rng.uniform([-20, -5, -20], [20, 5, 20], (1, 3)).astype(np.float32) + np.float32(1.5)
            be.set_source_positions([sid], pos)
            be.set_source_volumes([sid], np.array([rng.uniform(0.3, 1.0)], np.float32))
            alive.append(sid)

    add(n_start)
    outs, valids, fins = [], [], []
    for b in range(n_buffers):
        t = b * frames / 48000.0
        be.set_head([t, 0.0, 0.0], quat_yaw(0.4 * t))
        v, o = be.mix()
        outs.append(o); valids.append(v)
        fin = sorted(be.poll_finished())
        fins.append(fin)
        for f in fin:
            if f in alive:
                alive.remove(f)
        kill = [alive[i] for i in sorted(rng.choice(len(alive), size=min(per_buffer, len(alive)), replace=False))]
        for sid in kill:
            be.source_destroy(sid)
            alive.remove(sid)
        add(per_buffer)
    return np.stack(outs), valids, fins


def churn_golden():
    g = np.load(os.path.join(GOLDEN_DIR, "scenarios.npz"))
    off = g["churn__fin_off"]
    fins = [g["churn__fin_flat"][off[k]:off[k + 1]].tolist() for k in range(len(off) - 1)]
    return g["churn"], g["churn__valid"].tolist(), fins


def run_pipelined(sc, be, n_buffers):
    """Scenario.run through sbm_mix_submit / sbm_mix_collect with two buffers in flight: the events of buffer b are applied
    while buffer b - 1 is still being mixed."""
    sc.build(be)
    outs, valids = [], []
    for b in range(n_buffers):
        sc.step(be, b)
        be.mix_submit()
        if b >= 1:
            v, o = be.mix_collect()
            outs.append(o); valids.append(v)
    v, o = be.mix_collect()
    outs.append(o); valids.append(v)
    return np.stack(outs), valids


def rel_rms(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    den = np.sqrt(np.sum(b * b))
    return float(np.sqrt(np.sum((a - b) ** 2)) / den) if den > 0 else float(np.sqrt(np.sum(a * a)))


# ---------------------------------------------------------------------------------------------------
# Full-size scenes: BASELINE.json's own source counts against the oracle (tests/test_gpu_fullsize.py, bench.py's parity leg).
# The per-source arrays are drawn once for the WHOLE scene; any backend can build any subset of it (`sel`), so the oracle
# runs as P engines on disjoint subsets in P host threads (ctypes releases the GIL) whose outputs add up -- the mix is
# linear in the source set (one bus sum; the SH decode is a linear filter).  PCM comes from a pool of shared loops (what
# AudioBuffer sharing looks like, AudioEngine.h:122): 10^6 private loops would be 64 GB on the host.
# ---------------------------------------------------------------------------------------------------
FULL_SCENES = {
    # name: BASELINE config it is the full-size form of
    "c2": dict(n=4096, frames=512, kind=KIND_PAN, feeder="resample", taps=0, pool=256, loop=16384, buffers=4, seed=901),
    "c3": dict(n=65536, frames=1024, kind=KIND_HRTF, feeder="native", taps=128, pool=1024, loop=16384, buffers=4, seed=902),
    "c4_rank": dict(n=131072, frames=512, kind=KIND_PAN, feeder="doppler", taps=0, pool=2048, loop=16384, buffers=4, seed=903),
    "c4_1m": dict(n=1048576, frames=512, kind=KIND_PAN, feeder="doppler", taps=0, pool=4096, loop=16384, buffers=4, seed=904),
    "c5": dict(n=33792, frames=256, kind=KIND_HRTF, feeder="c5", taps=256, pool=512, loop=16384, buffers=6, seed=905, voices=1024),
}


class BigScene:
    """One of FULL_SCENES (or an override of it): seeded arrays for every source, replayable on a subset."""

    def __init__(self, n, frames, kind, feeder, taps, pool, loop, buffers, seed, voices=0, volume_toggle=True):
        rng = np.random.default_rng(seed)
        self.n, self.F, self.kind, self.feeder, self.taps, self.nb, self.voices = n, frames, kind, feeder, taps, buffers, voices
        self.volume_toggle = volume_toggle
        t = np.arange(loop, dtype=np.float64) / 48000.0
        f = rng.uniform(100.0, 4000.0, (pool, 1))
        self.pcm = (0.5 * np.sin(2 * np.pi * f * t) + 0.1 * rng.uniform(-1, 1, (pool, loop))).astype(np.float32)
        self.pos = positions(rng, n)
        self.vol = rng.uniform(0.25, 1.0, n).astype(np.float32)
        self.buf_of = rng.integers(0, pool, n).astype(np.int32)
        self.start = rng.integers(0, loop, n)
        v = rng.uniform(-30, 30, (n, 2))
        self.delta = (343.0 - v[:, 0]) / (343.0 - v[:, 1])
        self.mixf = rng.uniform(0.5, 1.0, n).astype(np.float32)
        self.rates = (rng.choice([44100, 24000, 16000, 48000], n, p=[0.5, 0.2, 0.15, 0.15]) if feeder == "resample" else np.full(n, 48000)).astype(np.int32)
        self.seed = seed
        self.hrir = (load_sh_hrir()[:, :taps] if taps else load_sh_hrir()) if kind == KIND_HRTF else None

    def oracle(self, double_bus=False):
        return OracleBackend(self.F, sh_hrir=self.hrir, double_bus=double_bus)

    def build(self, be, sel):
        sel = np.asarray(sel)
        bufs = np.array([be.buffer_create(p) for p in self.pcm], np.int32)
        create = be.sources_create if hasattr(be, "sources_create") else be.source_create
        first = create(int(self.kind), len(sel))
        ids = np.arange(first, first + len(sel), dtype=np.int32)
        if self.feeder == "doppler":
            be.sources_attach_mix1(ids, bufs[self.buf_of[sel]], self.start[sel].astype(np.float64), self.delta[sel], self.mixf[sel])
        elif self.feeder == "c5":
            ring = sel < self.voices
            for sid in ids[ring]:
                be.source_attach_ring(int(sid), 48000)
            be.sources_attach_buffer(ids[~ring], bufs[self.buf_of[sel[~ring]]], self.start[sel[~ring]], self.rates[sel[~ring]])
        else:
            be.sources_attach_buffer(ids, bufs[self.buf_of[sel]], self.start[sel], self.rates[sel])
        be.set_source_positions(ids, self.pos[sel])
        be.set_source_volumes(ids, self.vol[sel])
        return ids

    def step(self, be, ids, sel, b):
        """Events before buffer b: the listener moves and turns (gain ramps on every source), every 8th source of the scene is
        switched off for buffers 1-2 and back on (the -60 dB rules and the mixer-processor sweep), voices receive their packets."""
        t = b * self.F / 48000.0
        be.set_head([1.0 * t, 0.0, 0.0], quat_yaw(np.deg2rad(30.0) * t))
        if self.volume_toggle and b in (1, 3):
            m = (np.asarray(sel) % 8) == 3
            if m.any():
                be.set_source_volumes(ids[m], np.zeros(int(m.sum()), np.float32) if b == 1 else self.vol[np.asarray(sel)[m]])
        if self.feeder == "c5":
            ring = np.asarray(sel) < self.voices
            if ring.any():
                n_pk = int(np.ceil(self.F / 480.0)) + (1 if b == 0 else 0)
                vids, packets = [], []
                for sid, s in zip(ids[ring], np.asarray(sel)[ring]):
                    if b == 3 and s % 5 == 0:
                        continue                       # a dropped packet: ring underflow, zero-padded (AudioEngine.cpp:466-473)
                    prng = np.random.default_rng(self.seed * 7919 + int(s) * 64 + b)
                    vids.append(int(sid)); packets.append((0.3 * prng.uniform(-1, 1, n_pk * 480)).astype(np.float32))
                if hasattr(be, "sources_ring_push"):
                    be.sources_ring_push(vids, packets)
                else:
                    for sid, pk in zip(vids, packets):
                        be.source_ring_push(sid, pk)

    def run(self, be, sel=None, n_buffers=None):
        sel = np.arange(self.n) if sel is None else np.asarray(sel)
        ids = self.build(be, sel)
        outs, valids = [], []
        for b in range(n_buffers or self.nb):
            self.step(be, ids, sel, b)
            v, o = be.mix()
            outs.append(o.copy()); valids.append(v)
        return np.stack(outs), valids

    def run_oracle_partitioned(self, parts, double_bus=False, n_buffers=None):
        """The oracle on `parts` disjoint subsets of the scene in as many threads; returns the f64 sum of the partitions'
        outputs (and the per-buffer valid flag = any partition valid).  The subsets are contiguous BLOCKS of the scene: the only
        state the sources of an engine share -- "did the bus mixer receive any input in the previous buffer" (GainMixer's
        sweep, RA/dsp/gain_mixer.cc:80-96) -- must be what it is in the whole scene, i.e. no partition may consist solely of
        the sources that step() switches off together (every 8th: a stride-16 split would build two such partitions, whose
        mixers then skip a sweep the full scene performs -- measured 4e-4 rel RMS of pure harness artefact)."""
        import threading
        res = [None] * parts
        blocks = np.array_split(np.arange(self.n), parts)

        def work(p):
            res[p] = self.run(self.oracle(double_bus), blocks[p], n_buffers)

        th = [threading.Thread(target=work, args=(p,)) for p in range(parts)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        out = np.sum([r[0].astype(np.float64) for r in res], axis=0)
        valid = [int(any(r[1][b] for r in res)) for b in range(len(res[0][1]))]
        return out, valid


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def full_scene_parity(name, make_engine, parts=None, double_bus=None, **override):
    """rel RMS of an engine (make_engine(scene) -> backend with the MixEngine interface) against the oracle on one of
    FULL_SCENES.  Scenes up to ~1.3e5 sources are also compared with the single f32-bus oracle summing in id order (the
    reference's arithmetic); larger ones only with the partitioned fp64-bus oracle (at 10^6 sources the f32 sum's
    own order noise exceeds the 1e-5 bar)."""
    kw = dict(FULL_SCENES[name]); kw.update(override)
    sc = BigScene(**kw)
    parts = parts or min(host_cores(), 16)
    go, gv = sc.run(make_engine(sc))
    res = {"scene": name, "sources": sc.n, "frames": sc.F, "buffers": sc.nb}
    if double_bus is None:
        double_bus = sc.n > 200000
    oo, ov = sc.run_oracle_partitioned(parts, double_bus=double_bus)
    res["oracle"] = "%d partitions x %s bus, outputs summed in f64" % (parts, "fp64" if double_bus else "f32")
    res["rel_rms"] = rel_rms(go, oo)
    res["valid_equal"] = bool(gv == ov)
    return res
