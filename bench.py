#!/usr/bin/env python
"""bench.py -- source-samples/s of the spatial-audio mix hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N --steps K --warmup W] [--workload NAME]      # N > 1: launched by torchrun, one rank per GPU
  python bench.py --impl reference [...]                                  # the reference's CPU implementation, same metric

One "step" = one output buffer: every active source's window gathered, resampled / pitch-shifted, attenuated, panned
(or ambisonic-encoded + SH-HRIR decoded) and summed to interleaved stereo.  Default workload: the configuration
BASELINE.json's target is quoted on -- 1 048 576 concurrent doppler-resampled, distance-attenuated, panned sources,
512-frame buffers at 48 kHz, on ONE B200 (config 4's source set, which fits a single GPU: 64 GiB of private 16 384-sample
loops); with --gpus N every rank mixes its own 1 048 576 sources (weak scaling) and the 19 x 512 bus is summed across
ranks inside the kernels over NVLink peer memory (K3's last CTA pushes, k_peer_gather sums; --exchange allreduce: one NCCL
all-reduce on the engine stream before the finish kernel, the comparison arm).

Timed legs (each W warm-up steps, then BLOCKS blocks of exactly K timed steps; every block is bracketed by CUDA events on
the engine stream with barrier + synchronize on both sides and reduced with MAX over ranks; the line reports the MEDIAN block,
all block times are listed under "timing"):
  value : inputs resident in HBM; per step the listener pose is set and one buffer is mixed.
  e2e   : through the C ABI with HOST buffers: per step 4096 source positions for ids drawn AT RANDOM over all sources + the
          listener pose come from host arrays (sbm_set_source_positions / sbm_set_head -> lock-free setter queue -> one pinned
          upload block + scatter kernel inside sbm_mix) and the interleaved stereo buffer is copied back to host memory.
  e2e.all_move : the same with EVERY source's position sent every step (16 B per source of H2D).
With --gpus N the line also carries "strong_scaling": BASELINE config 4 as written -- 1 048 576 sources IN TOTAL sharded over
the N ranks --, "parity": the sharded mix through the peer exchange against the oracle (all ranks bit-identical), and
"per_rank_kernels_ms": every rank's own K1 / K2 / tail (a buffer ends with all ranks waiting for the slowest one).
"e2e.kernels_ms": the kernel times of the e2e leg (its movers ramp: K2 differs from the resident leg's).
Per-buffer PCM windows of the default workload total 2.2 GB >> 126 MB L2, so every step streams from HBM (no L2 flush needed;
config says so); workloads whose per-step windows fit in L2 are marked L2-resident and carry no HBM roofline fraction.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "source-samples/sec mixed @48kHz/512-frame buffers"
UNIT = "source-samples/s"
STATE_BYTES = {"pan": 88, "hq": 200}     # per-source state read+written per buffer (+ 8 B with room effects: the send's gain)

WORKLOADS = {
    # name: sources per GPU, frames, source kind, feeder, loop length, SH-HRIR taps, description
    "c4_1m_doppler_pan": dict(n=1048576, frames=512, kind="pan", feeder="doppler", loop=16384, taps=0,
                              desc="1 048 576 sources/GPU, doppler resample (delta in [0.92,1.10]) + distance atten + stereo pan, 512-frame @48kHz"),
    "c4_1m_doppler_pan_room": dict(n=1048576, frames=512, kind="pan", feeder="doppler", loop=16384, taps=0, room=True,
                                   desc="1 048 576 sources/GPU, doppler resample + atten + pan + ROOM EFFECTS (reflections send per source, Substrata's parcel settings), 512-frame @48kHz"),
    "256k_hq_room": dict(n=262144, frames=512, kind="hq", feeder="native", loop=16384, taps=256, room=True,
                         desc="262 144 sources/GPU, 3rd-order ambisonic encode + 256-tap SH-HRIR decode + ROOM EFFECTS, 512-frame @48kHz"),
    "c2_4096_resample_pan": dict(n=4096, frames=512, kind="pan", feeder="resample", loop=16384, taps=0,
                                 desc="4 096 sources, linear resample (44.1k/24k/16k/48k mix) + atten + pan, 512-frame @48kHz"),
    "c3_65536_hq_128tap": dict(n=65536, frames=1024, kind="hq", feeder="native", loop=16384, taps=128,
                               desc="65 536 sources, 3rd-order ambisonic encode + 128-tap SH-HRIR decode per ear, 1 024-frame @48kHz"),
    "c5_voice_ambient": dict(n=33792, frames=256, kind="hq", feeder="c5", loop=16384, taps=256, voices=1024,
                             desc="1 024 streamed voice sources (rings fed 480-sample packets, batched push) + 32 768 ambient loops, ambisonic encode + 256-tap SH-HRIR decode, 256-frame @48kHz"),
    "c1_16_pan": dict(n=16, frames=512, kind="pan", feeder="native", loop=48000, taps=0,
                      desc="16 mono looping sources, distance atten + stereo pan, 512-frame @48kHz"),
    "1m_native_pan": dict(n=1048576, frames=512, kind="pan", feeder="native", loop=16384, taps=0,
                          desc="1 048 576 native-rate looping sources/GPU, atten + pan, 512-frame @48kHz"),
    "256k_hq": dict(n=262144, frames=512, kind="hq", feeder="native", loop=16384, taps=256,
                    desc="262 144 sources/GPU, 3rd-order ambisonic encode + 256-tap SH-HRIR decode, 512-frame @48kHz"),
}
MOVERS_PER_STEP = 4096


def synth_params(n, loop, seed, feeder):
    """Per-source parameters of the synthetic workload: positions uniform in a 80 x 20 x 80 m box around
    the listener (rejecting d < 1.05 m), volume U[0.25,1], random read offsets, doppler factors from v in U[-30,30] m/s."""
    rng = np.random.default_rng(seed)
    pos = rng.uniform([-40, -10, -40], [40, 10, 40], (n, 3)).astype(np.float32)
    near = np.linalg.norm(pos, axis=1) < 1.05
    pos[near] += np.float32(2.0)
    vol = rng.uniform(0.25, 1.0, n).astype(np.float32)
    start = rng.integers(0, loop, n)
    v = rng.uniform(-30, 30, (n, 2))
    delta = (343.0 - v[:, 0]) / (343.0 - v[:, 1])          # AudioEngine.cpp:140-145
    mixf = rng.uniform(0.5, 1.0, n).astype(np.float32)
    rates = rng.choice([44100, 24000, 16000, 48000], n, p=[0.5, 0.2, 0.15, 0.15]) if feeder == "resample" else np.full(n, 48000)
    return dict(pos=pos, vol=vol, start=start, delta=delta, mixf=mixf, rates=rates.astype(np.int32))


def listener_pose(step, frames):
    """Listener translating 1 m/s along +x and yawing 30 deg/s, updated once per buffer."""
    t = step * frames / 48000.0
    yaw = np.deg2rad(30.0) * t
    return np.array([t, 0.0, 0.0], np.float32), np.array([0.0, np.sin(yaw / 2), 0.0, np.cos(yaw / 2)], np.float32)


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md): NVML polled every 2 ms from a thread
    (a timed leg lasts tens of milliseconds, shorter than one nvidia-smi -lms period); nvidia-smi is the fallback."""
    REASONS = [("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4),
               ("hw_power_brake", 0x80)]

    def __init__(self, gpu_index):
        self.idx, self.sm, self.mask, self.mx, self.stop_flag, self.t, self.err = gpu_index, [], 0, None, False, None, None

    def _visible_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.idx])
            except Exception:
                return self.idx
        return self.idx

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._visible_index())
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception as ex:
            self.err = "nvml: %s" % ex
            self.t = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception as ex:
                self.err = "nvml: %s" % ex
                break
            time.sleep(0.002)

    def stop(self):
        self.stop_flag = True
        if self.t:
            self.t.join(timeout=1)
        if not self.sm:
            try:
                q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
                    "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
                r = subprocess.run(["nvidia-smi", "-i", str(self._visible_index()), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=10).stdout.strip().split(",")
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                return {"sm_mhz": float(r[0]), "sm_max_mhz": float(r[1]), "reasons": [n for n, v in zip(names, r[2:]) if v.strip().lower().startswith("active")],
                        "samples": 1, "how": "one nvidia-smi query after the timed legs (%s)" % self.err}
            except Exception as ex:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable: %s / %s" % (self.err, ex)], "samples": 0}
        return {"sm_mhz": statistics.median(self.sm), "sm_min_mhz": min(self.sm), "sm_max_mhz": self.mx,
                "reasons": [n for n, bit in self.REASONS if self.mask & bit], "samples": len(self.sm), "how": "NVML polled every 2 ms across both timed legs"}


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation (oracle/_ref: vendored Resonance Audio + AudioResampler compiled in place,
# restated feeder) when it was built, else the oracle port.  Independent engines on disjoint source subsets, one per
# host thread (the reference graph itself is single-threaded: RA/api/resonance_audio_api.h:139-141).
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_run(wl, threads, sources_per_thread, buffers, warmup, target_seconds=0.0):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import harness
    kind = "reference" if harness.ref_library_path() else "port"
    F = wl["frames"]
    hr = harness.load_sh_hrir()
    if wl["taps"]:
        hr = hr[:, :wl["taps"]]
    k = harness.KIND_PAN if wl["kind"] == "pan" else harness.KIND_HRTF
    feeder = "native" if wl["feeder"] == "c5" else wl["feeder"]      # C5 on the CPU: its ambient loops (32 768 of the 33 792 sources)

    def make():
        if kind == "reference":
            return harness.RefBackend(F, hrir_taps=wl["taps"])
        return harness.OracleBackend(F, sh_hrir=hr if wl["kind"] == "hq" else None)

    engines, scen = [], []
    for t in range(threads):
        sc = harness.Scenario(n_sources=sources_per_thread, frames=F, seed=1000 + t, kind=k, feeder=feeder, loop_len=min(wl["loop"], 16384),
                              n_shared_buffers=min(sources_per_thread, 256))
        be = make()
        sc.build(be)
        engines.append(be); scen.append(sc)

    def work(t, b0, nb, out):
        t0 = time.perf_counter()
        for b in range(b0, b0 + nb):
            scen[t].step(engines[t], b)
            engines[t].mix()
        out[t] = time.perf_counter() - t0

    def run(b0, nb):
        out = [0.0] * threads
        th = [threading.Thread(target=work, args=(t, b0, nb, out)) for t in range(threads)]
        t0 = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        return time.perf_counter() - t0

    run(0, warmup)
    done = warmup
    if target_seconds:
        # bounded sample: probe a few buffers, then size the timed run for about `target_seconds` of wall time
        probe = run(done, 4)
        done += 4
        buffers = int(min(20000, max(8, target_seconds / max(probe / 4, 1e-6))))
    wall = run(done, buffers)
    samples = threads * sources_per_thread * F * buffers
    return dict(value=samples / wall, unit=UNIT, cores=threads, kind=kind,
                sample="%d threads x %d sources x %d buffers of %d frames (%s), independent engines (the reference graph is single-threaded), wall %.2f s" %
                       (threads, sources_per_thread, buffers, F, feeder + "/" + wl["kind"], wall),
                ms_per_step=1e3 * wall / buffers)


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args, wl, rank, world):
    if rank != 0:
        return
    threads = min(host_threads(), 64)
    per = 1024 if wl["n"] >= 1024 * threads else max(1, wl["n"] // threads)
    if wl["n"] < threads:
        threads, per = 1, wl["n"]
    r = cpu_reference_run(wl, threads, per, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload,
                       "description": "CPU reference arm: %d independent engines x %d sources = %d sources in total (NOT the GPU arm's %d per GPU: a bounded sample of the same "
                                      "workload type -- %s; throughput is per source-sample, so the two arms compare as rates)" %
                                      (threads, per, threads * per, wl["n"], wl["desc"]),
                       "sources_total": threads * per, "engines": threads, "frames": wl["frames"], "sample_rate": 48000},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def build_engine(wl, device, seed):
    import torch
    import substrata_b200 as sb
    n, F, loop = wl["n"], wl["frames"], wl["loop"]
    stride = loop + 16
    # synthetic PCM, generated on the device in chunks: 0.5 sin(2 pi f t) + 0.1 noise, f in U[100,4000] Hz
    pcm = torch.empty((n, stride), dtype=torch.float32, device="cuda:%d" % device)
    g = torch.Generator(device="cuda:%d" % device)
    g.manual_seed(seed)
    t = torch.arange(stride, device=pcm.device, dtype=torch.float32) / 48000.0
    chunk = 16384
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        f = torch.rand((hi - lo, 1), generator=g, device=pcm.device) * 3900.0 + 100.0
        pcm[lo:hi] = 0.5 * torch.sin(2 * np.pi * f * t) + 0.1 * (2 * torch.rand((hi - lo, stride), generator=g, device=pcm.device) - 1)
    torch.cuda.synchronize(device)
    hr = sb.load_sh_hrir(wl["taps"] or None) if wl["kind"] == "hq" else None
    eng = sb.MixEngine(F, sh_hrir=hr, max_sources=n, device=device)
    first_buf = eng.buffer_adopt_device(pcm.data_ptr(), loop, stride, n)
    p = synth_params(n, loop, seed, wl["feeder"])
    first = eng.source_create(sb.KIND_PAN if wl["kind"] == "pan" else sb.KIND_HRTF, n)
    ids = np.arange(first, first + n, dtype=np.int32)
    bufs = np.arange(first_buf, first_buf + n, dtype=np.int32)
    if wl["feeder"] == "doppler":
        eng.sources_attach_mix1(ids, bufs, p["start"].astype(np.float64), p["delta"], p["mixf"])
    elif wl["feeder"] == "c5":
        nv = wl["voices"]
        for sid in ids[:nv]:
            eng.source_attach_ring(int(sid), 48000, capacity=wl.get("ring_capacity", 96000))
        eng.sources_attach_buffer(ids[nv:], bufs[nv:], p["start"][nv:], p["rates"][nv:])
    else:
        eng.sources_attach_buffer(ids, bufs, p["start"], p["rates"])
    eng.set_source_positions(ids, p["pos"])
    eng.set_source_volumes(ids, p["vol"])
    if wl.get("room"):
        # inside a parcel (GUIClient.cpp:6886-6896 -> AudioEngine.cpp:746-768): the parcel's box, Substrata's wall coefficients and gain
        eng.set_room_hrir(sb.load_sh_hrir_order1())
        eng.enable_room_effects(True)
        eng.set_reflection_properties([0.0, 0.0, 0.0], [0, 0, 0, 1], [90.0, 24.0, 90.0], [0.8, 0.8, 0.3, 0.8, 0.8, 0.2], 0.0, 0.7)
    return eng, pcm, ids, p


def parity_sample(wl):
    """Relative RMS error of the CUDA path against the CPU oracle on a bounded sample of the same workload type."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import harness
    import substrata_b200 as sb
    F = wl["frames"]
    hr = harness.load_sh_hrir()
    hr = hr[:, :wl["taps"]] if wl["taps"] else hr
    k = harness.KIND_PAN if wl["kind"] == "pan" else harness.KIND_HRTF
    n = min(wl["n"], 2048)
    sc = harness.Scenario(n_sources=n, frames=F, seed=4242, kind=k, feeder="ring" if wl["feeder"] == "c5" else wl["feeder"], loop_len=4096,
                          batched_ring_push=True)
    oo, _ = sc.run(harness.OracleBackend(F, sh_hrir=hr), 4)
    go, _ = sc.run(sb.MixEngine(F, sh_hrir=hr if wl["kind"] == "hq" else None, max_sources=n), 4)
    return {"rel_rms": harness.rel_rms(go, oo), "sample": "%d sources x 4 buffers vs oracle (CPU port pinned to the compiled reference)" % n}


def parity_full_size(device):
    """rel RMS against the oracle at BASELINE.json's OWN sizes, one entry per configuration (tests/harness.py FULL_SCENES; the
    same runs are `-m gpu` tests in tests/test_gpu_fullsize.py): C2 4 096, C3 65 536 x 1024 frames, C4 one rank slice of 131 072
    (single f32-bus oracle AND fp64-bus oracle), C4 all 1 048 576 (fp64-bus oracle), C5 1 024 rings + 32 768 loops."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import harness
    import substrata_b200 as sb

    def make(sc):
        return sb.MixEngine(sc.F, sh_hrir=sc.hrir, max_sources=sc.n, device=device)

    out = {}
    for name in ("c2", "c3", "c5", "c4_1m"):
        t0 = time.perf_counter()
        try:
            r = harness.full_scene_parity(name, make)
            r["seconds"] = round(time.perf_counter() - t0, 1)
            out[name] = r
        except Exception as ex:
            out[name] = {"error": str(ex)[:300]}
    try:
        t0 = time.perf_counter()
        sc = harness.BigScene(**harness.FULL_SCENES["c4_rank"])
        go, gv = sc.run(make(sc))
        o32, v32 = sc.run(sc.oracle())
        o64, v64 = sc.run_oracle_partitioned(min(harness.host_cores(), 16), double_bus=True)
        out["c4_rank"] = {"scene": "c4_rank", "sources": sc.n, "frames": sc.F, "buffers": sc.nb, "rel_rms_vs_f32_bus_oracle": harness.rel_rms(go, o32),
                          "rel_rms_vs_fp64_bus_oracle": harness.rel_rms(go, o64), "f32_vs_fp64_oracle": harness.rel_rms(o32, o64),
                          "rel_rms": max(harness.rel_rms(go, o32), harness.rel_rms(go, o64)), "valid_equal": bool(gv == v32 == v64),
                          "oracle": "one engine, f32 bus, id order (the reference's arithmetic) + partitioned fp64 bus", "seconds": round(time.perf_counter() - t0, 1)}
    except Exception as ex:
        out["c4_rank"] = {"error": str(ex)[:300]}
    vals = [v["rel_rms"] for v in out.values() if "rel_rms" in v]
    out["worst_rel_rms"] = max(vals) if vals else None
    out["bar"] = 1e-5
    return out


def parity_multi_gpu(rank, world, local_rank, per_rank=16384):
    """The N > 1 data path against the oracle, inside the bench line: a doppler + pan scene of per_rank x world sources is
    sharded like the bench workload (one engine per rank, buses exchanged over NVLink peer memory inside K3 / the gather
    kernel); rank 0 compares its output with the oracle's render of the WHOLE scene, and every rank's output must be
    bit-identical to rank 0's."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import harness
    import torch
    import torch.distributed as dist
    import substrata_b200 as sb
    from substrata_b200.sharding import ShardedMixer
    kw = dict(harness.FULL_SCENES["c4_rank"]); kw.update(n=per_rank * world, pool=256, seed=977)
    sc = harness.BigScene(**kw)
    sel = np.arange(rank, sc.n, world)
    eng = sb.MixEngine(sc.F, max_sources=len(sel), device=local_rank)
    ids = sc.build(eng, sel)
    mixer = ShardedMixer(eng, device="cuda:%d" % local_rank, exchange="peer")
    outs, valids = [], []
    for b in range(sc.nb):
        sc.step(eng, ids, sel, b)
        v, o = mixer.mix()
        outs.append(o.copy()); valids.append(v)
    got = np.stack(outs)
    t = torch.from_numpy(got.copy()).cuda(local_rank)
    ref = t.clone()
    dist.broadcast(ref, 0)
    same = torch.tensor([1 if torch.equal(t, ref) else 0], device=t.device)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    res = None
    if rank == 0:
        oo, ov = sc.run_oracle_partitioned(min(harness.host_cores(), 16))
        res = {"sources_total": sc.n, "sources_per_rank": per_rank, "frames": sc.F, "buffers": sc.nb, "rel_rms": harness.rel_rms(got, oo),
               "valid_equal": bool(valids == ov), "all_ranks_bit_identical": bool(int(same.item()) == 1),
               "what": "sharded over %d ranks through the in-kernel peer exchange vs the oracle on the whole scene" % world}
    dist.barrier()
    eng.close()
    return res


ALSO = [("c2_4096_resample_pan", 200), ("c3_65536_hq_128tap", 100), ("c5_voice_ambient", 100), ("256k_hq", 50)]


def also_workloads(args):
    """The other single-GPU configurations of BASELINE.json (C2, C3, C5) and the 262 144-source ambisonic workload, each
    measured by a fresh `bench.py --workload ... --no-cpu` process on the same GPU right after the headline run; only the
    figures, the full lines come from running those commands directly."""
    out = {}
    for name, steps in ALSO:
        if name == args.workload:
            continue
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", name, "--steps", str(steps), "--warmup", "5", "--no-cpu",
                                "--pipeline", str(args.pipeline)], capture_output=True, text=True, timeout=300)
            d = json.loads(r.stdout.strip().splitlines()[-1])
            out[name] = {"value": d["value"], "unit": d["unit"], "ms_per_step": d["ms_per_step"], "e2e_value": d["e2e"]["value"],
                         "e2e_ms_per_step": d["e2e"]["ms_per_step"], "k2_roofline_frac": d["roofline"]["frac"], "k2_ms": d["roofline"]["kernel_ms"],
                         "l2_resident": d["roofline"].get("l2_resident"), "latency_us": d.get("latency_us"), "clocks": d.get("clocks"),
                         "realtime_budget_ms": d["config"]["realtime_budget_ms"], "description": d["config"]["description"]}
        except Exception as ex:
            out[name] = {"error": str(ex)[:200]}
    return out


L2_BYTES = 126e6


def run_gpu_arm(args, wl, rank, world, local_rank):
    import torch
    import substrata_b200 as sb
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng, pcm, ids, p = build_engine(wl, local_rank, 12345 + rank)
    F, n = wl["frames"], wl["n"]
    from substrata_b200.sharding import ShardedMixer
    mixer = ShardedMixer(eng, device="cuda:%d" % local_rank, exchange=args.exchange)
    out_host = np.zeros(2 * F, np.float32)
    eng.set_profiling(True)
    rng = np.random.default_rng(99 + rank)

    def sync_all():
        torch.cuda.synchronize(local_rank)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(local_rank)

    # C5: voice rings.  Packets of 480 samples per voice every 10 ms (ClientUDPHandlerThread.cpp:196), all voices in one
    # sbm_source_ring_push_batch call.  value leg: rings pre-filled before the timed region (inputs resident);
    # e2e leg: the packets cross PCIe inside the timed region, as they would behind a network thread.
    voices = wl.get("voices", 0)
    if voices:
        v_ids = np.ascontiguousarray(ids[:voices])
        v_offs = (np.arange(voices + 1, dtype=np.uint64) * np.uint64(480))
        v_payload = []
        for _ in range(4):      # packet payloads assembled in page-locked host memory, as a network thread feeding the engine would
            buf = sb.pinned_array(voices * 480)
            buf[:] = (0.3 * rng.uniform(-1, 1, voices * 480)).astype(np.float32)
            v_payload.append(buf)
        ring_state = {"pushed": 0, "consumed": 0}

        def feed_voices(upto_samples):
            while ring_state["pushed"] < upto_samples:
                eng.sources_ring_push_flat(v_ids, v_payload[(ring_state["pushed"] // 480) % 4], v_offs)
                ring_state["pushed"] += 480

    def prep_resident(s):
        hp, hq = listener_pose(s, F)
        eng.set_head(hp, hq)

    # e2e: the movers of a step are drawn at random over ALL ids (what a game produces: scattered slots, not one block)
    movers = min(MOVERS_PER_STEP, n)
    mv_ids = [np.ascontiguousarray(np.sort(rng.choice(ids, movers, replace=False)).astype(np.int32)) for _ in range(8)]
    mv_pos = [rng.uniform([-40, -10, -40], [40, 10, 40], (movers, 3)).astype(np.float32) + np.float32(1.5) for _ in range(8)]
    all_pos = [rng.uniform([-40, -10, -40], [40, 10, 40], (n, 3)).astype(np.float32) + np.float32(1.5) for _ in range(2)]

    def prep_e2e(s):
        hp, hq = listener_pose(s, F)
        if voices:
            feed_voices(ring_state["consumed"] + F)       # the packets that arrived since the previous buffer
        eng.set_source_positions(mv_ids[s % 8], mv_pos[s % 8])
        eng.set_head(hp, hq)

    def prep_all_move(s):
        hp, hq = listener_pose(s, F)
        if voices:
            feed_voices(ring_state["consumed"] + F)
        eng.set_source_positions(ids, all_pos[s % 2])
        eng.set_head(hp, hq)

    pipelined = args.pipeline >= 2 and (args.exchange == "peer" or world == 1)

    def timed(mx, en, prep_fn, s0, resident, steps, blocks):
        """W warm-up buffers, then `blocks` blocks of exactly `steps` timed buffers (CUDA events on the engine stream, barrier +
        synchronize on both sides of every block, MAX over ranks per block); the MEDIAN block is reported.  prep_fn(s) = the
        host-side events of buffer s (pose / position / packet updates); the buffer itself goes through sbm_mix, or --
        pipelined -- sbm_mix_submit with the previous buffer collected afterwards (two buffers in flight)."""
        stream = mx.stream
        total = args.warmup + steps * blocks
        if voices and resident:
            feed_voices(ring_state["consumed"] + total * F)      # resident inputs: outside the timed region
        s = s0
        for _ in range(args.warmup):
            prep_fn(s); s += 1
            mx.mix(out_host)
            if voices:
                ring_state["consumed"] += F
        acc = dict(k1=0.0, k2=0.0, k3=0.0, algo=0, rendered=0, h2d=0, d2h=0)
        l0 = en.stats()["kernel_launches"]

        def account():
            st = en.stats()
            acc["k1"] += st["last_params_kernel_ms"]; acc["k2"] += st["last_mix_kernel_ms"]; acc["k3"] += st["last_finish_kernel_ms"]
            acc["algo"] += st["last_algorithmic_bytes"]; acc["rendered"] += st["last_rendered_sources"]
            acc["h2d"] += st["last_h2d_bytes"]; acc["d2h"] += st["last_d2h_bytes"]

        block_ms = []
        for _ in range(blocks):
            sync_all()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for k in range(steps):
                prep_fn(s); s += 1
                if pipelined:
                    mx.submit()
                    if k >= 1:
                        mx.collect(out_host); account()
                else:
                    mx.mix(out_host); account()
                if voices:
                    ring_state["consumed"] += F
            if pipelined:
                mx.collect(out_host); account()
            e1.record(stream)
            sync_all()
            block_ms.append(e0.elapsed_time(e1))
        if world > 1:
            tt = torch.tensor(block_ms, device="cuda:%d" % local_rank)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            block_ms = [float(x) for x in tt.tolist()]
        K = steps * blocks
        med = statistics.median(block_ms)
        per_rank = None
        if world > 1:
            # every rank's own K1 / K2 / tail: the buffer ends with all ranks waiting for the slowest one's K3 (the exchange), so the
            # step is the SLOWEST rank's K1 + K2 -- and the faster ranks' wait shows up in their "tail"
            mine = torch.tensor([acc["k1"] / K, acc["k2"] / K, acc["k3"] / K], device="cuda:%d" % local_rank, dtype=torch.float64)
            allr = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine)
            per_rank = {"k1_params": [round(float(t[0]), 5) for t in allr], "k2_mix": [round(float(t[1]), 5) for t in allr],
                        "k3_reduce+exchange(wait)+k4_finish": [round(float(t[2]), 5) for t in allr]}
        return dict(ms=med / steps, per_rank=per_rank, block_ms_per_step=[round(x / steps, 5) for x in block_ms], k1=acc["k1"] / K, k2=acc["k2"] / K, k3=acc["k3"] / K,
                    algo=acc["algo"] / K, rendered=acc["rendered"] / K, h2d=acc["h2d"] / K, d2h=acc["d2h"] / K,
                    launches=(en.stats()["kernel_launches"] - l0) * steps // max(1, args.warmup + K), next_step=s)

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    B = args.blocks
    res = timed(mixer, eng, prep_resident, 0, True, args.steps, B)
    e2e = timed(mixer, eng, prep_e2e, res["next_step"], False, args.steps, B)
    clk = clocks.stop() if rank == 0 else None
    allmove = timed(mixer, eng, prep_all_move, e2e["next_step"], False, min(args.steps, 10), 1)

    # per-buffer latency of the synchronous call (small, latency-bound configurations: C1 / C2 / C5)
    latency = None
    if world == 1 and n <= 65536:
        lat = []
        s = allmove["next_step"]
        for k in range(200):
            prep_e2e(s + k)
            t0 = time.perf_counter()
            eng.mix(out_host)
            lat.append((time.perf_counter() - t0) * 1e6)
            if voices:
                ring_state["consumed"] += F
        lat.sort()
        latency = {"p50": round(lat[len(lat) // 2], 1), "p99": round(lat[int(len(lat) * 0.99)], 1), "max": round(lat[-1], 1), "buffers": len(lat),
                   "what": "wall clock of one synchronous sbm_mix through the C ABI (setters applied, upload block, K1-K4, result in host memory), us"}

    # BASELINE config 4 as written: 1 048 576 sources IN TOTAL, sharded over the ranks (strong scaling)
    strong = None
    if wl["feeder"] == "doppler" and args.strong and world > 1:
        n_s = wl["n"] // world
        eng_s = sb.MixEngine(F, max_sources=n_s, device=local_rank)
        stride = wl["loop"] + 16
        fb = eng_s.buffer_adopt_device(pcm.data_ptr(), wl["loop"], stride, n_s)
        first = eng_s.source_create(sb.KIND_PAN, n_s)
        ids_s = np.arange(first, first + n_s, dtype=np.int32)
        eng_s.sources_attach_mix1(ids_s, np.arange(fb, fb + n_s, dtype=np.int32), p["start"][:n_s].astype(np.float64), p["delta"][:n_s], p["mixf"][:n_s])
        eng_s.set_source_positions(ids_s, p["pos"][:n_s]); eng_s.set_source_volumes(ids_s, p["vol"][:n_s])
        eng_s.set_profiling(True)
        mixer_s = ShardedMixer(eng_s, device="cuda:%d" % local_rank, exchange=args.exchange)

        def prep_s(s):
            hp, hq = listener_pose(s, F)
            eng_s.set_head(hp, hq)

        mv_s = [np.ascontiguousarray(np.sort(rng.choice(ids_s, min(movers, n_s), replace=False)).astype(np.int32)) for _ in range(8)]

        def prep_s_e2e(s):
            hp, hq = listener_pose(s, F)
            eng_s.set_source_positions(mv_s[s % 8], mv_pos[s % 8][:len(mv_s[s % 8])])
            eng_s.set_head(hp, hq)

        rs = timed(mixer_s, eng_s, prep_s, 0, True, args.steps, B)
        rs_e = timed(mixer_s, eng_s, prep_s_e2e, rs["next_step"], False, args.steps, B)
        tot = n_s * world
        strong = {"scaling": "strong", "sources_total": tot, "sources_per_gpu": n_s, "value": tot * F / (rs["ms"] * 1e-3), "unit": UNIT, "ms_per_step": rs["ms"],
                  "e2e_value": tot * F / (rs_e["ms"] * 1e-3), "e2e_ms_per_step": rs_e["ms"], "block_ms_per_step": rs["block_ms_per_step"],
                  "kernels_ms": {"k1_params": rs["k1"], "k2_mix": rs["k2"], "k3_reduce+exchange+k4_finish": rs["k3"]}, "per_rank_kernels_ms": rs["per_rank"],
                  "what": "BASELINE config 4 as written: the first %d sources of every rank's set, %d in total; speed-up = this value / the 1-GPU line's value (same 1 048 576 sources)" % (n_s, tot)}
        eng_s.close()

    par_multi = None
    if world > 1 and not args.no_cpu:
        try:
            par_multi = parity_multi_gpu(rank, world, local_rank)
        except Exception as ex:
            par_multi = {"error": str(ex)[:300]}

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy, burst)"
        else:
            peak, peak_src = 6650.0, "B200_PROFILING.md fallback"
        S = STATE_BYTES[wl["kind"]]
        algo_bytes = res["algo"] + S * res["rendered"]
        achieved = algo_bytes / (res["k2"] * 1e-3) / 1e9 if res["k2"] > 0 else None
        traffic = None
        tp = os.path.join(ROOT, "profiles", "k2_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(args.workload)
        total = world * n * F
        l2_resident = res["algo"] < L2_BYTES
        l2_text = ("per-step PCM windows %.0f MB < 126 MB L2: this workload is L2-RESIDENT after the first buffer (no flush between steps); its kernel time is latency / launch bound "
                   "and is NOT quoted against the HBM roofline" % (res["algo"] / 1e6)) if l2_resident else \
                  ("per-step PCM windows %.0f MB >> 126 MB L2: inputs larger than L2, no flush" % (res["algo"] / 1e6))
        line = {
            "metric": METRIC, "value": total / (res["ms"] * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "description": wl["desc"], "sources_per_gpu": n, "frames": F, "sample_rate": 48000,
                       "pcm": "private %d-sample f32 loops resident in HBM (%.1f GiB/GPU)" % (wl["loop"], n * (wl["loop"] + 16) * 4 / 2**30),
                       "l2": l2_text,
                       "parallelism": "sources sharded, 1 rank/GPU" + ((", 18x%d bus exchanged over NVLink peer memory inside K3 / the gather kernel" % F if args.exchange == "peer"
                                                                        else ", NCCL all-reduce of the 18x%d bus" % F) if world > 1 else ""),
                       "realtime_budget_ms": 1e3 * F / 48000.0,
                       "buffers_in_flight": 2 if pipelined else 1},
            "timing": {"blocks": B, "steps_per_block": args.steps, "reported": "median block", "value_block_ms_per_step": res["block_ms_per_step"],
                       "e2e_block_ms_per_step": e2e["block_ms_per_step"]},
            "e2e": {"value": total / (e2e["ms"] * 1e-3), "unit": UNIT, "ms_per_step": e2e["ms"], "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                    "kernels_ms": {"k1_params": e2e["k1"], "k2_mix": e2e["k2"], "k3_reduce+collective+k4_finish": e2e["k3"]},
                    "what": ("%d source positions for ids drawn at random over all %d sources + listener pose from host arrays per step, stereo buffer back to host, through sbm_set_*/sbm_mix" % (movers, n)) +
                            ("; %d voice packets of 480 samples per 10 ms of audio through sbm_source_ring_push_batch" % voices if voices else ""),
                    "all_move": {"value": total / (allmove["ms"] * 1e-3), "ms_per_step": allmove["ms"], "h2d_bytes_per_step": allmove["h2d"], "d2h_bytes_per_step": allmove["d2h"],
                                 "steps": min(args.steps, 10), "what": "EVERY source's position (%d x 12 B from host arrays) sent every step" % n}},
            "gpu_launches": res["launches"],
            "roofline": {"bound": "hbm", "kernel": "k2_mix_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak if achieved else None) if not l2_resident else None, "l2_resident": l2_resident,
                         "traffic": traffic, "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": res["k2"], "peak_source": peak_src,
                         "other_kernels_ms": {"k1_params": res["k1"], "k3_reduce+collective+k4_finish": res["k3"]}, "per_rank_kernels_ms": res["per_rank"],
                         "whole_step": {"achieved": algo_bytes / (res["ms"] * 1e-3) / 1e9, "frac": (algo_bytes / (res["ms"] * 1e-3) / 1e9 / peak) if not l2_resident else None,
                                        "what": "the same algorithmic bytes over the whole buffer time (K1 + K2 + K3 + K4 + launch gaps)"}},
            "clocks": clk,
        }
        if latency:
            line["latency_us"] = latency
        if strong:
            line["strong_scaling"] = strong
        if par_multi is not None:
            line["parity"] = {"multi_gpu": par_multi}
        if world == 1 and not args.no_cpu:
            try:
                line["parity"] = parity_sample(wl)
            except Exception as ex:          # the oracle is test infrastructure; its absence must not hide the GPU numbers
                line["parity"] = {"error": str(ex)}
            if not args.no_full_parity:
                line["parity"]["full_size"] = parity_full_size(local_rank)
            threads = min(host_threads(), 64)
            line["cpu_baseline"] = {k: v for k, v in cpu_reference_run(wl, threads, min(1024, max(1, wl["n"] // threads)), 0, 2, target_seconds=12.0).items() if k != "ms_per_step"}
            if not args.no_also:
                line["also"] = also_workloads(args)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4_1m_doppler_pan", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity legs (and the secondary workloads)")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary workloads (C2 / C3 / C5 figures under \"also\")")
    ap.add_argument("--exchange", default="peer", choices=["peer", "allreduce"],
                    help="N > 1: bus exchange inside the kernels over NVLink peer memory (default) or an NCCL all-reduce between mix_begin and mix_finish")
    ap.add_argument("--pipeline", type=int, default=2, choices=[1, 2],
                    help="buffers in flight: 2 = sbm_mix_submit / sbm_mix_collect (default), 1 = synchronous sbm_mix")
    ap.add_argument("--blocks", type=int, default=5, help="timed blocks of --steps steps per leg; the median block is reported")
    ap.add_argument("--no-full-parity", action="store_true", help="skip the full-size oracle comparisons (parity.full_size)")
    ap.add_argument("--no-strong", dest="strong", action="store_false", help="skip the strong-scaling leg (1 048 576 sources in total over the ranks)")
    ap.add_argument("--sources", type=int, default=0, help="override the workload's sources per GPU (profiling under ncu only; the line says so)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = dict(WORKLOADS[args.workload])
    if args.sources:
        wl["n"] = args.sources
        wl["desc"] += " [--sources override: %d sources/GPU, NOT the named configuration]" % args.sources
    wl["ring_capacity"] = max(96000, (args.warmup + args.steps * args.blocks + 8) * wl["frames"] + 960)     # the resident leg pre-fills every timed buffer's packets
    if args.impl == "reference":
        run_reference_arm(args, wl, rank, world)
        return
    run_gpu_arm(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
