#!/usr/bin/env python
"""bench.py — the raw->display hot path on N B200s (one process per GPU).

metric: MP/s of the default darkroom graph (i-raw -> denoise -> hilite -> demosaic -> crop -> colour -> filmcurv ->
llap -> grade -> o-pfm, bin/default-darkroom.i-raw) on synthetic 61 MP Bayer stills (BASELINE.json configs[1]).
a "step" is one full pass of the graph over one still.  stills are independent units: with N > 1 every rank develops
its own stills (weak scaling, no data-path collective; torch.distributed is used for the barrier and max-over-ranks only).

  value      whole-job MP/s, input mosaic already resident in HBM, result left in HBM (kernel path only, CUDA events)
  e2e        same metric through the reference-facing C-ABI graph call with HOST buffers: pinned H2D of the u16 mosaic and
             D2H of the rgb f32 sink image (the PFM payload) inside the timed region
  roofline   dominant kernel: unique bytes in+out of the launch / its average CUDA-event duration vs measured HBM copy peak
  cpu_baseline / --impl reference: the CPU restatement of the reference's algorithm (oracle, OpenMP, all host cores) on the
             same full frame.  the reference's own Vulkan pipeline cannot be built or run here (no loader, ICD, glslang:
             profiles/r02_vulkan_probe.txt).
  fast       the kernel path again with the fast build of the kernels (SFU exp / pow, fused multiply-adds: round 1's
             arithmetic).  the headline runs the strict build, which is bit compatible with the restatement.
  bands      (N > 1, or --workload still201 --gpus N) BASELINE.json config 5: ONE 201 MP still split into bands over the N GPUs
             of the box with halo exchange over NVLink peer access, driven by rank 0 (strong scaling of one frame).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (width, height, source module, packed bpp)
    "still61": (9504, 6336, "i-raw", 0),
    "still24": (6000, 4000, "i-raw", 0),
    "mlv4k": (4096, 2160, "i-mlv", 14),
    "xtrans26": (6240, 4152, "i-raw", 0),     # BASELINE config 3 (x-trans: filters 9)
    "still201": (16384, 12288, "i-raw", 0),   # BASELINE config 5 on ONE GPU (the band split is not built)
}
WB = (2.0, 1.0, 1.5)
CAM = (0.8, 0.15, 0.05, 0.1, 0.85, 0.05, 0.02, 0.18, 0.8)


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed region (B200_PROFILING.md): NVML polled every 5 ms in this process
    (the timed region of a 61 MP run is ~150 ms, `nvidia-smi -lms` would see it once or twice); nvidia-smi as a fallback."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.sm, self.mx, self.reasons, self.stop_flag, self.proc, self.src = gpu, [], [], set(), False, None, "nvml"

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)))
            while not self.stop_flag:
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                try:
                    bits = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    bits = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in self.REASONS.items():
                    if bits & bit:
                        self.reasons.add(name)
                time.sleep(0.005)
            return
        except Exception:
            self.src = "nvidia-smi"
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                r = [x.strip() for x in line.split(",")]
                if r and r[0].replace(".", "").isdigit():
                    self.sm.append(float(r[0]))
                if len(r) > 1 and r[1].replace(".", "").isdigit():
                    self.mx.append(float(r[1]))
                for i in range(4):
                    if len(r) >= 6 and r[2 + i].lower().startswith("active"):
                        self.reasons.add(names[i])
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        self.join(timeout=1.0)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.src}


def oracle_cfg(O, w, h, strength):
    d = O.darkroom_defaults(w, h)
    for k in range(3):
        d.whitebalance[k] = WB[k]
    for k in range(9):
        d.cam_to_rec2020[k] = CAM[k]
    d.denoise.strength = strength
    d.noise_a, d.noise_b = 100.0, 2.0
    return d


def workload_name(name, W, H, mp, strength):
    """one string for both arms (the driver compares config.workload of the two lines)."""
    return "%s: %dx%d %s 14-bit still (%.1f MP), default darkroom graph incl. hilite + llap + grade, denoise strength %.2f" % (
        name, W, H, "x-trans" if name == "xtrans26" else "bayer rggb", mp, strength)


def cpu_reference_rate(sample_wh, steps, warmup, strength, xtrans=False):
    """times the oracle (CPU port of the reference's algorithm, OpenMP over all host cores) on one frame of that size."""
    from oracle import oracle_py as O
    from vkdt_b200 import synth
    w, h = sample_wh
    raw = synth.mosaic(w, h, seed=0x5EED0000, xtrans=xtrans)
    d = oracle_cfg(O, w, h, strength)
    if xtrans:
        d.filters = 9
    ow, oh = O.darkroom_out_size(d)
    threads = O.set_threads(len(os.sched_getaffinity(0)))   # all host cores this process may use, whatever OMP_NUM_THREADS says
    for _ in range(warmup):
        O.darkroom_run(d, raw)
    t0 = time.time()
    for _ in range(steps):
        O.darkroom_run(d, raw)
    dt = (time.time() - t0) / max(1, steps)
    return (w * h) / dt / 1e6, dt, threads



def ncu_traffic(label, key="bytes_per_launch"):
    """dram bytes per launch (or SM issue-slot utilisation) of this kernel from the committed ncu --set full capture
    (profiles/rNN_traffic.json), or None."""
    import glob
    files = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r*_traffic.json")))
    if not files:
        return None
    try:
        tr = json.load(open(files[-1]))[key]
    except (OSError, ValueError, KeyError):
        return None
    for k, v in tr.items():
        if k in label:
            return int(v) if key == "bytes_per_launch" else round(float(v), 1)
    return None


def run_workload(api, synth, torch, dist, args, rank, world, local_rank, W, H, src, bpp, strength, steps, warmup, sample_clocks=True,
                 mode=0, e2e=True, sink8=False, gloo=None):
    """both legs (kernel path with HBM-resident input, end to end with host buffers) of one workload on this rank.
    mode: api.MODE_STRICT (the headline) or api.MODE_FAST; e2e=False: kernel leg only."""
    # ---- synthetic input: a few distinct stills per rank, cycled (deterministic seeds per SURVEY §8d) ----
    xtrans = args.workload == "xtrans26" and src == "i-raw"
    big = W * H > 100e6
    nstills = 1 if big else 2
    if big:   # a 4x4 tiling of a 12.6 MP synthetic still: same statistics, generated in seconds
        raws = [np.ascontiguousarray(np.tile(synth.mosaic(W // 4, H // 4, seed=0x5EED0000 + rank * 1000 + i, wb=WB), (4, 4))) for i in range(nstills)]
    else:
        raws = [synth.mosaic(W, H, seed=0x5EED0000 + rank * 1000 + i, wb=WB, xtrans=xtrans) for i in range(nstills)]
    if bpp:
        payload = [synth.pack_bits_fast14(r) for r in raws]
        in_bytes = (W * H * bpp + 7) // 8
    else:
        payload = raws
        in_bytes = W * H * 2
    rp = api.raw_params(W, H, wb=WB, cam_to_rec2020=CAM, noise_a=100.0, noise_b=2.0, packed_bpp=bpp,
                        **({"filters": 9} if xtrans else {}))
    # pinned host staging for the e2e leg, device copies for the kernel leg
    host_in = []
    for p in payload:
        hp = api.host_alloc(p.nbytes + 64)
        C = api.C
        C.memmove(hp, p.ctypes.data, p.nbytes)
        host_in.append(hp)
    dev_in = []
    for p in payload:
        dp = api.dev_alloc(p.nbytes + 256)
        api.check(api.lib.vkb_memcpy_h2d(dp, p.ctypes.data, p.nbytes, None))
        dev_in.append(dp)
    api.check(api.lib.vkb_stream_sync(None))

    def make_graph():
        if sink8:   # the reference's default export: o-jpg behind colenc (sRGB primaries, rec709 curve), packed 8 bit rgb
            g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src=src), sink="o-jpg", prim=1, trc=1)
            g.set_sink_layout(api.SINK_RGB_UI8)
        else:
            g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src=src))
            g.set_sink_layout(api.SINK_RGB_F32)   # the PFM payload (r g b f32, 12 B/px) is what o-pfm puts into the file
        g.set_device(local_rank)
        if strength > 0:
            g.line("param:denoise:01:strength:%g" % strength)
        g.set_mode(mode)
        return g

    # ---- leg 1: kernel path, input resident in HBM, sink left in HBM ----
    g = make_graph()
    g.set_source(dev_in[0], rp, device=True)
    g.set_sink_buffer(None, 0)
    g.run()                              # builds the plan, allocates the pool
    ow, oh = g.sink_size()
    out_bytes = ow * oh * (3 if sink8 else 12)
    stream = g.stream()
    FR = api.RUN_RECORD
    for i in range(warmup):
        g.set_source(dev_in[i % nstills], rp, device=True)
        g.run(FR | api.RUN_UPLOAD)
    api.check(api.lib.vkb_stream_sync(api.C.c_void_p(stream)))
    sampler = ClockSampler(local_rank)
    if sample_clocks:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    api.lib.vkb_launch_count_reset()
    e0, e1 = api.Event(), api.Event()
    per_kernel = {}
    e0.record(stream)
    for i in range(steps):
        g.set_source(dev_in[i % nstills], rp, device=True)
        g.run(FR | api.RUN_UPLOAD | ((api.RUN_WAIT | api.RUN_PERF) if i == steps - 1 or i % 4 == 3 else 0))
        if i == steps - 1 or i % 4 == 3:      # per-launch events are valid after a synchronised run
            for label, ms, nbytes in g.perf_entries():
                a = per_kernel.setdefault(label, [0.0, 0, nbytes])
                a[0] += ms; a[1] += 1
    e1.record(stream)
    e1.sync()
    torch.cuda.synchronize()
    launches = api.launch_count()
    t_kernel_ms = e0.elapsed_ms(e1)
    if world > 1:
        t = torch.tensor([t_kernel_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_kernel_ms = float(t.item())
    clocks = sampler.stop() if sample_clocks else None
    pool_bytes = g.pool_bytes()
    if not e2e:
        for hp in host_in:
            api.host_free(hp)
        for dp in dev_in:
            api.dev_free(dp)
        g.close()
        return dict(t_kernel_ms=t_kernel_ms, launches=launches, per_kernel=per_kernel, clocks=clocks, in_bytes=in_bytes, out_bytes=out_bytes,
                    ow=ow, oh=oh, pool_bytes=pool_bytes, nstills=nstills)

    # ---- leg 2: end to end through the C-ABI with host buffers ----
    # two graph instances (each with its own pool, stream and pinned sink) ping-pong: while one frame's 722 MB result
    # drains over PCIe the next frame uploads and computes.  every frame still pays its full H2D + kernels + D2H.
    # stills: the 722 MB download bounds the frame and two instances keep the copy engine busy (three measured slower);
    # 4K video frames: compute and download are of similar length, a third instance absorbs the jitter
    NG = max(2, int(os.environ.get('VKB_E2E_GRAPHS', '2' if out_bytes > 400e6 else '3')))
    gs, host_out = [], []
    for k in range(NG):
        gk = make_graph()
        ho = api.host_alloc(out_bytes)
        gk.set_source(host_in[0], rp)
        gk.set_sink_buffer(ho, out_bytes)
        gk.run()
        gs.append(gk); host_out.append(ho)
    FE = api.RUN_RECORD | api.RUN_UPLOAD | api.RUN_DOWNLOAD
    for i in range(max(2, warmup)):
        gs[i % NG].set_source(host_in[i % nstills], rp)
        gs[i % NG].run(FE | api.RUN_WAIT)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e2e_steps = max(4, min(steps, 200))   # the K steps of the contract (the pipeline's fill and drain are inside the timed region)
    # the job: a sequence of world * e2e_steps frames (stills), frame f on rank f mod world (vkdt_b200/shard.py), every rank
    # keeps a small record per frame it developed, rank 0 gathers them in frame order afterwards: no data-path collective
    from vkdt_b200 import shard
    my_frames = shard.frames_for_rank(world * e2e_steps, rank, world)
    records = {}

    def landed(i):      # frame my_frames[i] is in host memory: note its first bytes
        records[my_frames[i]] = bytes((api.C.c_uint8 * 16).from_address(host_out[i % NG]))

    f0, f1 = api.Event(), api.Event()
    t0 = time.time()
    f0.record(gs[0].stream())
    for i in range(e2e_steps):
        gk = gs[i % NG]
        if i >= NG:
            gk.run(api.RUN_WAIT)          # frame i-NG has fully landed in host memory before its buffers are reused
            landed(i - NG)
        # (prefetching the next source with an upload-only run, executor.cpp up_stream, measured slower here: 14.9 against
        # 14.2 ms per 61 MP frame; the two instances interleave best when each frame's upload sits in front of its launches)
        gk.set_source(host_in[my_frames[i] % nstills], rp)
        gk.run(FE)
    for k in range(NG):
        gs[(e2e_steps + k) % NG].run(api.RUN_WAIT)
        if e2e_steps - NG + k >= 0:
            landed(e2e_steps - NG + k)
    t_wall_ms = (time.time() - t0) * 1e3
    f1.record(gs[0].stream())
    f1.sync()
    t_e2e_ms = max(f0.elapsed_ms(f1), t_wall_ms)
    if world > 1:
        t = torch.tensor([t_e2e_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e_ms = float(t.item())
    checksum = float(np.frombuffer((api.C.c_uint8 * 16).from_address(host_out[(e2e_steps - 1) % NG]), dtype=np.uint8 if sink8 else np.float32).sum())
    gathered = shard.gather_in_frame_order(records, world * e2e_steps, rank, world, group=gloo)   # rank 0: every frame, in order
    frames_gathered = len(gathered) if gathered is not None else 0

    for hp in host_in + host_out:
        api.host_free(hp)
    for dp in dev_in:
        api.dev_free(dp)
    for gk in gs + [g]:
        gk.close()
    return dict(t_kernel_ms=t_kernel_ms, t_e2e_ms=t_e2e_ms, e2e_steps=e2e_steps, launches=launches, per_kernel=per_kernel, clocks=clocks,
                in_bytes=in_bytes, out_bytes=out_bytes, ow=ow, oh=oh, pool_bytes=pool_bytes, checksum=checksum, nstills=nstills, frames_gathered=frames_gathered)



def run_in_flight(api, synth, local_rank, W, H, src, bpp, strength, steps, mode, n):
    """kernel path of n graph instances (own pool, own stream) fed round robin without waiting: what frames in flight buy on one
    GPU when a frame's pyramid tails leave most SMs idle (video).  device resident input; wall clock around stream waits."""
    import time
    raw = synth.mosaic(W, H, seed=0x5EED0000, wb=WB)
    buf = synth.pack_bits_fast14(raw) if bpp else raw
    rp = api.raw_params(W, H, wb=WB, cam_to_rec2020=CAM, noise_a=100.0, noise_b=2.0, packed_bpp=bpp)
    gs, keep = [], []
    for k in range(n):
        d = api.dev_alloc(buf.nbytes + 256)
        api.check(api.lib.vkb_memcpy_h2d(d, buf.ctypes.data, buf.nbytes, None))
        api.check(api.lib.vkb_stream_sync(None))
        keep.append(d)
        g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src=src))
        g.set_sink_layout(api.SINK_RGB_F32)
        g.set_device(local_rank)
        if strength > 0:
            g.line("param:denoise:01:strength:%g" % strength)
        g.set_mode(mode)
        g.set_source(d, rp, device=True)
        g.set_sink_buffer(None, 0)
        g.run()
        gs.append(g)
    best = 1e30
    for rep in range(3):
        for g in gs:
            g.run(api.RUN_WAIT)
        t0 = time.perf_counter()
        for i in range(steps):
            gs[i % n].run(api.RUN_RECORD | api.RUN_UPLOAD)
        for g in gs:
            g.run(api.RUN_WAIT)
        if rep:
            best = min(best, (time.perf_counter() - t0) / steps * 1e3)
    for g in gs:
        g.close()
    for d in keep:
        api.dev_free(d)
    return best


def run_bands(api, synth, devices, W, H, strength, steps, warmup, with_single=True):
    """BASELINE.json config 5: ONE still split into horizontal bands over `devices` (vkb_graph_set_bands), halo rows pulled
    between the devices' pools over NVLink peer access.  kernel leg: every device holds its source rows in its pool (uploaded
    once), the sink stays in HBM; end to end: every device uploads its source rows from and downloads its sink rows into one
    pinned host frame, inside the timed region.  times are device side: the slowest device's span (vkb_graph_band_elapsed_ms)."""
    n = len(devices)
    if W * H > 100e6:   # a 4x4 tiling of a 12.6 MP synthetic still: same statistics, generated in seconds
        raw = np.ascontiguousarray(np.tile(synth.mosaic(W // 4, H // 4, seed=0x5EED0000, wb=WB), (4, 4)))
    else:
        raw = synth.mosaic(W, H, seed=0x5EED0000, wb=WB)
    rp = api.raw_params(W, H, wb=WB, cam_to_rec2020=CAM, noise_a=100.0, noise_b=2.0)
    host_in = api.host_alloc(raw.nbytes + 64)
    api.C.memmove(host_in, raw.ctypes.data, raw.nbytes)

    def graph(devs):
        g = api.Graph(cfg_text=api.DARKROOM_CFG.format(src="i-raw"))
        g.line("param:denoise:01:strength:%g" % strength)
        g.set_sink_layout(api.SINK_RGB_F32)
        if len(devs) > 1:
            g.set_bands(devs)
        else:
            g.set_device(devs[0])
        g.set_source(host_in, rp)
        g.set_sink_buffer(None, 0)
        g.run()
        return g

    g = graph(devices)
    ow, oh = g.sink_size()
    out_bytes = ow * oh * 12
    FR = api.RUN_RECORD
    for _ in range(max(1, warmup)):
        g.run(FR)
    g.run(api.RUN_WAIT)
    api.lib.vkb_launch_count_reset()
    g.band_mark(0)
    for _ in range(steps):
        g.run(FR)
    g.band_mark(1)
    t_kernel = g.band_elapsed_ms()
    launches = api.launch_count()
    stats = g.band_stats()
    pool = g.pool_bytes()
    # end to end: host frame in, host frame out
    host_out = api.host_alloc(out_bytes)
    g.set_sink_buffer(host_out, out_bytes)
    FE = api.RUN_RECORD | api.RUN_UPLOAD | api.RUN_DOWNLOAD | api.RUN_WAIT
    g.run(FE)
    e2e_steps = max(3, min(steps, 8))
    t0 = time.time()
    for _ in range(e2e_steps):
        g.run(FE)
    t_e2e = (time.time() - t0) * 1e3
    checksum = float(np.frombuffer((api.C.c_float * 4).from_address(host_out), dtype=np.float32).sum())
    g.close()
    single = None
    if with_single:   # the same frame on ONE GPU, for the strong scaling efficiency
        g1 = graph(devices[:1])
        e0, e1 = api.Event(), api.Event()
        for _ in range(max(1, warmup)):
            g1.run(FR)
        g1.run(api.RUN_WAIT)
        st = g1.stream()
        e0.record(st)
        k1 = max(2, min(steps, 5))
        for _ in range(k1):
            g1.run(FR)
        e1.record(st)
        e1.sync()
        single = e0.elapsed_ms(e1) / k1
        g1.close()
    api.host_free(host_in)
    api.host_free(host_out)
    mp = W * H / 1e6
    out = {"workload": "still%d: ONE %dx%d bayer still (%.0f MP) split into %d horizontal bands, one per GPU, default darkroom graph incl. denoise 0.40 + hilite + llap + grade; "
                       "halo rows pulled over NVLink peer access, result bit identical to the one GPU frame (tests/test_bands_gpu.py)" % (round(mp), W, H, mp, n),
           "n_gpus": n, "scaling": "strong", "value": round(steps * mp / (t_kernel * 1e-3), 1), "unit": "MP/s", "ms_per_frame": round(t_kernel / steps, 3), "steps": steps,
           "e2e": {"value": round(e2e_steps * mp / (t_e2e * 1e-3), 1), "unit": "MP/s", "ms_per_frame": round(t_e2e / e2e_steps, 3), "steps": e2e_steps,
                   "h2d_bytes_per_step": int(raw.nbytes), "d2h_bytes_per_step": int(out_bytes), "timing": "wall clock around synchronous frames", "checksum": checksum},
           "nvlink_bytes_per_frame": stats["bytes_total"], "nvlink_bytes_busiest_gpu": stats["bytes_max_device"], "peer_copies_per_frame": stats["pulls"],
           "kernel_launches_per_frame": stats["launches"], "gpu_launches": int(launches), "pool_bytes_per_gpu": pool}
    if single:
        out["one_gpu_ms_per_frame"] = round(single, 3)
        out["speedup_vs_one_gpu"] = round(single / (t_kernel / steps), 3)
        out["efficiency"] = round(single / (t_kernel / steps) / n, 3)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="still61", choices=sorted(WORKLOADS))
    ap.add_argument("--denoise", type=float, default=None, help="denoise:strength (default: 0.4 once the wavelet kernels are built, else 0)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mlv", action="store_true", help="skip the extra MLV 4K frames/s measurement")
    ap.add_argument("--no-bands", action="store_true", help="N > 1: skip the 201 MP band split leg")
    ap.add_argument("--mode", default="strict", choices=["strict", "fast"],
                    help="strict (default): kernels bit compatible with the CPU restatement; fast: SFU transcendentals + fma")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W, H, src, bpp = WORKLOADS[args.workload]
    mp = W * H / 1e6

    if args.impl == "reference":
        # the reference's own CPU-side implementation of the path = the oracle port (vkdt-cli needs vulkan + glslang: absent,
        # profiles/r02_vulkan_probe.txt).  the FULL frame of the workload, bounded in the number of passes (one pass is ~15 s)
        if rank != 0:
            return 0
        strength = args.denoise if args.denoise is not None else 0.4
        steps_run, warm_run = max(1, min(args.steps, 2)), max(0, min(args.warmup, 1))
        rate, dt, cores = cpu_reference_rate((W, H), steps_run, warm_run, strength, xtrans=args.workload == "xtrans26")
        print(json.dumps({
            "impl": "reference", "metric": "MP/s raw->display graph", "value": round(rate, 3), "unit": "MP/s", "n_gpus": args.gpus,
            "steps": steps_run, "warmup": warm_run, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, W, H, mp, strength)},
            "cpu_baseline": {"value": round(rate, 3), "unit": "MP/s", "cores": cores, "kind": "port",
                             "sample": "the full %dx%d frame, %d timed pass(es) of the CPU oracle (OpenMP, %d threads) after %d warm-up; requested steps %d / warmup %d are bounded to keep the run within minutes" % (
                                 W, H, steps_run, cores, warm_run, args.steps, args.warmup)},
            "e2e": {"value": round(rate, 3), "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    import torch
    import torch.distributed as dist
    from vkdt_b200 import api, synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (vkdt_b200 has no CPU path)")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    # one process per GPU: run on the CPUs next to that GPU, so that the pinned staging buffers (first touch) and the
    # copy engine's traffic stay on the GPU's NUMA node.  the CPU baseline leg restores the full mask.
    full_affinity = os.sched_getaffinity(0)
    numa = "all cpus"
    try:
        import pynvml
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local_rank)
        try:    # CUDA_VISIBLE_DEVICES can renumber: go through the PCI address
            hdl = pynvml.nvmlDeviceGetHandleByPciBusId(("%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)).encode())
        except Exception:
            hdl = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        pynvml.nvmlDeviceSetCpuAffinity(hdl)
        numa = "%d cpus next to gpu %d" % (len(os.sched_getaffinity(0)), local_rank)
    except Exception as e:  # not fatal: only placement
        numa = "unpinned (%s)" % type(e).__name__
    api.init(local_rank)
    have_wavelet = ("denoise", "doub") in api.kernels()
    strength = args.denoise if args.denoise is not None else (0.4 if have_wavelet else 0.0)

    mode = api.MODE_FAST if args.mode == "fast" else api.MODE_STRICT
    gloo = dist.new_group(backend="gloo") if world > 1 else None
    band_devices = list(range(world if world > 1 else args.gpus))
    if os.environ.get("VKB_BENCH_BAND_DEVICES"):   # e.g. "0,0": bands sharing one GPU, to exercise the leg on a one GPU box
        band_devices = [int(x) for x in os.environ["VKB_BENCH_BAND_DEVICES"].split(",")]
    if args.workload == "still201" and len(band_devices) > 1:
        # the band split IS the workload: rank 0 drives all the GPUs of the box, the other ranks stand by
        if rank == 0:
            api.set_mode(mode)
            Bd = run_bands(api, synth, band_devices, W, H, strength, args.steps, args.warmup)
            line = {"metric": "MP/s raw->display graph", "value": Bd["value"], "unit": "MP/s", "n_gpus": len(band_devices), "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": Bd["ms_per_frame"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                    "dtype": "f32", "data": "synthetic", "config": {"workload": Bd["workload"], "mode": args.mode,
                    "timing": "pool of %.1f GB per GPU, far beyond the 126 MB L2" % (Bd["pool_bytes_per_gpu"] / 1e9)},
                    "e2e": Bd["e2e"], "gpu_launches": Bd["gpu_launches"], "bands": Bd}
            print(json.dumps(line))
        if world > 1:
            dist.barrier(group=gloo)
            dist.destroy_process_group()
        return 0

    mlv_W, mlv_H, mlv_src, mlv_bpp = WORKLOADS["mlv4k"]
    R = run_workload(api, synth, torch, dist, args, rank, world, local_rank, W, H, src, bpp, strength, args.steps, args.warmup, mode=mode, gloo=gloo)
    # the other build of the kernels, kernel path only, for the record
    other = api.MODE_STRICT if mode == api.MODE_FAST else api.MODE_FAST
    F = run_workload(api, synth, torch, dist, args, rank, world, local_rank, W, H, src, bpp, strength, max(4, min(args.steps, 10)), 3,
                     sample_clocks=False, mode=other, e2e=False)
    M = None
    if args.workload != "mlv4k" and not args.no_mlv:
        M = run_workload(api, synth, torch, dist, args, rank, world, local_rank, mlv_W, mlv_H, mlv_src, mlv_bpp, strength,
                         max(40, min(10 * args.steps, 200)), 3, sample_clocks=False, mode=mode, gloo=gloo)
    M8 = S8 = None
    if not args.no_mlv:
        if args.workload != "mlv4k":
            M8 = run_workload(api, synth, torch, dist, args, rank, world, local_rank, mlv_W, mlv_H, mlv_src, mlv_bpp, strength,
                              max(40, min(10 * args.steps, 200)), 3, sample_clocks=False, mode=mode, sink8=True, gloo=gloo)
        S8 = run_workload(api, synth, torch, dist, args, rank, world, local_rank, W, H, src, bpp, strength, max(4, min(args.steps, 10)), 3,
                          sample_clocks=False, mode=mode, sink8=True, gloo=gloo)
    Bd = None
    if world > 1 and not args.no_bands:
        # config 5 rides along in the scaling runs: one 201 MP still over all the GPUs of the box, driven by rank 0
        torch.cuda.empty_cache()
        dist.barrier(group=gloo)
        if rank == 0:
            try:
                api.set_mode(mode)
                bw, bh, _, _ = WORKLOADS["still201"]
                Bd = run_bands(api, synth, band_devices, bw, bh, strength, max(5, min(args.steps, 10)), 3)
            except Exception as e:   # report, do not lose the headline line
                Bd = {"error": str(e)}
        dist.barrier(group=gloo)
    t_kernel_ms, t_e2e_ms, e2e_steps, launches, per_kernel, clocks = R["t_kernel_ms"], R["t_e2e_ms"], R["e2e_steps"], R["launches"], R["per_kernel"], R["clocks"]
    in_bytes, out_bytes, ow, oh, checksum, nstills = R["in_bytes"], R["out_bytes"], R["ow"], R["oh"], R["checksum"], R["nstills"]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel ----
    pk, pk_kind = peaks()
    top = max(per_kernel.items(), key=lambda kv: kv[1][0])
    top_label, (top_ms_sum, top_n, top_bytes) = top
    top_ms = top_ms_sum / top_n
    achieved = top_bytes / (top_ms * 1e-3) / 1e9
    total_ms = sum(v[0] / v[1] for v in per_kernel.values())
    graph_alg_bytes = in_bytes + out_bytes
    roof = {"bound": "hbm", "kernel": top_label, "achieved": round(achieved, 1), "peak": pk["hbm_gbs"], "peak_kind": pk_kind + " copy bandwidth",
            "unit": "GB/s", "frac": round(achieved / pk["hbm_gbs"], 4), "traffic": ncu_traffic(top_label),
            "sm_issue_pct": ncu_traffic(top_label, "sm_issue_pct"),
            "note": "instruction bound, not HBM bound: see sm_issue_pct (ncu smsp__issue_active of this kernel) and DESIGN.md section 3; the strict build spends its "
                    "issue slots on libm-exact exp / pow in fp64 (profiles/r02_summary.md)",
            "algorithmic_bytes_per_launch": top_bytes, "avg_launch_ms": round(top_ms, 4),
            "share_of_step": round(top_ms / total_ms, 4),
            "graph": {"algorithmic_bytes_per_step": graph_alg_bytes, "achieved_gbs": round(graph_alg_bytes / (t_kernel_ms / args.steps * 1e-3) / 1e9, 1),
                      "frac": round(graph_alg_bytes / (t_kernel_ms / args.steps * 1e-3) / 1e9 / pk["hbm_gbs"], 4)},
            "kernels": {k: {"ms": round(v[0] / v[1], 4), "gbs": round(v[2] / (v[0] / v[1] * 1e-3) / 1e9, 1)} for k, v in
                        sorted(per_kernel.items(), key=lambda kv: -kv[1][0])[:12]}}
    line = {
        "metric": "MP/s raw->display graph", "value": round(world * args.steps * mp / (t_kernel_ms * 1e-3), 2), "unit": "MP/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(t_kernel_ms / args.steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, W, H, mp, strength), "sink": "rgb f32 %dx%d (PFM payload, 12 B/px)" % (ow, oh),
                   "mode": args.mode + (": kernels bit compatible with the CPU restatement (libm-exact transcendentals, no fma), output identical to the oracle at full size"
                                        if args.mode == "strict" else ": SFU transcendentals + fma (round 1's arithmetic)"),
                   "timing": "inputs and intermediates (%.0f MB pool) exceed the 126 MB L2; %d distinct stills cycled" % (R["pool_bytes"] / 1e6, nstills),
                   "parallelism": "independent stills per GPU, no collective"},
        "e2e": {"value": round(world * e2e_steps * mp / (t_e2e_ms * 1e-3), 2), "unit": "MP/s", "h2d_bytes_per_step": in_bytes,
                "d2h_bytes_per_step": out_bytes, "ms_per_step": round(t_e2e_ms / e2e_steps, 3), "steps": e2e_steps, "checksum": checksum,
                "frames_gathered_in_order": R["frames_gathered"]},
        "gpu_launches": int(launches) * world, "launches_per_step": int(launches // max(1, args.steps)),  # every rank launches the same sequence
        "clocks": clocks, "roofline": roof, "pool_bytes": R["pool_bytes"],
    }
    other_name = "strict" if args.mode == "fast" else "fast"
    fsteps = max(4, min(args.steps, 10))
    ftop_label, (ftop_sum, ftop_n, ftop_bytes) = max(F["per_kernel"].items(), key=lambda kv: kv[1][0])
    ftop_ms = ftop_sum / ftop_n
    line[other_name] = {"what": "the same kernel path with the %s build of the kernels" % other_name,
                        "value": round(world * fsteps * mp / (F["t_kernel_ms"] * 1e-3), 2), "unit": "MP/s", "ms_per_step": round(F["t_kernel_ms"] / fsteps, 4),
                        "roofline": {"bound": "hbm", "kernel": ftop_label, "achieved": round(ftop_bytes / (ftop_ms * 1e-3) / 1e9, 1), "peak": pk["hbm_gbs"], "unit": "GB/s",
                                     "frac": round(ftop_bytes / (ftop_ms * 1e-3) / 1e9 / pk["hbm_gbs"], 4), "algorithmic_bytes_per_launch": ftop_bytes, "avg_launch_ms": round(ftop_ms, 4)},
                        "kernels": {k: round(v[0] / v[1], 4) for k, v in sorted(F["per_kernel"].items(), key=lambda kv: -kv[1][0])[:8]}}
    # BASELINE.json north_star: ">= 60 % of the HBM roofline on the fused pointwise chain" (crop + colour + filmcurv in one launch:
    # 8 B/px in, 8 B/px out): both builds, live CUDA events of this run
    def pointwise_roof(per_kernel):
        for k, v in per_kernel.items():
            if "b200_pointw" in k and v[1] > 0:
                ms = v[0] / v[1]
                gbs = v[2] / (ms * 1e-3) / 1e9
                return {"kernel": k, "avg_launch_ms": round(ms, 4), "algorithmic_bytes_per_launch": v[2], "achieved": round(gbs, 1), "unit": "GB/s",
                        "peak": pk["hbm_gbs"], "frac": round(gbs / pk["hbm_gbs"], 4)}
        return None
    line["pointwise_chain"] = {("fast" if args.mode == "fast" else "strict"): pointwise_roof(per_kernel), other_name: pointwise_roof(F["per_kernel"]),
                               "target": "frac >= 0.6 (north_star); the strict build pays for libm-exact pow / exp in its tone curve, the fast build is the HBM-bound form"}
    if Bd:
        line["bands"] = Bd
    if M:
        msteps = max(40, min(10 * args.steps, 200))
        line["mlv4k"] = {"workload": "MLV 4096x2160 14-bit packed frames, default darkroom graph, frame f -> rank f mod N",
                         "frames_per_s": round(world * msteps / (M["t_kernel_ms"] * 1e-3), 1), "ms_per_frame": round(M["t_kernel_ms"] / msteps, 3),
                         "e2e_frames_per_s": round(world * M["e2e_steps"] / (M["t_e2e_ms"] * 1e-3), 1),
                         "h2d_bytes_per_frame": M["in_bytes"], "d2h_bytes_per_frame": M["out_bytes"], "n_gpus": world}
    if M and world == 1:
        # for the record: the reference's default-darkroom.i-mlv leaves denoise at strength 0 (a noop), and a clip's frames are
        # independent: two instances in flight fill the SMs that one frame's pyramid tails leave idle
        v = {}
        for key, st, n in (("in_flight_2", strength, 2), ("denoise_off", 0.0, 1), ("denoise_off_in_flight_2", 0.0, 2)):
            ms = run_in_flight(api, synth, local_rank, mlv_W, mlv_H, mlv_src, mlv_bpp, st, 80, mode, n)
            v[key] = {"frames_per_s": round(1e3 / ms, 1), "ms_per_frame": round(ms, 3)}
        line["mlv4k"]["variants"] = v
    if M8:
        msteps = max(40, min(10 * args.steps, 200))
        line["mlv4k"]["sink_8bit"] = {"what": "the reference's default export (o-jpg: colenc sRGB / rec709 curve, packed rgb ui8, 3 B/px) instead of the f32 PFM payload",
                                      "frames_per_s": round(world * msteps / (M8["t_kernel_ms"] * 1e-3), 1), "e2e_frames_per_s": round(world * M8["e2e_steps"] / (M8["t_e2e_ms"] * 1e-3), 1),
                                      "d2h_bytes_per_frame": M8["out_bytes"]}
    if S8:
        s8steps = max(4, min(args.steps, 10))
        line["e2e_8bit"] = {"what": "the same stills through the reference's default export (o-jpg: colenc + packed rgb ui8 sink, 3 B/px)",
                            "value": round(world * S8["e2e_steps"] * mp / (S8["t_e2e_ms"] * 1e-3), 2), "unit": "MP/s", "d2h_bytes_per_step": S8["out_bytes"],
                            "kernel_path_value": round(world * s8steps * mp / (S8["t_kernel_ms"] * 1e-3), 2)}
    line["config"]["host_placement"] = numa
    if not args.no_cpu_baseline:
        os.sched_setaffinity(0, full_affinity)
        big = W * H > 100e6
        sample = (W // 4, H // 4) if big else (W, H)
        rate, dt, cores = cpu_reference_rate(sample, 1, 0, strength, xtrans=args.workload == "xtrans26")
        line["cpu_baseline"] = {"value": round(rate, 3), "unit": "MP/s", "cores": cores, "kind": "port",
                                "sample": "%s %dx%d frame of the workload, 1 pass of the CPU oracle (OpenMP, %d threads), %.1f s" % (
                                    "a quarter size" if big else "the full", sample[0], sample[1], cores, dt)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
