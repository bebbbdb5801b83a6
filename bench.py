#!/usr/bin/env python
"""bench.py — Mrays/s (primary + secondary) of the nrays render hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config C3] [--extra C2,C4,C5|none]

A "step" is one frame: scene::render (src/scene.rs:29-116) of BASELINE.json's headline workload,
crytek_sponza.scene at 1920x1080x4spp (config C3; the mesh is the seeded synthetic stand-in, the
real asset is not in the reference tree).  At N > 1 (torchrun, one rank per GPU) the SAME frame is
tile-sharded over the ranks -> strong scaling; the exchange is fused into the resolve kernel.

  value     whole-job Mrays/s with everything resident in HBM (nrb_render_device / tile shards)
  e2e       same metric through the host-facing C-ABI call nrb_render: camera arguments in, image
            copied back to pinned host memory inside the timed region
  roofline  the dominant kernel (trace_kernel alone: NrbStats separates the tail kernel) — algorithmic bytes /
            CUDA-event launch time vs the measured HBM peak, next to the frame-level figure of SURVEY 8d and the
            ncu-measured DRAM traffic of profiles/traffic.json (N = 1 only)
  configs   the other BASELINE.json configurations in the same run (fewer steps): C2, C4 and C5 at N = 1; C5 (the
            configuration BASELINE defines for 8 GPUs) at every N > 1
  cpu_baseline / --impl reference
            the CPU oracle (C++ f64 restatement of the reference path; the Rust reference itself cannot be built
            here) on all host cores, on the FULL frame of the same config
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mrays/s (primary+secondary) at 1920x1080; HBM GB/s vs peak"
UNIT = "Mrays/s"
SAMPLE_BANDS = 9       # cpu sample of the big configs (C4, C5): 9 bands of 12 rows spread over the frame
SAMPLE_ROWS = 12
FULL_FRAME_CPU = ("C1", "C2", "C3")   # configs whose whole frame the CPU arm renders per step (C3: ~6 s on 16 cores)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3")
    ap.add_argument("--extra", default="auto", help="other configs reported in the `configs` block: auto | none | C2,C4,...")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (test infrastructure) timed as the reference's CPU path — the one place bench.py executes oracle/
# ---------------------------------------------------------------------------------------------------------------------
_CPU_CACHE = {}


def cpu_frame(cfg, threads=0):
    """One CPU "step": the whole frame for FULL_FRAME_CPU configs, else SAMPLE_BANDS x SAMPLE_ROWS rows of it (scene + BVT
    built once, outside the timing, like Scene::new is outside render).  Returns (rays, seconds, cores, full)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_lib as O
    from nrays_b200 import configs, make_camera

    if cfg not in _CPU_CACHE:
        scene, camdesc, c = configs.build_flat(cfg)
        w, h = c["width"], c["height"]
        cam = make_camera(w, h, c["spp"], c["window"], camdesc.eye, camdesc.projection((w, h)), seed=0)
        _CPU_CACHE[cfg] = (O.OracleScene(scene.flat, 64), cam, w, h, np.zeros((w * h, 3), np.float32))
    osc, cam, w, h, out = _CPU_CACHE[cfg]
    cores = os.cpu_count() if threads <= 0 else threads
    full = cfg in FULL_FRAME_CPU
    rays, secs = 0, 0.0
    if full:
        t0 = time.perf_counter()
        _, st = osc.render(cam, cores, 0, 0, out)
        secs += time.perf_counter() - t0
        rays += st.rays_total
    else:
        for b in range(SAMPLE_BANDS):
            y0 = max(0, int((b + 0.5) * h / SAMPLE_BANDS) - SAMPLE_ROWS // 2)
            rows = min(SAMPLE_ROWS, h - y0)
            t0 = time.perf_counter()
            _, st = osc.render(cam, cores, y0 * w, rows * w, out)
            secs += time.perf_counter() - t0
            rays += st.rays_total
    return rays, secs, cores, full


def sample_label(cfg_id, cfg_name, h):
    if cfg_id in FULL_FRAME_CPU:
        return "the full frame of %s (every pixel, all spp)" % cfg_name
    return "%d bands x %d rows (%.0f %% of the frame's pixels, all spp) of %s" % (
        SAMPLE_BANDS, SAMPLE_ROWS, 100.0 * SAMPLE_BANDS * SAMPLE_ROWS / float(h), cfg_name)


def workload_config(cfg):
    """The part of `config` both arms share (the driver compares the arms' configs)."""
    return {"workload": cfg["name"], "resolution": [cfg["width"], cfg["height"]], "spp": cfg["spp"], "window": cfg["window"]}


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU path (oracle restatement; kind "port") on all host cores, same config."""
    if rank != 0:
        return
    from nrays_b200 import configs

    cfg = configs.CONFIGS[args.config]
    for _ in range(args.warmup):
        cpu_frame(args.config)
    t_total, rays_total = 0.0, 0
    cores = os.cpu_count()
    for _ in range(args.steps):
        rays, secs, cores, _full = cpu_frame(args.config)
        t_total += secs
        rays_total += rays
    value = rays_total / t_total / 1e6
    label = sample_label(args.config, cfg["name"], cfg["height"])
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(1, args.steps), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(cfg),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": label,
                         "note": "C++ f64 restatement of the reference path (oracle/), threads = os.cpu_count() like "
                                 "num_cpus::get() (src/scene.rs:49); the Rust reference cannot be built in this image "
                                 "(no rustc/cargo, ncollide3d un-vendored)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def measure_config(ctx, cfg_id, steps, warmup):
    """Device-resident and end-to-end timing of one BASELINE config on the ranks of ctx.  Returns a dict on rank 0."""
    import numpy as np
    import torch

    from nrays_b200 import _abi as A
    from nrays_b200 import _lib, configs, dist, make_camera

    lib, rank, world, local_rank, dev, td, stream = ctx.lib, ctx.rank, ctx.world, ctx.local_rank, ctx.dev, ctx.td, ctx.stream
    cfg = configs.CONFIGS[cfg_id]
    scene, camdesc, _ = configs.build(cfg_id, device=local_rank)
    w, h, spp, window = cfg["width"], cfg["height"], cfg["spp"], cfg["window"]
    proj = camdesc.projection((w, h))
    dist.use_stream(scene, stream.cuda_stream)

    out = torch.empty(h * w * 3, dtype=torch.float32, device=dev)
    tpr = dist.tiles_per_rank(w, h, world)
    packed = torch.zeros((tpr, 16, 16, 3), dtype=torch.float32, device=dev) if world > 1 else None
    # N > 1: every rank's resolve kernel stores its finished tiles straight into rank 0's image over NVLink (CUDA IPC
    # mapping); if the mapping cannot be made the ranks fall back to ONE NCCL all-gather of packed tiles + un-tile.
    peer = None
    if world > 1 and os.environ.get("NRB_BENCH_EXCHANGE", "peer") == "peer":
        try:
            peer = dist.PeerImage(w, h, rank, world, local_rank)
            if rank == 0:
                out = peer.tensor()
        except RuntimeError as e:
            if rank == 0:
                print("bench: peer image unavailable (%s), using the all-gather exchange" % e, file=sys.stderr)
            peer = None
    host_ptr = lib.nrb_host_alloc(h * w * 3 * 4)  # pinned destination of the e2e image
    pinned_t = torch.empty(h * w * 3, dtype=torch.float32).pin_memory() if world > 1 else None

    # End-to-end at N > 1: the destination of scene::render is a HOST image; every rank DMAs its tile columns straight into
    # one shared pinned host image over its own PCIe link (dist.SharedHostImage + nrb_render_tiles_to_host).  Fallback:
    # the device image on rank 0 + one device->host copy of the whole frame (with the reader-done fence of PeerImage).
    shost = None
    if world > 1 and os.environ.get("NRB_BENCH_E2E", "shared") == "shared" and w % 16 == 0 and (w // 16) % world == 0:
        try:
            shost = dist.SharedHostImage(w, h, rank, world, local_rank)
        except Exception as e:
            if rank == 0:
                print("bench: shared host image unavailable (%s), e2e copies the frame from rank 0" % e, file=sys.stderr)
            shost = None

    def cam_for(step):
        return make_camera(w, h, spp, window, camdesc.eye, proj, seed=step)

    def step_device(step):
        cam = cam_for(step)
        if world == 1:
            return dist.render_device(scene, cam, out), 0
        if peer is not None:
            st = dist.render_tiles_to_image(scene, cam, rank, world, peer.ptr)
            peer.sync()   # one 4-byte all-reduce: rank 0's next work sees every rank's pixels
            return st, 0
        st, _n = dist.render_tiles_device(scene, cam, rank, world, packed)
        g = dist.all_gather_tiles(packed, world)
        if rank == 0:
            dist.untile_device(g, world, w, h, out, device=local_rank, stream=stream.cuda_stream)
            return st, 1
        return st, 0

    def step_e2e(step):
        cam = cam_for(step)
        if world == 1:
            st = A.NrbStats()
            _lib.check(lib.nrb_render(scene.handle, C.byref(cam), C.cast(host_ptr, C.POINTER(C.c_float)), C.byref(st)))
            return st, 0
        if shost is not None:
            st = dist.render_tiles_to_host(scene, cam, rank, world, shost.addr)
            shost.sync()   # every rank's DMA has landed in the host image when this all-reduce completes
            return st, 0
        st, extra = step_device(step)
        if rank == 0:
            pinned_t.copy_(out, non_blocking=True)
        if peer is not None:
            peer.release()   # reader-done fence: nobody stores frame i+1 into the image while rank 0 still copies frame i
        if rank == 0:
            stream.synchronize()
        return st, extra

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
            torch.cuda.synchronize()

    def timed(fn, n, first_step):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        stats = []
        barrier()
        for i in range(n):
            ctx.flush.zero_()  # evict L2 between timed frames (outside the event pair)
            evs[i][0].record(stream)
            st, extra = fn(first_step + i)
            evs[i][1].record(stream)
            stats.append((st.as_dict(), extra))
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        return ms, stats

    for i in range(warmup):
        step_device(1000 + i)
    verified = None
    if world > 1:
        # untimed check of the sharded frame against this GPU's own full-frame render of the same camera
        step_device(999)
        barrier()
        ref_img = None
        if rank == 0:
            ref_img = torch.empty(h * w * 3, dtype=torch.float32, device=dev)
            dist.render_device(scene, cam_for(999), ref_img)
            verified = bool((out - ref_img).abs().max().item() < 1e-4)
        barrier()
        if shost is not None:
            # the shared host image must hold the same frame; if it does not, every rank drops back to the copy path
            ok_t = torch.ones(1, dtype=torch.int32, device=dev)
            try:
                step_e2e(999)
                barrier()
                if rank == 0 and not bool(np.abs(shost.array - ref_img.cpu().numpy()).max() < 1e-4):
                    ok_t.zero_()
            except Exception as e:
                print("bench: rank %d: shared host image path failed (%s)" % (rank, e), file=sys.stderr)
                ok_t.zero_()
            td.all_reduce(ok_t, op=td.ReduceOp.MIN)
            if int(ok_t.item()) == 0:
                shost.close()
                shost = None
            barrier()
        del ref_img
    for i in range(max(1, min(warmup, 2))):
        step_e2e(2000 + i)

    ms_dev, stats_dev = timed(step_device, steps, 0)
    ms_e2e, stats_e2e = timed(step_e2e, steps, 0)

    def reduce(ms, stats):
        rays = sum(s["rays_total"] for s, _ in stats)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        r = torch.tensor([float(rays)], dtype=torch.float64, device=dev)
        if world > 1:
            td.all_reduce(t, op=td.ReduceOp.MAX)
            td.all_reduce(r, op=td.ReduceOp.SUM)
        return float(t.item()), float(r.item())

    keys = ("rays_primary", "rays_reflect", "rays_refract", "rays_shadow", "rays_shadow_culled", "rays_tail")
    tcounts = torch.tensor([float(sum(s[k] for s, _ in stats_dev)) for k in keys], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(tcounts, op=td.ReduceOp.SUM)
    counts_all = dict(zip(keys, [float(x) for x in tcounts.tolist()]))   # whole job (all ranks)
    ms_dev_max, rays_dev = reduce(ms_dev, stats_dev)
    ms_e2e_max, rays_e2e = reduce(ms_e2e, stats_e2e)
    launches = sum(s["kernel_launches"] + extra for s, extra in stats_dev)
    lt = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(lt, op=td.ReduceOp.SUM)

    res = None
    if rank == 0:
        peak, peak_src = peaks()
        s0 = stats_dev[0][0]
        n_steps = len(stats_dev)
        geom_bytes = s0["scene_bytes"]   # nodes + leaf-ordered triangles (+ uvs, texels) as resident on the device
        # ---- (1) the dominant kernel alone: trace_kernel on THIS rank (NrbStats.ms_trace / launches_trace exclude the tail
        # kernel).  Compulsory HBM bytes: a closest-hit ray reads 32 B of queue entry and writes a 16 B hit record (wave 0
        # generates its rays on chip: 16 B only); a shadow ray reads 48 B and issues one 16 B RED; the geometry (64 B nodes
        # + 48 B triangles) is read from HBM once per frame (it stays in L2 between the frame's launches).
        n_primary_k = sum(s["rays_primary"] for s, _ in stats_dev)
        n_closest_k = sum(s["rays_primary"] + s["rays_reflect"] + s["rays_refract"] - s["rays_tail"] for s, _ in stats_dev)
        n_shadow_k = sum(s["rays_shadow"] - s["rays_shadow_culled"] for s, _ in stats_dev)
        ms_k = sum(s["ms_trace"] for s, _ in stats_dev)
        l_k = sum(s["launches_trace"] for s, _ in stats_dev)
        bytes_k = 16 * n_primary_k + 48 * (n_closest_k - n_primary_k) + 64 * n_shadow_k + n_steps * geom_bytes
        achieved = bytes_k / (ms_k * 1e-3) / 1e9 if ms_k > 0 else 0.0
        # ---- (2) frame-level figure of SURVEY 8d over the WHOLE job: 160 B per ray + 16 B per primary + the scene once per rank and frame
        n_all = (counts_all["rays_primary"] + counts_all["rays_reflect"] + counts_all["rays_refract"] + counts_all["rays_shadow"] -
                 counts_all["rays_shadow_culled"])
        frame_bytes = 160 * n_all + 16 * counts_all["rays_primary"] + world * n_steps * s0["scene_bytes"]
        # ---- (3) ncu-measured DRAM traffic of the same kernel (profiles/traffic.json, keyed by config and N; captures exist
        # for N = 1 only — a shard's launches move different bytes, so at N > 1 this is null rather than a guess)
        traffic, ncu_note = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                ent = tj.get(cfg_id, {}).get(str(world))
                if ent:
                    traffic, ncu_note = ent.get("trace_kernel"), ent.get("limiter")
            except Exception:
                traffic = None
        per_launch_ms = ms_k / max(1, l_k)
        roofline = {
            "bound": "hbm", "kernel": "trace_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
            "headline": "frac = algorithmic bytes of trace_kernel launches / their CUDA-event time / measured HBM peak; "
                        "`frame` is SURVEY 8d's whole-frame figure, `traffic_*` the ncu-measured DRAM bytes",
            "bytes_per_launch": bytes_k / max(1, l_k), "ms_per_launch": per_launch_ms, "launches_per_step": l_k / float(n_steps),
            "kernel_share_of_step": ms_k / ms_dev if ms_dev > 0 else None,
            "limiter_ncu": ncu_note,   # what actually bounds the kernel (ncu, profiles/): not HBM
            # BASELINE's "HBM GB/s vs peak": ncu DRAM bytes per launch over the live launch time
            "traffic_gbs": (traffic / (per_launch_ms * 1e-3) / 1e9) if (traffic and ms_k > 0) else None,
            "traffic_frac": (traffic / (per_launch_ms * 1e-3) / 1e9 / peak) if (traffic and ms_k > 0) else None,
            "frame": {"algorithmic_bytes_per_step": frame_bytes / n_steps,
                      "achieved": frame_bytes / (ms_dev_max * 1e-3) / 1e9, "peak": peak * world,
                      "frac": frame_bytes / (ms_dev_max * 1e-3) / 1e9 / (peak * world)},
            "other_kernels_ms_per_step": {"tail_kernel": sum(s["ms_tail"] for s, _ in stats_dev) / n_steps,
                                          "shade_kernel": sum(s["ms_shade_kernel"] for s, _ in stats_dev) / n_steps,
                                          "trace_kernel": ms_k / n_steps},
        }
        cfgd = workload_config(cfg)
        cfgd.update({
            "l2": "flushed between timed steps (256 MiB memset outside the event pair)",
            "sharding": "none" if world == 1 else (
                "16x16 tiles round-robin over %d ranks; " % world +
                ("each rank's resolve kernel stores its tiles into rank 0's image over NVLink (CUDA IPC) + one 4-byte all-reduce"
                 if peer is not None else "one NCCL all-gather of packed tiles + un-tile")),
            "sharded_frame_equals_single_gpu_frame": verified,
            "node_format": int(scene.build_info().node_format),   # device node record (include/nrays_b200.h: NrbBuildInfo)
            "rays_per_step": rays_dev / steps,
            "rays_reference_per_step": (counts_all["rays_primary"] + counts_all["rays_reflect"] + counts_all["rays_refract"] +
                                        counts_all["rays_shadow"]) / steps,
            "note": "value counts BVH queries actually performed; the reference also casts the light samples of "
                    "zero weight (rays_reference_per_step), which this library skips"})
        res = {
            "value": rays_dev / (ms_dev_max * 1e-3) / 1e6, "unit": UNIT, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_dev_max / steps, "config": cfgd,
            "e2e": {"value": rays_e2e / (ms_e2e_max * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": C.sizeof(A.NrbCamera) * world,
                    "d2h_bytes_per_step": w * h * 3 * 4, "ms_per_step": ms_e2e_max / steps,
                    "path": ("nrb_render into pinned host memory" if world == 1 else
                             ("every rank DMAs its tile columns into one shared pinned host image over its own PCIe link"
                              if shost is not None else "device image on rank 0, then one device->host copy of the frame"))},
            "gpu_launches": int(lt.item()),
            "roofline": roofline,
        }
    lib.nrb_host_free(host_ptr)
    if shost is not None:
        barrier()
        shost.close()
    if peer is not None:
        out = None
        barrier()
        peer.close()
    scene.close()
    del out, packed
    torch.cuda.empty_cache()
    return res


def run_ours(args, rank, local_rank, world):
    import torch

    from nrays_b200 import _lib, configs

    ctx = Ctx()
    ctx.lib = _lib.load()  # raises if the CUDA library is missing: no CPU fallback
    if ctx.lib.nrb_device_count() <= local_rank:
        raise RuntimeError("bench.py needs a CUDA device per rank (found %d)" % ctx.lib.nrb_device_count())
    torch.cuda.set_device(local_rank)
    ctx.rank, ctx.world, ctx.local_rank = rank, world, local_rank
    ctx.dev = torch.device("cuda", local_rank)
    ctx.td = None
    if world > 1:
        import torch.distributed as td

        td.init_process_group("nccl", device_id=ctx.dev)
        ctx.td = td
    ctx.stream = torch.cuda.current_stream()
    ctx.flush = torch.empty(256 << 20, dtype=torch.uint8, device=ctx.dev)  # > 126 MB L2

    if args.extra == "auto":
        extra = ["C2", "C4", "C5"] if world == 1 else ["C5"]
    elif args.extra == "none":
        extra = []
    else:
        extra = [c for c in args.extra.split(",") if c]
    extra = [c for c in extra if c != args.config and c in configs.CONFIGS]

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.4)   # let nvidia-smi start sampling before the timed regions
    main_res = measure_config(ctx, args.config, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None   # sampled across both timed regions of the headline config

    others = {}
    for c in extra:
        r = measure_config(ctx, c, max(3, min(args.steps, 5)), 3)
        if rank == 0:
            others[c] = {k: r[k] for k in ("value", "unit", "steps", "warmup", "ms_per_step", "e2e", "gpu_launches", "config")}
            others[c]["roofline"] = {k: r["roofline"][k] for k in ("kernel", "achieved", "peak", "frac", "traffic", "ms_per_launch",
                                                                   "launches_per_step", "kernel_share_of_step", "frame",
                                                                   "other_kernels_ms_per_step")}

    if rank == 0:
        cfg = configs.CONFIGS[args.config]
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu_frame(args.config)   # builds the oracle scene, pages everything in
            rays, secs, cores, _full = cpu_frame(args.config)
            reps = 1
            while secs < 10.0 and reps < 64:
                r2, s2, cores, _full = cpu_frame(args.config)
                rays, secs, reps = rays + r2, secs + s2, reps + 1
            cpu = {"value": rays / secs / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": sample_label(args.config, cfg["name"], cfg["height"]) + " x%d (%.1f s wall on %d threads)" % (reps, secs, cores)}
        line = {
            "metric": METRIC, "value": main_res["value"], "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": main_res["config"], "e2e": main_res["e2e"], "gpu_launches": main_res["gpu_launches"],
            "roofline": main_res["roofline"], "cpu_baseline": cpu, "clocks": clocks,
            "configs": others,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        ctx.td.destroy_process_group()


def main():
    args = parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1 and args.impl == "ours":
        # convenience: relaunch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29517"), os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup), "--config", args.config,
               "--extra", args.extra]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
