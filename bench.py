#!/usr/bin/env python3
"""bench.py — headline benchmark of the B200-native TinyRayTracing hot path.

Workload (BASELINE.json configs[1]): the fixed 16 Mi random-ray closest-hit batch against the cornell-box BVH
(the Cornell shell `test/back` — cornell-box.obj itself is a missing blob in the reference checkout), 1 B200.
A "step" is one pass of the closest-hit path over the whole batch.

  value     Mrays/s with rays and results resident in HBM (CUDA events on the launching stream)
  e2e       the same metric through the C-ABI call a host program makes (trt_trace_closest with pinned host
            buffers): H2D of the rays and D2H of ids + distances inside the timed region
  roofline  dominant kernel (k_closest_persistent) against the resource that binds it on this workload — the SM issue
            slots times the lanes active per issued instruction (the 2 kB scene is served from L1; ncu metrics of the
            shipped build in profiles/r02_ncu_metrics.json, recomputable from ms_per_step) — with the DRAM traffic ncu
            measured beside it, and the SURVEY §8d HBM-denominated figure (B = 48 + 32*A + 48*T bytes per ray, A, T from
            profiles/algorithmic_work.json) kept as roofline.survey_hbm for cross-config comparison
  other_scenes  the same kernel on staircase and on the 10 M-triangle stress mesh (BASELINE config 5, the one config
            whose scene exceeds L2), each with its ncu DRAM traffic against the measured HBM peak
  cpu_baseline / --impl reference: the UNMODIFIED reference traverseBVH (oracle/_ref/libref.so) on the box's
            host cores over a bounded sample of the same rays

N > 1 (torchrun, one rank per GPU): the scene is replicated, every rank traces its own 16 Mi batch (weak
scaling, no data-path collective: ray batches shard by index); time = max over ranks.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_RAYS = 16 << 20
SCENE = "back"
WORKLOAD = ("16Mi-ray closest-hit batch (25% camera / 50% uniform-in-AABB / 25% cosine bounce), cornell-box shell "
            "example-scenes-cg22/test/back (26 tris; cornell-box.obj is a missing blob)")


def env_int(k, d):
    return int(os.environ.get(k, d))


def bench_config():
    """`config` of the JSON line: the workload, spelled identically by both arms (--impl b200 / reference) so that the
    driver can tell they ran the same thing; what differs between the arms (rays actually traced per step, threads,
    traversal mode) is under `arm`."""
    return {"workload": WORKLOAD, "scene": "example-scenes-cg22/test/back", "rays_per_batch": N_RAYS,
            "ray_seed": "0x5EED0001 + rank", "resolution_of_camera_rays": "512x512",
            "l2": "inputs (%d MB rays + %d MB results per batch) exceed the 126 MB L2" % (N_RAYS * 24 >> 20, N_RAYS * 8 >> 20)}


def load_ncu_metrics(key):
    """ncu metrics of the shipped build for one workload (tools/ncu_metrics.py -> profiles/r02_ncu_metrics.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_ncu_metrics.json")) as f:
            return json.load(f)[key]
    except Exception:
        return None


def sm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["sm_max_mhz"])
    except Exception:
        return 1965.0


def issue_roofline(key, rays, ms, sm_mhz, n_sm=148):
    """The kernel against the SM issue slots: a B200 SM issues at most 4 warp instructions per cycle (one per
    sub-partition), each serving up to 32 lanes.  warp instructions / ray and lanes / instruction come from the committed
    ncu capture of the shipped build; the rate is this run's.  frac = issue-slot utilisation x lanes / 32 = the fraction
    of the machine's thread-instruction throughput that does work."""
    m = load_ncu_metrics(key)
    if not m:
        return None
    clock = (sm_mhz or sm_peak()) * 1e6
    rays_per_s = rays / (ms * 1e-3)
    slots = n_sm * 4 * clock
    issue = m["warp_inst_per_ray"] * rays_per_s / slots
    achieved = m["warp_inst_per_ray"] * m["lanes_per_inst"] * rays_per_s
    traffic = (m["dram_bytes_read"] + m["dram_bytes_write"]) * (rays / float(m["rays_per_launch"]))
    peak_gbs, which = measured_peak_gbs()
    return {"bound": "issue", "achieved": achieved / 1e12, "peak": slots * 32 / 1e12, "unit": "Tthread-inst/s",
            "frac": achieved / (slots * 32), "traffic": traffic, "kernel": m.get("kernel", "k_closest_persistent"),
            "issue_slot_frac": issue, "lanes_per_inst": m["lanes_per_inst"], "warp_inst_per_ray": m["warp_inst_per_ray"],
            "ncu_issue_active_pct": m.get("issue_active_pct"), "sm_mhz_used": clock / 1e6,
            "dram": {"achieved": traffic / (ms * 1e-3) / 1e9, "peak": peak_gbs, "unit": "GB/s",
                     "frac": traffic / (ms * 1e-3) / 1e9 / peak_gbs, "peak_source": which + " (MEASURED_PEAKS.json hbm_gbs)",
                     "note": "ncu dram__bytes_read.sum + dram__bytes_write.sum of the same launch, scaled to this batch"},
            "source": "profiles/r02_ncu_metrics.json: " + m.get("capture", "")}


def load_algorithmic_work(scene):
    p = os.path.join(ROOT, "profiles", "algorithmic_work.json")
    try:
        with open(p) as f:
            w = json.load(f)[scene]
        return float(w["A"]), float(w["T"])
    except Exception:
        return None


def load_traffic(scene, rays_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu capture, per launch."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)[scene]
        return (t["dram_bytes_read"] + t["dram_bytes_write"]) * (rays_per_launch / float(t["rays_per_launch"]))
    except Exception:
        return None


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  Polls NVML in-process from a thread (a query takes
    well under a millisecond, so even a 10 ms timed region gets several samples); if NVML cannot be opened it falls
    back to the `nvidia-smi -lms` loop of the B200_PROFILING.md recipe, whose first sample takes ~100 ms to arrive."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.nvml, self.handle, self.samples, self.run, self.keep = None, None, [], False, False
        try:
            import pynvml

            pynvml.nvmlInit()
            handle = None
            try:  # the CUDA ordinal is not the NVML index under CUDA_VISIBLE_DEVICES: go through the UUID
                import torch

                uuid = str(torch.cuda.get_device_properties(index).uuid)
                handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)
            self.nvml, self.handle = pynvml, handle
        except Exception:
            self.nvml = None

    def _poll(self):
        nv, h = self.nvml, self.handle
        while self.run:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    why = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    why = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                if self.keep:
                    self.samples.append((mhz, why))
                else:
                    self.last_before = (mhz, why)
            except Exception:
                pass
            time.sleep(0.0005)

    def start(self):
        if self.nvml:
            self.run = True
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def mark(self):
        """Samples before this point (warm-up) are dropped; if none arrives after it, the last one before is kept."""
        self.mark_at = len(self.lines)
        self.keep = True

    def stop(self):
        if self.nvml:
            self.run = False
            self.t.join(timeout=2)
            nv = self.nvml
            samples = self.samples or ([self.last_before] if hasattr(self, "last_before") else [])
            bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                    "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            reasons = sorted(k for k, b in bits.items() if any(w & b for _, w in samples))
            try:
                mx = float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))
            except Exception:
                mx = None
            sm = [float(m) for m, _ in samples]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": reasons,
                    "samples": len(self.samples), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        start = getattr(self, "mark_at", 0)
        lines = self.lines[start:] if len(self.lines) > start else self.lines[-1:]
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for nme, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


def materialize_scene(tmp):
    from tinyraytracing_b200 import scenes

    return scenes.materialize(SCENE, tmp, width=512, height=512)


# ------------------------------------------------------------------------------------------------ reference arm
def reference_rays(ref, n, seed):
    """Same ray population as the GPU arm, surface points supplied by the reference's own traversal."""
    from tinyraytracing_b200 import workloads

    cam12 = ref.camera()
    w, h = ref.image_size()
    cam = dict(eye=cam12[0:3], llc=cam12[3:6], horizontal=cam12[6:9], vertical=cam12[9:12], width=w, height=h)
    boxes, _ = ref.bvh_flatten()

    def tracer(rays):
        t, ids, pn, hp = ref.trace(rays, want_pn=True)
        return ids, hp, pn

    return workloads.fixed_ray_batch(n, cam, (boxes[0, :3], boxes[0, 3:]), tracer, seed=seed)


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refbridge

    if not refbridge.available():
        emit(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref.so not built (run __graft_entry__.build() in the build container)"}))
        return 0
    with tempfile.TemporaryDirectory() as tmp:
        f = materialize_scene(tmp)
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        os.dup2(devnull, 1)  # the reference loaders print to stdout
        try:
            ref = refbridge.RefScene(f["xml"], f["obj"], f["mtl"], f["basedir"])
        finally:
            os.dup2(saved, 1)
        cores = os.cpu_count() or 1
        # bounded sample: size it so that one step takes about 2 s on this host
        pilot = reference_rays(ref, 1 << 18, seed=0x5EED0001)
        t0 = time.perf_counter()
        ref.trace(pilot, threads=cores)
        rate = len(pilot) / (time.perf_counter() - t0)
        n = int(min(N_RAYS, max(1 << 18, rate * 2.0)))
        rays = reference_rays(ref, n, seed=0x5EED0001)
        for _ in range(args.warmup):
            ref.trace(rays[: max(1, n // 8)], threads=cores)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ref.trace(rays, threads=cores)
        dt = time.perf_counter() - t0
        mrays = n * args.steps / dt / 1e6
        sample = "%d rays of the %d-ray batch per step" % (n, N_RAYS)
        emit(json.dumps({
            "impl": "reference", "metric": "Mrays/s (closest-hit)", "value": mrays, "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(), "arm": {"rays_per_step": n, "host_threads": cores},
            "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
    return 0


def bind_near_gpu(index):
    """Run this process on the CPUs of the NUMA node the GPU hangs off, so that the page-locked ray / result buffers
    allocated next live in that node's memory: with one process per GPU the host legs of eight e2e pipelines otherwise
    cross the socket interconnect (round 1: 52 GB/s H2D per GPU at N = 1, 15.6 GB/s at N = 8).  Returns what it did and
    the previous affinity (restored before the CPU baseline, which wants every core)."""
    old = os.sched_getaffinity(0)
    try:
        import torch

        p = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read())
        if node < 0:
            return {"numa_node": None, "note": "the platform reports no NUMA node for %s" % bdf}, old
        cpus = set()
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        cpus &= old
        if cpus:
            os.sched_setaffinity(0, cpus)
        nodes = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
        return {"numa_node": node, "numa_nodes": nodes, "cpus_bound": len(cpus), "pci": bdf}, old
    except Exception as e:  # containers without sysfs, non-Linux: leave the affinity alone
        return {"numa_node": None, "note": "not bound: %s" % e}, old


def link_ceiling(n, steps, world, barrier):
    """What the host link allows for one e2e step: plain page-locked cudaMemcpyAsync of the same byte counts (24 B in,
    8 B out per ray), both directions at once on two streams, every rank at the same time — no kernel, no chunking."""
    import torch
    import torch.distributed as dist

    h_in = torch.empty(n * 24, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n * 8, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n * 24, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros(n * 8, dtype=torch.uint8, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def once():
        with torch.cuda.stream(s_in):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s_out):
            h_out.copy_(d_out, non_blocking=True)
        s_in.synchronize()
        s_out.synchronize()

    for _ in range(2):
        once()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        once()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    if world > 1:
        tm = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm.item())
    return {"ms_per_step": ms, "h2d_gbs_per_gpu": n * 24 / (ms * 1e-3) / 1e9, "d2h_gbs_per_gpu": n * 8 / (ms * 1e-3) / 1e9,
            "what": "pinned cudaMemcpyAsync of the same bytes, H2D and D2H concurrently, all %d ranks at once (max over ranks)" % world}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist

    import tinyraytracing_b200 as trt
    from tinyraytracing_b200 import workloads

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = trt.load_library()
    numa, old_affinity = bind_near_gpu(local)
    with tempfile.TemporaryDirectory() as tmp:
        f = materialize_scene(tmp)
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        os.dup2(devnull, 1)  # the loaders keep the reference's stdout chatter
        try:
            host = trt.HostScene.load(f["xml"], f["obj"], f["mtl"], f["basedir"])
        finally:
            os.dup2(saved, 1)
        dev = trt.DeviceScene(host, local)

        def tracer(rays):
            ids, t = dev.trace_closest(rays)
            hp, pn = dev.hit_attributes(rays, ids, t)
            return ids, hp, pn

        n = args.rays
        rays = workloads.fixed_ray_batch(n, host.camera(), host.root_box(), tracer, seed=0x5EED0001 + rank)
        # pinned host buffers for the e2e leg, device buffers for the resident leg
        p_rays, p_id, p_t = lib.trt_host_alloc(n * 24), lib.trt_host_alloc(n * 4), lib.trt_host_alloc(n * 4)
        C.memmove(p_rays, rays.ctypes.data, n * 24)
        d_rays = torch.from_numpy(rays).cuda()
        d_id = torch.empty(n, dtype=torch.int32, device="cuda")
        d_t = torch.empty(n, dtype=torch.float32, device="cuda")
        stream = torch.cuda.current_stream()
        sp = stream.cuda_stream

        def step_resident():
            dev.trace_closest_async(d_rays.data_ptr(), n, d_id.data_ptr(), d_t.data_ptr(), args.flags, sp)

        def step_e2e():
            dev.trace_closest_ptr(p_rays, n, p_id, p_t, args.flags)

        def barrier():
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        def timed(fn, events):
            clocks = ClockSampler(local)
            clocks.start()  # nvidia-smi needs a moment before its first sample: start it ahead of the warm-up
            if events:
                warm_up_gpu(fn)  # the SM clock back at its maximum after the host-side ray generation (see warm_up_gpu)
            for _ in range(args.warmup):
                fn()
            barrier()
            dev.reset_stats()
            clocks.mark()
            if events:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(args.steps):
                    fn()
                e1.record(stream)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
            else:
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    fn()
                torch.cuda.synchronize()
                ms = (time.perf_counter() - t0) * 1e3
            ck = clocks.stop()
            launches = dev.stats()["kernel_launches"]
            barrier()
            if world > 1:
                tm = torch.tensor([ms], dtype=torch.float64, device="cuda")
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                ms = float(tm.item())
            return ms, ck, launches

        ms_res, clocks, launches = timed(step_resident, events=True)
        ms_e2e, clocks_e2e, _ = timed(step_e2e, events=False)
        link = link_ceiling(n, max(3, min(args.steps, 10)), world, barrier)
        try:
            os.sched_setaffinity(0, old_affinity)  # the CPU legs below want every core
        except Exception:
            pass

        # sanity of what was timed: results of both legs agree and really hit geometry
        ids_host = np.ctypeslib.as_array(C.cast(p_id, C.POINTER(C.c_int32)), (n,))
        assert np.array_equal(ids_host, d_id.cpu().numpy()), "resident and e2e legs disagree"
        hit_fraction = float((ids_host >= 0).mean())

        total_rays = float(n) * world * args.steps
        value = total_rays / (ms_res * 1e-3) / 1e6
        e2e = total_rays / (ms_e2e * 1e-3) / 1e6
        out = {
            "metric": "Mrays/s (closest-hit)", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(),
            "arm": {"rays_per_step_per_gpu": n, "hit_fraction": hit_fraction,
                    "traversal": {0: "default", 2: "exhaustive", 4: "reftopo"}.get(args.flags, str(args.flags))},
            "e2e": {"value": e2e, "unit": "Mrays/s", "h2d_bytes_per_step": n * 24 * world, "d2h_bytes_per_step": n * 8 * world,
                    "ms_per_step": ms_e2e / args.steps, "timer": "host wall clock around the blocking C-ABI call",
                    "link": link, "link_frac": link["ms_per_step"] / (ms_e2e / args.steps), "host_placement": numa,
                    "bound": "host link: %.1f GB/s H2D + %.1f GB/s D2H per GPU (24 B in, 8 B out per ray); the kernel itself "
                             "needs %.1f %% of the step" % (n * 24 / (ms_e2e / args.steps * 1e-3) / 1e9,
                                                           n * 8 / (ms_e2e / args.steps * 1e-3) / 1e9,
                                                           100.0 * ms_res / ms_e2e)},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        work = load_algorithmic_work(SCENE)
        peak, which = measured_peak_gbs()
        roof = issue_roofline(SCENE, n, ms_res / args.steps, clocks.get("sm_mhz")) or {
            "bound": "issue", "achieved": None, "peak": None, "unit": "Tthread-inst/s", "frac": None, "traffic": load_traffic(SCENE, n),
            "note": "profiles/r02_ncu_metrics.json missing: no ncu capture of this build to read the instruction counts from"}
        if work:
            A, T = work
            bytes_per_ray = 48 + 32 * A + 48 * T
            achieved = bytes_per_ray * n / (ms_res / args.steps * 1e-3) / 1e9  # per launch = per step on one GPU
            roof["survey_hbm"] = {"achieved": achieved, "peak": peak, "unit": "GB/s", "ratio": achieved / peak,
                                  "algorithmic_bytes_per_launch": bytes_per_ray * n, "bytes_per_ray": bytes_per_ray, "A": A, "T": T,
                                  "layout": layout_roofline(dev, rays, n, ms_res / args.steps, peak),
                                  "note": "SURVEY §8d's cross-config figure: algorithmic bytes of the ordered-pruned walk of the "
                                          "REFERENCE topology / measured HBM copy bandwidth. NOT a roofline fraction for this "
                                          "scene: the 2 kB scene is served from L1, so the ratio exceeds 1 by construction; the "
                                          "DRAM the kernel really moves is roofline.dram"}
        out["roofline"] = roof
        if world > 1:
            one = e2e_one_process(args, dev, lib, rays, n, rank, world, local, barrier)
            if rank == 0:
                out["e2e"]["one_process"] = one
        if not args.no_render:
            out["render"] = render_measurements(args, tmp, rank, world, local, barrier)
            if rank == 0 and world == 1:
                out["cornell_standin"] = standin_measurements(args)
        if rank == 0 and world == 1 and not args.no_cpu:
            t_host = np.ctypeslib.as_array(C.cast(p_t, C.POINTER(C.c_float)), (n,))
            out["cpu_baseline"] = cpu_baseline(f, rays, ids_host, t_host, host)
        if rank == 0 and world == 1 and not args.no_render:
            out["other_scenes"] = other_scene_measurements(args, tmp)
        if args.extra and rank == 0 and world == 1:
            out["extra"] = extra_measurements(args, tmp)
        for p in (p_rays, p_id, p_t):
            lib.trt_host_free(p)
        dev.close()
    if rank == 0:
        emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


def layout_roofline(dev, rays, n, ms_per_step, peak):
    """The same figure with the bytes the DEFAULT traversal actually touches per ray (its own tree, counted on the
    GPU by trt_trace_counters): 32 B ray I/O + 128 B per 4-wide node visited + 48 B per triangle tested.  The SURVEY
    figure above is defined on the reference topology, so it exceeds what this layout moves."""
    w = dev.trace_counters(rays[:: max(1, n >> 20)])
    b = 32 + 128 * w["nodes"] + 48 * w["tris"]
    achieved = b * n / (ms_per_step * 1e-3) / 1e9
    return {"bytes_per_ray": b, "achieved": achieved, "frac": achieved / peak, "work_per_ray": w,
            "note": "served from L1 (hit rate ~90 %): the kernel is issue / L1-latency bound, see profiles/"}


def cpu_baseline(files, rays, gpu_ids=None, gpu_t=None, host=None):
    """The unmodified reference traversal (oracle/_ref) on the host cores, bounded sample of the same rays.  The same
    pass doubles as BASELINE config 2's check on the bench's own batch: the GPU's triangle ids and distance bits (from
    the timed e2e leg) against what the reference's traverseBVH returned for every ray of the sample."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    cores = os.cpu_count() or 1
    import refbridge

    if refbridge.available():
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        os.dup2(devnull, 1)
        try:
            ref = refbridge.RefScene(files["xml"], files["obj"], files["mtl"], files["basedir"])
        finally:
            os.dup2(saved, 1)
        trace, kind = (lambda r: ref.trace(r, threads=cores)), "reference"  # -> (t, canonical post-build id)
    else:
        import oraclelib

        orc = oraclelib.OracleScene(oraclelib.parsed_scene(SCENE, 512, 512))
        trace, kind = (lambda r: orc.trace(r, threads=cores)[::-1]), "port"  # the oracle returns (id, t)
    pilot = rays[: 1 << 18]
    t0 = time.perf_counter()
    trace(pilot)
    rate = len(pilot) / (time.perf_counter() - t0)
    n = int(min(len(rays), max(1 << 18, rate * 10.0)))  # about 10 s of CPU work
    t0 = time.perf_counter()
    ref_t, ref_id = trace(rays[:n])
    dt = time.perf_counter() - t0
    out = {"value": n / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": kind,
           "sample": "first %d rays of the %d-ray batch, %d OpenMP threads" % (n, len(rays), cores)}
    if gpu_ids is not None and gpu_t is not None:
        ids, t = np.asarray(gpu_ids[:n]), np.asarray(gpu_t[:n])
        same = ids == np.asarray(ref_id)
        if host is not None and not same.all():
            # the reference identifies a triangle by content: geometrically identical triangles are one identity
            v = host.triangles()["v"]
            bad = np.flatnonzero(~same)
            both = (ids[bad] >= 0) & (np.asarray(ref_id)[bad] >= 0)
            eq = np.zeros(len(bad), bool)
            eq[both] = (v[ids[bad][both]] == v[np.asarray(ref_id)[bad][both]]).all(axis=1)
            same[bad] = eq
        out["parity"] = {"rays_compared": int(n), "triangle_ids_equal": bool(same.all()),
                         "id_mismatches": int((~same).sum()),
                         "distance_bits_equal": bool(np.array_equal(t.view(np.uint32), np.asarray(ref_t).view(np.uint32))),
                         "against": "the unmodified reference traverseBVH" if kind == "reference" else "the oracle port"}
    return out


def e2e_one_process(args, dev, lib, rays, n, rank, world, local, barrier):
    """The e2e leg once more with ONE process feeding all N GPUs (trt_trace_closest_multi: the batch shards by ray index
    over replicas of the scene, one host thread per GPU) instead of N processes: N x 16 Mi rays in page-locked memory of
    rank 0's process, the other ranks off their GPUs.  Says whether the host-side ceiling of `e2e.link` belongs to the
    box or to the process layout."""
    import torch.distributed as dist

    import tinyraytracing_b200 as trt

    barrier()
    store = dist.distributed_c10d._get_default_store()
    if rank != 0:
        store.wait(["trt_e2e_one_process_done"])
        return None
    total = n * world
    p_rays, p_id, p_t = lib.trt_host_alloc(total * 24), lib.trt_host_alloc(total * 4), lib.trt_host_alloc(total * 4)
    out = None
    try:
        if p_rays and p_id and p_t:
            for i in range(world):
                C.memmove(p_rays + i * n * 24, rays.ctypes.data, n * 24)
            devs = [dev] + [dev.replicate(i) for i in range(world) if i != local]
            handles = (C.c_void_p * world)(*[d.h for d in devs])
            call = lambda: lib.trt_trace_closest_multi(handles, world, p_rays, total, p_id, p_t, args.flags)
            for _ in range(2):
                assert call() == 0, lib.trt_last_error()
            times = []
            for _ in range(5):
                t0 = time.perf_counter()
                assert call() == 0, lib.trt_last_error()
                times.append((time.perf_counter() - t0) * 1e3)
            ms = float(np.median(times))
            ids = np.ctypeslib.as_array(C.cast(p_id, C.POINTER(C.c_int32)), (total,))
            same = all(np.array_equal(ids[i * n:(i + 1) * n], ids[:n]) for i in range(1, world))
            out = {"value": total / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_call": ms, "rays_per_call": total,
                   "shards_agree": bool(same), "what": "trt_trace_closest_multi: one process, %d GPUs, one host thread per GPU" % world}
            for d in devs[1:]:
                d.close()
    finally:
        for p in (p_rays, p_id, p_t):
            if p:
                lib.trt_host_free(p)
        store.set("trt_e2e_one_process_done", "1")
    return out


def warm_up_gpu(launch, seconds=0.2):
    """Keep the GPU busy for a fifth of a second before a short timed region: after a second or two of host-side work
    (scene load, layout build) the SM clock has dropped to idle and takes tens of milliseconds of load to come back — three
    sub-millisecond warm-up launches are over before it has (a 4 Mi-ray staircase batch then reads 4.1 instead of 6.4 Grays/s)."""
    import torch

    t_end = time.perf_counter() + seconds
    while time.perf_counter() < t_end:
        for _ in range(4):
            launch()
        torch.cuda.synchronize()


def frame_hashes(img, rgb8=None):
    """What a reader needs to tell that two runs produced the same picture: the 8-bit gamma-packed frame is identical
    for any GPU count (the float64 sums differ in their last bits with the order of the per-GPU partial sums)."""
    import hashlib

    if rgb8 is None:  # imshow's formula, main.cpp:30-38
        rgb8 = np.clip(np.power(img, np.float64(np.float32(1.0) / np.float32(2.2))) * 255, 0, 255).astype(np.uint8)
    return {"frame_rgb8_sha256": hashlib.sha256(np.ascontiguousarray(rgb8).tobytes()).hexdigest()[:16],
            "frame_f64_sha256": hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest()[:16],
            "image_mean": float(img.mean())}


def render_measurements(args, tmp, rank, world, local, barrier):
    """BASELINE configs 1 / 3 / 4: full renders with the samples sharded over the GPUs and ONE sum-reduce of the
    accumulation buffers, resolve on GPU 0.  spp/s = spp / device time including the reduce and the resolve.
    N = 1: trt_render.  N > 1: INSIDE the library (trt_render_multi: rank 0's process drives all N GPUs — one host thread
    per GPU, scene replicated device to device, ncclCommInitAll + ncclReduce(sum, f64), and the library's own peer-memory
    reduce + resolve kernel measured beside it) while the other ranks wait on the rendezvous store without touching their
    GPUs; the round-1 path (one process per GPU, torch.distributed reduce) is timed beside it on config 3."""
    import torch
    import torch.distributed as dist

    import tinyraytracing_b200 as trt
    from tinyraytracing_b200 import scenes
    from tinyraytracing_b200.distributed import render_on_gpus

    cfgs = [("config1_cornell_shell", "back", 512, 512, 16), ("config3_veach_mis", "veach-mis", 1280, 720, args.spp3)]
    if world == 8 or args.config4:
        cfgs.append(("config4_staircase", "staircase", 1920, 1080, args.spp4))
    store = dist.distributed_c10d._get_default_store() if world > 1 else None
    out = {}
    for key, name, w, h, spp in cfgs:
        f = scenes.materialize(name, os.path.join(tmp, "r_%s_%d" % (name, rank)), width=w, height=h)
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        os.dup2(devnull, 1)
        try:
            host = trt.HostScene.load(f["xml"], f["obj"], f["mtl"], f["basedir"])
        finally:
            os.dup2(saved, 1)
        dev = trt.DeviceScene(host, local)
        frame = dev.pinned_image()  # the frame lands in page-locked host memory (valid until dev.close())
        # timed repetitions: the median is reported (a one-off stall of several ms shows up now and then in the first
        # render after another process's activity, and in the few-ms renders of config 1 at N > 1), the list is printed
        reps = {"config1_cornell_shell": 9, "config3_veach_mis": 3 if world == 1 else 7}.get(key, 1)
        rec = {"scene": name, "width": w, "height": h, "spp": spp}
        if world == 1:
            dev.render(spp, seed=2, out=frame)  # warm-up at full size: allocates the wavefront buffers
            times = []
            for _ in range(reps):
                dev.reset_stats()
                img = dev.render(spp, seed=1, out=frame)
                times.append(dev.stats()["last_render_ms"])
            st = dev.stats()
            ms = float(np.median(times))
            rays = st["rays_closest"] + st["rays_shadow"]
            rec.update({"ms": ms, "ms_min": float(min(times)), "spp_per_s": spp / (ms * 1e-3), "mrays_per_s": rays / (ms * 1e-3) / 1e6,
                        "rays_closest": int(st["rays_closest"]), "rays_shadow": int(st["rays_shadow"]),
                        "kernel_launches": int(st["kernel_launches"]), "timed_renders_ms": [round(x, 3) for x in times],
                        "path": "trt_render (one GPU)"})
            rec.update(frame_hashes(img))
        else:
            # ---- round-1 path, config 3 only: one process per GPU, torch.distributed reduce over NCCL
            if key == "config3_veach_mis":
                render_on_gpus(dev, spp, seed=2, out=frame)
                tt = []
                for _ in range(reps):
                    barrier()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    img = render_on_gpus(dev, spp, seed=1, out=frame)
                    e1.record()
                    torch.cuda.synchronize()
                    tt.append(e0.elapsed_time(e1))
                t = torch.tensor(tt, dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                if rank == 0:
                    rec["one_process_per_gpu"] = {"ms": float(np.median(t.tolist())), "spp_per_s": spp / (float(np.median(t.tolist())) * 1e-3),
                                                  "reduce": "torch.distributed.reduce(SUM, f64) over NCCL", **frame_hashes(img)}
            barrier()
            # ---- the product path: rank 0 drives all N GPUs through trt_render_multi; the other ranks keep off their GPUs
            if rank == 0:
                devs = [dev] + [dev.replicate(i) for i in range(world) if i != local]
                for label, flags in (("nccl", 0), ("peer", trt.RENDER_PEER_REDUCE)):
                    trt.render_multi(devs, spp, seed=2, flags=flags, out=frame)  # warm-up (communicators, peer mappings)
                    times = []
                    for _ in range(reps):
                        for d in devs:
                            d.reset_stats()
                        img, rgb = trt.render_multi(devs, spp, seed=1, flags=flags, out=frame, want_rgb8=True)
                        times.append(dev.stats()["last_render_ms"])
                    ms = float(np.median(times))
                    sts = [d.stats() for d in devs]
                    rays = sum(s_["rays_closest"] + s_["rays_shadow"] for s_ in sts)
                    # (under torchrun the other ranks' processes keep their CUDA contexts on the GPUs this process now drives:
                    # the driver time-slices between them, and a few-ms render now and then waits out a slice — ms_min
                    # is the undisturbed figure, ms the median as everywhere else)
                    r = {"ms": ms, "ms_min": float(min(times)), "spp_per_s": spp / (ms * 1e-3), "mrays_per_s": rays / (ms * 1e-3) / 1e6,
                         "timed_renders_ms": [round(x, 3) for x in times], **frame_hashes(img, rgb)}
                    if label == "nccl":
                        rec.update(r)
                        rec.update({"rays_closest": int(sum(s_["rays_closest"] for s_ in sts)),
                                    "rays_shadow": int(sum(s_["rays_shadow"] for s_ in sts)),
                                    "kernel_launches": int(sum(s_["kernel_launches"] for s_ in sts)),
                                    "path": "trt_render_multi: one process, %d GPUs, scene replicated device to device, samples "
                                            "[i*spp/N, (i+1)*spp/N) per GPU, ncclCommInitAll + one ncclReduce(sum, f64, W*H*3), "
                                            "resolve on GPU 0" % world})
                    else:
                        rec["peer_reduce"] = dict(r, path="same, with the library's own kernel: GPU 0 reads the peers' buffers "
                                                          "over NVLink, sums in rank order and resolves in one pass")
                for d in devs[1:]:
                    d.close()
                store.set("trt_render_done_" + key, "1")
            else:
                store.wait(["trt_render_done_" + key])
        dev.close()
        barrier()
        if rank == 0 and world == 1 and not args.no_cpu:
            if key == "config3_veach_mis":
                rec["cpu_reference"] = cpu_reference_render(name, w, h, tmp, full=False)
            elif key == "config1_cornell_shell":
                rec["cpu_reference"] = cpu_reference_render(name, w, h, tmp, full=True, spp=spp)
            if "spp_per_s_at_config" in rec.get("cpu_reference", {}):
                rec["vs_cpu_reference"] = rec["spp_per_s"] / rec["cpu_reference"]["spp_per_s_at_config"]
        out[key] = rec
    return out


def other_scene_measurements(args, tmp):
    """The dominant kernel on the scenes where it is NOT served from L1: staircase (31 k triangles, 1.9 MB) and the
    10 M-triangle stress mesh of BASELINE config 5 (590 MB of nodes + triangles: the one config whose scene exceeds the
    126 MB L2).  4 Mi config-2 rays each, device-resident, CUDA events; each with the DRAM traffic ncu measured for the
    same launch against the measured HBM peak."""
    import torch

    import tinyraytracing_b200 as trt
    from tinyraytracing_b200 import scenes, workloads

    out = {}
    n = 4 << 20

    def measure(host, key, label):
        dev = trt.DeviceScene(host, 0)

        def tracer(rays):
            ids, t = dev.trace_closest(rays)
            hp, pn = dev.hit_attributes(rays, ids, t)
            return ids, hp, pn

        rays = workloads.fixed_ray_batch(n, host.camera(), host.root_box(), tracer)
        d_rays = torch.from_numpy(rays).cuda()
        d_id = torch.empty(n, dtype=torch.int32, device="cuda")
        d_t = torch.empty(n, dtype=torch.float32, device="cuda")
        sp = torch.cuda.current_stream().cuda_stream
        warm_up_gpu(lambda: dev.trace_closest_async(d_rays.data_ptr(), n, d_id.data_ptr(), d_t.data_ptr(), args.flags, sp))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dev.trace_closest_async(d_rays.data_ptr(), n, d_id.data_ptr(), d_t.data_ptr(), args.flags, sp)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        st = dev.stats()
        rec = {"what": label, "tris": int(host.n_tris), "ref_depth": int(st["ref_depth"]), "rays_per_launch": n, "ms_per_launch": ms,
               "closest_hit_mrays": n / (ms * 1e-3) / 1e6, "hit_fraction": float((d_id >= 0).float().mean()),
               "work_per_ray": dev.trace_counters(rays[:: max(1, n >> 19)]), "rays_strict": int(st["rays_strict"])}
        roof = issue_roofline(key, n, ms, None)
        if roof:
            rec["roofline"] = roof
        dev.close()
        return rec

    f = scenes.materialize("staircase", os.path.join(tmp, "o_staircase"), width=1280, height=720)
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)
    try:
        host = trt.HostScene.load(f["xml"], f["obj"], f["mtl"], f["basedir"])
    finally:
        os.dup2(saved, 1)
    out["staircase"] = measure(host, "staircase", "example-scenes-cg22/staircase (largest loadable cg22 scene)")
    host.close()
    if not args.no_stress:
        t0 = time.perf_counter()
        m = workloads.stress_mesh(2236)
        cam = m["camera"]
        host = trt.HostScene.from_arrays(m["v9"], m["mtl"], m["materials"], m["lights"], cam["eye"], cam["lookat"], cam["up"],
                                         cam["fovy"], 3840, 2160, vn9=m["vn9"])
        out["stress_10m"] = measure(host, "stress_10m", "BASELINE config 5: procedurally tessellated 10 M-triangle mesh (SURVEY §8d-5)")
        out["stress_10m"]["host_build_s"] = round(host.build_seconds, 2)
        out["stress_10m"]["setup_s"] = round(time.perf_counter() - t0, 1)
        host.close()
    return out


def standin_measurements(args):
    """Configs 1/2 on the labelled STAND-IN for the missing cornell-box.obj (Cornell shell + 100k-triangle displaced
    sphere, tinyraytracing_b200.scenes.cornell_with_standin): closest-hit Mrays/s of a 4 Mi config-2 batch and the
    512x512 16-spp render of config 1, reference behaviour (RR only) and with the added max-depth-5 truncation."""
    import torch

    import tinyraytracing_b200 as trt
    from tinyraytracing_b200 import scenes, workloads

    a = scenes.cornell_with_standin()
    host = trt.HostScene.from_arrays(a["v9"], a["mtl"], a["materials"], a["lights"], a["eye"], a["lookat"], a["up"], a["fovy"],
                                     a["width"], a["height"], vn9=a["vn9"], vt6=a["vt6"])
    dev = trt.DeviceScene(host, 0)

    def tracer(rays):
        ids, t = dev.trace_closest(rays)
        hp, pn = dev.hit_attributes(rays, ids, t)
        return ids, hp, pn

    n = 4 << 20
    rays = workloads.fixed_ray_batch(n, host.camera(), host.root_box(), tracer)
    d_rays = torch.from_numpy(rays).cuda()
    d_id = torch.empty(n, dtype=torch.int32, device="cuda")
    d_t = torch.empty(n, dtype=torch.float32, device="cuda")
    sp = torch.cuda.current_stream().cuda_stream
    warm_up_gpu(lambda: dev.trace_closest_async(d_rays.data_ptr(), n, d_id.data_ptr(), d_t.data_ptr(), args.flags, sp))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dev.trace_closest_async(d_rays.data_ptr(), n, d_id.data_ptr(), d_t.data_ptr(), args.flags, sp)
    e1.record()
    torch.cuda.synchronize()
    out = {"what": "STAND-IN, not reference geometry: test/back shell + displaced sphere", "tris": int(host.n_tris),
           "ref_depth": dev.stats()["ref_depth"], "closest_hit_mrays": 10 * n / (e0.elapsed_time(e1) * 1e-3) / 1e6}
    for key, md in (("render_512x512_16spp", 0), ("render_512x512_16spp_maxdepth5", 5)):
        dev.render(16, seed=1, max_depth=md)
        dev.reset_stats()
        dev.render(16, seed=1, max_depth=md)
        st = dev.stats()
        out[key] = {"ms": st["last_render_ms"], "spp_per_s": 16 / (st["last_render_ms"] * 1e-3),
                    "mrays_per_s": (st["rays_closest"] + st["rays_shadow"]) / (st["last_render_ms"] * 1e-3) / 1e6}
    dev.close()
    host.close()
    return out


def cpu_reference_render(name, w, h, tmp, full=False, spp=16):
    """The UNMODIFIED reference program (oracle/_ref/ref_cpu, stdin protocol of main.cpp:46-55) on the host cores.
    full: the config as it stands (config 1 is the reference's own CPU-runnable case).  Otherwise a bounded sample of the
    config: reduced resolution (same aspect) and spp = host cores so that its `omp parallel for` over samples
    (main.cpp:79-81) has work for every core, scaled linearly in pixel-samples.  Wall clock around the process (load +
    BVH build included, small for these scenes)."""
    from tinyraytracing_b200 import scenes

    exe = os.path.join(ROOT, "oracle", "_ref", "ref_cpu")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/ref_cpu not built"}
    cores = os.cpu_count() or 1
    if full:
        sw, sh = w, h
    else:
        sw, sh, spp = w // 5, h // 5, max(8, min(cores, 50))
    d = os.path.join(tmp, "cpu_ref_" + name)
    f = scenes.materialize(name, d, width=sw, height=sh)
    inp = "%s\n%s\n%s\n%s\n%d\n" % (f["basedir"], f["mtl"], f["xml"], f["obj"], spp)
    t0 = time.perf_counter()
    r = subprocess.run([exe], input=inp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, text=True, timeout=600)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        return {"unavailable": "ref_cpu exited %d" % r.returncode}
    pxs = sw * sh * spp / dt
    return {"kind": "reference", "cores": cores, "threads": min(50, spp),
            "sample": "%dx%d, %d spp, %.1f s wall%s" % (sw, sh, spp, dt, " (the whole config)" if full else ""),
            "pixel_samples_per_s": pxs, "spp_per_s_at_config": pxs / (w * h)}


def extra_measurements(args, tmp):
    """Other BASELINE configs, reported beside the headline (not the bench line's value)."""
    import torch

    import tinyraytracing_b200 as trt
    from tinyraytracing_b200 import scenes, workloads

    out = {}
    for name, (w, h, spp) in {"back": (512, 512, 16), "veach-mis": (1280, 720, 16), "staircase": (1280, 720, 8)}.items():
        f = scenes.materialize(name, os.path.join(tmp, "x_" + name), width=w, height=h)
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        os.dup2(devnull, 1)
        try:
            host = trt.HostScene.load(f["xml"], f["obj"], f["mtl"], f["basedir"])
        finally:
            os.dup2(saved, 1)
        dev = trt.DeviceScene(host, 0)

        def tracer(rays):
            ids, t = dev.trace_closest(rays)
            hp, pn = dev.hit_attributes(rays, ids, t)
            return ids, hp, pn

        n = 4 << 20
        rays = workloads.fixed_ray_batch(n, host.camera(), host.root_box(), tracer)
        d_rays = torch.from_numpy(rays).cuda()
        d_id = torch.empty(n, dtype=torch.int32, device="cuda")
        d_t = torch.empty(n, dtype=torch.float32, device="cuda")
        sp = torch.cuda.current_stream().cuda_stream
        for _ in range(2):
            dev.trace_closest_async(d_rays.data_ptr(), n, d_id.data_ptr(), d_t.data_ptr(), args.flags, sp)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            dev.trace_closest_async(d_rays.data_ptr(), n, d_id.data_ptr(), d_t.data_ptr(), args.flags, sp)
        e1.record()
        torch.cuda.synchronize()
        rec = {"closest_hit_mrays": 3 * n / (e0.elapsed_time(e1) * 1e-3) / 1e6, "tris": host.n_tris}
        dev.render(1, seed=1)  # warm-up (allocates the wavefront buffers)
        dev.reset_stats()
        dev.render(spp, seed=1)
        st = dev.stats()
        rays_total = st["rays_closest"] + st["rays_shadow"]
        rec["render"] = {"width": w, "height": h, "spp": spp, "ms": st["last_render_ms"],
                         "spp_per_s": spp / (st["last_render_ms"] * 1e-3),
                         "mrays": rays_total / (st["last_render_ms"] * 1e-3) / 1e6,
                         "rays_closest": st["rays_closest"], "rays_shadow": st["rays_shadow"],
                         "kernel_launches": st["kernel_launches"]}
        out[name] = rec
        dev.close()
    return out


_REAL_STDOUT = None


def quiet_stdout():
    """Everything except the ONE JSON line goes to stderr: libraries (NCCL's version banner, the loaders' chatter
    kept from the reference) write to fd 1 from C code."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (line + "\n").encode())


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rays", type=int, default=N_RAYS)
    ap.add_argument("--flags", type=int, default=0, help="TRT_TRACE_* traversal flags (0 = default layout)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--extra", action="store_true", help="also measure the other scenes and the render path")
    ap.add_argument("--no-render", action="store_true", help="skip the config-1/3/4 render measurements and the other scenes")
    ap.add_argument("--no-stress", action="store_true", help="skip the 10 M-triangle stress mesh (about 40 s of host-side build)")
    ap.add_argument("--config4", action="store_true", help="also render config 4 (staircase 1920x1080) when N != 8")
    ap.add_argument("--spp3", type=int, default=256, help="spp of config 3 (veach-mis 1280x720)")
    ap.add_argument("--spp4", type=int, default=1024, help="spp of config 4 (staircase 1920x1080)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    return run_reference(args) if args.impl == "reference" else run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
