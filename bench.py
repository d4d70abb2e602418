#!/usr/bin/env python3
"""Benchmark of the path-tracing hot path: Mrays/s (incl. secondary rays) on BASELINE.json's headline configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload mesh1m|cornell] [--mode job|pass]

--mode job (default; what the driver's BENCH / SCALE runs execute): one STEP = the WHOLE job of the workload, end to
end, the way main.cpp:181-236 runs it -- every pixel sample of the frame (BASELINE config 5: 995,328-triangle mesh,
1920 wide, max-depth 3, -m 4, 128 spp; config 4 with --workload cornell: 3480 wide), split over the N ranks by pixel
sample (rank g renders samples i = g mod N, scene replicated), ONE ncclReduce(sum) of the W*H*4 float accumulation
buffers onto rank 0 through the library's own communicator (trn_comm_* / trn_render_rank), and -- in the e2e leg --
ONE device->host copy of the image plus the host tone map (/pps, exposure, gamma). Total work is fixed as N grows:
"scaling": "strong". reduce / D2H / tone-map times are reported separately.

--mode pass (kernel A/B during development): one STEP = one pixel-sample pass over the frame per rank (weak scaling),
the round-1 measurement.

A "ray" is Stats::num_rays (pathtracer.cpp:17-21): shadow queries are traversed but not counted.
The line printed by rank 0 follows the driver's contract; DESIGN.md "Measurement" says how each field is made.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    "mesh1m": dict(n=288, width=1920, max_depth=3, mc_samples=4, pixel_samples=128,
                   desc="S-mesh1M displaced cube-sphere, 995328 triangles, width 1920, max-depth 3, -m 4, 128 spp"),
    "cornell": dict(n=0, width=3480, max_depth=3, mc_samples=4, pixel_samples=128,
                    desc="cornell_box (36 triangles), width 3480, max-depth 3, -m 4, 128 spp"),
}


def load_scene(name):
    from turner_b200 import scenes
    w = WORKLOADS[name]
    return scenes.cubesphere(w["n"]) if name == "mesh1m" else scenes.fixture("cornell_box")


def make_config_dict(name, mode):
    w = WORKLOADS[name]
    if mode == "job":
        return {"workload": w["desc"], "mode": "job",
                "step": "the whole job: all %d pixel samples of the frame, split over the ranks by pixel sample; one "
                        "ncclReduce(sum) of W*H*4 f32 onto rank 0; e2e adds one D2H of the image and the host tone map" % w["pixel_samples"],
                "sample_split": "rank g of N renders samples i = g (mod N); scene replicated",
                "l2": "inputs larger than L2: every step streams several GB of ray waves (48 B per ray) through a 126 MB L2; "
                      "a 256 MiB memset between steps flushes it as well",
                "seed": 1}
    return {"workload": w["desc"], "mode": "pass", "step": "one pixel-sample pass over the full frame per rank",
            "sample_split": "rank g renders sample (step*N+g) mod %d; one reduce(sum) of W*H*4 f32 at the end" % w["pixel_samples"],
            "l2": "flushed between steps (256 MiB memset inside the timed region)", "seed": 1}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region. Polls NVML (what nvidia-smi reads: clocks.sm, clocks.max.sm,
    clocks_event_reasons.*) every 5 ms from a thread; falls back to the `nvidia-smi -lms` line of B200_PROFILING.md when
    pynvml is unavailable."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.index = index
        self.samples = []  # (t, sm_mhz, reasons_bitmask)
        self.stop_flag = False
        self.thread = None
        self.proc = None
        self.max_mhz = None
        self.source = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.source = "nvidia-smi"
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        while not self.stop_flag:
            try:
                mhz = float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                why = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((time.perf_counter(), mhz, why))
            except Exception:
                pass
            time.sleep(0.005)

    def _pump(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                mhz, mx = float(f[0]), float(f[1])
            except ValueError:
                continue
            self.max_mhz = mx
            why = 0
            for bit, col in ((0x8, 3), (0x40, 4), (0x20, 5), (0x4, 6)):
                if f[col].lower().startswith("active"):
                    why |= bit
            self.samples.append((time.perf_counter(), mhz, why))

    def stop(self, t0=None, t1=None):
        """summary over the samples taken in [t0, t1] (perf_counter stamps of the timed region)"""
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        if self.thread:
            self.thread.join(timeout=2)
        inside = [s for s in self.samples if t0 is None or (t0 <= s[0] <= t1)]
        window = "timed region"
        if not inside and self.samples:  # region shorter than one sampling period: nearest samples
            mid = 0.5 * ((t0 or 0) + (t1 or 0))
            inside = sorted(self.samples, key=lambda s: abs(s[0] - mid))[:3]
            window = "nearest samples"
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no clock samples (%s)" % self.source], "samples": 0}
        mask = 0
        for s_ in inside:
            mask |= s_[2]
        return {"sm_mhz": float(np.median([s_[1] for s_ in inside])), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(v for k, v in self.REASONS.items() if mask & k), "samples": len(inside),
                "source": self.source, "window": window}


def alg_bytes(inner, leaf_nodes, tri_tests, queries):
    """B_alg = 8*n_inner + 8*n_leafnodes + 64*n_tri_tests + (32 ray in + 16 hit out) per query"""
    return 8 * inner + 8 * leaf_nodes + 64 * tri_tests + 48 * queries


def requested_bytes(pc, queries):
    """bytes the pooled schedule itself requests through L1 (trn_stats.trace_pooled, include/turner_b200.h): 16 per walk
    step (node pair), 4*16 per chunk (the four plane records behind the leaf's references, one 64-byte block) + 16 for the id
    vector of a chunk with a pre-filter survivor (at most one per exact test), 32 per exact test (+32 with the cold record),
    16 per stack push / pop (local memory), 32 ray in + 16 hit out per query"""
    steps, chunks, tris, exact, cold, push, pop = [int(x) for x in list(pc)[:7]]
    return 16 * steps + 64 * chunks + 16 * min(chunks, exact) + 32 * exact + 32 * cold + 16 * (push + pop) + 48 * queries


def cpu_reference_run(name, sc, nodes, box, steps=1, warmup=0, budget_s=100.0):
    """the reference's own CPU implementation (oracle/_ref: its kdtree.cpp + pathtracer.cpp) on all host cores, on a
    bounded sample of the workload AT ITS OWN WIDTH: every s-th image row (a row is the reference's unit of work,
    main.cpp:194), one pixel sample; s is chosen from a probe so that the K+W steps take about budget_s seconds.
    Falls back to the oracle port when oracle/_ref is not built."""
    from oracle import bindings as ob
    w = WORKLOADS[name]
    cores = os.cpu_count() or 1
    W = w["width"]
    vals, rays_total, t_total = [], 0, 0.0
    nsteps = max(1, steps + warmup)
    if ob.ref_available():
        kind = "reference"
        # with `nodes`: the tree is loaded through the reference's own serialize() hook (what main.cpp:147-152 does with
        # kdtree.cache); it is node-for-node the tree its builder makes (tests/test_host.py), which takes ~60 s at 1M
        r = (ob.RefScene(sc["vertices"], sc["normals"], sc["diffuse"], nodes=nodes, box=box) if nodes is not None
             else ob.RefScene(sc["vertices"], sc["normals"], sc["diffuse"]))
        cam = ob.ref_camera(sc)
        cfg = ob.ref_config(sc, W, w["max_depth"], w["mc_samples"], 1, num_threads=cores)
        H = int(W / 1.0)
        # probe: 1/64 of the rows -> rays/s estimate -> row stride for the budget
        cfg.row_begin, cfg.row_stride = 17, 64
        _, _, _, st = r.render(cam, cfg)
        probe_s = max(st.runtime_ms / 1e3, 1e-3)
        full_s = probe_s * 64.0
        stride = int(min(64, max(1, np.ceil(full_s * nsteps / budget_s))))
        for it in range(warmup + steps):
            cfg.row_begin, cfg.row_stride = it % stride, stride
            _, _, _, st = r.render(cam, cfg)
            if it >= warmup:
                rays_total += st.num_rays
                t_total += st.runtime_ms / 1e3
        sample = ("same scene / width %d / max-depth %d / -m %d; one step = 1 of %d pixel samples on every %s image row "
                  "(%d of %d rows; Mrays/s does not depend on either)"
                  % (W, w["max_depth"], w["mc_samples"], w["pixel_samples"],
                     "" if stride == 1 else "%d-th" % stride, (H + stride - 1) // stride, H))
    else:
        kind = "port"
        o = (ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"], nodes=nodes, box=box) if nodes is not None
             else ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"]))
        Wc = max(64, W // 4)
        cfg = ob.make_cfg(sc, Wc, w["max_depth"], w["mc_samples"], 1, num_threads=cores)
        for it in range(warmup + steps):
            _, _, st = o.render(cfg)
            if it >= warmup:
                rays_total += st.num_rays
                t_total += st.runtime_ms / 1e3
        sample = "oracle port; same scene / max-depth / -m, width %d instead of %d, 1 pixel sample per step" % (Wc, W)
    value = rays_total / max(t_total, 1e-9) / 1e6
    return {"value": value, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample,
            "rays": int(rays_total), "seconds": t_total}


def gather_peaks(api, device):
    """measured ceilings of random 16-byte gathers (the traversal kernels' access shape) on this device"""
    out = {}
    for key, size, mode in (("l1_gbs", 32 << 10, 0), ("l2_gbs", 32 << 20, 1), ("l2_96m_gbs", 96 << 20, 1), ("hbm_gbs", 4 << 30, 1)):
        try:
            out[key] = api.measure_gather_peak(size, mode, device)
        except Exception as e:  # noqa: BLE001
            out[key] = None
            out[key + "_error"] = str(e)
    out["how"] = ("trn_measure_gather_peak: independent random 16-byte __ldg gathers, one line per lane, 8 in flight per "
                  "thread, best of 3 after a warm-up run; working set 32 KiB (lives in every SM's L1) / 32 MiB (L2) / 96 MiB (about the "
                  "scene's hot set: L2 with misses) / 4 GiB (HBM)")
    return out


def emit(line):
    """the ONE JSON line of the driver's contract, on the process's real stdout"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    # stdout carries exactly one JSON line: whatever libraries print there (NCCL's version banner, loader chatter) goes to
    # stderr -- file descriptor 1 is pointed at stderr for the run, the line is written to the saved descriptor
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mesh1m", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="job", choices=["job", "pass"])
    ap.add_argument("--kd-builder", default="auto", choices=["auto", "gpu", "host"],
                    help="gpu: device-built kd-tree (trn_scene_create_gpu); host: the reference's tree, node for node; auto: gpu "
                         "for the mesh (renders 44 %% faster on it), host for cornell_box (its reference tree is one leaf, which "
                         "the GPU prefers: 4899 vs 3362 Mrays/s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 4 if args.mode == "job" else 8
    w = WORKLOADS[args.workload]
    if args.kd_builder == "auto":
        args.kd_builder = "gpu" if args.workload == "mesh1m" else "host"
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = make_config_dict(args.workload, args.mode)
    scaling = "strong" if args.mode == "job" else "weak"

    if args.impl == "reference":
        # the reference arm: rank 0 alone times the reference's CPU path; other ranks exit
        if rank != 0:
            return 0
        sc = load_scene(args.workload)
        # nothing of turner_b200 on this arm: the reference builds its own kd-tree with its own builder (~1 min at 1M
        # triangles, untimed like "Loading time" in the reference's report) and renders with its own trace()
        base = cpu_reference_run(args.workload, sc, None, None, steps=max(1, args.steps), warmup=min(args.warmup, 1))
        line = {"impl": "reference", "metric": "Mrays/s (incl. secondary)", "value": base["value"], "unit": "Mrays/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * base["seconds"] / max(1, args.steps), "higher_is_better": True, "scaling": scaling,
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": base["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    import torch
    import torch.distributed as dist
    from turner_b200 import api, dist as tdist

    if api.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        tdist.init_process_group("nccl")
    os.environ.setdefault("TRN_BUILD_THREADS", str(max(1, (os.cpu_count() or 1) // world)))

    sc = load_scene(args.workload)
    scene = api.Scene.from_dict(sc, builder=args.kd_builder, device=local_rank)
    config["kd_builder"] = ("device build (binned SAH, trn_scene_create_gpu): hits equal the reference tree's, checked below"
                            if args.kd_builder == "gpu" else "host build: the reference's tree, node for node")
    pps = w["pixel_samples"]
    cam, cfg = api.make_config(sc, w["width"], max_depth=w["max_depth"], mc_samples=w["mc_samples"], pixel_samples=pps, seed=1)
    H, W = cfg.height, cfg.width
    flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
    stream = torch.cuda.current_stream()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # polls through warm-up and the timed region; only samples inside the region are reported

    def allreduce_max(x):
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allreduce_sum(x):
        t = torch.tensor([x], device="cuda", dtype=torch.int64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())

    tot = dict(rays=0, shadow=0, launches=0, ms_trace=0.0, ms_shadow=0.0, ms_shade=0.0, ms_other=0.0, trace_launches=0,
               trace_queries=0, shadow_launches=0)
    extra = {}

    if args.mode == "job":
        # the library's own communicator: rank 0 makes the NCCL id, torch.distributed only carries its 128 bytes
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if world > 1:
            if rank == 0:
                idt.copy_(torch.from_numpy(api.Comm.unique_id()))
            dist.broadcast(idt, src=0)
        comm = api.Comm(idt.cpu().numpy(), world, rank, local_rank)  # one rank: no NCCL involved
        cfg.sample_begin, cfg.sample_stride = 0, 1
        host_pinned = torch.zeros(H, W, 4, dtype=torch.float32).pin_memory() if rank == 0 else None
        host_img = host_pinned.numpy() if rank == 0 else None
        final_img = np.zeros((H, W, 4), np.float32) if rank == 0 else None

        def job(out):
            """one whole job through the C ABI; out = host image buffer (rank 0) or None (result stays on the device)"""
            return scene.render_rank(comm, cam, cfg, out=out)

        for i in range(args.warmup):
            job(None)
        torch.cuda.synchronize()
        # ---- value: K jobs, result resident on rank 0's device; barrier + synchronize on both sides, CUDA events
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        t_region0 = time.perf_counter()
        dev_ms, red_ms = 0.0, 0.0
        for i in range(args.steps):
            st = job(None)
            flush.zero_()
            tot["rays"] += st.rays
            tot["shadow"] += st.shadow_rays
            tot["launches"] += st.launches + (1 if world > 1 else 0)
            tot["trace_launches"] += st.trace_launches
            tot["trace_queries"] += st.trace_queries
            tot["shadow_launches"] += st.shadow_launches
            dev_ms += st.ms_render
            red_ms += st.ms_reduce
        ev1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t_region1 = time.perf_counter()
        ms = ev0.elapsed_time(ev1)
        clocks = sampler.stop(t_region0, t_region1) if rank == 0 else None
        ms_max = allreduce_max(ms)
        rays_all = allreduce_sum(tot["rays"])
        value = rays_all / (ms_max / 1e3) / 1e6
        extra["job"] = {"render_ms_per_job_rank0": (dev_ms - red_ms) / args.steps, "reduce_ms_per_job_rank0": red_ms / args.steps,
                        "reduce_bytes": int(H * W * 16) if world > 1 else 0,
                        "samples_per_rank": (pps + world - 1) // world}
        # ---- e2e: the same K jobs with the image delivered: + D2H into pinned host memory + host tone map (rank 0)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_rays, d2h_ms, tm_ms, red2_ms = 0, 0.0, 0.0, 0.0
        for i in range(args.steps):
            st = job(host_img)
            e2e_rays += st.rays
            d2h_ms += st.ms_d2h
            red2_ms += st.ms_reduce
            if rank == 0:
                ta = time.perf_counter()
                api.tonemap(host_img, pps, out=final_img)
                tm_ms += 1e3 * (time.perf_counter() - ta)
            if world > 1:
                dist.barrier()  # the job is done when rank 0 holds the final image
        torch.cuda.synchronize()
        dt = allreduce_max(time.perf_counter() - t0)
        e2e = {"value": allreduce_sum(e2e_rays) / dt / 1e6, "unit": "Mrays/s",
               "h2d_bytes_per_step": int(api.C.sizeof(api.Camera) + api.C.sizeof(api.RenderConfig)),
               "d2h_bytes_per_step": int(H * W * 16),
               "ms_per_step": 1e3 * dt / args.steps, "reduce_ms": red2_ms / args.steps, "d2h_ms": d2h_ms / args.steps,
               "tonemap_ms": tm_ms / args.steps,
               "note": "trn_render / trn_render_rank: camera+config in, ONE ncclReduce, ONE D2H of the W*H*4 f32 image into "
                       "pinned host memory, host tone map (trn_tonemap), all inside the timed region; scene resident"}
        # per-kernel durations of one job share (rank 0's), one event pair per launch (shadow waves then run serially)
        prof_steps = 1

        def profiled():
            accum = torch.zeros(H, W, 4, device="cuda", dtype=torch.float32)
            c2 = api.copy_config(cfg)
            c2.sample_begin, c2.sample_stride = rank, world
            return scene.render_device(cam, c2, accum.data_ptr(), stream.cuda_stream, device=local_rank, want_stats=True)
    else:
        accum = torch.zeros(H, W, 4, device="cuda", dtype=torch.float32)

        def step(i, stats=True):
            cfg.sample_begin = (i * world + rank) % pps
            cfg.sample_stride = pps  # exactly one sample index per step
            return scene.render_device(cam, cfg, accum.data_ptr(), stream.cuda_stream, device=local_rank, want_stats=stats)

        for i in range(args.warmup):
            step(i)
        torch.cuda.synchronize()
        accum.zero_()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        t_region0 = time.perf_counter()
        for i in range(args.steps):
            st = step(args.warmup + i)
            flush.zero_()
            tot["rays"] += st.rays
            tot["shadow"] += st.shadow_rays
            tot["launches"] += st.launches
            tot["trace_launches"] += st.trace_launches
            tot["trace_queries"] += st.trace_queries
            tot["shadow_launches"] += st.shadow_launches
        tdist.reduce_accum(accum, root=0)
        ev1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t_region1 = time.perf_counter()
        ms = ev0.elapsed_time(ev1)
        clocks = sampler.stop(t_region0, t_region1) if rank == 0 else None
        ms_max = allreduce_max(ms)
        rays_all = allreduce_sum(tot["rays"])
        value = rays_all / (ms_max / 1e3) / 1e6
        # ---- e2e: frames through the asynchronous host-buffer call (D2H of frame k next to the render of k+1)
        bufs = [torch.zeros(H, W, 4, dtype=torch.float32).pin_memory() for _ in range(2)]
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        e2e_rays, pending = 0, None
        for i in range(args.steps):
            cfg.sample_begin = ((args.warmup + i) * world + rank) % pps
            cfg.sample_stride = pps
            jobh = scene.render_async(cam, cfg, bufs[i & 1].numpy(), device=local_rank)
            if pending is not None:
                e2e_rays += pending.wait()[1].rays
            pending = jobh
        e2e_rays += pending.wait()[1].rays
        torch.cuda.synchronize()
        dt = allreduce_max(time.perf_counter() - t0)
        e2e = {"value": allreduce_sum(e2e_rays) / dt / 1e6, "unit": "Mrays/s",
               "h2d_bytes_per_step": int(api.C.sizeof(api.Camera) + api.C.sizeof(api.RenderConfig)),
               "d2h_bytes_per_step": int(H * W * 16),
               "note": "trn_render_async/trn_wait: camera+config in, W*H*4 f32 image out to pinned host memory every step"}
        prof_steps = args.steps

        def profiled():
            scratch = torch.zeros(H, W, 4, device="cuda", dtype=torch.float32)
            agg = None
            for i in range(args.steps):
                cfg.sample_begin = ((args.warmup + i) * world + rank) % pps
                cfg.sample_stride = pps
                st_ = scene.render_device(cam, cfg, scratch.data_ptr(), stream.cuda_stream, device=local_rank, want_stats=True)
                flush.zero_()
                if agg is None:
                    agg = st_
                else:
                    for k in ("ms_trace", "ms_shadow", "ms_shade", "ms_other", "ms_render", "trace_launches", "trace_queries",
                              "shadow_rays", "shadow_launches"):
                        setattr(agg, k, getattr(agg, k) + getattr(st_, k))
            return agg

    if rank != 0:
        if world > 1:
            dist.barrier()  # rank 0 finishes its single-GPU roofline / baseline legs first
            dist.destroy_process_group()
        return 0

    # ---- the reference-identical host tree: build time beside the device build, primary hits of the two trees compared,
    # and the tree the CPU baseline traverses (loaded through the reference's own serialize() hook)
    kd = {"builder": args.kd_builder, "build_ms": scene.info.build_ms, "height": int(scene.height),
          "leaf_refs": int(scene.info.num_leaf_refs), "cut_nodes": int(scene.info.num_cut_nodes)}
    host_scene = scene
    if args.kd_builder == "gpu" and not (args.no_cpu_baseline and args.no_roofline):
        host_scene = api.Scene.from_dict(sc)
        kd["host_build_ms"] = host_scene.info.build_ms
        kd["host_build_threads"] = int(os.environ.get("TRN_BUILD_THREADS", "0")) or (os.cpu_count() or 1)
        c4 = api.copy_config(cfg)
        c4.pixel_samples, c4.sample_begin, c4.sample_stride = 1, 0, 1
        ia, ra_ = scene.primary_hits(cam, c4, device=local_rank)
        ib, rb_ = host_scene.primary_hits(cam, c4, device=local_rank)
        kd["primary_hits_compared"] = int(ia.size)
        kd["primary_id_differences_vs_reference_tree"] = int((ia != ib).sum())
        kd["primary_rst_bit_differences"] = int((ra_.view(np.uint32) != rb_.view(np.uint32)).any(-1).sum())
    roofline = None
    if not args.no_roofline:
        # ---- per-kernel durations for the roofline leg: one CUDA-event pair per kernel launch on the launching stream.
        # Untimed for `value`: while every launch carries its own events the shadow waves are not overlapped with the
        # next closest-hit wave, so the spans add up and can be compared with the serialised ncu launch list in profiles/.
        api.set_profiling(True)
        pst = profiled()
        torch.cuda.synchronize()
        api.set_profiling(False)
        # ---- what the kernels visit / request: instrumented twins on one pixel-sample pass (untimed)
        api.set_counting(True)
        c3 = api.copy_config(cfg)
        c3.sample_begin, c3.sample_stride = 0, pps
        scratch = torch.zeros(H, W, 4, device="cuda", dtype=torch.float32)
        cst = scene.render_device(cam, c3, scratch.data_ptr(), stream.cuda_stream, device=local_rank, want_stats=True)
        # the reference-shaped counters that define B_alg are taken on the REFERENCE's tree (host build); the
        # reference view of a device-built tree has its empty-space cuts dropped and says nothing about either program
        rst = cst if host_scene is scene else host_scene.render_device(cam, c3, scratch.data_ptr(), stream.cuda_stream,
                                                                     device=local_rank, want_stats=True)
        api.set_counting(False)
        torch.cuda.synchronize()
        del scratch
        q, sq = max(cst.trace_queries, 1), max(cst.shadow_rays, 1)
        rq, rsq = max(rst.trace_queries, 1), max(rst.shadow_rays, 1)
        pooled = sum(cst.trace_pooled) > 0
        peak_hbm, peak_src = measured_peaks()
        peaks = gather_peaks(api, local_rank)
        b_alg = alg_bytes(rst.trace_inner, rst.trace_leaf_nodes, rst.trace_tri_tests, rst.trace_queries) / rq
        b_req = requested_bytes(cst.trace_pooled, cst.trace_queries) / q if pooled else None
        ms_trace = max(pst.ms_trace, 1e-9)
        # the bytes the dominant kernel REQUESTS per second (all of them pass the L1 tag/data pipe as divergent 16-byte
        # gathers) against the measured ceiling of exactly that access shape served from L1
        if pooled and peaks.get("l1_gbs"):
            achieved = b_req * pst.trace_queries / ms_trace / 1e6
            peak, bound, unit_note = peaks["l1_gbs"], "l1tex", "requested bytes (trn_stats.trace_pooled) / kernel time"
        else:
            achieved = b_alg * pst.trace_queries / ms_trace / 1e6
            peak, bound, unit_note = peak_hbm, "hbm", "algorithmic bytes / kernel time"
        traffic, kernel_name, ncu_fig = None, ("trace_pooled_kernel<0>" if pooled else "trace_persistent_ww_kernel<0,false>"), None
        unit = "GB/s"
        flat_records = int(getattr(pst, "flat_records", 0))
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        flat = None
        if flat_records > 0:
            # One-leaf scene: the brute-force kernel (traverse_flat.cuh). Its triangle records are kernel parameters (constant
            # bank) and its rays stream through once, so no memory level binds it: it is bound by the FP32 pipes. Algorithmic
            # work of a query = its scan: every record costs 3 FMUL + 7 FFMA + 1 FADD + 1 MUFU (18 flop on the FMA pipe; the 9
            # compares run on the ALU pipe beside it) -- the exact tests of the ~2 candidates per ray are not counted.
            nt = (flat_records + 3) // 4 * 4
            kernel_name = "trace_flat_kernel<0,%d>" % nt
            flop_q = 18.0 * flat_records
            achieved = flop_q * pst.trace_queries / ms_trace / 1e9
            peak = 148 * 128 * 2 * sm_mhz / 1e6
            bound, unit = "fp32", "TFLOP/s"
            unit_note = ("scan flops (18 per record x %d records per query) / kernel time against the FFMA peak 148 SMs x 128 lanes x 2 x "
                         "SM clock under load (MEASURED_PEAKS.json holds no fp32 figure)" % flat_records)
            flat = {"records_per_query": flat_records, "flop_per_query": flop_q,
                    # the FMA pipe issues one warp instruction per 2 cycles per SM sub-partition (B300_MICROARCH.md); the scan
                    # alone needs 11 FMA-pipe instructions per record
                    "fma_pipe_busy_frac_scan_only": 11.0 * flat_records / 32.0 * 2.0 * pst.trace_queries / (ms_trace * 1e-3)
                                                    / (148 * 4 * sm_mhz * 1e6),
                    "queries_per_s": pst.trace_queries / (ms_trace * 1e-3)}
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(prof):
            try:
                ent = json.load(open(prof)).get(args.workload, {})
                traffic = ent.get("trace_closest_dram_bytes_per_launch")
                ncu_fig = ent.get("ncu")
            except Exception:
                traffic = None
        fp32_rate = 37.0 * (rst.trace_tri_tests / rq) * pst.trace_queries / ms_trace / 1e6
        roofline = {
            "bound": bound, "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": unit,
            "frac": achieved / peak, "traffic": traffic, "how": unit_note, "brute_force": flat,
            "peaks_measured": peaks,
            "requested_bytes_per_query": b_req,
            "requested_per_query": ({k: v / q for k, v in zip(("walk_steps", "chunks", "tri_pretests", "exact_tests", "cold_records",
                                                               "stack_pushes", "stack_pops", "leaves", "walk_steps_at_cuts"), list(cst.trace_pooled)[:9])} if pooled else None),
            "l2_gather_frac": (achieved / peaks["l2_gbs"]) if (pooled and peaks.get("l2_gbs")) else None,
            # secondary, SURVEY 8(d)'s definition: algorithmic bytes of the reference-shaped schedule against the HBM copy
            # peak. The scene is L2-resident, so this is a normalised work rate, not a DRAM utilisation (it can exceed 1).
            "hbm_normalised": {"alg_bytes_per_query": b_alg, "achieved": b_alg * pst.trace_queries / ms_trace / 1e6,
                               "peak": peak_hbm, "peak_source": peak_src,
                               "frac": b_alg * pst.trace_queries / ms_trace / 1e6 / peak_hbm},
            "ncu": ncu_fig,
            "per_query": {"inner": rst.trace_inner / rq, "leaf_nodes": rst.trace_leaf_nodes / rq, "tri_tests": rst.trace_tri_tests / rq,
                          "note": "reference-shaped early-exit schedule on the reference's tree: the n_* of B_alg"},
            "per_query_actual": {"inner": cst.trace_actual_inner / q, "leaf_nodes": cst.trace_actual_leaf_nodes / q,
                                 "tri_tests": cst.trace_actual_tri_tests / q,
                                 "note": "visits of the per-ray schedule on the device layout (empty-space cuts kept)"},
            "launches": int(pst.trace_launches), "avg_launch_ms": pst.ms_trace / max(pst.trace_launches, 1),
            "share_of_step": pst.ms_trace / max(pst.ms_render, 1e-9),
            "profiled_ms_per_step": pst.ms_render / prof_steps,
            "kernel_ms": {k: getattr(pst, k) for k in ("ms_trace", "ms_shadow", "ms_shade", "ms_other")},
            "shadow_kernel": {
                "alg_bytes_per_query": alg_bytes(rst.shadow_inner, rst.shadow_leaf_nodes, rst.shadow_tri_tests, rst.shadow_rays) / rsq,
                "requested_bytes_per_query": (requested_bytes(cst.shadow_pooled, cst.shadow_rays) / sq) if pooled else None,
                "achieved_requested": (requested_bytes(cst.shadow_pooled, cst.shadow_rays) / sq * pst.shadow_rays
                                       / max(pst.ms_shadow, 1e-9) / 1e6) if pooled else None},
            "fp32_tri_test_rate_gflops": fp32_rate,
            # FP32 FMA peak (148 SMs x 128 lanes x 2 flop x SM clock under load); the tests run WITHOUT FMA contraction
            # (bit-exactness), so 50 % is the ceiling of this ratio
            "fp32_peak_gflops": 148 * 128 * 2 * sm_mhz / 1e3,
            "fp32_frac": fp32_rate / (148 * 128 * 2 * sm_mhz / 1e3), "fp32_clock_mhz": sm_mhz,
        }

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        b = cpu_reference_run(args.workload, sc, host_scene.nodes(), np.array(host_scene.info.box, np.float32), budget_s=20.0)
        cpu_baseline = {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")}

    scene.refresh_info()
    line = {
        "metric": "Mrays/s (incl. secondary)", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
        "rays_per_step": rays_all / args.steps, "shadow_rays_per_step_rank0": tot["shadow"] / args.steps,
        "kd_build_ms": scene.info.build_ms, "kd_height": int(scene.height), "triangles": int(scene.num_triangles), "kd": kd,
        "time_to_image_ms": scene.info.build_ms + scene.info.upload_ms + e2e.get("ms_per_step", ms_max / args.steps),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(tot["launches"]), "roofline": roofline,
        "cpu_baseline": cpu_baseline,
    }
    line.update(extra)
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
