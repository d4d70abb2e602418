#!/usr/bin/env python3
"""Benchmark of the path-tracing hot path: Mrays/s (incl. secondary rays) on BASELINE.json's headline config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload mesh1m|cornell]

Workload (N=1 and weak scaling): BASELINE config 5 -- the ~1M-triangle displaced cube-sphere (995,328 triangles),
width 1920, aspect 1, max-depth 3, -m 4, 128 pixel samples. One STEP = one pixel-sample pass over the whole
1920x1920 frame (3.69 M primaries and their full ray trees) on every rank; rank g renders sample index
(step*N + g) mod 128, i.e. the sample split of SURVEY 8(e); ONE reduce(sum) of the W*H*4 float accumulation
buffers onto rank 0 closes the timed region (torch.distributed / NCCL). A "ray" is Stats::num_rays
(pathtracer.cpp:17-21): shadow queries are traversed but not counted.

The line printed by rank 0 follows the driver's contract; see DESIGN.md "Measurement" for how each field is made.
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
    # name: (scene builder, width, max_depth, mc_samples, pixel_samples, cpu-sample width)
    "mesh1m": dict(n=288, width=1920, max_depth=3, mc_samples=4, pixel_samples=128, cpu_width=480,
                   desc="S-mesh1M displaced cube-sphere, 995328 triangles, width 1920, max-depth 3, -m 4, 128 spp"),
    "cornell": dict(n=0, width=3480, max_depth=3, mc_samples=4, pixel_samples=128, cpu_width=256,
                    desc="cornell_box (36 triangles), width 3480, max-depth 3, -m 4, 128 spp"),
}


def load_scene(name):
    from turner_b200 import scenes
    w = WORKLOADS[name]
    return scenes.cubesphere(w["n"]) if name == "mesh1m" else scenes.fixture("cornell_box")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region. Polls NVML (what nvidia-smi reads: clocks.sm, clocks.max.sm,
    clocks_event_reasons.*) every 5 ms from a thread -- the timed region of a short run is ~100 ms, below nvidia-smi's
    own start-up time; falls back to the `nvidia-smi -lms` line of B200_PROFILING.md when pynvml is unavailable."""

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
            # honour CUDA_VISIBLE_DEVICES-free boxes: LOCAL_RANK == NVML index on the driver's single-node runs
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
        if not inside and self.samples:  # region shorter than one sampling period: nearest samples (GPU under load: warm-up precedes)
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


def cpu_reference_run(name, sc, nodes, box, steps=1, warmup=0):
    """the reference's own CPU implementation (oracle/_ref: its kdtree.cpp + pathtracer.cpp) on the host cores,
    on a bounded sample of the workload; falls back to the oracle port when oracle/_ref is not built."""
    from oracle import bindings as ob
    w = WORKLOADS[name]
    cores = os.cpu_count() or 1
    W = w["cpu_width"]
    # ~10 s of host work for one step of the mesh at 16 cores; a run of K steps keeps its total near 30 s
    cpu_pps = max(1, min(16, 48 // max(1, steps + warmup)))
    sample = ("same scene / max-depth %d / -m %d, width %d instead of %d, %d of %d pixel samples per step "
              "(Mrays/s does not depend on either)" % (w["max_depth"], w["mc_samples"], W, w["width"], cpu_pps, w["pixel_samples"]))
    vals, rays_total, t_total = [], 0, 0.0
    if ob.ref_available():
        kind = "reference"
        # the tree is loaded through the reference's own serialize() hook (what main.cpp:147-152 does with
        # kdtree.cache); it is node-for-node the tree its builder makes (tests/test_host.py), which takes ~60 s at 1M
        r = (ob.RefScene(sc["vertices"], sc["normals"], sc["diffuse"], nodes=nodes, box=box) if nodes is not None
             else ob.RefScene(sc["vertices"], sc["normals"], sc["diffuse"]))
        cam = ob.ref_camera(sc)
        cfg = ob.ref_config(sc, W, w["max_depth"], w["mc_samples"], cpu_pps, num_threads=cores)
        for it in range(warmup + steps):
            _, _, _, st = r.render(cam, cfg)
            if it >= warmup:
                rays_total += st.num_rays
                t_total += st.runtime_ms / 1e3
    else:
        kind = "port"
        o = (ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"], nodes=nodes, box=box) if nodes is not None
             else ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"]))
        cfg = ob.make_cfg(sc, W, w["max_depth"], w["mc_samples"], cpu_pps, num_threads=cores)
        for it in range(warmup + steps):
            _, _, st = o.render(cfg)
            if it >= warmup:
                rays_total += st.num_rays
                t_total += st.runtime_ms / 1e3
    value = rays_total / max(t_total, 1e-9) / 1e6
    return {"value": value, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample,
            "rays": int(rays_total), "seconds": t_total}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mesh1m", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": w["desc"], "step": "one pixel-sample pass over the full frame per rank",
              "sample_split": "rank g renders sample (step*N+g) mod %d; one reduce(sum) of W*H*4 f32 at the end" % w["pixel_samples"],
              "l2": "flushed between steps (256 MiB memset inside the timed region)", "seed": 1,
              "kernel_timing": "value: CUDA events around the K timed steps, no per-kernel events; roofline.kernel_ms: the same K "
                               "steps once more with one event pair per launch (shadow/closest-hit waves then run serially)"}

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
                "ms_per_step": 1e3 * base["seconds"] / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": base["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
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
    scene = api.Scene.from_dict(sc)
    pps = w["pixel_samples"]
    cam, cfg = api.make_config(sc, w["width"], max_depth=w["max_depth"], mc_samples=w["mc_samples"], pixel_samples=pps, seed=1)
    H, W = cfg.height, cfg.width
    accum = torch.zeros(H, W, 4, device="cuda", dtype=torch.float32)
    flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
    stream = torch.cuda.current_stream()

    def step(i, stats=True):
        cfg.sample_begin = (i * world + rank) % pps
        cfg.sample_stride = pps  # exactly one sample index per step
        return scene.render_device(cam, cfg, accum.data_ptr(), stream.cuda_stream, device=local_rank, want_stats=stats)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # polls through warm-up and the timed region; only samples inside the region are reported
    # ---- warm-up (untimed): scene upload, wave buffers, jitter table, clocks
    for i in range(max(args.warmup, 3)):
        step(i)
    torch.cuda.synchronize()

    # ---- timed region: exactly K steps + the final reduce, barrier + synchronize on both sides
    accum.zero_()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    t_region0 = time.perf_counter()
    tot = dict(rays=0, shadow=0, launches=0, ms_trace=0.0, ms_shadow=0.0, ms_shade=0.0, ms_other=0.0, trace_launches=0,
               trace_queries=0, shadow_launches=0)
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
    # ---- per-kernel durations for the roofline leg: the same K steps once more with one CUDA-event pair per kernel launch
    # on the launching stream. Untimed for `value`: while every launch carries its own events the shadow waves are not
    # overlapped with the next closest-hit wave (two streams), so the spans add up to the step and the shares can be
    # compared with the serialised ncu launch list in profiles/.
    api.set_profiling(True)
    scratch = torch.zeros_like(accum)
    ms_profiled = 0.0
    for i in range(args.steps):
        cfg.sample_begin = ((args.warmup + i) * world + rank) % pps
        cfg.sample_stride = pps
        st = scene.render_device(cam, cfg, scratch.data_ptr(), stream.cuda_stream, device=local_rank, want_stats=True)
        flush.zero_()
        for k in ("ms_trace", "ms_shadow", "ms_shade", "ms_other"):
            tot[k] += getattr(st, k)
        ms_profiled += st.ms_render
    torch.cuda.synchronize()
    api.set_profiling(False)
    del scratch
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    r = torch.tensor([tot["rays"]], device="cuda", dtype=torch.int64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
    ms_max, rays_all = float(t.item()), int(r.item())
    value = rays_all / (ms_max / 1e3) / 1e6

    # ---- e2e: the same metric through the host-buffer C-ABI call (trn_render: frame parameters in, image out to
    # host memory inside the timed region)
    host_pinned = torch.zeros(H, W, 4, dtype=torch.float32).pin_memory()  # pinned host memory for the per-step D2H
    host_img = host_pinned.numpy()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_rays = 0
    for i in range(args.steps):
        cfg.sample_begin = ((args.warmup + i) * world + rank) % pps
        cfg.sample_stride = pps
        _, st = scene.render(cam, cfg, device=local_rank, out=host_img)
        e2e_rays += st.rays
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    te = torch.tensor([dt], device="cuda", dtype=torch.float64)
    re = torch.tensor([e2e_rays], device="cuda", dtype=torch.int64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(re, op=dist.ReduceOp.SUM)
    e2e = {"value": int(re.item()) / float(te.item()) / 1e6, "unit": "Mrays/s",
           "h2d_bytes_per_step": int(api.C.sizeof(api.Camera) + api.C.sizeof(api.RenderConfig)),
           "d2h_bytes_per_step": int(host_img.nbytes),
           "note": "trn_render(): camera+config in, W*H*4 f32 image out to host memory every step; scene resident"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (closest-hit kd traversal): algorithmic bytes per launch from the
    # instrumented twin run once on the same steps' rays (untimed), duration from the event pairs above
    api.set_counting(True)
    cnt = dict(inner=0, leaf=0, tri=0, q=0, s_inner=0, s_leaf=0, s_tri=0, s_q=0, a_inner=0, a_leaf=0, a_tri=0, sa_inner=0,
               sa_leaf=0, sa_tri=0)
    probe_steps = min(args.steps, 2)
    for i in range(probe_steps):
        st = step(args.warmup + i)
        cnt["inner"] += st.trace_inner
        cnt["leaf"] += st.trace_leaf_nodes
        cnt["tri"] += st.trace_tri_tests
        cnt["q"] += st.trace_queries
        cnt["s_inner"] += st.shadow_inner
        cnt["s_leaf"] += st.shadow_leaf_nodes
        cnt["s_tri"] += st.shadow_tri_tests
        cnt["s_q"] += st.shadow_rays
        cnt["a_inner"] += st.trace_actual_inner
        cnt["a_leaf"] += st.trace_actual_leaf_nodes
        cnt["a_tri"] += st.trace_actual_tri_tests
        cnt["sa_inner"] += st.shadow_actual_inner
        cnt["sa_leaf"] += st.shadow_actual_leaf_nodes
        cnt["sa_tri"] += st.shadow_actual_tri_tests
    api.set_counting(False)
    torch.cuda.synchronize()
    peak, peak_src = measured_peaks()
    bytes_per_query = alg_bytes(cnt["inner"], cnt["leaf"], cnt["tri"], cnt["q"]) / max(cnt["q"], 1)
    trace_bytes = bytes_per_query * tot["trace_queries"]
    achieved = trace_bytes / max(tot["ms_trace"], 1e-9) / 1e6  # GB/s
    traffic, kernel_name, ncu_fig = None, "trace_pooled_kernel<0>", None
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            ent = json.load(open(prof)).get(args.workload, {})
            traffic = ent.get("trace_closest_dram_bytes_per_launch")
            kernel_name = ent.get("kernel", kernel_name)
            ncu_fig = ent.get("ncu")
        except Exception:
            traffic = None
    roofline = {
        "bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
        # what really bounds the kernel (the scene is L2-resident, so frac above is a normalised work rate that can exceed 1):
        # L1 wavefront and issue-slot utilisation from the committed ncu capture of the same kernel
        "ncu": ncu_fig,
        "alg_bytes_per_query": bytes_per_query,
        "per_query": {"inner": cnt["inner"] / max(cnt["q"], 1), "leaf_nodes": cnt["leaf"] / max(cnt["q"], 1),
                      "tri_tests": cnt["tri"] / max(cnt["q"], 1)},
        "per_query_actual": {"inner": cnt["a_inner"] / max(cnt["q"], 1), "leaf_nodes": cnt["a_leaf"] / max(cnt["q"], 1),
                             "tri_tests": cnt["a_tri"] / max(cnt["q"], 1),
                             "note": "visits of the production kernel on the device layout (empty-space cuts kept)"},
        "launches": tot["trace_launches"], "avg_launch_ms": tot["ms_trace"] / max(tot["trace_launches"], 1),
        "share_of_step": tot["ms_trace"] / max(ms_profiled, 1e-9),
        "profiled_pass_ms_per_step": ms_profiled / max(args.steps, 1),
        "kernel_ms": {k: tot[k] for k in ("ms_trace", "ms_shadow", "ms_shade", "ms_other")},
        "shadow_per_query": {"ref": [cnt["s_inner"] / max(cnt["s_q"], 1), cnt["s_leaf"] / max(cnt["s_q"], 1), cnt["s_tri"] / max(cnt["s_q"], 1)],
                             "actual": [cnt["sa_inner"] / max(cnt["s_q"], 1), cnt["sa_leaf"] / max(cnt["s_q"], 1), cnt["sa_tri"] / max(cnt["s_q"], 1)]},
        "shadow_kernel": {"alg_bytes_per_query": alg_bytes(cnt["s_inner"], cnt["s_leaf"], cnt["s_tri"], cnt["s_q"]) / max(cnt["s_q"], 1),
                          "achieved": alg_bytes(cnt["s_inner"], cnt["s_leaf"], cnt["s_tri"], cnt["s_q"]) / max(cnt["s_q"], 1)
                          * tot["shadow"] / max(tot["ms_shadow"], 1e-9) / 1e6},
        "fp32_tri_test_rate_gflops": 37.0 * (cnt["tri"] / max(cnt["q"], 1)) * tot["trace_queries"] / max(tot["ms_trace"], 1e-9) / 1e6,
    }
    # FP32 intersection-test rate against the FP32 FMA peak (148 SMs x 128 lanes x 2 flop x SM clock under load); the
    # tests run WITHOUT FMA contraction (bit-exactness), so 50 % is the ceiling of this ratio
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    roofline["fp32_peak_gflops"] = 148 * 128 * 2 * sm_mhz / 1e3
    roofline["fp32_frac"] = roofline["fp32_tri_test_rate_gflops"] / roofline["fp32_peak_gflops"]
    roofline["fp32_clock_mhz"] = sm_mhz

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        b = cpu_reference_run(args.workload, sc, scene.nodes(), np.array(scene.info.box, np.float32))
        cpu_baseline = {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")}

    line = {
        "metric": "Mrays/s (incl. secondary)", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
        "rays_per_step_per_gpu": tot["rays"] / args.steps, "shadow_rays_per_step_per_gpu": tot["shadow"] / args.steps,
        "queries_per_s_M": (tot["rays"] + tot["shadow"]) * world / (ms_max / 1e3) / 1e6,
        "kd_build_ms": scene.info.build_ms, "kd_height": int(scene.height), "triangles": int(scene.num_triangles),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(tot["launches"]), "roofline": roofline,
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
