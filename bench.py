#!/usr/bin/env python
"""bench.py -- the reference's headline metric on BASELINE.json configs[1]:
120 000-triangle random-cube scene (create_n_cubes(10 000), src/testbase.rs:608-615), 1 M create_ray rays
(src/testbase.rs:687-691), f32/3D.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one pass of the hot path's query side over one ray batch: batched Bvh::traverse of 1 M rays
against the device-resident tree, producing the CSR hit lists (at N > 1 every rank traverses its own
1 M-ray shard of the seed chain and the per-rank hit lists are all-gathered over NCCL inside the step:
weak scaling).  `value` = rays of all ranks / max-over-ranks mean step time, inputs resident in HBM.
The build side (Bvh::build + flatten, Mprims/s) is timed in the same run and reported under "build".
`e2e` is the same traversal through the host-pointer C-ABI call (pinned host rays in, host CSR out).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CUBES = 10_000           # 120 000 triangles
N_RAYS = 1_000_000
BUILD_PRIM_VISITS = 2_144_236     # P of the 120k scene (sum over internal nodes of their shape count)
METRIC = "traversal_Mrays_per_s"
UNIT = "Mrays/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        inside = [l for (t, l) in self.lines if self.t0 is not None and self.t0 - 0.15 <= t <= (self.t1 or t) + 0.15]
        window = "timed region"
        if not inside:                     # region shorter than the sampling period: fall back to the whole (loaded) run
            inside, window = [l for (_, l) in self.lines], "warm-up + timed region"
        for l in inside:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm), "window": window}


# ------------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU path (the C++ restatement in oracle/: the Rust crate cannot be built in this
    image), all host threads, on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O

    hw = O.hardware_threads()
    shapes = O.create_n_cubes(N_CUBES)
    res = O.build(shapes, threads=hw)
    sample = 1_000_000
    rays, _ = O.create_rays(sample)
    cands = sorted({t for t in (hw, hw // 2, hw // 4, 32, 16) if 1 <= t <= hw})
    best = {t: min(O.traverse(res.nodes, shapes, rays, O.MODE_RECURSIVE, threads=t).seconds for _ in range(max(args.warmup, 1))) for t in cands}
    threads = min(best, key=best.get)       # fastest thread count on this box (warm-up doubles as the sweep)
    dt = 0.0
    for _ in range(args.steps):      # Bvh::traverse, rays split over all cores; time = thread create .. join inside C++
        dt += O.traverse(res.nodes, shapes, rays, O.MODE_RECURSIVE, threads=threads).seconds
    dt /= args.steps
    tb = [O.build(shapes, threads=threads).seconds for _ in range(3)]
    value = sample / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: create_n_cubes(10000) = 120000 triangles, create_ray rays from seed 0, Bvh::traverse (recursive)",
                   "rays_per_step": sample, "note": "C++ restatement of the reference (oracle/), no Rust toolchain in the image"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": f"{sample} rays/step x {args.steps} steps, rays split evenly over {threads} threads"},
        "build": {"value": len(shapes) / min(tb) / 1e6, "unit": "Mprims/s", "cores": threads, "what": "Bvh::build_par analogue (fork-join, grain 64), best of 3"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench.py --gpus N ...")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    from bvh_b200 import api, capi, scenes
    from bvh_b200.dist import ShardedTraversal
    from bvh_b200.dtypes import RAY3F

    ctx = api.Context(local)
    stream = torch.cuda.Stream(dev)            # everything timed runs on this one stream (torch events see it)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    # ---- inputs: generated on the host once, resident in HBM before anything is timed -------------------
    aabbs = scenes.create_n_cubes_aabbs(N_CUBES)
    n = len(aabbs)
    d_aabbs = torch.from_numpy(aabbs.view(np.uint8).reshape(-1)).to(dev)
    o, d = scenes.ray_endpoints(N_RAYS, first_ray=rank * N_RAYS)          # rank r owns rays [r*1M, (r+1)*1M) of the seed chain
    d_o, d_d = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
    d_rays = torch.empty(N_RAYS * RAY3F.itemsize, dtype=torch.uint8, device=dev)
    capi.check(capi.lib().bvhgpu_rays_new_dev_f32x3(ctx._h, d_o.data_ptr(), d_d.data_ptr(), N_RAYS, d_rays.data_ptr()))   # Ray::new on the device
    cap = 8 * N_RAYS
    d_off = torch.empty(N_RAYS + 1, dtype=torch.int32, device=dev)
    d_hits = torch.empty(cap, dtype=torch.int32, device=dev)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)    # > 126 MB L2

    bvh = api.Bvh.build_dev(d_aabbs.data_ptr(), n, ctx=ctx)
    bvh.flatten()
    ctx.synchronize()
    ctx.set_option("profile", 1)

    sharded = ShardedTraversal(bvh, N_RAYS, 2 * N_RAYS * world) if world > 1 else None

    def step():
        if sharded is None:
            bvh.traverse_dev(d_rays.data_ptr(), N_RAYS, d_off.data_ptr(), d_hits.data_ptr(), cap)
        else:              # walk + the path's one exchange step (hit lists stored into every rank's global CSR over NVLink P2P)
            sharded.step(d_rays.data_ptr(), N_RAYS)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        step()
    barrier()
    launches0 = ctx.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    walk_ms = []
    if rank == 0:                          # nvidia-smi needs a moment for its first sample: keep the GPU loaded until it has one
        t_wait = time.time()
        while not sampler.lines and time.time() - t_wait < 5.0:
            bvh.traverse_dev(d_rays.data_ptr(), N_RAYS, d_off.data_ptr(), d_hits.data_ptr(), cap)   # local work only (no peer handshake)
            torch.cuda.synchronize(dev)
    barrier()
    sampler.mark_begin()
    for k in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (outside the event pair)
        ev[k][0].record(stream)
        step()
        ev[k][1].record(stream)
        if rank == 0:
            walk_ms.append(ctx.get_metric("walk_ms"))
    barrier()
    sampler.mark_end()
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop()
    step_ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    t = torch.tensor([step_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms = float(t.item())
    value = world * N_RAYS / (step_ms * 1e-3) / 1e6

    if world > 1:
        if rank == 0:
            line = _base_line(args, value, step_ms, launches, clocks)
            line["config"]["parallelism"] = (f"ray batch sharded over {world} GPUs (1M rays each), tree replicated; all-gather of the CSR hit lists fused into the "
                                             "traversal: totals via peer mailboxes, emit kernel stores into every rank's global CSR over NVLink P2P (CUDA IPC), inside the step")
            print(json.dumps(line), flush=True)
        sharded.close()
        dist.destroy_process_group()
        return

    # ---- single GPU extras: roofline of the dominant kernel, build, e2e, cpu baseline -------------------
    visits, hits_total = _stats_after_sync_traverse(bvh, d_rays, d_off, d_hits, cap)
    walk = sum(walk_ms) / len(walk_ms)
    alg_bytes = N_RAYS * 36 + visits * 32 + N_RAYS * 4 + hits_total * 4        # DESIGN.md "algorithmic bytes, traversal"
    peak, peak_src = _peaks()
    achieved = alg_bytes / (walk * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "pass 1 of the traversal: coherence_probe_kernel + walk_persistent_kernel<float,false,false> (walk_count_kernel gated off for this incoherent batch)", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": _ncu_traffic(), "peak_source": peak_src, "bytes_per_launch": alg_bytes, "kernel_ms": walk,
                "node_visits_per_ray": visits / N_RAYS, "note": "tree (7.7 MB) is L2-resident by construction; bytes are algorithmic, not DRAM"}

    # build: Bvh::build (+ flatten) of the 120k scene, device-resident AABBs, events around each build
    ctx.set_option("profile", 0)
    bt = []
    for k in range(3 + 10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        b2 = api.Bvh.build_dev(d_aabbs.data_ptr(), n, ctx=ctx)
        b2.flatten_dev()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        b2.free()
        if k >= 3:
            bt.append(e0.elapsed_time(e1))
    bt.sort()
    build_ms = bt[len(bt) // 2]

    # e2e: the host-pointer C-ABI call, pinned host buffers, H2D + traverse + D2H inside the timed region
    h_rays = torch.empty(N_RAYS * RAY3F.itemsize, dtype=torch.uint8).pin_memory()
    h_rays.copy_(d_rays.cpu())
    h_off = torch.empty(N_RAYS + 1, dtype=torch.int32).pin_memory()
    h_hits = torch.empty(cap, dtype=torch.int32).pin_memory()
    import ctypes as C
    tot = C.c_size_t(0)
    fn = capi.lib().bvhgpu_traverse_f32x3

    def e2e_step():
        capi.check(fn(bvh._h, capi.TRAVERSE_BVH, h_rays.data_ptr(), N_RAYS, h_off.data_ptr(), h_hits.data_ptr(), cap, C.byref(tot)))

    for _ in range(3):
        e2e_step()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()                       # synchronous call: returns when offsets + hits are in host memory
    torch.cuda.synchronize(dev)
    e2e_s = (time.perf_counter() - t0) / args.steps
    e2e = {"value": N_RAYS / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": N_RAYS * 36, "d2h_bytes_per_step": (N_RAYS + 1) * 4 + int(tot.value) * 4,
           "ms_per_step": e2e_s * 1e3}

    line = _base_line(args, value, step_ms, launches, clocks)
    line["roofline"] = roofline
    line["e2e"] = e2e
    # algorithmic bytes of the build (SURVEY 8d / DESIGN 4.1): n*S_aabb + P*(S_aabb+8+8) + (2n-1)*S_node + n*8, P = sum over internal
    # nodes of their range size = 2 144 236 for this scene (oracle counter; asserted in tests/test_oracle_goldens.py)
    build_bytes = n * 24 + BUILD_PRIM_VISITS * (24 + 8 + 8) + (2 * n - 1) * 64 + n * 8
    line["build"] = {"value": n / (build_ms * 1e-3) / 1e6, "unit": "Mprims/s", "ms": build_ms, "what": "Bvh::build (exact SAH, bit-identical) + flatten, 120000 shapes, AABBs resident in HBM, median of 10",
                     "roofline": {"bound": "hbm", "achieved": build_bytes / (build_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": build_bytes / (build_ms * 1e-3) / 1e9 / peak,
                                  "bytes_per_build": build_bytes, "note": "latency-bound at this size: ~9 multi-warp tree levels + warp-serial subtrees on a 30 MB L2-resident working set (DESIGN.md 4.1)"}}
    lt = []
    for k in range(3 + 10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        b2 = api.Bvh.build_dev(d_aabbs.data_ptr(), n, ctx=ctx, mode=capi.BUILD_LBVH)
        b2.flatten_dev()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        b2.free()
        if k >= 3:
            lt.append(e0.elapsed_time(e1))
    lt.sort()
    line["build_lbvh"] = {"value": n / (lt[len(lt) // 2] * 1e-3) / 1e6, "unit": "Mprims/s", "ms": lt[len(lt) // 2],
                          "what": "BVHGPU_BUILD_LBVH (Morton/Karras, same node layout, identical hit sets, different topology) + flatten"}
    line["cpu_baseline"] = _cpu_baseline()
    print(json.dumps(line), flush=True)


def _stats_after_sync_traverse(bvh, d_rays, d_off, d_hits, cap):
    bvh.traverse_dev(d_rays.data_ptr(), N_RAYS, d_off.data_ptr(), d_hits.data_ptr(), cap, want_total=True)
    return bvh.traverse_stats()


def _ncu_traffic():
    """dram bytes per launch of the walk kernel from the committed ncu capture (profiles/), if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("walk_count_kernel_dram_bytes_per_launch")
        except Exception:
            return None
    return None


def _base_line(args, value, step_ms, launches, clocks):
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: create_n_cubes(10000) = 120000 triangles, 1M create_ray rays per GPU from seed 0, batched Bvh::traverse -> CSR hit lists",
                   "rays_per_gpu": N_RAYS, "shapes": 12 * N_CUBES, "l2": "512 MB flush write between timed iterations", "builder": "exact_sah"},
        "gpu_launches": launches, "clocks": clocks,
    }


def _cpu_baseline():
    from oracle import oracle as O

    hw = O.hardware_threads()
    shapes = O.create_n_cubes(N_CUBES)
    res = O.build(shapes, threads=hw)
    sample = 1_000_000 if hw >= 16 else 250_000
    rays, _ = O.create_rays(sample)
    # "all the host threads it can use": pick the thread count that is fastest on this box (SMT / NUMA can make hw slower)
    cands = sorted({t for t in (hw, hw // 2, hw // 4, 32, 16) if 1 <= t <= hw})
    best = {t: min(O.traverse(res.nodes, shapes, rays, O.MODE_RECURSIVE, threads=t).seconds for _ in range(2)) for t in cands}
    threads = min(best, key=best.get)
    reps, t0, secs = 0, time.perf_counter(), 0.0
    while reps < 3 or time.perf_counter() - t0 < 10.0:
        secs += O.traverse(res.nodes, shapes, rays, O.MODE_RECURSIVE, threads=threads).seconds
        reps += 1
        if reps >= 40:
            break
    dt = secs / reps
    b1 = O.build(shapes, threads=1).seconds
    tb = [O.build(shapes, threads=t).seconds for t in cands for _ in range(2)]
    return {"value": sample / dt / 1e6, "unit": UNIT, "cores": threads, "kind": "port", "host_threads_available": hw,
            "thread_sweep_s": {str(k): round(v, 5) for k, v in best.items()},
            "sample": f"{sample} of the 1M rays x {reps} reps, Bvh::traverse (recursive), rays split evenly over {threads} threads",
            "build_Mprims_per_s_1thread": len(shapes) / b1 / 1e6, "build_Mprims_per_s_all_threads": len(shapes) / min(tb) / 1e6}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
