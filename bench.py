#!/usr/bin/env python
"""bench.py -- frames/s (and Mtriangles/s) of gel's per-frame render path on N B200s, beside the reference's CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg2|cfg5|cfg4|cfg1] [--impl reference]

A "step" is one pass of the render path (transform -> setup/binning -> tile raster) over one batch of views of
the synthetic workload.  Default workload = BASELINE.json configs[2], the HBM-roofline case the north-star target
is quoted on: 999 698-triangle sphere + 2048^2 texture at 3840x2160, 64 views per step per GPU (weak scaling:
views are independent, each rank renders its own block, no collective on the render path -- SURVEY.md §8(e)).

  value       whole-job frames/s, device-timed (CUDA events on the library's stream around the kernels of every
              step), mesh/texture/views resident in HBM, frames left in HBM; max over ranks
  e2e         same metric through the C ABI with HOST buffers: views from pinned host memory, every frame's
              pixels copied back to pinned host memory inside the timed region (wall clock, max over ranks)
  roofline    dominant kernel (direct_raster_kernel<0> in the direct pipeline, raster_kernel in the tile pipeline):
              algorithmic bytes per launch / its CUDA-event duration vs the measured HBM peak (MEASURED_PEAKS.json);
              step_* = the same bytes over ALL the step's kernels (the figure the north-star target is about)
  cpu_baseline  the reference's CPU path on this box's host cores on a bounded sample (rank 0, N=1 only)
`--impl reference` times the reference's own CPU implementation (oracle/_ref = unmodified main.c built headless,
else the oracle port) frames-parallel on all host cores and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (mesh kind, mesh args, texture, xres, yres, views per step per GPU, description)
    "cfg1": ("sphere", 50, 256, 800, 600, 64, "5 000-tri sphere + 256^2 texture @ 800x600 (the reference's window)"),
    "cfg2": ("sphere", 50, 256, 1920, 1080, 360, "5 000-tri sphere + 256^2 texture, 360 rotated views @ 1920x1080"),
    "cfg3": ("sphere", 707, 2048, 3840, 2160, 64, "999 698-tri sphere + 2048^2 texture @ 3840x2160 (HBM-roofline case)"),
    "cfg4": ("overdraw", 100_000, 256, 1920, 1080, 8, "200 000 small overlapping tris (z ties) @ 1920x1080"),
    "cfg5": ("sphere", 50, 256, 1920, 1080, 8192, "5 000-tri sphere, 8192 rotated views @ 1920x1080 sharded over the GPUs"),
}


# ---- distributed plumbing (torch.distributed is plumbing only; the render path has no collective) ----------

def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def _dev():
    import torch
    d = _dist()
    return torch.device("cuda", torch.cuda.current_device()) if d is not None and d.get_backend() == "nccl" else torch.device("cpu")


def barrier():
    d = _dist()
    if d is not None:
        d.barrier()


def max_over_ranks(x: float) -> float:
    import torch
    d = _dist()
    if d is None:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=_dev())
    d.all_reduce(t, op=d.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x: float) -> float:
    import torch
    d = _dist()
    if d is None:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=_dev())
    d.all_reduce(t, op=d.ReduceOp.SUM)
    return float(t.item())


def gather_view_values(mine: np.ndarray, nviews: int, world: int, rank: int):
    """Optional epilogue (outside any timed region): per-view 64-bit values of every rank's block -> rank 0, in view order."""
    import torch
    import gel_b200
    d = _dist()
    if d is None:
        return np.asarray(mine)
    width = max(gel_b200.shard_views(nviews, world, r)[1] - gel_b200.shard_views(nviews, world, r)[0] for r in range(world))
    pad = torch.zeros(width, dtype=torch.int64, device=_dev())
    pad[: len(mine)] = torch.from_numpy(np.asarray(mine).astype(np.int64)).to(_dev())
    out = [torch.zeros_like(pad) for _ in range(world)]
    d.all_gather(out, pad)
    if rank != 0:
        return None
    parts = []
    for r in range(world):
        lo, hi = gel_b200.shard_views(nviews, world, r)
        parts.append(out[r][: hi - lo].cpu().numpy())
    return np.concatenate(parts)


# ---- clocks ------------------------------------------------------------------------------------------------

class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        rows = [r for (t, r) in self.rows if t0 - 0.02 <= t <= t1 + 0.05] or [r for (_, r) in self.rows[-3:]]
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); smax = max(smax, float(f[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax or None, "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc:
            self.proc.terminate()


# ---- workload ----------------------------------------------------------------------------------------------

def build_inputs(name: str, workdir: str):
    """Writes the workload's OBJ + BMP (the same files the reference's loaders read) and loads them through the
    product's host flow.  Returns dict(tv, tn, tt, tex, obj, bmp)."""
    import gel_b200
    from gel_b200 import synth
    kind, arg, texn, *_ = WORKLOADS[name]
    obj = os.path.join(workdir, f"{kind}{arg}.obj")
    bmp = os.path.join(workdir, f"tex{texn}.bmp")
    # several ranks may get here at once: each writes its own temporary and renames it into place (atomic; the content
    # is deterministic, so whichever rename lands last leaves the same file)
    if not os.path.exists(obj):
        text = synth.sphere_obj_text(arg, arg) if kind == "sphere" else synth.overdraw_obj_text(arg)
        tmp = f"{obj}.tmp{os.getpid()}"
        with open(tmp, "w") as f:
            f.write(text)
        os.replace(tmp, obj)
    if not os.path.exists(bmp):
        tmp = f"{bmp}.tmp{os.getpid()}"
        with open(tmp, "wb") as f:
            f.write(synth.texture_bmp_bytes(texn))
        os.replace(tmp, bmp)
    tv, tn, tt = gel_b200.load_obj(obj)
    return {"tv": tv, "tn": tn, "tt": tt, "tex": gel_b200.load_bmp(bmp), "obj": obj, "bmp": bmp}


def step_bases(name: str, nviews: int, offset: int = 0):
    """View sweep of the workload: xt_k = 2*pi*k/n, yt = 0; cfg4 uses small jitters around 0."""
    import gel_b200
    from gel_b200 import synth
    if name == "cfg4":
        ang = np.stack([0.02 * np.sin(np.arange(nviews) + offset), 0.01 * np.cos(np.arange(nviews) + offset)], 1).astype(np.float32)
    else:
        total = WORKLOADS[name][5] if name in ("cfg2", "cfg5") else max(nviews, 64)
        ang = synth.view_angles(total)[(np.arange(nviews) + offset) % total]
    return gel_b200.view_bases(ang)


def algorithmic_bytes(ntri, xres, yres, lit):
    """B_alg = 96*T + 8*W*H + 4*L per frame (inputs once, colour + z once, one texel per lit pixel)."""
    return 96.0 * ntri + 8.0 * xres * yres + 4.0 * lit


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# ---- the reference's CPU path ------------------------------------------------------------------------------

def run_cpu_reference(name: str, inputs, steps: int, warmup: int, frames_per_step: int, budget_s: float = 150.0):
    """Frames-parallel CPU run of the reference path on all host cores.
    kind "reference": P copies of oracle/_ref/gel_ref_<res> (the unmodified main.c, headless), each rendering
    (warmup+steps)*F scripted-mouse frames after a rendezvous; a step = frames [s*F, (s+1)*F) of every process, its
    time = the slowest process's summed slock->sunlock time.  kind "port": oracle/libgeloracle.so on P threads."""
    import oracle
    _, _, _, xres, yres, _, _ = WORKLOADS[name]
    cores = os.cpu_count() or 1
    exe = oracle.ref_binary(xres, yres)
    ntri = inputs["tv"].shape[0]
    # size the sample from a one-frame probe so the whole run stays inside the budget
    t0 = time.time()
    oracle.render(inputs["tv"], inputs["tn"], inputs["tt"], inputs["tex"], xres, yres, oracle.view_basis(0.3, 0.0))
    t_frame = max(time.time() - t0, 1e-4)
    total_steps = steps + warmup
    F = int(max(1, min(frames_per_step, budget_s / (2.0 * t_frame * total_steps))))
    if exe is not None:
        with tempfile.TemporaryDirectory() as td:
            procs = []
            for p in range(cores):
                env = dict(os.environ, GELSHIM_FRAMES=str(total_steps * F), GELSHIM_DX=str(-(7 + p % 11)), GELSHIM_DY="0", GELSHIM_BARRIER=f"{td}:{cores}")
                procs.append(subprocess.Popen([exe, inputs["obj"], inputs["bmp"]], env=env, stdout=subprocess.PIPE, text=True))
            per_proc = []
            for pr in procs:
                out, _ = pr.communicate()
                per_proc.append([json.loads(l)["render_ms"] for l in out.splitlines() if l.startswith("{")])
        step_ms = [max(sum(pp[s * F:(s + 1) * F]) for pp in per_proc) for s in range(total_steps)]
        kind = "reference"
        sample = f"{cores} processes of the unmodified reference (oracle/_ref, strict fp32 -O2), {F} scripted-mouse frames each per step, render time = slock..sunlock"
    else:
        step_ms = []
        nv = cores * F
        for s in range(total_steps):
            r = oracle.render_views(inputs["tv"], inputs["tn"], inputs["tt"], inputs["tex"], xres, yres, step_bases(name, nv, s * nv),
                                    nthreads=cores, pixels=False)
            step_ms.append(r["seconds"] * 1e3)
        kind = "port"
        sample = f"oracle port (oracle/ref_cpu.c, strict fp32 -O2) on {cores} threads, {F} views per thread per step"
    timed = step_ms[warmup:]
    frames = cores * F * len(timed)
    fps = frames / (sum(timed) * 1e-3)
    return {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample, "cpu": cpu_model(),
            "mtri_per_s": fps * ntri / 1e6, "ms_per_step": sum(timed) / len(timed), "frames_per_step": cores * F,
            "single_thread_ms_per_frame": t_frame * 1e3}


# ---- main --------------------------------------------------------------------------------------------------

def measure_gpu(r, name, inputs, views, steps, warmup, rank_offset, flush):
    """Device-timed steps with frames resident in HBM.  Returns (sum device ms, stats of last step, launches, stage ms sums)."""
    import torch
    bases = [step_bases(name, views, rank_offset + s * views) for s in range(max(steps, warmup))]
    for s in range(warmup):
        r.render(bases[s % len(bases)], pixels=False)
    torch.cuda.synchronize(); barrier(); torch.cuda.synchronize()
    t0 = time.time()
    dev_ms, launches, stage = 0.0, 0, {"ms_transform": 0.0, "ms_bin": 0.0, "ms_raster": 0.0, "ms_dominant": 0.0}
    for s in range(steps):
        if flush is not None:
            flush.zero_(); torch.cuda.synchronize()
        out = r.render(bases[s % len(bases)], pixels=False)
        st = r.stats()
        dev_ms += out["device_ms"]; launches += st["kernels_launched"]
        for k in stage:
            stage[k] += st[k]
    torch.cuda.synchronize(); barrier(); torch.cuda.synchronize()
    return dev_ms, st, launches, stage, (t0, time.time())


def measure_e2e(r, name, views, steps, warmup, rank_offset, pinned, sink=False):
    """sink=False: the drop-in call (sideways XRGB frames, 4 bytes per pixel over PCIe); sink=True: the device frame
    sink (gelcu_render_rgb8: upright 24-bit frames, 3 bytes per pixel)."""
    import torch
    bases = [step_bases(name, views, rank_offset + s * views) for s in range(max(steps, warmup))]
    call = (lambda b: r.render_rgb8(b, rgb_out=pinned.array[:views])) if sink else (lambda b: r.render(b, pixel_out=pinned.array[:views]))
    for s in range(min(warmup, 3)):
        call(bases[s % len(bases)])
    torch.cuda.synchronize(); barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(steps):
        call(bases[s % len(bases)])
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    st = r.stats()
    barrier()
    return wall, st


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="gel_b200", choices=["gel_b200", "reference"])
    ap.add_argument("--views", type=int, default=0, help="views per step per GPU (default: the workload's)")
    ap.add_argument("--batch", type=int, default=0, help="library batch_views option (0 = default)")
    ap.add_argument("--no-extra", action="store_true", help="skip the cfg2/cfg5 side measurements")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--opt", action="append", default=[], help="library tunable name=value (gelcu_set_option), repeatable")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank, local_rank, world = dist_env()
    name = args.workload
    kind, marg, texn, xres, yres, default_views, desc = WORKLOADS[name]
    workdir = os.path.join(tempfile.gettempdir(), "gel_b200_bench")
    os.makedirs(workdir, exist_ok=True)

    if args.impl == "reference":
        if rank != 0:
            return 0
        inputs = build_inputs(name, workdir)
        cb = run_cpu_reference(name, inputs, args.steps, args.warmup, frames_per_step=4)
        line = {"impl": "reference", "metric": "frames/s", "value": cb["value"], "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "mtri_per_s": cb["mtri_per_s"],
                "config": {"workload": f"{name}: {desc}", "triangles": int(inputs["tv"].shape[0]), "resolution": f"{xres}x{yres}",
                           "frames_per_step": cb["frames_per_step"]},
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "cpu")},
                "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import gel_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the render path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        inputs = build_inputs(name, workdir)
    barrier()
    if rank != 0:
        inputs = build_inputs(name, workdir)
    ntri = int(inputs["tv"].shape[0])
    views = args.views or (default_views if name not in ("cfg2", "cfg5") else (gel_b200.shard_views(default_views, world, rank)[1] - gel_b200.shard_views(default_views, world, rank)[0]))
    rank_offset = gel_b200.shard_views(default_views, world, rank)[0] if name in ("cfg2", "cfg5") else rank * views
    scaling = "strong" if name in ("cfg2", "cfg5") and not args.views else "weak"

    r = gel_b200.Renderer(xres, yres, device=local_rank)
    r.set_mesh(inputs["tv"], inputs["tn"], inputs["tt"])
    r.set_texture(inputs["tex"])
    if args.batch:
        r.set_option("batch_views", args.batch)
    for o in args.opt:
        r.set_option(o.split("=")[0], int(o.split("=")[1]))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    # lit pixels L of this rank's first view (for B_alg), read back once outside any timed region
    r.render(step_bases(name, 1, rank_offset), pixels=False)
    _, zb = r.read_frame(0)
    lit = int((zb != np.finfo(np.float32).min).sum())
    b_alg = algorithmic_bytes(ntri, xres, yres, lit)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    dev_ms, st, launches, stage, (tw0, tw1) = measure_gpu(r, name, inputs, views, args.steps, args.warmup, rank_offset, flush)
    clocks = sampler.window(tw0, tw1) if sampler else None
    worst_ms = max_over_ranks(dev_ms)
    total_views = sum_over_ranks(views) * args.steps
    fps = total_views / (worst_ms * 1e-3)
    raster_ms = max_over_ranks(stage["ms_dominant"])
    pipeline = {1: "tile", 2: "direct"}.get(int(st.get("pipeline", 1)), "tile")
    dominant = "direct_raster_kernel<0>" if pipeline == "direct" else "raster_kernel"

    # end to end through the C ABI with host buffers (pixels of every frame come back to pinned host memory)
    e2e_views = min(views, max(1, (2 << 30) // (4 * xres * yres)))
    pinned = gel_b200.PinnedBuffer((e2e_views, xres * yres), np.uint32)
    e2e_steps = max(1, min(args.steps, 5))
    wall, est = measure_e2e(r, name, e2e_views, e2e_steps, args.warmup, rank_offset, pinned)
    e2e_fps = sum_over_ranks(e2e_views) * e2e_steps / max_over_ranks(wall)
    pinned.free()
    # the same through the device frame sink: 24-bit upright frames, 25 % fewer PCIe bytes
    pinned = gel_b200.PinnedBuffer((e2e_views, yres, xres, 3), np.uint8)
    wall8, est8 = measure_e2e(r, name, e2e_views, e2e_steps, args.warmup, rank_offset, pinned, sink=True)
    e2e8_fps = sum_over_ranks(e2e_views) * e2e_steps / max_over_ranks(wall8)
    pinned.free()

    peak, peak_src = hbm_peak()
    launches_per_step_raster = st["batches"]
    raster_launch_ms = raster_ms / (args.steps * launches_per_step_raster)
    views_per_launch = views / launches_per_step_raster
    achieved = b_alg * views_per_launch / (raster_launch_ms * 1e-3) / 1e9
    step_achieved = b_alg * views / (worst_ms / args.steps * 1e-3) / 1e9

    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dominant)
        if tj and tj.get("workload") == name:
            traffic = tj["bytes_per_frame"] * views_per_launch        # ncu --set full capture, per launch like `achieved`
    except (OSError, ValueError):
        pass

    extra = {}
    if not args.no_extra and name == "cfg3":
        for other in ("cfg2", "cfg5"):
            oin = build_inputs(other, workdir)
            _, _, _, ox, oy, ov, odesc = WORKLOADS[other]
            lo, hi = gel_b200.shard_views(ov, world, rank)
            with gel_b200.Renderer(ox, oy, device=local_rank) as r2:
                r2.set_mesh(oin["tv"], oin["tn"], oin["tt"]); r2.set_texture(oin["tex"])
                ms2, st2, l2, _, _ = measure_gpu(r2, other, oin, hi - lo, 3, 3, lo, flush)
            w2 = max_over_ranks(ms2)
            extra[other] = {"workload": odesc, "scaling": "strong", "views_total": ov, "frames_per_s": ov * 3 / (w2 * 1e-3),
                            "mtri_per_s": ov * 3 * oin["tv"].shape[0] / (w2 * 1e-3) / 1e6, "ms_per_pass": w2 / 3, "gpu_launches": int(l2)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = run_cpu_reference(name, inputs, steps=2, warmup=1, frames_per_step=4, budget_s=25.0)
    if sampler:
        sampler.stop()
    r.close()

    if rank == 0:
        line = {
            "metric": "frames/s", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": worst_ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "mtri_per_s": fps * ntri / 1e6,
            "config": {"workload": f"{name}: {desc}", "triangles": ntri, "unique_vertices": int(st["unique_vertices"]), "resolution": f"{xres}x{yres}",
                       "views_per_step_per_gpu": views, "parallelism": f"views sharded over {world} GPU(s), mesh+texture replicated, no collective",
                       "l2": "flushed between steps (256 MiB write); a step also writes %.1f GB of frames" % (views * 8.0 * xres * yres / 1e9),
                       "library_batches_per_step": int(st["batches"]), "pipeline": pipeline, "bin_entries_per_view": st["bin_entries"] / max(1, st["views"]), "lit_pixels": lit},
            "stage_ms_per_step": {k: v / args.steps for k, v in stage.items()},
            # achieved / frac are quoted over ALL kernels of the step (the conservative reading: B_alg is the traffic of the
            # whole pass, and the pass is several kernels); the dominant kernel's own time and share are alongside.
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": step_achieved, "peak": peak, "unit": "GB/s", "frac": step_achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "alg_bytes_per_frame": b_alg, "frames_per_launch": views_per_launch,
                         "basis": "B_alg x frames / CUDA-event time of every kernel of the step",
                         "dominant_kernel_ms_per_launch": raster_launch_ms, "dominant_kernel_share_of_step": raster_ms / worst_ms,
                         "dominant_kernel_alone": {"achieved": achieved, "frac": achieved / peak, "note": "B_alg x frames / the dominant kernel's time alone"}},
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": int(est["h2d_bytes"]), "d2h_bytes_per_step": int(est["d2h_bytes"]),
                    "views_per_step_per_gpu": e2e_views, "note": "views from host, every frame's pixels copied to pinned host memory; PCIe-bound"},
            "e2e_rgb8_sink": {"value": e2e8_fps, "unit": "frames/s", "h2d_bytes_per_step": int(est8["h2d_bytes"]), "d2h_bytes_per_step": int(est8["d2h_bytes"]),
                              "note": "same, through gelcu_render_rgb8: frames un-rotated and packed to 24 bits on the device before the copy"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if cpu:
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "cpu", "single_thread_ms_per_frame")}
        if extra:
            line["other_workloads"] = extra
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
