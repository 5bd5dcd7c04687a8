#!/usr/bin/env python
"""bench.py -- frames/s (and Mtriangles/s) of gel's per-frame render path on N B200s, beside the reference's CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg2|cfg5|cfg4|cfg1] [--impl reference]

A "step" is one pass of the render path (transform -> setup/binning -> raster -> shade) over one batch of views of the
synthetic workload.  Default workload = BASELINE.json configs[2], the HBM-roofline case the north-star target is quoted
on: 999 698-triangle sphere + 2048^2 texture at 3840x2160, 64 views per step per GPU (weak scaling: views are
independent, each rank renders its own block, no collective on the render path -- SURVEY.md 8(e)).

  value         whole-job frames/s, device-timed (CUDA events on the library's stream around the kernels of every step),
                mesh / texture / views resident in HBM, frames left in HBM; max over ranks
  e2e           the same metric through the C ABI with HOST buffers, wall clock, max over ranks: views from host memory,
                every frame delivered complete into pinned host frames inside the timed region.  The call is
                gelcu_render_region: only each view's screen region crosses PCIe, the caller's reused frame slots are
                kept complete by the dirty-rectangle contract (include/gelcu.h).  e2e_variants lists the same through
                gelcu_render (whole frames copied, the round-1 figure), the 24-bit frame sink, and sink + region.
  roofline      algorithmic bytes B_alg = 96 T + 8 W H + 4 L per frame x frames / CUDA-event time of ALL
                kernels of the step, against the measured HBM peak (MEASURED_PEAKS.json); the dominant kernel's own time
                and share are alongside
  cpu_baseline  the UNMODIFIED reference (oracle/_ref, main.c built headless) on this box's host cores, bounded sample
                (rank 0, N=1 only): all cores frames-parallel with the as-shipped -Ofast flags (value) and the strict
                build, plus 1-thread ms/frame of both (BASELINE.md 3 i-iii)
  other_workloads  the other four BASELINE.json configs, each with frames/s, roofline fraction, e2e and a CPU leg; cfg1 also
                as single-view latency through gelcu_render into a pageable buffer (the INTEGRATION.md drop-in call)
  parity_ranks  ranks whose per-view device checksums of the 8192-view cfg-5 list were gathered and compared on rank 0 with
                a one-GPU render of the whole list and with an oracle sample
`--impl reference` times the reference's own CPU implementation of the same workload (oracle/_ref) frames-parallel on all
host cores and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import hashlib
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
    "cfg4": ("overdraw", 100_000, 256, 1920, 1080, 64, "200 000 small overlapping tris (z ties) @ 1920x1080"),
    "cfg5": ("sphere", 50, 256, 1920, 1080, 8192, "5 000-tri sphere, 8192 rotated views @ 1920x1080 sharded over the GPUs"),
}
SWEEP = {"cfg1": 64, "cfg2": 360, "cfg3": 64, "cfg5": 8192}          # views of the workload's full rotation, xt_k = 2 pi k / n
KERNEL_SOURCES = ["gel_math.h", "gel_kernels.cuh", "gel_band.cuh", "gel_direct.cuh", "gel_sink.cuh", "gel_mesh.cuh", "gelcu.cu"]


def workload_config(name: str, ntri: int):
    """Identical in both arms (`--impl gel_b200` and `--impl reference`): only what defines the workload."""
    _, _, texn, xres, yres, _, desc = WORKLOADS[name]
    views = ("xt = 0.02 sin k, yt = 0.01 cos k (jitter around the front view)" if name == "cfg4"
             else f"rotation sweep xt_k = 2 pi k / {SWEEP[name]}, yt = 0 (the reference arm reaches it with integer mouse steps of 0.005 rad)")
    return {"workload": f"{name}: {desc}", "triangles": int(ntri), "resolution": f"{xres}x{yres}", "texture": f"{texn}x{texn}", "views": views,
            "l2": "GPU arm: L2 flushed between steps (256 MiB write) and a step writes more frame bytes than L2 holds; CPU arm: not applicable"}


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


def shard_span(nviews: int, world: int, rank: int):
    """Same partition as gel_b200.shard_views (kept here too so that the reference arm never imports the product)."""
    return nviews * rank // world, nviews * (rank + 1) // world


def gather_view_values(mine: np.ndarray, nviews: int, world: int, rank: int):
    """Epilogue, outside any timed region: per-view 64-bit values of every rank's block -> rank 0, in view order."""
    import torch
    d = _dist()
    if d is None:
        return np.asarray(mine)
    width = max(shard_span(nviews, world, r)[1] - shard_span(nviews, world, r)[0] for r in range(world))
    pad = torch.zeros(width, dtype=torch.int64, device=_dev())
    pad[: len(mine)] = torch.from_numpy(np.ascontiguousarray(mine).view(np.int64).copy()).to(_dev())
    out = [torch.zeros_like(pad) for _ in range(world)]
    d.all_gather(out, pad)
    if rank != 0:
        return None
    parts = []
    for r in range(world):
        lo, hi = shard_span(nviews, world, r)
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

def write_inputs(name: str, workdir: str):
    """Writes the workload's OBJ + BMP -- the files the reference's own loaders read -- and returns their paths.
    Pure Python (gel_b200.synth): no library of the product is loaded here."""
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
    return obj, bmp


def build_inputs(name: str, workdir: str):
    """The files of write_inputs, loaded through the product's host flow.  Returns dict(tv, tn, tt, tex, obj, bmp)."""
    import gel_b200
    obj, bmp = write_inputs(name, workdir)
    tv, tn, tt = gel_b200.load_obj(obj)
    return {"tv": tv, "tn": tn, "tt": tt, "tex": gel_b200.load_bmp(bmp), "obj": obj, "bmp": bmp}


def step_angles(name: str, nviews: int, offset: int = 0) -> np.ndarray:
    """(xt, yt) float32 of the workload's views [offset, offset + nviews)."""
    k = np.arange(nviews) + offset
    if name == "cfg4":
        return np.stack([0.02 * np.sin(k), 0.01 * np.cos(k)], 1).astype(np.float32)
    total = SWEEP[name]
    xt = (2.0 * np.pi * (k % total) / total).astype(np.float32)
    return np.stack([xt, np.zeros_like(xt)], 1)


def step_bases(name: str, nviews: int, offset: int = 0):
    import gel_b200
    return gel_b200.view_bases(step_angles(name, nviews, offset))


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


def kernel_sources_sha() -> str:
    h = hashlib.sha256()
    for n in KERNEL_SOURCES:
        with open(os.path.join(ROOT, "gel_b200", "csrc", n), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def recorded_traffic(kernel: str, name: str):
    """DRAM bytes per frame of `kernel` from the committed `ncu --set full` capture (profiles/traffic.json), or None with
    the reason when the capture was taken on other kernel sources than the ones this run executes."""
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except (OSError, ValueError):
        return None, "no profiles/traffic.json"
    rec = next((r for r in tj.get("records", []) if r.get("kernel") == kernel and r.get("workload") == name), None)
    if rec is None:
        return None, "no capture of this kernel on this workload"
    if tj.get("kernel_sources_sha") != kernel_sources_sha():
        return None, "stale: captured on other kernel sources (%s), not reported" % tj.get("kernel_sources_sha")
    return float(rec["bytes_per_frame"]), rec.get("from", "")


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# ---- the reference's CPU path ------------------------------------------------------------------------------

def mouse_script(name: str, first: int, frames: int):
    """Integer mouse steps (dx, dy per frame transition) that walk the reference through the workload's views
    [first, first + frames): xt -= 0.005 dx, yt += 0.005 dy (main.c:408-409), frame 0 is always (0, 0) (main.c:501), so
    the first step jumps to the first view.  Angles are the nearest multiples of 0.005 rad."""
    ang = step_angles(name, frames, first).astype(np.float64)
    tx = np.rint(-ang[:, 0] / 0.005).astype(np.int64)
    ty = np.rint(ang[:, 1] / 0.005).astype(np.int64)
    dx = np.diff(np.concatenate([[0], tx])); dy = np.diff(np.concatenate([[0], ty]))
    return "".join(f"{int(a)} {int(b)}\n" for a, b in zip(dx, dy))


def run_ref_processes(exe, obj, bmp, name, nproc, frames_each, td):
    """nproc copies of the reference binary, each rendering frame 0 (the fixed start, untimed) + frames_each views of the
    workload after a rendezvous; returns per-process lists of slock..sunlock milliseconds (frame 0 dropped)."""
    procs = []
    for p in range(nproc):
        script = os.path.join(td, f"script{p}.txt")
        with open(script, "w") as f:
            f.write(mouse_script(name, p * frames_each, frames_each))
        env = dict(os.environ, GELSHIM_FRAMES=str(frames_each + 1), GELSHIM_SCRIPT=script)
        if nproc > 1:
            env["GELSHIM_BARRIER"] = f"{td}:{nproc}"
        procs.append(subprocess.Popen([exe, obj, bmp], env=env, stdout=subprocess.PIPE, text=True))
    out = []
    for pr in procs:
        text, _ = pr.communicate()
        out.append([json.loads(l)["render_ms"] for l in text.splitlines() if l.startswith("{")][1:])
    for f in os.listdir(td):
        if f.startswith("ready."):
            os.unlink(os.path.join(td, f))
    return out


def run_cpu_reference(name: str, obj: str, bmp: str, ntri: int, steps: int, warmup: int, frames_per_step: int, budget_s: float = 60.0, single: bool = True):
    """The reference's CPU path on the host cores, bounded sample.
    kind "reference": P processes of oracle/_ref/gel_ref_[shipped_]<res> (the unmodified main.c, headless; for resolutions
    other than 800x600 with the resolution literal of main.c:499 substituted), each rendering (warmup+steps)*F views of the
    workload; a step = views [s*F, (s+1)*F) of every process, its time = the slowest process's summed slock..sunlock time.
    kind "port": oracle/libgeloracle.so on P threads (only when oracle/_ref is absent)."""
    import oracle
    _, _, _, xres, yres, _, _ = WORKLOADS[name]
    cores = os.cpu_count() or 1
    strict, shipped = oracle.ref_binary(xres, yres), oracle.ref_binary(xres, yres, shipped=True)
    total_steps = steps + warmup
    res = {"unit": "frames/s", "cores": cores, "cpu": cpu_model()}
    if strict is None:
        tv, tn, tt = oracle.load_obj(obj); tex = oracle.load_bmp(bmp)
        t0 = time.time()
        oracle.render(tv, tn, tt, tex, xres, yres, oracle.view_basis(0.3, 0.0))
        t_frame = max(time.time() - t0, 1e-4)
        F = int(max(1, min(frames_per_step, budget_s / (2.0 * t_frame * total_steps))))
        step_ms = []
        for s in range(total_steps):
            bases = np.stack([oracle.view_basis(a, b) for a, b in step_angles(name, cores * F, s * cores * F)])
            step_ms.append(oracle.render_views(tv, tn, tt, tex, xres, yres, bases, nthreads=cores, pixels=False)["seconds"] * 1e3)
        timed = step_ms[warmup:]
        fps = cores * F * len(timed) / (sum(timed) * 1e-3)
        res.update(value=fps, kind="port", mtri_per_s=fps * ntri / 1e6, ms_per_step=sum(timed) / len(timed), frames_per_step=cores * F,
                   sample=f"oracle port (oracle/ref_cpu.c, strict fp32 -O2) on {cores} threads, {F} views per thread per step",
                   single_thread_ms_per_frame={"strict_port": t_frame * 1e3})
        return res
    with tempfile.TemporaryDirectory() as td:
        # 1 thread, the faithful single-threaded reference: median slock..sunlock of 5 views (BASELINE.md 3 i, ii)
        one = {}
        for label, exe in (("strict_O2", strict), ("shipped_Ofast", shipped)):
            if exe is not None and (single or label == "shipped_Ofast"):
                ms = run_ref_processes(exe, obj, bmp, name, 1, 5, td)[0]
                one[label] = float(np.median(ms)) if ms else None
        t_frame = max((one.get("shipped_Ofast") or one.get("strict_O2") or 100.0) * 1e-3, 1e-4)
        F = int(max(1, min(frames_per_step, budget_s / (2.5 * t_frame * total_steps))))
        allcores = {}
        for label, exe in (("shipped_Ofast", shipped), ("strict_O2", strict)):
            if exe is None or (label == "strict_O2" and not single):
                continue
            per_proc = run_ref_processes(exe, obj, bmp, name, cores, total_steps * F, td)
            step_ms = [max(sum(pp[s * F:(s + 1) * F]) for pp in per_proc) for s in range(total_steps)]
            timed = step_ms[warmup:]
            allcores[label] = {"frames_per_s": cores * F * len(timed) / (sum(timed) * 1e-3), "ms_per_step": sum(timed) / len(timed)}
    head = "shipped_Ofast" if "shipped_Ofast" in allcores else "strict_O2"
    fps = allcores[head]["frames_per_s"]
    what = "main.c unmodified" if (xres, yres) == (800, 600) else "main.c with the resolution literal ssetup(800, 600) substituted"
    res.update(value=fps, kind="reference", mtri_per_s=fps * ntri / 1e6, ms_per_step=allcores[head]["ms_per_step"], frames_per_step=cores * F,
               sample=(f"{cores} processes of the reference ({what}; oracle/_ref, as-shipped flags -Ofast for the value, strict -O2 -ffp-contract=off "
                       f"alongside), each rendering {F} views of the workload per step via scripted mouse input, time = slock..sunlock of the slowest process"),
               all_cores={k: round(v["frames_per_s"], 3) for k, v in allcores.items()}, single_thread_ms_per_frame=one)
    return res


# ---- GPU measurements ----------------------------------------------------------------------------------------

def measure_device(r, name, views, steps, warmup, rank_offset, flush, hashes=False):
    """Device-timed steps with frames resident in HBM.  Returns (sum device ms, stats of the last step, launches, stage ms sums,
    (wall t0, t1), hashes of the last step or None)."""
    import torch
    bases = [step_bases(name, views, rank_offset + s * views) for s in range(max(steps, warmup))]
    for s in range(warmup):
        r.render(bases[s % len(bases)], pixels=False)
    torch.cuda.synchronize(); barrier(); torch.cuda.synchronize()
    t0 = time.time()
    dev_ms, launches, stage = 0.0, 0, {"ms_transform": 0.0, "ms_bin": 0.0, "ms_raster": 0.0, "ms_dominant": 0.0}
    out = None
    for s in range(steps):
        if flush is not None:
            flush.zero_(); torch.cuda.synchronize()
        out = r.render(bases[s % len(bases)], pixels=False, hashes=hashes)
        st = r.stats()
        dev_ms += out["device_ms"]; launches += st["kernels_launched"]
        for k in stage:
            stage[k] += st[k]
    torch.cuda.synchronize(); barrier(); torch.cuda.synchronize()
    return dev_ms, st, launches, stage, (t0, time.time()), (out["hash"] if hashes else None)


E2E_MODES = {
    "region": "gelcu_render_region: only each view's screen region is copied (strided copies), frames kept complete in reused host slots",
    "full": "gelcu_render: every frame's pixels copied whole (4 bytes per pixel; the round-1 figure)",
    "rgb8": "gelcu_render_rgb8: frames un-rotated and packed to 24 bits on the device, copied whole",
    "region_rgb8": "gelcu_render_region with the 24-bit frame sink: region only, 3 bytes per pixel",
}


def measure_e2e(r, name, views, steps, warmup, rank_offset, mode):
    """Wall-clock steps through the C ABI with host buffers: views from host memory, every frame complete in pinned host frames."""
    import torch
    import gel_b200
    xres, yres = r.xres, r.yres
    rgb = mode in ("rgb8", "region_rgb8")
    pinned = gel_b200.PinnedBuffer((views, yres, xres, 3) if rgb else (views, xres * yres), np.uint8 if rgb else np.uint32)
    pinned.array[...] = 0
    rects = np.tile(np.array([0, 0, -1, -1], np.int32), (views, 1))       # zeroed frames hold nothing
    bases = [step_bases(name, views, rank_offset + s * views) for s in range(max(steps, warmup))]
    if mode == "full":
        call = lambda b: r.render(b, pixel_out=pinned.array)
    elif mode == "rgb8":
        call = lambda b: r.render_rgb8(b, rgb_out=pinned.array)
    else:
        call = lambda b: r.render_region(b, pinned.array, rects, rgb8=rgb)
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
    pinned.free()
    return wall, st


def single_view_latency(r, name, frames=200):
    """The INTEGRATION.md drop-in call: ONE view per gelcu_render call into a pageable pixel buffer (SDL's streaming texture
    memory in the reference, main.c:504), successive views of the sweep; wall-clock per call."""
    bases = step_bases(name, frames + 10, 0)
    px = np.empty((1, r.xres * r.yres), np.uint32)
    for k in range(10):
        r.render(bases[k:k + 1], pixel_out=px)
    t = []
    for k in range(10, frames + 10):
        t0 = time.perf_counter()
        r.render(bases[k:k + 1], pixel_out=px)
        t.append(time.perf_counter() - t0)
    t = np.array(t) * 1e3
    return {"ms_per_frame_median": float(np.median(t)), "ms_per_frame_p99": float(np.percentile(t, 99)), "frames_per_s": float(1e3 / np.median(t)),
            "frames": frames, "call": "gelcu_render(ctx, &view, 1, pixel /*pageable*/, NULL, NULL, NULL), one call per frame"}


def bench_workload(name, world, rank, local_rank, workdir, steps, warmup, flush, opts, views_override=0, batch=0, e2e_modes=("region",),
                   want_hashes=False, latency=False, sampler=None):
    """Everything measured on the GPUs for one workload.  Returns (report dict on every rank, aux dict)."""
    import gel_b200
    kind, marg, texn, xres, yres, default_views, desc = WORKLOADS[name]
    inputs = build_inputs(name, workdir)
    ntri = int(inputs["tv"].shape[0])
    strong = name in ("cfg2", "cfg5") and not views_override
    if strong:
        lo, hi = shard_span(default_views, world, rank)
        views, rank_offset = hi - lo, lo
    else:
        views = views_override or default_views
        rank_offset = rank * views
    r = gel_b200.Renderer(xres, yres, device=local_rank)
    r.set_mesh(inputs["tv"], inputs["tn"], inputs["tt"])
    r.set_texture(inputs["tex"])
    if batch:
        r.set_option("batch_views", batch)
    for o in opts:
        r.set_option(o.split("=")[0], int(o.split("=")[1]))

    # lit pixels L of this rank's first view (for B_alg), read back once outside any timed region
    r.render(step_bases(name, 1, rank_offset), pixels=False)
    _, zb = r.read_frame(0)
    lit = int((zb != np.finfo(np.float32).min).sum())
    b_alg = algorithmic_bytes(ntri, xres, yres, lit)

    dev_ms, st, launches, stage, (tw0, tw1), _ = measure_device(r, name, views, steps, warmup, rank_offset, flush)
    # per-view device checksums for the rank-parity check: a separate, untimed pass (the checksum variants of the kernels hash every pixel)
    hashes = r.render(step_bases(name, views, rank_offset), pixels=False, hashes=True)["hash"] if want_hashes else None
    clocks = sampler.window(tw0, tw1) if sampler else None
    worst_ms = max_over_ranks(dev_ms)
    total_views = sum_over_ranks(views) * steps
    fps = total_views / (worst_ms * 1e-3)
    dom_ms = max_over_ranks(stage["ms_dominant"])
    pipeline = {1: "tile", 2: "direct"}.get(int(st.get("pipeline", 1)), "tile")
    band = not any(o.replace(" ", "") == "raster_mode=0" for o in (opts or []))
    dominant = "direct_raster_kernel<0>" if pipeline == "direct" else "raster_band_kernel" if band else "raster_kernel"

    peak, peak_src = hbm_peak()
    nlaunch = max(1, int(st["batches"]))
    dom_launch_ms = dom_ms / (steps * nlaunch)
    views_per_launch = views / nlaunch
    step_achieved = b_alg * views / (worst_ms / steps * 1e-3) / 1e9             # per GPU: this rank's bytes over the slowest rank's time
    dom_achieved = b_alg * views_per_launch / (dom_launch_ms * 1e-3) / 1e9 if dom_launch_ms > 0 else 0.0
    traffic, traffic_note = recorded_traffic(dominant, name)

    e2e = {}
    e2e_views = min(views, max(1, (2 << 30) // (4 * xres * yres)))
    e2e_steps = max(1, min(steps, 5))
    for mode in e2e_modes:
        wall, est = measure_e2e(r, name, e2e_views, e2e_steps, warmup, rank_offset, mode)
        e2e[mode] = {"value": sum_over_ranks(e2e_views) * e2e_steps / max_over_ranks(wall), "unit": "frames/s",
                     "h2d_bytes_per_step": int(est["h2d_bytes"]), "d2h_bytes_per_step": int(est["d2h_bytes"]),
                     "views_per_step_per_gpu": e2e_views, "call": E2E_MODES[mode]}
    lat = single_view_latency(r, name) if latency and rank == 0 else None
    r.close()

    report = {
        "value": fps, "ms_per_step": worst_ms / steps, "mtri_per_s": fps * ntri / 1e6, "scaling": "strong" if strong else "weak",
        "config": workload_config(name, ntri),
        "run": {"views_per_step_per_gpu": views, "unique_vertices": int(st["unique_vertices"]), "library_batches_per_step": int(st["batches"]),
                "pipeline": pipeline, "bin_entries_per_view": st["bin_entries"] / max(1, st["views"]), "lit_pixels": lit,
                "parallelism": f"views sharded over {world} GPU(s), mesh + texture replicated, no collective",
                "frame_bytes_written_per_step_gb": views * 8.0 * xres * yres / 1e9},
        "stage_ms_per_step": {k: v / steps for k, v in stage.items()},
        # achieved / frac are quoted over ALL kernels of the step (B_alg is the traffic of the whole pass, and the pass is several
        # kernels); the dominant kernel's own time and share are alongside.
        "roofline": {"bound": "hbm", "kernel": dominant, "achieved": step_achieved, "peak": peak, "unit": "GB/s", "frac": step_achieved / peak,
                     "traffic": traffic * views_per_launch if traffic else None, "traffic_note": traffic_note, "peak_source": peak_src,
                     "alg_bytes_per_frame": b_alg, "frames_per_launch": views_per_launch,
                     "basis": "B_alg x frames / CUDA-event time of every kernel of the step, per GPU",
                     "dominant_kernel_ms_per_launch": dom_launch_ms, "dominant_kernel_share_of_step": dom_ms / worst_ms if worst_ms else None,
                     "dominant_kernel_alone": {"achieved": dom_achieved, "frac": dom_achieved / peak, "note": "B_alg x frames / the dominant kernel's time alone -- not the figure of merit"}},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if e2e:
        first = e2e_modes[0]
        report["e2e"] = e2e[first]
        if len(e2e) > 1:
            report["e2e_variants"] = {k: v for k, v in e2e.items() if k != first}
    if lat:
        report["single_view_latency"] = lat
    return report, {"inputs": inputs, "hashes": hashes, "views": views, "rank_offset": rank_offset, "ntri": ntri}


def check_rank_parity(world, rank, local_rank, workdir, aux):
    """Under N ranks: the cfg-5 per-view checksums every rank computed for its block are gathered; rank 0 renders the WHOLE
    8192-view list on its own GPU and requires equality, then checks a 64-view sample against the CPU oracle (the checker).
    A frame depends only on mesh, texture and (xt, yt) (main.c:509-522), so the blocks must agree with the single-GPU list."""
    import gel_b200
    n = WORKLOADS["cfg5"][5]
    mine = aux["hashes"]
    cols = [gather_view_values(mine[:, k], n, world, rank) for k in (0, 1)]
    ok, detail = True, ""
    if rank == 0:
        inputs = aux["inputs"]
        got = np.stack(cols, 1).view(np.uint64)
        with gel_b200.Renderer(1920, 1080, device=local_rank) as r:
            r.set_mesh(inputs["tv"], inputs["tn"], inputs["tt"]); r.set_texture(inputs["tex"])
            whole = r.render(step_bases("cfg5", n, 0), pixels=False, hashes=True)["hash"]
        bad = int((whole != got).any(axis=1).sum())
        import oracle
        sample = np.arange(0, n, n // 64)
        bases = step_bases("cfg5", n, 0)[sample]
        ref = oracle.render_views(inputs["tv"], inputs["tn"], inputs["tt"], inputs["tex"], 1920, 1080, bases,
                                  nthreads=os.cpu_count() or 1, pixels=False, hashes=True)["hash"]
        bad_oracle = int((ref != got[sample]).any(axis=1).sum())
        ok = bad == 0 and bad_oracle == 0
        detail = f"{n} views: {bad} differ from the one-GPU render, {bad_oracle} of {len(sample)} sampled views differ from the oracle"
    return ok, detail


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="gel_b200", choices=["gel_b200", "reference"])
    ap.add_argument("--views", type=int, default=0, help="views per step per GPU (default: the workload's)")
    ap.add_argument("--batch", type=int, default=0, help="library batch_views option (0 = default)")
    ap.add_argument("--no-extra", action="store_true", help="skip the other four workloads")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--e2e", default="region,full,rgb8,region_rgb8", help="end-to-end modes to time for the main workload (first = headline)")
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
        obj, bmp = write_inputs(name, workdir)
        ntri = sum(1 for l in open(obj) if l.startswith("f "))
        cb = run_cpu_reference(name, obj, bmp, ntri, args.steps, args.warmup, frames_per_step=4, budget_s=90.0)
        line = {"impl": "reference", "metric": "frames/s", "value": cb["value"], "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "mtri_per_s": cb["mtri_per_s"],
                "config": workload_config(name, ntri),
                "run": {"frames_per_step": cb["frames_per_step"]},
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "cpu", "all_cores", "single_thread_ms_per_frame") if k in cb},
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
        write_inputs(name, workdir)
        if not args.no_extra:
            for other in WORKLOADS:
                write_inputs(other, workdir)
    barrier()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    sampler = ClockSampler(local_rank) if rank == 0 else None

    main_rep, main_aux = bench_workload(name, world, rank, local_rank, workdir, args.steps, args.warmup, flush, args.opt, views_override=args.views,
                                        batch=args.batch, e2e_modes=tuple(m for m in args.e2e.split(",") if m), sampler=sampler,
                                        want_hashes=(name == "cfg5" and not args.views), latency=(name == "cfg1"))
    extra, parity = {}, None
    cfg5_aux = main_aux if (name == "cfg5" and not args.views) else None
    if not args.no_extra:
        for other in ("cfg1", "cfg2", "cfg4", "cfg5"):
            if other == name:
                continue
            rep, aux = bench_workload(other, world, rank, local_rank, workdir, 3, 3, flush, [], e2e_modes=("region",), sampler=None,
                                      want_hashes=(other == "cfg5"), latency=(other == "cfg1"))
            if other == "cfg5":
                cfg5_aux = aux
            aux["report"] = rep
            extra[other] = (rep, aux)
    if cfg5_aux is not None and cfg5_aux["hashes"] is not None:
        ok, detail = check_rank_parity(world, rank, local_rank, workdir, cfg5_aux)
        if rank == 0 and not ok:
            raise SystemExit(f"parity check across ranks FAILED: {detail}")
        parity = {"parity_ranks": world, "parity_detail": detail}

    if rank == 0 and world == 1 and not args.no_cpu:
        main_rep["cpu_baseline"] = run_cpu_reference(name, main_aux["inputs"]["obj"], main_aux["inputs"]["bmp"], main_aux["ntri"], steps=2, warmup=1,
                                                     frames_per_step=4, budget_s=20.0)
        for other, (rep, aux) in extra.items():
            rep["cpu_baseline"] = run_cpu_reference(other, aux["inputs"]["obj"], aux["inputs"]["bmp"], aux["ntri"], steps=2, warmup=1,
                                                    frames_per_step=4, budget_s=6.0, single=(other == "cfg1"))
    if sampler:
        sampler.stop()

    if rank == 0:
        line = {"metric": "frames/s", "value": main_rep.pop("value"), "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": main_rep.pop("ms_per_step"), "higher_is_better": True, "scaling": main_rep.pop("scaling"), "vs_baseline": None,
                "dtype": "f32", "data": "synthetic"}
        line.update(main_rep)
        if parity:
            line.update(parity)
        if extra:
            keep = ("value", "ms_per_step", "mtri_per_s", "scaling", "config", "run", "stage_ms_per_step", "roofline", "e2e", "single_view_latency", "cpu_baseline", "gpu_launches")
            line["other_workloads"] = {k: {f: rep[f] for f in keep if f in rep} for k, (rep, _) in extra.items()}
            for k in line["other_workloads"]:
                line["other_workloads"][k]["unit"] = "frames/s"
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
