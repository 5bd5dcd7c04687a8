#!/usr/bin/env python
"""Pure device -> host copy sweep: what the host side of `e2e` can absorb with 1, 2, 4, 8 GPUs copying at once.

    python scripts/d2h_sweep.py [--mb 512] [--reps 6] > gpurun_out/d2h_sweep.json

One process, one host thread + one CUDA stream per GPU (the copies are asynchronous; the threads only issue them and wait).
For every concurrency level G in {1, 2, 4, 8} (as far as the box has GPUs) and every variant it reports the aggregate and the
per-GPU GB/s:
    pinned        cudaHostAlloc (default flags), contiguous cudaMemcpyAsync
    pinned_wc     cudaHostAllocWriteCombined destination
    pinned_2d     cudaMemcpy2DAsync of a 1752-row x 7008-byte rectangle out of an 8640-byte pitch (a cfg-3 view's region)
    pageable      plain malloc'd destination (what gelcu_render gets from the reference's SDL texture memory)
No kernel of the product runs here: this isolates the PCIe / host-memory ceiling that `e2e` at N = 8 runs into.
"""
import argparse
import ctypes
import json
import os
import threading
import time

import numpy as np
import torch


def cudart():
    for name in ("libcudart.so", "libcudart.so.12"):
        try:
            return ctypes.CDLL(name)
        except OSError:
            pass
    import glob
    cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) + glob.glob("/usr/local/cuda/lib64/libcudart.so*")
    return ctypes.CDLL(cands[0])


RT = cudart()
RT.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
RT.cudaFreeHost.argtypes = [ctypes.c_void_p]
RT.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
RT.cudaMemcpy2DAsync.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
RT.cudaSetDevice.argtypes = [ctypes.c_int]
D2H = 2


def run_level(gpus, variant, nbytes, reps):
    """Every GPU in `gpus` copies `nbytes` to the host `reps` times, all at once.  Returns (aggregate GB/s, per-GPU list)."""
    bufs, times = {}, {}
    pitch, width, rows = 8640, 7008, 1752
    frame = 3840 * pitch
    if variant == "pinned_2d":
        nbytes = (nbytes // frame) * frame or frame
    for g in gpus:
        RT.cudaSetDevice(g)
        src = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{g}")
        src.fill_(g + 1)
        if variant == "pageable":
            host = np.empty(nbytes, np.uint8); host[::4096] = 0
            ptr = host.ctypes.data
        else:
            p = ctypes.c_void_p()
            rc = RT.cudaHostAlloc(ctypes.byref(p), nbytes, 4 if variant == "pinned_wc" else 0)
            assert rc == 0, rc
            host, ptr = p, p.value
        bufs[g] = (src, host, ptr, torch.cuda.Stream(device=g))
    start = threading.Barrier(len(gpus) + 1)

    def worker(g):
        src, host, ptr, stream = bufs[g]
        RT.cudaSetDevice(g)
        torch.cuda.set_device(g)
        s = ctypes.c_void_p(stream.cuda_stream)

        def issue():
            if variant == "pinned_2d":
                for f in range(nbytes // frame):
                    RT.cudaMemcpy2DAsync(ptr + f * frame + 1040 * pitch + 200 * 4, pitch, src.data_ptr() + f * frame + 1040 * pitch + 200 * 4, pitch, width, rows, D2H, s)
            else:
                RT.cudaMemcpyAsync(ptr, src.data_ptr(), nbytes, D2H, s)
        issue(); stream.synchronize()                                   # warm-up (page faults of the destination)
        start.wait()
        t0 = time.perf_counter()
        for _ in range(reps):
            issue()
        stream.synchronize()
        times[g] = time.perf_counter() - t0

    ths = [threading.Thread(target=worker, args=(g,)) for g in gpus]
    [t.start() for t in ths]
    start.wait()
    [t.join() for t in ths]
    moved = (width * rows * (nbytes // frame) if variant == "pinned_2d" else nbytes) * reps
    per = [moved / times[g] / 1e9 for g in gpus]
    agg = moved * len(gpus) / max(times.values()) / 1e9
    for g in gpus:
        src, host, ptr, stream = bufs[g]
        if variant != "pageable":
            RT.cudaFreeHost(host)
    return agg, per


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=512)
    ap.add_argument("--reps", type=int, default=6)
    args = ap.parse_args()
    n = torch.cuda.device_count()
    out = {"gpus_on_box": n, "bytes_per_copy": args.mb << 20, "reps": args.reps, "cpus": os.cpu_count(), "levels": []}
    for G in (1, 2, 4, 8):
        if G > n:
            break
        for variant in ("pinned", "pinned_wc", "pinned_2d", "pageable"):
            if variant == "pageable" and G > 2:
                continue
            agg, per = run_level(list(range(G)), variant, args.mb << 20, args.reps)
            out["levels"].append({"gpus": G, "variant": variant, "aggregate_gbs": round(agg, 2), "per_gpu_gbs": [round(x, 2) for x in per]})
            print(json.dumps(out["levels"][-1]), flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
