#!/usr/bin/env python
"""bench.py — Mbases/s depth-counted on a synthetic 30x WGS-shaped alignment stream (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one pass of the depth hot path over one contig: segments -> per-base depth (on chip only)
-> per-window sums + callable-class runs (gl_depth_begin / add_segments / reduce).
Workload at N=1 = BASELINE config[1]: synthetic 30x chr20 (64,444,167 bp, 150 bp reads), W=500.
At N>1 every rank processes its own chr20-sized contig (contigs shard across GPUs with no
data-path collective) -> weak scaling; value = N * bases / max-over-ranks device time.

value  : inputs already resident in HBM in the engine's segment format (packed8) before the timed region;
         int32_resident: the same from plain int32 (start,end) device arrays.
e2e    : the one-call C-ABI entry gl_depth_region_packed8 (the feeder's short-read format, 2 B/segment) with
         PINNED HOST buffers: H2D of the segments, all kernels and D2H of window sums + runs are inside the
         timed region, every step.  e2e_packed16 / e2e_int32: the same through the 4 B and 8 B/segment entries.
roofline / cpu_baseline: see DESIGN.md §Measurement.
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

W = 500
MINCOV = 4
MAXMEAN = 0
STEP = 10_000_000           # depth/depth.go:48 (a multiple of W=500)
METRIC = "Mbases/s depth-counted (synth 30x WGS)"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            t = [x.strip() for x in ln.split(",")]
            if len(t) < 6:
                continue
            try:
                sm.append(float(t[0])); mx.append(float(t[1]))
            except ValueError:
                continue
            for nm, v in zip(names, t[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(kernel):
    """dram__bytes_read+write per launch of `kernel` from the committed ncu --set full capture (profiles/), or None"""
    try:
        j = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_full_summary.json")))
        return int(j[kernel]["traffic_bytes"])
    except Exception:
        return None


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def cpu_depth_pass(orc, s, e, L, threads, chunks):
    """The reference's CPU path for one contig, chunk-parallel like `goleft depth -p`:
    per 10 Mb chunk, per-base counting (the samtools child) + the callback's window/class walk + BED text."""
    from concurrent.futures import ThreadPoolExecutor
    order = np.argsort(s, kind="stable") if not (np.diff(s) >= 0).all() else None
    if order is not None:
        s, e = s[order], e[order]
    maxlen = int((e - s).max()) if s.size else 0

    def one(ch):
        cs, ce = ch
        lo = np.searchsorted(s, cs - maxlen, "left")
        hi = np.searchsorted(s, ce, "left")
        d = orc.pileup_diff(s[lo:hi], e[lo:hi], cs, ce)
        hd, ca = orc.walk_chunk("chr20", cs, ce, W, MINCOV, MAXMEAN, d)
        return len(hd) + len(ca)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        nbytes = sum(ex.map(one, chunks))
    return time.perf_counter() - t0, nbytes


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; neither go nor samtools
    exists in this image, so oracle/_ref cannot be built), all host threads, bounded sample."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from goleft_b200 import synth
    from oracle import loader as orc
    L = synth.CHR20_LEN
    s, e = synth.chr20_like()
    threads = os.cpu_count() or 1
    chunks = orc.gen_chunks(L, W)
    times = []
    for i in range(args.warmup + args.steps):
        dt, _ = cpu_depth_pass(orc, s, e, L, threads, chunks)
        if i >= args.warmup:
            times.append(dt)
    t = float(np.mean(times))
    val = L / t / 1e6
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Mbases/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
           "config": {"workload": "depth: synthetic 30x chr20 (64,444,167 bp, 150 bp reads), W=500, 1 sample",
                      "window": W, "mincov": MINCOV, "segments": int(s.size)},
           "cpu_baseline": {"value": val, "unit": "Mbases/s", "cores": min(threads, len(chunks)), "kind": "port",
                            "sample": "whole chr20 contig, 7 chunks of 10 Mb, one thread per chunk (the reference's own unit of "
                                      "parallelism, depth.go:132,392); per-base counting + window/class walk + BED text "
                                      "(samtools text printing/parsing excluded)"},
           "e2e": {"value": val, "unit": "Mbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    rank, world, local = dist_env()
    dist = None
    os.environ.setdefault("NCCL_DEBUG", "WARN")          # keep stdout to the one JSON line
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from goleft_b200 import capi, synth
    if capi.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; goleft_b200 has no CPU fallback")

    L = synth.CHR20_LEN
    t0 = time.time()
    s, e = synth.chr20_like(contig_index=19 + rank)
    nseg = int(s.size)
    log(f"[rank {rank}] synth: {nseg} segments over {L} bp in {time.time() - t0:.1f}s")

    ctx = capi.Ctx(local)
    d_s, d_e = ctx.dev_array(s), ctx.dev_array(e)
    n_win = (L - 1) // W + 1
    ctx.flush_l2()

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()

    def step_resident_int32():
        ctx.depth_begin(0, L)
        ctx.depth_add_segments_device(d_s, d_e, nseg)
        ctx.depth_reduce(W, MINCOV, MAXMEAN, STEP)

    # the engine's native segment format (packed8: 64-slot blocks, uint8 start delta + uint8 length), resident in HBM
    qa, qd, ql = capi.pack_segments8(s, e)
    d_qa, d_qd, d_ql = ctx.dev_array(qa), ctx.dev_array(qd), ctx.dev_array(ql)

    def step_resident():
        ctx.depth_begin(0, L)
        ctx.depth_add_segments_packed8_device(d_qa, d_qd, d_ql, qa.size)
        ctx.depth_reduce(W, MINCOV, MAXMEAN, STEP)

    # pinned host buffers for the end-to-end arm
    h_s, h_e = ctx.pinned_empty(nseg, np.int32), ctx.pinned_empty(nseg, np.int32)
    h_s[:] = s
    h_e[:] = e
    run_cap = L // 16 + 4096
    o_sum, o_rs, o_rc = ctx.pinned_empty(n_win, np.int64), ctx.pinned_empty(run_cap, np.int32), ctx.pinned_empty(run_cap, np.uint8)

    def step_e2e_int32():
        return ctx.depth_region(0, L, h_s, h_e, W, MINCOV, MAXMEAN, STEP, out=(o_sum, o_rs, o_rc))

    # the feeder's native compact format (packed16: 4 B/segment instead of 8), also in pinned host memory
    pa, po, pl = capi.pack_segments16(s, e)
    h_a, h_o, h_l = ctx.pinned_empty(pa.size, np.int32), ctx.pinned_empty(po.size, np.uint16), ctx.pinned_empty(pl.size, np.uint16)
    h_a[:] = pa
    h_o[:] = po
    h_l[:] = pl
    packed_bytes = int(pa.nbytes + po.nbytes + pl.nbytes)

    def step_e2e_p16():
        return ctx.depth_region_packed16(0, L, h_a, h_o, h_l, W, MINCOV, MAXMEAN, STEP, out=(o_sum, o_rs, o_rc))

    # the same packed8 words in pinned host memory for the end-to-end arm
    h8_a, h8_d, h8_l = ctx.pinned_empty(qa.size, np.int32), ctx.pinned_empty(qd.size, np.uint8), ctx.pinned_empty(ql.size, np.uint8)
    h8_a[:] = qa
    h8_d[:] = qd
    h8_l[:] = ql
    packed8_bytes = int(qa.nbytes + qd.nbytes + ql.nbytes)

    def step_e2e():
        return ctx.depth_region_packed8(0, L, h8_a, h8_d, h8_l, W, MINCOV, MAXMEAN, STEP, out=(o_sum, o_rs, o_rc))

    # ---- warm-up (also sizes every grow-only buffer)
    for _ in range(args.warmup):
        step_resident()
        step_resident_int32()
    ws, r0, rc = step_e2e()
    n_runs = int(r0.size)
    for _ in range(max(0, args.warmup - 1)):
        step_e2e()
    for _ in range(args.warmup):
        step_e2e_int32()
        step_e2e_p16()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- timed: device-resident.  The per-step working set (88 MB of segments) fits the 126 MB L2, so L2 is
    #      flushed (256 MiB memset) before every timed step; each step is timed with CUDA events on the ctx
    #      stream and the K step times are summed.
    barrier()
    l0 = ctx.launch_count()
    ms = 0.0
    for _ in range(args.steps):
        ctx.flush_l2()
        ctx.timer_start()
        step_resident()
        ms += ctx.timer_stop_ms()
    launches = ctx.launch_count() - l0
    barrier()
    ms_int32 = 0.0
    for _ in range(args.steps):
        ctx.flush_l2()
        ctx.timer_start()
        step_resident_int32()
        ms_int32 += ctx.timer_stop_ms()
    barrier()

    # ---- timed: end to end through the C ABI with pinned host buffers (H2D + kernels + D2H inside)
    barrier()
    ms_e2e = 0.0
    for _ in range(args.steps):
        ctx.flush_l2()
        ctx.sync()
        te0 = time.perf_counter()
        ctx.timer_start()
        step_e2e()
        dev = ctx.timer_stop_ms()
        ms_e2e += max(dev, (time.perf_counter() - te0) * 1e3)   # synchronous call: host wall time bounds it from above
    barrier()
    ms_e2e_int32 = 0.0
    for _ in range(args.steps):
        ctx.flush_l2()
        ctx.sync()
        te0 = time.perf_counter()
        ctx.timer_start()
        step_e2e_int32()
        dev = ctx.timer_stop_ms()
        ms_e2e_int32 += max(dev, (time.perf_counter() - te0) * 1e3)
    barrier()
    ms_e2e_p16 = 0.0
    for _ in range(args.steps):
        ctx.flush_l2()
        ctx.sync()
        te0 = time.perf_counter()
        ctx.timer_start()
        step_e2e_p16()
        dev = ctx.timer_stop_ms()
        ms_e2e_p16 += max(dev, (time.perf_counter() - te0) * 1e3)
    barrier()

    # ---- per-kernel live timing for the roofline: CUDA events on the launching stream around every
    #      kernel of the same step (library-side, gl_profile_*), averaged over the repetitions
    def kernel_times(reps, step=None):
        step = step or step_resident
        ctx.profile_enable(True)
        ctx.profile_read()
        acc = {}
        for _ in range(reps):
            ctx.flush_l2()
            step()
            for nm, t in ctx.profile_read():
                acc.setdefault(nm, []).append(t)
        ctx.profile_enable(False)
        per_step = {k: float(np.sum(v)) / reps for k, v in acc.items()}
        return per_step
    reps = max(5, min(args.steps, 20))
    k_ms = kernel_times(reps)
    path = ctx.depth_last_path()
    k_ms_int32 = kernel_times(reps, step_resident_int32)
    # the general (scatter) path on the same input, for the record
    ctx.depth_set_path(2)
    for _ in range(2):
        step_resident_int32()
    ms_general = 0.0
    for _ in range(reps):
        ctx.flush_l2()
        ctx.timer_start()
        step_resident_int32()
        ms_general += ctx.timer_stop_ms() / reps
    k_ms_general = kernel_times(reps, step_resident_int32)
    ctx.depth_set_path(0)
    clocks = sampler.stop() if rank == 0 else None

    os.environ.setdefault("NCCL_DEBUG", "WARN")
    if dist is not None:
        import torch
        t = torch.tensor([ms, ms_e2e, ms_e2e_int32, ms_e2e_p16, ms_int32], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_e2e_int32, ms_e2e_p16, ms_int32 = (float(x) for x in t)

    if rank == 0:
        peak, peak_src = peaks()
        ms_step = ms / args.steps
        val = world * L / (ms_step * 1e-3) / 1e6
        e2e_val = world * L / (ms_e2e / args.steps * 1e-3) / 1e6
        # ALGORITHMIC bytes per launch: what each kernel must move at minimum.
        # fused path: K_fused reads every (start,end) once (8 B/segment) and writes sums (8 B/window) and
        # runs (5 B/run); K_index reads the same 8 B/segment and writes the 4 B/cell offset table.
        # general path: memset 4L, scatter 8 B/segment read + two int32 updates, scan 4L read.
        ncells = L // 256 + 66
        p8_bytes = int(qa.nbytes + qd.nbytes + ql.nbytes)
        alg = {"depth_fused8_kernel": p8_bytes + 8 * n_win + 5 * n_runs,
               "depth_tileidx8_kernel": int(qa.nbytes) + 8 * ((L - 1) // 4096 + 1),
               "depth_fused_kernel": 8 * nseg + 8 * n_win + 5 * n_runs,
               "depth_index_kernel": 8 * nseg + 4 * ncells,
               "depth_scan_kernel": 4 * L + 8 * n_win + 5 * n_runs,
               "depth_scatter_kernel": 16 * nseg,
               "memset_diff": 4 * L}
        dom = max((k for k in k_ms if k in alg), key=lambda k: k_ms[k])
        achieved = alg[dom] / (k_ms[dom] * 1e-3) / 1e9
        step_bytes = sum(alg[k] for k in k_ms if k in alg)
        survey_bytes = 8 * nseg + 8 * L + 12 * n_win + 9 * n_runs       # SURVEY.md §8(d) formula (HBM difference array)
        gdom = max((k for k in k_ms_general if k in alg), key=lambda k: k_ms_general[k])
        out = {"metric": METRIC, "value": val, "unit": "Mbases/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "int32", "data": "synthetic",
               "config": {"workload": "depth: synthetic 30x chr20 (64,444,167 bp, 150 bp reads), W=500, 1 sample per GPU",
                          "window": W, "mincov": MINCOV, "run_break": STEP, "segments_per_gpu": nseg,
                          "windows_per_gpu": n_win, "runs_per_gpu": n_runs, "contigs": world,
                          "resident_format": "packed8 (64-slot blocks: int32 anchor + uint8 start delta + uint8 length per segment), %d bytes" % p8_bytes,
                          "path": {1: "fused (sorted int32 segments -> smem difference tiles)", 2: "general (HBM difference array)",
                                   3: "packed8 (packed words -> smem difference tiles)"}.get(path, str(path)),
                          "l2": "L2 flushed (256 MiB memset) before every timed step; per-step CUDA-event times summed"},
               "e2e": {"value": e2e_val, "unit": "Mbases/s", "h2d_bytes_per_step": packed8_bytes,
                       "d2h_bytes_per_step": 8 * n_win + 5 * n_runs, "ms_per_step": ms_e2e / args.steps,
                       "call": "gl_depth_region_packed8 (the feeder's short-read segment format: uint8 start delta + uint8 length, pinned host buffers)"},
               "e2e_packed16": {"value": world * L / (ms_e2e_p16 / args.steps * 1e-3) / 1e6, "unit": "Mbases/s",
                                "h2d_bytes_per_step": packed_bytes, "d2h_bytes_per_step": 8 * n_win + 5 * n_runs,
                                "ms_per_step": ms_e2e_p16 / args.steps,
                                "call": "gl_depth_region_packed16 (uint16 offset + uint16 length per segment, any read length)"},
               "e2e_int32": {"value": world * L / (ms_e2e_int32 / args.steps * 1e-3) / 1e6, "unit": "Mbases/s",
                             "h2d_bytes_per_step": 8 * nseg, "d2h_bytes_per_step": 8 * n_win + 5 * n_runs,
                             "ms_per_step": ms_e2e_int32 / args.steps,
                             "call": "gl_depth_region (plain int32 start/end arrays, pinned host buffers)"},
               "gpu_launches": int(launches),
               "roofline": {"bound": "hbm", "kernel": dom, "unit": "GB/s", "peak": peak, "peak_source": peak_src,
                            # contract definition: SURVEY.md §8(d)'s algorithmic bytes x the units one launch processes.
                            # The dominant kernel (K_fused) performs that whole per-base pipeline (difference array, scan,
                            # window reduce, class runs) on chip, so this is its EFFECTIVE bandwidth ...
                            "achieved": survey_bytes / (k_ms[dom] * 1e-3) / 1e9 if dom in ("depth_fused_kernel", "depth_fused8_kernel") else achieved,
                            "frac": (survey_bytes / (k_ms[dom] * 1e-3) / 1e9 if dom in ("depth_fused_kernel", "depth_fused8_kernel") else achieved) / peak,
                            "alg_bytes_per_launch": survey_bytes if dom in ("depth_fused_kernel", "depth_fused8_kernel") else alg[dom],
                            "basis": "SURVEY.md 8(d): 8*N_seg + 8*L + 12*ceil(L/W) + 9*runs (HBM difference-array pipeline); "
                                     "effective bandwidth of the kernel that does that pipeline's work",
                            # ... and these are the bytes the kernel itself has to move, with the ncu DRAM traffic beside them:
                            "own": {"alg_bytes_per_launch": alg[dom], "achieved": achieved, "frac": achieved / peak,
                                    "note": "packed segment words in + 8 B/window + 5 B/run out: the 8 B/base difference array never "
                                            "exists in HBM, so the kernel is latency/issue-bound, not HBM-bound "
                                            "(ncu: profiles/r01_ncu_full_summary.json)"},
                            "traffic": ncu_traffic(dom), "kernel_ms": k_ms,
                            "step": {"survey_alg_bytes": survey_bytes, "achieved": survey_bytes / (ms_step * 1e-3) / 1e9,
                                     "frac": survey_bytes / (ms_step * 1e-3) / 1e9 / peak,
                                     "own_alg_bytes": step_bytes, "own_achieved": step_bytes / (ms_step * 1e-3) / 1e9}},
               "int32_resident": {"ms_per_step": ms_int32 / args.steps, "value": world * L / (ms_int32 / args.steps * 1e-3) / 1e6,
                                  "kernel_ms": k_ms_int32,
                                  "note": "same region from plain int32 (start,end) device arrays: K_index + K_fused + K_gather"},
               "general_path": {"ms_per_step": ms_general, "value": world * L / (ms_general * 1e-3) / 1e6,
                                "kernel_ms": k_ms_general, "dominant": gdom,
                                "achieved": alg[gdom] / (k_ms_general[gdom] * 1e-3) / 1e9,
                                "frac": alg[gdom] / (k_ms_general[gdom] * 1e-3) / 1e9 / peak},
               "clocks": clocks}
        if not args.no_cpu_baseline and world == 1:         # rank 0 at N=1 only (the other ranks' feeders share the host cores at N>1)
            from oracle import loader as orc       # cpu_baseline leg: the oracle port timed on this box's host cores
            threads = os.cpu_count() or 1
            chunks = orc.gen_chunks(L, W)
            reps, tot = 0, 0.0
            while reps < 3 and tot < 20.0:
                dt, _ = cpu_depth_pass(orc, s, e, L, threads, chunks)
                tot += dt; reps += 1
            out["cpu_baseline"] = {"value": L / (tot / reps) / 1e6, "unit": "Mbases/s", "cores": min(threads, len(chunks)), "kind": "port",
                                   "sample": f"{reps} x whole chr20 contig (7 chunks of 10 Mb, one thread per chunk = the reference's "
                                             "unit of parallelism): per-base counting + window/class walk + BED text; "
                                             "samtools text print/parse excluded"}
        print(json.dumps(out), flush=True)

    d_s.free(); d_e.free(); d_qa.free(); d_qd.free(); d_ql.free()
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
