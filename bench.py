#!/usr/bin/env python
"""bench.py -- descriptor pairs/sec of the keypoint-matching hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3]

A step = one pass of the hot path (match.cpp:638-652: every image pair of a group through
ComputeMatches) over one synthetic keypoint group.  Workload at N = 1 is BASELINE.json
configs[1]: 10 images x 20 000 keypoints, random SURF3D-format descriptors (`iid` set,
SURVEY.md 8d), -d 1 -d2 1.  For N > 1 the scaling is WEAK: every rank matches its own
10 x 20k group (image pairs are independent units, sharded by group, no data-path collective);
only the compacted match lists are gathered to rank 0 over NCCL.

One JSON line on rank 0:  value = descriptor pairs/s with keypoints resident in HBM;
e2e = the same through the C ABI from pinned HOST buffers (H2D upload + prep + match + D2H);
roofline = tensor-core FLOP rate of the scoring kernel (96 FLOP per descriptor pair);
cpu_baseline = the verbatim reference match.cpp (oracle/_ref/match_ref) on this box's cores.
"""
from __future__ import annotations

import argparse
import collections
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (images per group, keypoints per image, set, dist, ratio)
    "c1": (2, 5000, "bank", 0.22, 1.0),
    "c2": (10, 20000, "iid", 1.0, 1.0),
    "c3": (50, 50000, "iid", 1.0, 1.0),
    "c4": (200, 20000, "iid", 1.0, 0.8),
}
FLOP_PER_PAIR = 96.0  # 2 * D, D = 48


def measured_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of one scoring-kernel launch, from the committed ncu capture
    of this workload (profiles/); None for workloads that were not captured."""
    p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if not os.path.exists(p):
        return None
    j = json.load(open(p))
    return float(j["dram_bytes_read"] + j["dram_bytes_write"]) if j.get("workload") == workload else None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["bf16_tflops"]), float(j["hbm_gbs"]), "measured"
    return 1590.0, 6650.0, "fallback"


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference binary on a bounded sample


def cpu_sample_group(tmpdir, kind, n_images, n_points):
    from frog_b200 import synth
    return synth.write_group(tmpdir, kind, n_images, n_points, fmt="bin")


def run_reference(list_path, dist, ratio, threads):
    from frog_b200 import pairsbin
    from oracle import oracle
    out = os.path.join(os.path.dirname(list_path), "ref_pairs.bin")
    t0 = time.time()
    res = oracle.run_ref_binary([list_path, "-o", out, "-d", dist, "-d2", ratio], threads=threads)
    wall = time.time() - t0
    # the reference prints " : <seconds>s" after each phase; the third one is "Pairing" (match.cpp:655)
    secs = [float(x) for x in re.findall(r" : ([0-9.eE+-]+)s$", res.stdout, flags=re.M)]
    pairing = secs[2] if len(secs) >= 3 else wall
    # Loaded (phantom-inclusive, match.cpp:179-208) keypoint counts come from the pairs.bin the reference
    # wrote: its per-image stdout lines are printed from OpenMP threads without a lock and interleave.
    n = np.array([p.shape[0] for p in pairsbin.parse(out).points], np.float64)
    pairs = (n.sum() ** 2 - (n ** 2).sum()) / 2.0  # sum_{i<j} N_i N_j
    return pairs, pairing, wall


def sample_size(cores):
    """Bounded CPU sample: ~3 s of wall-clock per run at ~6.4e7 descriptor pairs/s/core,
    with at least as many image pairs as cores (the reference parallelises over image pairs only)."""
    n_img = int(np.ceil(np.sqrt(15.0 * cores))) + 1
    return max(8, min(64, n_img)), 5000


def cpu_baseline(workload):
    _, _, kind, dist, ratio = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    sample_images, sample_points = sample_size(cores)
    tmp = tempfile.mkdtemp(prefix="fm_cpu_")
    try:
        lst = cpu_sample_group(tmp, kind, sample_images, sample_points)
        pairs, pairing, _ = run_reference(lst, dist, ratio, cores)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return {"value": pairs / pairing, "unit": "descriptor pairs/s", "cores": cores, "kind": "reference",
            "sample": f"{sample_images} images x {sample_points} keypoints ({kind}), {sample_images * (sample_images - 1) // 2} "
                      f"image pairs, reference's own Pairing timer, -nt {cores}",
            "seconds": pairing}


def bench_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_img, n_pts, kind, dist, ratio = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    s_img, s_pts = sample_size(cores)
    tmp = tempfile.mkdtemp(prefix="fm_ref_")
    try:
        lst = cpu_sample_group(tmp, kind, s_img, s_pts)
        times, pairs = [], 0.0
        for it in range(args.warmup + args.steps):
            pairs, pairing, _ = run_reference(lst, dist, ratio, cores)
            if it >= args.warmup:
                times.append(pairing)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    t = float(np.mean(times))
    val = pairs / t
    sample = f"{s_img} images x {s_pts} keypoints ({kind}) per step = {pairs:.3g} descriptor pairs"
    line = {
        "impl": "reference", "metric": "descriptor pairs/sec (keypoint matching, match.cpp ComputeMatches)",
        "value": val, "unit": "descriptor pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args.workload, args.gpus), "reference_sample": sample},
        "cpu_baseline": {"value": val, "unit": "descriptor pairs/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": "descriptor pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(w, gpus):
    n_img, n_pts, kind, dist, ratio = WORKLOADS[w]
    s = f"{w}: {n_img} images x {n_pts} keypoints, {kind} SURF3D descriptors (D=48), -d {dist:g} -d2 {ratio:g}, all {n_img * (n_img - 1) // 2} image pairs"
    if gpus > 1:
        s += f"; one such group per GPU ({gpus} groups), match lists gathered to rank 0 over NCCL"
    return s


# --------------------------------------------------------------------------------------------------


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for l in self.proc.stdout:
            self.lines.append((time.time(), l.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                mx = float(f[1])
                if t0 <= ts <= t1 + 0.1:
                    sm.append(float(f[0]))
                    for nm, v in zip(names, f[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(nm)
            except ValueError:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def bench_ours(args):
    import torch
    import torch.distributed as dist

    from frog_b200 import capi, synth
    from frog_b200 import dist as fdist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the matcher has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # stdout carries exactly ONE JSON line: native libraries that write to fd 1 (NCCL prints its version banner
    # there) are pointed at stderr, the line itself goes to a private duplicate of the original stdout
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if world > 1:
        # NCCL on a high-priority stream: the list transfer of one group runs while the next group's scoring kernel
        # still has thread blocks waiting to be dispatched, and must not queue behind them
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    n_img, n_pts, kind, thr, ratio = WORKLOADS[args.workload]

    # ---- synthetic group of this rank, in pinned host memory -----------------------------------
    kps = [synth.make(kind, n_pts, rank * n_img + i) for i in range(n_img)]
    host = []
    for k in kps:
        d = torch.from_numpy(k.desc).pin_memory()
        s = torch.from_numpy(k.scale).pin_memory()
        l = torch.from_numpy(k.lap).pin_memory()
        host.append((d, s, l))
    h2d_bytes = sum(d.numel() * 4 + s.numel() * 4 + l.numel() * 4 for d, s, l in host)
    pf = [i for i in range(n_img) for j in range(i + 1, n_img)]
    ps = [j for i in range(n_img) for j in range(i + 1, n_img)]
    desc_pairs_rank = float(sum(kps[i].n * kps[j].n for i, j in zip(pf, ps)))

    # a real (non-default) stream: handle 0 would mean "the context's own stream" to fm_set_stream
    stream = torch.cuda.Stream(device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    rows_rank = sum(kps[j].n for j in ps)  # one outer-loop row per keypoint of image `second`, per image pair

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def upload_all(mm):
        for i, (d, s, l) in enumerate(host):
            mm.upload_raw(i, d.data_ptr(), s.data_ptr(), l.data_ptr(), d.shape[0], d.shape[1])

    # ---- resident: keypoints stay in HBM, K asynchronous fm_match calls back to back ---------------
    # Every call is queued with FM_FLAG_ASYNC (no host synchronisation inside the timed region); for N > 1 the
    # compacted lists go GPU-to-GPU to rank 0 in fixed-capacity buffers right behind the kernels that wrote them,
    # and the stream waits for that transfer one step later (it overlaps the next group's scoring).
    m = capi.Matcher(local)
    m.set_stream(stream.cuda_stream)
    upload_all(m)
    m.synchronize()
    recv = fdist.FixedGather(len(pf), 2 * rows_rank, dev, slots=2) if world > 1 else None
    pending, retired = collections.deque(), []

    def retire(keep):
        while len(pending) > keep:
            res, works = pending.popleft()
            for w in works:
                w.wait()  # stream-level: the compute stream waits for the NCCL transfer before the buffers are reused
            retired.append((res.stats(), res.total))
            res.free()

    step_no = [0]

    def step_resident():
        retire(1)
        res = m.match(pf, ps, thr, ratio, device_only=True, asynchronous=True)
        works = []
        if world > 1:
            cptr, pptr = res.device_pointers()
            works = recv.start(fdist.as_torch_u32(cptr, len(pf), dev), fdist.as_torch_u32(pptr, 2 * rows_rank, dev),
                               step_no[0] % 2)
        step_no[0] += 1
        pending.append((res, works))

    def timed_resident(steps, warmup):
        for _ in range(warmup):
            step_resident()
        retire(0)
        retired.clear()
        evs = []
        barrier()
        t0 = time.time()
        for k in range(steps + 1):
            if k < steps:
                flush.zero_()  # evict the group (38 MB at c2) from the 126 MB L2 between timed steps
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            if k < steps:
                step_resident()
            else:
                retire(0)  # the last transfers, inside a bracket of their own
            b.record(stream)
            evs.append((a, b))
        barrier()
        t1 = time.time()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), t0, t1

    sampler = ClockSampler(local) if rank == 0 else None
    with torch.cuda.stream(stream):
        ms_res, t0, t1 = timed_resident(args.steps, args.warmup)
    clocks = sampler.stop(t0, t1) if sampler else None
    stats = [o[0] for o in retired]
    matches = retired[-1][1]
    m.close()

    # ---- end to end: host buffers in, host lists out, through the C ABI ---------------------------
    # Two contexts on two streams, used alternately as a three-stage pipeline: while one runs group k's prep +
    # kernels, the other first hands group k-1's lists to pinned host memory (D2H; for N > 1 after the NCCL gather
    # to rank 0) and then uploads group k+1 (H2D).  Every step's H2D and D2H is inside the timed region, which is
    # ONE CUDA-event bracket around all K steps (L2 flush writes included).
    streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
    ctxs = [capi.Matcher(local), capi.Matcher(local)]
    for mm, st in zip(ctxs, streams):
        mm.set_stream(st.cuda_stream)
    e2e_gather = fdist.FixedGather(len(pf), 2 * rows_rank, dev, slots=2) if world > 1 else None
    e2e_no = [0]
    pinned_out = [(torch.empty(len(pf), dtype=torch.int32).pin_memory(), torch.empty(2 * rows_rank, dtype=torch.int32).pin_memory())
                  for _ in range(world if (world > 1 and rank == 0) else 0)]

    def e2e_start(mm, st):
        with torch.cuda.stream(st):
            flush.zero_()
            # prep + kernels queued, call returns; lists go to pinned host memory in e2e_finish (world == 1)
            return mm.match(pf, ps, thr, ratio, device_only=world > 1, asynchronous=True)

    host_ms = collections.defaultdict(float)  # host wall-clock per e2e phase (this rank), timed steps only

    class phase:
        def __init__(self, name):
            self.name = name

        def __enter__(self):
            self.t = time.perf_counter()

        def __exit__(self, *a):
            host_ms[self.name] += (time.perf_counter() - self.t) * 1e3

    # N > 1: list transfers of a finished group run beside the next group's upload, on streams of their own
    drain = [torch.cuda.Stream(device=dev) for _ in range(world if rank == 0 else 1)]

    def e2e_finish(res, st):
        if world == 1:
            with phase("wait_fetch"):
                res.wait()  # counts, then the lists to pinned host memory (D2H on the context's stream)
            d2h = res.total * 8 + res.n_pairs * 4
            res.free()
            return d2h
        with phase("wait_kernels"):
            res.wait()  # the group's kernels are done; its lists sit in the result's own device buffers
        with phase("gather_d2h"):
            cptr, pptr = res.device_pointers()
            pairs_cap = fdist.as_torch_u32(pptr, 2 * rows_rank, dev)
            slot = e2e_no[0] % 2
            e2e_no[0] += 1
            with torch.cuda.stream(drain[0]):
                works = e2e_gather.start(fdist.as_torch_u32(cptr, res.n_pairs, dev), pairs_cap, slot, per_peer=True)
            d2h = 0
            if rank == 0:
                # rank 0: its own lists, then every peer's counts and (capacity-sized) list buffer to pinned host memory,
                # each peer on a stream of its own so that its D2H copy starts as soon as ITS lists have arrived
                with torch.cuda.stream(drain[0]):
                    n = 2 * res.total
                    pinned_out[0][1][:n].copy_(pairs_cap[:n], non_blocking=True)
                    d2h += n * 4 + res.n_pairs * 4
                for r in range(1, world):
                    with torch.cuda.stream(drain[r]):
                        for w in works[r]:
                            w.wait()
                        pinned_out[r][0].copy_(e2e_gather.counts[slot][r], non_blocking=True)
                        pinned_out[r][1].copy_(e2e_gather.pairs[slot][r], non_blocking=True)
                    d2h += (len(pf) + 2 * rows_rank) * 4
                for st_r in drain:
                    st_r.synchronize()  # lists are on the host
            else:
                with torch.cuda.stream(drain[0]):
                    for w in works[0]:
                        w.wait()
                drain[0].synchronize()  # lists have left this GPU
        res.free()
        return d2h

    def timed_e2e(steps, warmup):
        d2h, total, prev = [], warmup + steps, None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctxs[0].clear()
        upload_all(ctxs[0])
        for k in range(total):
            cur, st = ctxs[k % 2], streams[k % 2]
            if k == warmup:
                if prev is not None:
                    d2h.append(e2e_finish(*prev))
                    prev = None
                barrier()  # every stream of every rank is idle: start the clock
                host_ms.clear()
                e0.record(st)
                cur.clear()
                upload_all(cur)  # pipeline fill: the first timed group's own upload is inside the region
            with phase("start_prep"):
                res = e2e_start(cur, st)  # group k: prep + kernels queued
            if k + 1 < total and k + 1 != warmup:
                # group k+1: H2D from pinned host memory, under group k's kernels.  Its context is the one that ran
                # group k-1; that group's lists live in the result's own buffers, not in the image arena.
                nxt = ctxs[(k + 1) % 2]
                with phase("clear"):
                    nxt.clear()
                with phase("upload_enqueue"):
                    upload_all(nxt)
            if prev is not None:
                d2h.append(e2e_finish(*prev))  # group k-1: lists to (rank 0's) pinned host memory, under group k's kernels
            prev = (res, st)
        d2h.append(e2e_finish(*prev))
        e1.record(streams[(total - 1) % 2])
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), d2h[warmup:]

    ms_e2e, d2h_list = timed_e2e(args.steps, max(1, args.warmup))
    d2h_bytes = float(np.mean(d2h_list))
    for mm in ctxs:
        mm.close()

    # whole-job aggregates
    agg = torch.tensor([desc_pairs_rank, float(h2d_bytes), d2h_bytes, float(sum(s["kernel_launches"] for s in stats)),
                        float(matches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    total_pairs, h2d_all, d2h_all, launches, matches_all = agg.tolist()

    if rank == 0:
        peak_tf, peak_gbs, which = peaks()
        sc_ms = float(np.mean([s["ms_score"] / max(1, s["score_launches"]) for s in stats]))
        pairs_per_launch = float(np.mean([s["descriptor_pairs"] / max(1, s["score_launches"]) for s in stats]))
        scored_per_launch = float(np.mean([s["scored_pairs"] / max(1, s["score_launches"]) for s in stats]))
        achieved = FLOP_PER_PAIR * pairs_per_launch / (sc_ms * 1e-3) / 1e12 if sc_ms > 0 else 0.0
        executed = 2.0 * 64.0 * scored_per_launch / (sc_ms * 1e-3) / 1e12 if sc_ms > 0 else 0.0
        s0 = stats[-1]
        line = {
            "metric": "descriptor pairs/sec (keypoint matching, match.cpp ComputeMatches)",
            "value": total_pairs * args.steps / (ms_res * 1e-3), "unit": "descriptor pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_res / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16 operands / f32 accumulate + exact f32 rescoring",
            "data": "synthetic",
            "config": {"workload": workload_name(args.workload, world), "l2": "256 MiB flush buffer written between timed steps",
                       "timing": "value: CUDA events per step on the launch stream, summed over steps (asynchronous fm_match calls, no host "
                                 "synchronisation between steps), max over ranks; e2e: one CUDA-event bracket around all steps, two contexts "
                                 "used alternately so a group's H2D upload overlaps the previous group's kernels",
                       "matches_per_step": matches_all},
            "e2e": {"value": total_pairs * args.steps / (ms_e2e * 1e-3), "unit": "descriptor pairs/s",
                    "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": d2h_all, "ms_per_step": ms_e2e / args.steps,
                    "rank0_host_ms_per_step": {k: round(v / args.steps, 4) for k, v in host_ms.items()}},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved / peak_tf, "traffic": measured_traffic(args.workload), "peak_source": which + " (cuBLAS bf16 burst)",
                         "kernel": "fm::score_kernel<false>", "kernel_ms": sc_ms,
                         "algorithmic_flop_per_launch": FLOP_PER_PAIR * pairs_per_launch,
                         "executed_tflops": executed, "scored_fraction": scored_per_launch / max(1.0, pairs_per_launch)},
            # stream compaction (HBM-bound): 4 B read per row by each of the count and scatter passes,
            # 8 B written per match, over the CUDA-event time of the three compaction kernels
            "compaction": {"bound": "hbm", "unit": "GB/s", "peak": peak_gbs, "peak_source": which,
                           "achieved": (8.0 * s0["rows"] + 8.0 * matches) / max(s0["ms_compact"] * 1e-3, 1e-12) / 1e9,
                           "frac": (8.0 * s0["rows"] + 8.0 * matches) / max(s0["ms_compact"] * 1e-3, 1e-12) / 1e9 / peak_gbs,
                           "kernel_ms": s0["ms_compact"]},
            "phases_ms": {k: s0[k] for k in ("ms_total", "ms_score", "ms_rescore", "ms_exact", "ms_compact", "ms_prep")},
            "rows_exact_frac": s0["rows_exact"] / max(1, s0["rows"]), "candidates_per_row": s0["candidates"] / max(1, s0["rows"]),
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            try:
                line["cpu_baseline"] = cpu_baseline(args.workload)
            except Exception as e:  # the baseline must never take the GPU numbers down with it
                line["cpu_baseline"] = {"value": None, "unit": "descriptor pairs/s", "cores": os.cpu_count(), "kind": "reference",
                                        "sample": f"failed: {e}"}
        print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        bench_reference(args)
    else:
        bench_ours(args)


if __name__ == "__main__":
    main()
