#!/usr/bin/env python3
"""bench.py -- stage-1 Smith-Waterman GCUPS of the B200 strip-wavefront engine (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (CUDA, through the C ABI)
  python bench.py --impl reference [--gpus N] ...              the reference's own CPU path (oracle/_ref)

One "step" = one complete stage-1 pass with block pruning (best score + end coordinate, exact tie-break) over ONE
synthetic pair that does not depend on N (strong scaling): by default the BASELINE config-3 pair (23M x 25M,
Drosophila-chromosome shape, the config the metric is quoted on at 1/2/4/8 GPUs) generated at scale 0.4 = 9.2M x 10M,
the largest scale at which the driver's 25 steps fit its 870 s per-N limit on one GPU; `--workload cfg3 --scale 1`
runs the full pair (profiles/r02_cfg3_*.json), `--workload cfg5` the 249M x 228M target.  N = 1 runs the single-GPU
persistent kernel; N > 1 the block-cyclic chain: column chunks dealt round-robin to the GPUs, slice borders stored
straight into the next GPU's memory from inside the strip kernel, jobs scheduled by on-device events, running best
shared by peer atomics (no collective on the data path; NCCL only carries the barrier and the final 12-byte bests).
value = m*n*K / (max over ranks of the timed region), like the reference counts GCUPS (sw_stage1.cpp:444-448).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

METRIC = "GCUPS (stage-1 SW, device-timed)"
BENCH_WORKLOAD = ("cfg3", 0.4)          # default pair of every arm and every N


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_REAL_STDOUT = None


def protect_stdout():
    """The driver reads ONE JSON line from stdout.  Libraries print there too (NCCL writes its version banner to
    stdout when NCCL_DEBUG=VERSION is set in the environment), so file descriptor 1 is pointed at stderr for the whole
    run and the JSON line goes to a private duplicate of the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def resolve_workload(args):
    name, scale = BENCH_WORKLOAD
    if args.workload:
        name, scale = args.workload, 1.0
    if args.scale is not None:
        scale = args.scale
    return name, scale


def make_workload(args):
    """The pair depends on (--workload, --scale) only -- never on the number of GPUs."""
    import synth
    name, scale = resolve_workload(args)
    return synth.make_config(name, scale)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index
        self.marks = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, windows):
        sm, mx, reasons = [], 0.0, set()
        for ts, line in self.rows:
            if not any(t0 <= ts <= t1 for t0, t1 in windows):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def inst_per_cell(kernel_used):
    """SASS thread-instructions per DP cell of the dominant kernel, from the committed ncu summary."""
    for name in ("r02_inst_per_cell.json", "r01_inst_per_cell.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            with open(p) as f:
                d = json.load(f)
            d["_file"] = "profiles/" + name
            return d.get("s16x2" if kernel_used == 2 else "s32"), d
    return None, {}


# ------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation (oracle/_ref, built from /root/reference sources)
# ------------------------------------------------------------------------------------------------------------
def ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "oracle_cpu_block")
    return p if os.path.exists(p) else None


def run_reference_sample(a, b, rows, cols, cores, workdir):
    """Time the reference CPU path (CPUBlockProcessor via AbstractBlockAligner, --fork over `cores` processes)
    on the top-left rows x cols sample of the workload.  Returns (seconds of stage 1, cells)."""
    import synth
    fa, fb = os.path.join(workdir, "A.fa"), os.path.join(workdir, "B.fa")
    synth.write_fasta(fa, a[:rows], "bench_A")
    synth.write_fasta(fb, b[:cols], "bench_B")
    exe = ref_binary()
    cmd = [exe, f"--work-dir={os.path.join(workdir, 'w')}", "--clear", "--verbose=0", "--stage-1", "--no-flush"]
    if cores > 1:
        cmd.append(f"--fork={cores}")
    else:
        cmd.append("--no-block-pruning")      # --fork force-disables pruning (libmasa.cpp:1318-1321): keep both modes comparable
    cmd += [fa, fb]
    t0 = time.perf_counter()
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=workdir)
    dt = time.perf_counter() - t0
    return dt, rows * cols


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if ref_binary() is None:
        emit({"impl": "reference", "unavailable": "oracle/_ref/oracle_cpu_block not built (reference mount absent at build time)"})
        return 0
    cores = max(1, min(os.cpu_count() or 1, 32))
    a, b = make_workload(args)
    m, n = a.size, b.size
    # bounded sample: ~1.2e9 cells per core per step (about 5-8 s at the reference's ~0.2 GCUPS/core)
    side = int(min(m, n, (1.2e9 * cores) ** 0.5))
    side = max(2000, side)
    times = []
    with tempfile.TemporaryDirectory() as td:
        for it in range(args.warmup + args.steps):
            dt, cells = run_reference_sample(a, b, side, side, cores, td)
            if it >= args.warmup:
                times.append(dt)
            log(f"[reference] step {it}: {side}x{side} in {dt:.2f}s = {cells/dt/1e9:.3f} GCUPS on {cores} cores")
    total = sum(times)
    val = side * side * len(times) / total / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total / len(times) * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": workload_name(args, m, n), "sample": f"top-left {side}x{side} cells per step, every cell computed "
                   "(--fork disables block pruning in the reference, libmasa.cpp:1318-1321): compare with the GPU arm's "
                   "config.gcups_computed_cells, not with its whole-matrix value"},
        "cpu_baseline": {"value": val, "unit": "GCUPS", "cores": cores, "kind": "reference",
                         "sample": f"oracle/_ref/oracle_cpu_block --stage-1 --fork={cores} on the top-left {side}x{side} of the workload (process wall time)"},
        "e2e": {"value": val, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


def workload_name(args, m, n):
    name, scale = resolve_workload(args)
    shape = {"cfg1": "1M x 1M pair", "cfg2": "bacterial-genome-shaped 5M x 5M pair", "cfg3": "Drosophila-chromosome-shaped 23M x 25M pair",
             "cfg4": "chr21-shaped 48M x 46M pair", "cfg5": "chr1-shaped 249M x 228M pair"}[name]
    sc = "" if scale == 1.0 else f" generated at scale {scale}"
    return (f"{name}: BASELINE {shape}{sc} = {m}x{n} (tools/synth.py, the same pair at every GPU count), "
            "SW stage-1 with block pruning: best score + end coordinate")


def stage_split(path):
    """Per-stage wall time (ms) from MASA-Core's own GLOBAL STATISTICS block of the work directory (libmasa.cpp timer events:
    SEQUENCES = FASTA load, STAGE1..STAGE6).  Stages 1-3 run B200Aligner, stage 4 and stage 5 the GPU substitutes."""
    out = {}
    try:
        for line in open(path):
            f = line.split()
            if len(f) >= 2 and f[0].endswith(":") and f[0][:-1] in ("SEQUENCES", "INIT", "STAGE1", "STAGE2", "STAGE3", "STAGE4", "STAGE5", "STAGE6", "TOTAL"):
                out[f[0][:-1].lower()] = float(f[1])
    except (OSError, ValueError):
        pass
    return out


def full_alignment_leg(td):
    """Second half of the BASELINE metric: wall time of a complete alignment (stages 1-6) of the config-2 pair (5M x 5M)
    through build/cudalign, the drop-in binary (host/B200Aligner behind MASA-Core's own CLI).  N = 1 only."""
    import synth
    exe = os.path.join(ROOT, "build", "cudalign")
    if not os.path.exists(exe):
        return {"unavailable": "build/cudalign not built (needs the reference headers at build time)"}
    a, b = synth.make_config("cfg2")
    fa, fb = os.path.join(td, "cfg2_A.fa"), os.path.join(td, "cfg2_B.fa")
    synth.write_fasta(fa, a, "synth_cfg2_A"); synth.write_fasta(fb, b, "synth_cfg2_B")
    wd = os.path.join(td, "w")
    t0 = time.perf_counter()
    p = subprocess.run([exe, f"--work-dir={wd}", "--clear", "--verbose=0", "--ram-size=8G", fa, fb], cwd=td,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    dt = time.perf_counter() - t0
    out = {"workload": "cfg2: 5M x 5M synthetic pair, SW stages 1-6, build/cudalign --ram-size=8G (process start -> alignment.00.txt closed)",
           "wall_s": dt, "rc": p.returncode}
    xp = os.path.join(wd, "crosspoints", "crosspoint_01.00")
    if os.path.exists(xp):
        out["stage1_crosspoint"] = open(xp).read().split("\n")[1]
    txt = os.path.join(wd, "alignment.00.txt")
    out["alignment_txt_bytes"] = os.path.getsize(txt) if os.path.exists(txt) else 0
    out["stage_ms"] = stage_split(os.path.join(wd, "statistics"))
    if p.returncode != 0:
        out["tail"] = p.stdout[-400:]
    return out


def reference_gpu_leg(td, timeout_s=120):
    """Optional kernel-vs-kernel context (SURVEY.md 8d "Reference GPU baseline"): stage 1 of the BASELINE cfg1 pair (1M x 1M, which
    under-fills a B200: 977 strips on 2368 warp slots) and of cfg3 at scale 0.1 (2.3M x 2.5M)
    through the reference's OWN CUDA aligner built for sm_100 (oracle/_ref/cudalign_ref_gpu: R/src/CUDAligner.cu kernels and
    host loop unmodified, texture references replaced by __ldg pointers -- oracle/build_ref_gpu.sh) and through build/cudalign,
    same FASTA files, same MASA-Core driver, same flags.  Time = MASA-Core's ALIGN timer around alignPartition.  N = 1 only;
    never part of `value`."""
    import re
    import synth
    ref = os.path.join(ROOT, "oracle", "_ref", "cudalign_ref_gpu")
    new = os.path.join(ROOT, "build", "cudalign")
    if not (os.path.exists(ref) and os.path.exists(new)):
        return {"unavailable": "oracle/_ref/cudalign_ref_gpu or build/cudalign not built (both need the reference mount at build time)"}
    res = {"reference_build": "R/src/CUDAligner.cu + CUDAligner.cpp unmodified except texture references -> __ldg (oracle/build_ref_gpu.sh), -arch=sm_100",
           "pairs": []}
    for cfg, scale in (("cfg1", 1.0), ("cfg3", 0.1)):
        a, b = synth.make_config(cfg, scale)
        fa, fb = os.path.join(td, f"{cfg}_A.fa"), os.path.join(td, f"{cfg}_B.fa")
        synth.write_fasta(fa, a, f"synth_{cfg}_A"); synth.write_fasta(fb, b, f"synth_{cfg}_B")
        out = {"workload": f"{cfg} x {scale}: {a.size} x {b.size} synthetic pair, SW stage 1 with block pruning (--stage-1 --no-flush), time = MASA-Core's ALIGN timer"}
        for key, exe in (("reference_gpu", ref), ("b200", new)):
            wd = os.path.join(td, f"w_{cfg}_{key}")
            p = subprocess.run([exe, f"--work-dir={wd}", "--clear", "--verbose=0", "--stage-1", "--no-flush", fa, fb], cwd=td,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout_s)
            if p.returncode != 0:
                out[key] = {"rc": p.returncode, "tail": p.stdout[-300:]}
                continue
            st = open(os.path.join(wd, "statistics_01.00")).read()
            ms = float(re.search(r"ALIGN:\s+([0-9.]+)", st).group(1))
            out[key] = {"align_ms": ms, "gcups": a.size * b.size / ms / 1e6,
                        "stage1_crosspoint": open(os.path.join(wd, "crosspoints", "crosspoint_01.00")).read().split("\n")[1]}
        if "gcups" in out.get("reference_gpu", {}) and "gcups" in out.get("b200", {}):
            out["same_result"] = out["reference_gpu"]["stage1_crosspoint"] == out["b200"]["stage1_crosspoint"]
            out["speedup"] = out["b200"]["gcups"] / out["reference_gpu"]["gcups"]
        res["pairs"].append(out)
    return res


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"],
                    help="BASELINE pair to run instead of the default bench pair (cfg3 at scale 0.4)")
    ap.add_argument("--scale", type=float, default=None, help="scale of the generated pair (default: 0.4 for the bench pair, 1.0 with --workload)")
    ap.add_argument("--kernel", default="auto", choices=["auto", "s32", "s16x2"])
    ap.add_argument("--chunk-cols", type=int, default=0, help="N > 1: column chunk width (0 = automatic, < 0 = one contiguous slice per GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full-alignment", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true", help="skip the optional reference-GPU-kernel context leg (N = 1)")
    ap.add_argument("--no-e2e", action="store_true", help="records of the very large pairs: skip the end-to-end leg")
    ap.add_argument("--no-pruning", action="store_true", help="compute every cell (the reference's --no-block-pruning)")
    args = ap.parse_args()
    protect_stdout()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    from __graft_entry__ import load_package
    b200 = load_package()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            log(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks")
            return 2
    if not torch.cuda.is_available():
        log("bench.py: no CUDA device; the product path has no CPU fallback")
        return 2
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    a, b = make_workload(args)
    m, n = a.size, b.size
    kern = {"auto": b200.KERNEL_AUTO, "s32": b200.KERNEL_S32, "s16x2": b200.KERNEL_S16X2}[args.kernel]
    al = b200.Aligner(device=local, kernel=kern)
    if world > 1:
        al.mgpu_setup(dist, rank, world, m, n, args.chunk_cols)
    al.set_sequences(a, b)                                         # sequences resident in HBM for the `value` leg

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    prune = not args.no_pruning and args.kernel != "s32"

    def one_step(e2e):
        if e2e:
            al.set_sequences(a, b)                                 # H2D of both sequences from host memory
        # block pruning on, as in the reference's stage 1 (C/stage1/sw_stage1.cpp:219-225); GCUPS counts the whole
        # matrix like the reference does (sw_stage1.cpp:444-448), the cells actually computed are reported too
        if world > 1:
            return al.align_partition(0, 0, m, n, want_best_score=True, use_callbacks=False, mgpu=True, prune=prune,
                                      chunk_cols=args.chunk_cols)
        return al.align_partition(0, 0, m, n, want_best_score=True, use_callbacks=False, prune=prune)

    sampler = ClockSampler(local)
    sampler.start()
    res = None
    for _ in range(args.warmup):
        flush.zero_()
        barrier()
        res = one_step(False)
    launches0 = al.kernel_launches()
    windows, step_times, dev_ms = [], [], []
    for _ in range(args.steps):
        flush.zero_()                                              # L2 flush between timed iterations (untimed)
        barrier()
        w0 = time.time(); t0 = time.perf_counter()
        res = one_step(False)
        barrier()
        t1 = time.perf_counter(); w1 = time.time()
        step_times.append(t1 - t0); dev_ms.append(res["device_ms"]); windows.append((w0, w1))
    launches = al.kernel_launches() - launches0
    # e2e leg: same metric through the public C ABI with HOST buffers (sequence upload + result read-back timed)
    e2e_times = []
    for _ in range(0 if args.no_e2e else max(2, min(args.steps, (args.steps + 3) // 4))):
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        res_e = one_step(True)
        barrier()
        e2e_times.append(time.perf_counter() - t0)
        if tuple(res_e["best"]) != tuple(res["best"]):
            log("bench.py: the end-to-end step returned a different best cell"); return 3
    sampler.stop()

    total = sum(step_times)
    e2e_total = sum(e2e_times)
    my = dict(best=tuple(res["best"]), cells=int(res["cells"]), device_ms=sum(dev_ms) / max(len(dev_ms), 1),
              warp_busy=res["warp_busy"], warps=res["warps"])
    if dist is not None:
        t = torch.tensor([total, e2e_total, sum(dev_ms) / 1e3], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total, e2e_total, dev_total = (float(x) for x in t.tolist())
        per = [None] * world
        dist.all_gather_object(per, my)
    else:
        dev_total = sum(dev_ms) / 1e3
        per = [my]
    best = b200.merge_best([p["best"] for p in per])
    cells = sum(p["cells"] for p in per)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    K = len(step_times)
    value = m * n * K / total / 1e9
    clocks = sampler.summary(windows)
    peaks, peak_src = measured_peaks()
    kernel_used = res["kernel_used"]
    ipc, ipc_doc = inst_per_cell(kernel_used)
    f_mhz = clocks["sm_mhz"] or peaks.get("sm_max_mhz", 1965.0)
    # raw rate of the dominant kernel per GPU: cells actually COMPUTED / kernel time (CUDA events on the launch stream)
    kernel_gcups = cells * K / dev_total / 1e9 / args.gpus
    roofline = {"bound": "int-issue", "unit": "GCUPS", "achieved": kernel_gcups, "traffic": ipc_doc.get("dram_bytes_per_launch"),
                "traffic_def": ipc_doc.get("dram_def"),
                "achieved_def": "cells actually computed (block pruning skips the rest) / strip-kernel time from CUDA events on its launch stream, per GPU (max over GPUs of the kernel time)"}
    if ipc:
        peak = 148 * 4 * 32 * f_mhz * 1e6 / ipc / 1e9
        roofline.update({"peak": peak, "frac": kernel_gcups / peak, "inst_per_cell": ipc,
                         "peak_def": f"148 SMs x 4 schedulers x 32 lanes x {f_mhz:.0f} MHz (median SM clock sampled under load) / {ipc} SASS thread-instr per cell ({ipc_doc.get('_file')})"})
        alu = ipc_doc.get("alu_slots_per_cell_s16x2" if kernel_used == 2 else "alu_slots_per_cell_s32")
        if alu:
            apeak = 148 * 64 * f_mhz * 1e6 / alu / 1e9
            roofline.update({"alu_pipe_peak": apeak, "alu_pipe_frac": kernel_gcups / apeak,
                             "alu_pipe_def": f"148 SMs x 64 lanes/clk (measured VIADDMNMX/VIMNMX issue rate, profiles/r01_pipe_rates.txt) / {alu} ALU-pipe slots per cell"})
    # HBM is not the bound: algorithmic border traffic (16 B per column per strip + 1 B of seq1) vs the measured copy peak
    strips = res["strips"]
    alg_bytes = strips * n * 17.0 / args.gpus
    roofline["hbm"] = {"achieved_gbs": alg_bytes * K / dev_total / 1e9, "peak_gbs": peaks.get("hbm_gbs"), "peak_source": peak_src}

    mean_cells = cells / float(args.gpus)
    line = {
        "metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": args.gpus, "steps": K, "warmup": args.warmup,
        "ms_per_step": total / K * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int16x2" if kernel_used == 2 else "int32", "data": "synthetic",
        "config": {"workload": workload_name(args, m, n), "m": m, "n": n, "recurrence": "SW affine +1/-3/-3/-2",
                   "l2": "256 MiB flush buffer written between timed iterations", "kernel": "s16x2" if kernel_used == 2 else "s32",
                   "strips": strips, "best": list(best), "block_pruning": bool(prune), "cells_computed": cells,
                   "cells_computed_frac": cells / float(m * n), "gcups_computed_cells": cells * K / total / 1e9,
                   "multi_gpu": None if world == 1 else {"scheme": "block-cyclic column chunks, P2P border stores + event-driven job queues in the strip kernel",
                                                          "chunks": res["chunks"], "chunk_cols": res["chunk_cols"]},
                   "per_gpu": [{"cells_computed": p["cells"], "kernel_ms_per_step": p["device_ms"], "best": list(p["best"]),
                                "warp_busy": p["warp_busy"], "resident_warps": p["warps"]} for p in per],
                   "per_gpu_cells_max_over_mean": max(p["cells"] for p in per) / mean_cells if mean_cells else None,
                   "published_other_hw": "README: 5Mx5M 48.98 GCUPS on GTX 560 Ti (all stages); 249Mx228M 82,822 GCUPS on 512xV100"},
        "device_ms_per_step": dev_total / K * 1e3,
        "clocks": clocks,
        "gpu_launches": int(launches),
        "roofline": roofline,
    }
    if e2e_times:
        line["e2e"] = {"value": m * n * len(e2e_times) / e2e_total / 1e9, "unit": "GCUPS", "steps": len(e2e_times),
                       # pure A/C/G/T sequences cross PCIe 2-bit packed (16 bases per 32-bit word), once per GPU
                       "h2d_bytes_per_step": 4 * ((m + 15) // 16 + (n + 15) // 16) * args.gpus, "d2h_bytes_per_step": int(strips * 16 + 32) * args.gpus}
    if world == 1 and not args.no_full_alignment:
        with tempfile.TemporaryDirectory() as td:
            line["full_alignment"] = full_alignment_leg(td)
        line["full_alignment_s"] = line["full_alignment"].get("wall_s")
    if world == 1 and not args.no_reference_gpu:
        try:
            with tempfile.TemporaryDirectory() as td:
                line["reference_gpu_kernel"] = reference_gpu_leg(td)
        except Exception as e:                                    # context only: never costs the bench line
            line["reference_gpu_kernel"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    if world == 1 and not args.no_cpu_baseline and ref_binary() is not None:
        side = min(m, n, 60_000)
        with tempfile.TemporaryDirectory() as td:
            dt, c = run_reference_sample(a, b, side, side, 1, td)
        line["cpu_baseline"] = {"value": c / dt / 1e9, "unit": "GCUPS", "cores": 1, "kind": "reference",
                                "sample": f"oracle/_ref/oracle_cpu_block --stage-1 --no-block-pruning (reference CPUBlockProcessor, 1 core) on the top-left {side}x{side} of the workload, {dt:.1f}s"}
    elif world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        side = min(m, n, 20_000)
        t0 = time.perf_counter()
        O.full_matrix(a[:side], b[:side], O.SW, want_last_col=False)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": side * side / dt / 1e9, "unit": "GCUPS", "cores": 1, "kind": "port",
                                "sample": f"oracle/gotoh_oracle.c scalar port on the top-left {side}x{side}, {dt:.1f}s"}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
