#!/usr/bin/env python
"""bench.py -- headline benchmark of the SA tableau-search hot path (BASELINE.json metric).

Workload (BASELINE.json configs[4], the configuration the metric is quoted on): query D2PHLB1 (n1 = 19, LTYPE=T
LORDER=T LSOLN=F) against a synthetic 100 000-structure database (bootstrap of the reference's 586 real structures,
seed 20240502, size-sorted), 128 restarts x 100 moves, production (Philox) streams.  One "step" = one query searched
against the whole database.  With N GPUs the fixed database is split into N cost-weighted shards, one process per
GPU, no data-path collective ("strong" scaling); value = all structures searched / max-over-ranks device time.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm
  python bench.py --impl reference [...]                      the reference's own `-c` CPU path, all host cores

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

DB_SIZE = 100_000
DB_SEED = 20240502
QUERY = "D2PHLB1"
RESTARTS = 128
MOVES = 100
# reference-width algorithmic on-chip bytes per move-eval for D2PHLB1 (n1=19, LORDER=T)
ALGO_BYTES_PER_MOVE = 148.0
SM_COUNT = 148
SMEM_BYTES_PER_CLK_PER_SM = 128
ISSUE_SLOTS_PER_CLK_PER_SM = 4


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text()), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_db(S):
    base = S.Database.read_packed(ROOT / "tests" / "golden" / "small586.satsdb")
    return base.bootstrap(DB_SIZE, DB_SEED, True)


def query_db(S):
    qs = S.Database.read_packed(ROOT / "tests" / "golden" / "queries.satsdb")
    return qs.select([qs.find(QUERY)])


# ------------------------------------------------------------------------------------------------ CPU arms
REF_BIN = ROOT / "oracle" / "_ref" / "cudaSaTabsearch_ref"


def stratified_sample(S, db, count):
    idx = np.linspace(0, len(db) - 1, count).round().astype(np.int32)
    return db.select(idx)


def _reference_processes(td, procs, per_proc):
    """Runs `cudaSaTabsearch_ref -c` once per prepared input file in `td`, side by side.  Returns (structures/s, detail):
    the rate counts every process's structures against the SLOWEST process's own 'host execution time'."""
    t0 = time.perf_counter()
    ps = [subprocess.Popen([str(REF_BIN), "-c", "-r", str(RESTARTS)], stdin=open(os.path.join(td, "in%d" % p)),
                           stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, cwd=td, text=True) for p in range(procs)]
    search_ms = []
    for pr in ps:
        err = pr.communicate()[1]
        if pr.returncode != 0:
            raise RuntimeError("reference binary failed: " + err[-500:])
        search_ms.append(sum(float(x) for x in re.findall(r"host execution time ([0-9.]+) ms", err)))
    wall = time.perf_counter() - t0
    slowest = max(search_ms) / 1e3            # the reference's own timer around sa_tabsearch_host (search only)
    return per_proc * procs / slowest, {"wall_s": wall, "slowest_search_s": slowest}


def run_reference_cpu(S, db, per_proc: int, procs: int):
    """cpu_baseline leg of our own arm: P independent `cudaSaTabsearch -c` processes on P shards of a size-stratified
    sample (the reference is single threaded with a process-global drand48 stream: BASELINE.md section 3)."""
    sample = stratified_sample(S, db, per_proc * procs)
    qs = query_db(S)
    with tempfile.TemporaryDirectory() as td:
        qs.write_ascii(os.path.join(td, "q.ascii"))
        qtext = Path(td, "q.ascii").read_text()
        for p in range(procs):
            shard = sample.select(np.arange(p, per_proc * procs, procs, dtype=np.int32))
            shard.write_ascii(os.path.join(td, "db%d.ascii" % p))
            Path(td, "in%d" % p).write_text("db%d.ascii\nT T F\n%s" % (p, qtext))
        return _reference_processes(td, procs, per_proc)


def run_reference_gpu(S, db, count: int):
    """The reference's own GPU kernel (sa_tabsearch_gpu<<<128,128>>>, cuRAND XORWOW, --use_fast_math) rebuilt for sm_100a
    and run on this box against a size-stratified sample, timed by its own stderr timer around the kernel."""
    sample = stratified_sample(S, db, count)
    qs = query_db(S)
    with tempfile.TemporaryDirectory() as td:
        qs.write_ascii(os.path.join(td, "q.ascii"))
        sample.write_ascii(os.path.join(td, "db.ascii"))
        Path(td, "in").write_text("db.ascii\nT T F\n" + Path(td, "q.ascii").read_text())
        runs = []
        for _ in range(3):            # a fresh process starts on an idle (down-clocked) GPU: keep the best of three
            pr = subprocess.run([str(REF_BIN), "-r", str(RESTARTS)], stdin=open(os.path.join(td, "in")), stdout=subprocess.DEVNULL,
                                stderr=subprocess.PIPE, cwd=td, text=True, timeout=600)
            if pr.returncode != 0:
                return None
            ms = [float(x) for x in re.findall(r"GPU execution time ([0-9.]+) ms", pr.stderr)]
            if ms:
                runs.append(sum(ms))
    if not runs:
        return None
    ms = [min(runs)]
    return {"value": count / (sum(ms) / 1e3), "unit": "structures/s", "kernel_ms": sum(ms), "kernel_ms_all_runs": runs,
            "sample": "%d structures (size-stratified sample of the 100k synthetic db), the reference's sa_tabsearch_gpu<<<128,128>>> "
                      "rebuilt with -arch=sm_100a and its own --use_fast_math, timed by its own 'GPU execution time'; best of 3 runs" % count}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def bootstrap_picks(n_src: int, orders: np.ndarray, count: int, seed: int):
    """The synthetic database of SURVEY 8(d) without the product library: the same xorshift64* draws and the same stable
    sort by order as sats_db_bootstrap (csrc/sats_host.cpp), so the reference arm searches the very structures our arm
    does.  Returns (source entry per synthetic entry, name number per synthetic entry)."""
    mask = (1 << 64) - 1
    x = seed or 0x9E3779B97F4A7C15
    pick = np.empty(count, np.int64)
    for k in range(count):
        x ^= x >> 12
        x ^= (x << 25) & mask
        x ^= x >> 27
        pick[k] = (((x * 0x2545F4914F6CDD1D) & mask) >> 33) % n_src
    pos = np.argsort(orders[pick], kind="stable")
    return pick[pos], pos


def reference_arm(args):
    """The reference's own `-c` CPU path on all host cores.  Nothing of the product is loaded in this process: the
    synthetic db is re-derived with numpy (bootstrap_picks) from the committed 586-structure fixture and written with the
    test-side ASCII writer (tests/_refio.py)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not REF_BIN.exists():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/cudaSaTabsearch_ref not built"}))
        return
    sys.path.insert(0, str(ROOT / "tests"))
    from _refio import Structure, read_packed, write_ascii_db, write_query_input
    base = read_packed(ROOT / "tests" / "golden" / "small586.satsdb")
    query = {s.name: s for s in read_packed(ROOT / "tests" / "golden" / "queries.satsdb")}[QUERY]
    src, num = bootstrap_picks(len(base), np.array([s.n for s in base]), DB_SIZE, DB_SEED)
    cores = host_cores()
    per_proc = 600                       # ~0.7 s of single-core work per process per step
    sample = np.linspace(0, DB_SIZE - 1, per_proc * cores).round().astype(np.int64)       # size-stratified
    vals = []
    with tempfile.TemporaryDirectory() as td:
        for p in range(cores):
            ents = [Structure("s%06d" % (num[k] % 1000000), base[src[k]].tab, base[src[k]].dmat) for k in sample[p::cores]]
            write_ascii_db(os.path.join(td, "db%d.ascii" % p), ents)
            write_query_input(os.path.join(td, "in%d" % p), "db%d.ascii" % p, True, False, [query])
        for it in range(args.warmup + args.steps):
            v, _ = _reference_processes(td, cores, per_proc)
            if it >= args.warmup:
                vals.append(v)
    value = float(np.mean(vals))
    sample_txt = ("%d processes x %d structures (size-stratified sample of the 100k synthetic db), reference -c path, "
                  "rate = structures / slowest process's own 'host execution time'" % (cores, per_proc))
    full = None
    fj = ROOT / "profiles" / "r02_reference_full100k.json"
    if fj.exists():                      # one recorded run of the WHOLE 100k db (P shards, slowest), pinning the extrapolation
        full = json.loads(fj.read_text())
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "structures/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * per_proc * cores / value,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32+f32", "data": "synthetic",
        "config": CONFIG, "move_evals_per_s": value * RESTARTS * MOVES,
        "cpu_baseline": {"value": value, "unit": "structures/s", "cores": cores, "kind": "reference", "sample": sample_txt},
        "e2e": {"value": value, "unit": "structures/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "recorded_full_db_run": full,
    }
    print(json.dumps(line))


METRIC = "db structures searched/sec at 128 restarts (SA move-evals/sec = value x 128 x 100)"
CONFIG = {"workload": "D2PHLB1 (n1=19, T T F) vs synthetic 100k-structure db (bootstrap of 586 real structures, seed "
                      "20240502, size-sorted), 128 restarts x 100 moves, Philox streams; db sharded over the GPUs "
                      "(cost-weighted LPT partition)",
          "db_structures": DB_SIZE, "restarts": RESTARTS, "queries_per_step": 1,
          "l2": "L2 flushed between timed steps (256 MiB write)"}


# ------------------------------------------------------------------------------------------------ our arm
def gather_plan(idx_all: np.ndarray, db_size: int) -> np.ndarray:
    """idx_all: the ranks' padded entry-index vectors laid end to end (original db index of every gathered slot, -1 =
    padding).  Returns `take` with scores_by_original_index = gathered[take]; checks that the shards tile the database."""
    valid = np.nonzero(idx_all >= 0)[0]
    if len(valid) != db_size or len(np.unique(idx_all[valid])) != db_size:
        raise RuntimeError("the shards do not tile the database: %d slots for %d structures" % (len(valid), db_size))
    take = np.empty(db_size, np.int64)
    take[idx_all[valid]] = valid
    return take


def ours(args):
    import torch
    import cuda_satabsearch_b200 as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = max(world, 1)
    if world == 1 and args.gpus > 1:
        raise SystemExit("launch N>1 with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    db = synthetic_db(S)
    qs = query_db(S)
    sr = S.Searcher(db, local, rank, n_gpus)
    n_local = sr.entries
    p = S.default_params(lorder=1, lsoln=0, restarts=RESTARTS, rng_mode=S.RNG_PHILOX, seed=1234)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    scores = np.zeros((1, len(db)), np.int32)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sr.upload(qs)
    for _ in range(max(args.warmup, 3)):
        sr.launch(p, 0, timed=True)
    # ---- device-timed kernel-only steps (inputs resident in HBM)
    clocks = ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()
    l0 = sr.launches
    dev_ms = 0.0
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        dev_ms += sr.launch(p, 0, timed=True)
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    launches = sr.launches - l0
    # ---- end to end: host query in, ALL scores in one host buffer out.
    # N = 1: the public one-shot call sats_search() (query H2D, kernels, scores D2H, scatter to original order).
    # N > 1: every rank uploads the query and searches its shard; the shards' int32 score vectors are then gathered to
    #        rank 0 over NCCL straight from the searchers' device buffers ("one tiny gather after"), brought into
    #        original db order on the device and copied to the host once -- a step ends with the full 100k-score row on rank 0.
    gather_ms = 0.0
    if dist is None:
        barrier()
        e0 = time.perf_counter()
        for _ in range(args.steps):
            sr.search(qs, p, scores=scores)
        barrier()
        e2e_ms = (time.perf_counter() - e0) * 1e3
        d2h_bytes = 4 * DB_SIZE
    else:
        class _Dev:                      # a raw device pointer as a CUDA array (the C ABI hands out plain pointers)
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}

        counts = torch.zeros(n_gpus, dtype=torch.int64, device=dev)
        counts[rank] = n_local
        dist.all_reduce(counts)
        cap = int(counts.max())
        idx_pad = torch.full((cap,), -1, dtype=torch.int32, device=dev)
        idx_pad[:n_local] = torch.from_numpy(sr.entry_index()).to(dev)
        idx_all = [torch.empty_like(idx_pad) for _ in range(n_gpus)] if rank == 0 else None
        dist.gather(idx_pad, idx_all, dst=0)
        if rank == 0:
            take = gather_plan(torch.cat(idx_all).cpu().numpy(), DB_SIZE)      # original index -> position in the gathered buffer
            take_dev = torch.from_numpy(take).to(dev)
            gathered = torch.empty(n_gpus * cap, dtype=torch.int32, device=dev)
            ordered = torch.empty(DB_SIZE, dtype=torch.int32, device=dev)
            host = torch.empty(DB_SIZE, dtype=torch.int32).pin_memory()
            scores = host.numpy().reshape(1, DB_SIZE)            # the step's result buffer: pinned, filled by one D2H copy
        else:
            gathered = torch.empty(n_gpus * cap, dtype=torch.int32, device=dev)
        pad = torch.zeros(cap, dtype=torch.int32, device=dev)
        view = {}                        # device pointer of the searcher's score buffer -> torch view of it

        def step(timed_gather):
            sr.upload(qs)
            sr.launch(p, 0)
            sr.sync()
            g0 = time.perf_counter()
            ptr, _, n, _ = sr.device_results()
            if ptr not in view:
                view[ptr] = torch.as_tensor(_Dev(ptr, n), device=dev)
            pad[:n].copy_(view[ptr])
            dist.all_gather_into_tensor(gathered, pad)          # one NCCL collective; every shard is ~4 B x D / N
            if rank == 0:
                torch.index_select(gathered, 0, take_dev, out=ordered)      # un-permute to original db order on the device
                host.copy_(ordered, non_blocking=True)
                torch.cuda.synchronize()
            else:
                torch.cuda.synchronize()
            return (time.perf_counter() - g0) * 1e3 if timed_gather else 0.0

        step(False)
        barrier()
        e0 = time.perf_counter()
        for _ in range(args.steps):
            gather_ms += step(True)
        barrier()
        e2e_ms = (time.perf_counter() - e0) * 1e3
        d2h_bytes = 4 * DB_SIZE
        if rank == 0:                    # the assembled row must be the unsharded result: spot-check against one local search
            ref_row = np.full((1, len(db)), np.iinfo(np.int32).min, np.int32)
            sr.search(qs, p, scores=ref_row)
            mine = sr.entry_index()
            assert np.array_equal(scores[0, mine], ref_row[0, mine]) and scores.min() > np.iinfo(np.int32).min
    clk = clocks.stop() if rank == 0 else None

    rank_ms = [dev_ms / args.steps]
    if dist is not None:
        every = [torch.zeros(1, device=dev, dtype=torch.float64) for _ in range(n_gpus)]
        dist.all_gather(every, torch.tensor([dev_ms / args.steps], device=dev, dtype=torch.float64))
        rank_ms = [float(x) for x in every]            # device-timed ms per step of every rank: the spread is the load imbalance
        t = torch.tensor([dev_ms, e2e_ms, float(launches), wall_ms], device=dev, dtype=torch.float64)
        mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms, e2e_ms, wall_ms = float(mx[0]), float(mx[1]), float(mx[3])
        launches = int(sm[2])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks, peak_kind = load_peaks()
    ms_per_step = dev_ms / args.steps
    value = DB_SIZE / (ms_per_step / 1e3)
    moves = value * RESTARTS * MOVES
    e2e_value = DB_SIZE / (e2e_ms / args.steps / 1e3)
    sm_mhz = peaks.get("sm_max_mhz", 1965.0)
    smem_peak = n_gpus * SM_COUNT * SMEM_BYTES_PER_CLK_PER_SM * sm_mhz * 1e6 / 1e9      # GB/s at max clock
    achieved = moves * ALGO_BYTES_PER_MOVE / 1e9
    prof = {}
    pj = ROOT / "profiles" / "latest.json"
    if pj.exists():
        prof = json.loads(pj.read_text())
    orders = db.orders().astype(np.int64)
    blob_bytes = float(((80 + 8 * orders * (orders + 1) + 15) // 16 * 16).sum())      # every entry blob is read once per query
    roofline = {
        "bound": "smem", "achieved": achieved, "peak": smem_peak, "unit": "GB/s", "frac": achieved / smem_peak,
        "traffic": prof.get("dram_bytes_per_step"),
        "note": "north star: the bound is shared-memory bandwidth / SM issue slots, not HBM or tensor cores. achieved = "
                "move-evals/s x 148 B (SURVEY 8d reference-width on-chip bytes per move, n1=19 LORDER=T); peak = N x 148 SMs x "
                "128 B/clk x %.0f MHz (max SM clock, %s); traffic = ncu dram bytes per step (%s)" % (
                    sm_mhz, peak_kind, prof.get("capture", "no capture")),
        "issue": {"warp_inst_per_move": prof.get("warp_inst_per_move"),
                  "avg_active_threads_per_inst": prof.get("avg_active_threads_per_inst"),
                  "ncu_issue_active_pct": prof.get("issue_active_pct_time_weighted"),
                  "ncu_smem_wavefront_pct_of_peak": prof.get("smem_wavefront_pct_of_peak_time_weighted"),
                  "achieved_warp_inst_per_s": moves * prof["warp_inst_per_move"] if prof.get("warp_inst_per_move") else None,
                  "peak_warp_inst_per_s": n_gpus * SM_COUNT * ISSUE_SLOTS_PER_CLK_PER_SM * sm_mhz * 1e6,
                  "frac": (moves * prof["warp_inst_per_move"] / (n_gpus * SM_COUNT * ISSUE_SLOTS_PER_CLK_PER_SM * sm_mhz * 1e6))
                  if prof.get("warp_inst_per_move") else None,
                  "note": "warp instructions per move-eval from the committed ncu capture x live move-evals/s, over 148 SMs x 4 "
                          "schedulers x max SM clock"},
        "hbm": {"algorithmic_bytes_per_step": blob_bytes, "algorithmic_gbs": blob_bytes / (ms_per_step / 1e3) / 1e9,
                "peak_gbs": peaks.get("hbm_gbs"), "peak_kind": peak_kind,
                "frac": blob_bytes / (ms_per_step / 1e3) / 1e9 / (n_gpus * peaks.get("hbm_gbs", 6650.0))},
    }
    cpu = None
    if n_gpus == 1 and REF_BIN.exists() and not args.no_cpu:
        v, det = run_reference_cpu(S, db, 4000, 1)
        cpu = {"value": v, "unit": "structures/s", "cores": 1, "kind": "reference",
               "sample": "4000 structures (size-stratified sample of the 100k synthetic db), reference -c path, one "
                         "process; rate = structures / its own 'host execution time' (%.1f s)" % det["slowest_search_s"]}
    ref_gpu = None
    if n_gpus == 1 and REF_BIN.exists() and not args.no_cpu:
        try:
            ref_gpu = run_reference_gpu(S, db, 20000)
        except Exception as exc:                       # never let the side measurement break the bench line
            ref_gpu = {"error": str(exc)[:200]}
    line = {
        "metric": METRIC, "value": value, "unit": "structures/s", "n_gpus": n_gpus, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int32+f32", "data": "synthetic", "config": CONFIG,
        "move_evals_per_s": moves, "wall_ms_per_step_incl_l2_flush": wall_ms / args.steps,
        "e2e": {"value": e2e_value, "unit": "structures/s",
                "h2d_bytes_per_step": int(n_gpus * ((576 + 8 * 19 * 19 + 15) // 16 * 16 + 12)),      # query blob + its offset / size words, per GPU
                "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": e2e_ms / args.steps,
                "gather_ms_per_step": (gather_ms / args.steps) if n_gpus > 1 else None,
                "path": "sats_search(): host query in, host scores out" if n_gpus == 1 else
                        "per rank upload + launch; NCCL all-gather of the shards' device score vectors; rank 0 un-permutes them "
                        "to original db order on the device and copies the 100k scores to pinned host memory "
                        "(gather_ms_per_step = that tail, rank 0)"},
        "gpu_launches": launches, "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
        "reference_gpu_same_box": ref_gpu,
        "local_entries_rank0": n_local, "rank_ms_per_step": rank_ms,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
