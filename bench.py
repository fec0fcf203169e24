#!/usr/bin/env python
"""Benchmark of the all2all common-k-mer counting path (BASELINE.json metric: k-mer-pair updates/s).

    python bench.py --gpus N --steps K --warmup W            # our arm (sm_100a kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the UNMODIFIED reference binary on the host cores

Workload (config.workload).  N = 1: BASELINE.json configs[1] — 1,000 synthetic 5 Mbp bacterial genomes, k=18, f=1.0,
dense all2all — written by the stand-alone generator kmer-db_b200/bin/kdbx-synth (host/synth.cpp: 4 clusters x 250,
every genome a copy of a random earlier cluster member with 0.5 % substitutions).  BOTH arms read the same .db file.
A "step" is one full all2all over that database.  The unit of work U (updates) is a property of the database:
U = sum_p l_p(2 n_p - l_p - 1)/2 = the number of `row[col] += w` executions of the reference's dense path
(SURVEY.md §8d).

  value    K*U / device time of K steps, inputs (the raw trie) resident in HBM, result left in HBM;
           device time = CUDA events on the library's stream around each step, summed, max over ranks.
  e2e      same metric through the public C-ABI calls with HOST buffers: every step copies the trie
           from pinned host memory (kdbx_load_patterns) and reads the matrix back.
  roofline the scatter-add kernel.  hbm: 12 algorithmic bytes per update (4 B id + 4 B cell read + 4 B cell write,
           SURVEY.md §8d) x U per launch / measured average launch duration, against the measured HBM copy bandwidth
           in MEASURED_PEAKS.json — an HBM-EQUIVALENT figure: the accumulators live in shared memory, so it is not a
           bound and exceeds 1.  smem_atomic: the physical bound — shared-memory reductions issued per second against
           the conflict-free red.shared.add.u32 peak of the microbenchmark (profiles/r01_microbench_atomics.txt).
  cpu_baseline  the reference binary (oracle/_ref/kmer-db all2all -t <all cores>) once on the SAME .db, its CSV kept
           and compared byte for byte with the CSV written from the GPU result (parity_checked).
  --impl reference   the unmodified reference binary on the host cores.  The whole configs[1] database runs once in every
           invocation (full_config_run; its CSV is kept for the GPU arm's cmp).  If warmup + steps such runs fit 60 % of
           --ref-budget-s (420 s; the driver gives a run of this script 870 s and asks for 25 runs of 30 s) they all are
           runs of the whole database (same_config: true); otherwise every step is a bounded sample — the same generator,
           genomes and clusters at 1/5 (1/10 ...) of the genome length — and the line carries both rates and their quotient
           (steps_rate_over_full_config_rate).  At N > 1 the weak-scaling database is N times larger: rank 0 then times the
           configs[1] database (1/N of the workload, said in cpu_baseline.sample); updates/s is a rate.

N > 1 (torchrun, one rank per GPU): the database is SHARDED.  kdbxh_partitioner cuts the trie into N sub-tries (pieces
of its depth-first preorder balanced on a cost model, plus the ancestor chain of each piece with num_kmers = 0; the
matrix is linear in num_kmers, so the parts' matrices add up); every rank stages only its own part, declares the band of
sample ids it covers (kdbx_set_sample_window), runs the complete single-GPU pipeline on it and ONE
ncclReduceScatter(uint32, sum) INSIDE the library (kdbx_all2all_dense_reduce_scatter*) leaves every rank with its block
of the packed triangle.  torch.distributed only carries the NCCL unique id, the barriers and the max over ranks.
  --scaling weak  (default)  N x 1000 genomes in 4N clusters of UNEQUAL sizes (cluster_skew 0.3): the shape of
                  BASELINE.json configs[2]; the partitioner has to cut inside clusters.
  --scaling strong           the configs[1] database cut into N parts.
In both, rank 0 generates and partitions once (host, untimed: it is the layout of the sharded database, like
`build`), writes the parts next to the database and every rank reads its own.
"""
import argparse
import datetime
import json
import os
import re
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
PKG = ROOT / "kmer-db_b200"
sys.path.insert(0, str(PKG))

ALGO_BYTES_PER_UPDATE = 12
# conflict-free red.shared.add.u32 rate of one B200, ids in registers (kmer-db_b200/tools/microbench.cu,
# profiles/r01_microbench_atomics.txt: "1 CTA/SM x 32 warps ... consecutive columns")
SMEM_ATOMIC_PEAK = 5.097e12


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--samples", type=int, default=1000, help="genomes per GPU")
    ap.add_argument("--clusters", type=int, default=4, help="clusters per GPU's share of the genomes")
    ap.add_argument("--genome-kmers", type=int, default=5_000_000)
    ap.add_argument("--k", type=int, default=18)
    ap.add_argument("--mu", type=float, default=0.005)
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--skew", type=float, default=0.3, help="cluster size spread of the multi-GPU weak-scaling database")
    ap.add_argument("--chunk-ids", type=int, default=0)
    ap.add_argument("--tile-cols", type=int, default=0)
    ap.add_argument("--unit-updates", type=int, default=0)
    ap.add_argument("--tile-rows", type=int, default=0)
    ap.add_argument("--scatter-threads", type=int, default=0)
    ap.add_argument("--upload-chunk-mb", type=int, default=0, help="payload bytes per chunk of the asynchronous upload (kdbx_config::upload_chunk_bytes); 0 = default")
    ap.add_argument("--chunked-lists", action="store_true", help="force the chunked parent-chain expansion")
    ap.add_argument("--list-form", choices=["auto", "ids", "boundaries"], default="auto",
                    help="form of the full sample lists (kdbx.h: KDBX_FLAG_ID_LISTS / KDBX_FLAG_BOUNDARY_LISTS); auto = the library decides")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak")
    ap.add_argument("--no-window", action="store_true", help="N>1: do not declare the parts' sample windows")
    ap.add_argument("--check-reference", action="store_true",
                    help="N>1: run the reference binary on the whole (N times larger) database once for the CSV cmp, however long it takes")
    ap.add_argument("--keep-cache", action="store_true", help="N>1: keep the weak-scaling database and the parts under --cache-dir")
    ap.add_argument("--ref-budget-s", type=float, default=float(os.environ.get("KDBX_REF_BUDGET_S", "420")),
                    help="--impl reference: wall-clock budget of the whole invocation; steps become bounded samples when runs of the whole database do not fit")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cache-dir", default=os.environ.get("KDBX_CACHE", "/tmp/kdbx_cache"))
    return ap.parse_args()


# ---- the workload: one .db file per (shape, seed), written by the stand-alone generator -----------------------
def db_shape(a, world):
    """(samples, clusters, skew) of the database the run is about."""
    if world > 1 and a.scaling == "weak":
        return a.samples * world, a.clusters * world, a.skew
    return a.samples, a.clusters, 0.0


def db_path(a, samples, clusters, skew, genome_kmers=None):
    sk = f"_sk{skew:g}" if skew else ""
    return Path(a.cache_dir) / f"synth_n{samples}_c{clusters}_L{genome_kmers or a.genome_kmers}_k{a.k}_mu{a.mu}_s{a.seed}{sk}.db"


def ensure_db(a, samples, clusters, skew, genome_kmers=None):
    """Path and totals of the database; generated by bin/kdbx-synth (a host program: no CUDA) when missing."""
    path = db_path(a, samples, clusters, skew, genome_kmers)
    meta = Path(str(path) + ".json")
    if not (path.exists() and meta.exists()):
        exe = PKG / "bin" / "kdbx-synth"
        if not exe.exists():
            raise RuntimeError(f"{exe} missing: run `make -C {PKG}` (or __graft_entry__.build())")
        path.parent.mkdir(parents=True, exist_ok=True)
        t0 = time.time()
        subprocess.run([str(exe), "-o", str(path), "-n", str(samples), "-c", str(clusters), "-L", str(genome_kmers or a.genome_kmers), "-k", str(a.k),
                        "-mu", repr(a.mu), "-seed", str(a.seed), "-skew", repr(skew)], check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL)
        print(f"[bench] generated {path.name} in {time.time() - t0:.1f} s", file=sys.stderr)
    return path, json.loads(meta.read_text())


def workload_name(a, samples, clusters, skew, genome_kmers=None):
    L = genome_kmers or a.genome_kmers
    s = (f"{samples} synthetic {L / 1e6:g} Mbp genomes, {clusters} clusters"
         + (f" of unequal sizes (spread {skew:g})" if skew else "") + f", mu={a.mu}, k={a.k}, f=1.0, dense all2all")
    return s + (" (BASELINE.json configs[1])" if (samples, clusters) == (1000, 4) and L == 5_000_000 else
                " (BASELINE.json configs[2] shape)" if samples > 1000 else "")


def common_config(a, meta, samples, clusters, skew, genome_kmers=None):
    """The part of `config` that names the workload: identical in both arms."""
    return {"workload": workload_name(a, samples, clusters, skew, genome_kmers), "database": db_path(a, samples, clusters, skew, genome_kmers).name,
            "num_samples": int(meta["num_samples"]), "num_patterns": int(meta["num_patterns"]),
            "updates_per_step": int(meta["updates"]), "sum_n": int(meta["sum_n"]), "sum_l": int(meta["sum_l"])}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).
    nvidia-smi is started well before the timed region (its NVML start-up holds driver locks for tens of
    milliseconds and stalled whole steps when it began inside the region); only samples taken between
    begin() and stop() are reported."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines = []
        self.t_begin = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def wait_first_sample(self, timeout=20.0):
        t0 = time.perf_counter()
        while self.p and not self.lines and time.perf_counter() - t0 < timeout:
            time.sleep(0.05)

    def begin(self):
        self.t_begin = time.perf_counter()

    def stop(self):
        t_end = time.perf_counter()
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)  # one more period, so that a short region still gets its closing sample
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0 = self.t_begin if self.t_begin is not None else 0.0
        for ts, ln in self.lines:
            if ts < t0 or ts > t_end + 0.12:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ref_csv_path(db):
    return Path(str(db) + ".ref.csv")


def run_reference_binary(db, threads, keep_csv=True):
    """Times the reference's own all2all (the span it prints as 'OK (x seconds)' after
    'Calculating matrix of common k-mers...', src/console_all2all.cpp:31-36).  The CSV it writes stays next to the
    database (<db>.ref.csv): the GPU arm compares its own CSV with it."""
    exe = ROOT / "oracle" / "_ref" / "kmer-db"
    if not exe.exists():
        raise RuntimeError("oracle/_ref/kmer-db missing (run oracle/build_ref.sh where /root/reference is mounted)")
    out = ref_csv_path(db)
    tmp = Path(str(out) + ".tmp%d" % os.getpid())
    r = subprocess.run([str(exe), "all2all", "-t", str(threads), str(db), str(tmp)], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference all2all failed: " + r.stderr[-300:])
    txt = r.stdout + r.stderr
    m = re.search(r"Calculating matrix of common k-mers\.\.\..*?OK \(([0-9.eE+-]+) seconds\)", txt, re.S)
    if not m:
        raise RuntimeError("could not parse the reference's timing line")
    if keep_csv:
        os.replace(tmp, out)
    else:
        tmp.unlink(missing_ok=True)
    return float(m.group(1))


def files_identical(a, b, chunk=1 << 24):
    if os.path.getsize(a) != os.path.getsize(b):
        return False
    with open(a, "rb") as fa, open(b, "rb") as fb:
        while True:
            x, y = fa.read(chunk), fb.read(chunk)
            if x != y:
                return False
            if not x:
                return True


def ncu_capture(kernel):
    """figures of the committed ncu --set full capture of this kernel (profiles/scatter_traffic.json), if there is one"""
    p = ROOT / "profiles" / "scatter_traffic.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            if d.get("kernel", "k_scatter_add") == kernel:
                return d
        except Exception:
            return {}
    return {}


REF_SAMPLE_FRACTIONS = (5, 10, 20, 50, 100)   # the bounded sample of a step: the same database shape at 1/5 ... 1/100 of the genome length


def reference_arm(a, rank, world):
    """The unmodified reference on the host cores.  Loads none of this repository's libraries.

    The whole database runs ONCE in every invocation (full_config_run; its CSV is what the GPU arm compares with).  When
    warmup + steps runs of the whole database fit --ref-budget-s they all are such runs (same_config: true).  Otherwise —
    the driver's 25 runs of configs[1] take a quarter of an hour on 16 cores and its limit per run of this script is
    870 s — every step is a bounded sample: the same generator, the same number of genomes and clusters (so the same row
    lengths and list shapes) at a fraction of the genome length, i.e. a fraction of the patterns; the line then carries
    both rates (the steps' and the whole database's) and their quotient."""
    if rank != 0:
        return
    t_begin = time.perf_counter()
    cores = os.cpu_count() or 1
    samples, clusters, skew = db_shape(a, world)
    scale_note = ""
    if world > 1 and a.scaling == "weak":
        # the N-GPU database is N times configs[1]; the reference runs configs[1] itself
        base = (a.samples, a.clusters, 0.0)
        path, meta = ensure_db(a, *base)
        scale_note = (f"the {workload_name(a, *base)} database = about 1/{world} of the {world}-GPU workload "
                      f"({samples} genomes, {clusters} clusters): runs of the full one do not fit the driver's limit; ")
        cfg = {"workload": workload_name(a, samples, clusters, skew), "database": db_path(a, samples, clusters, skew).name,
               "num_samples": samples, "sample_of_workload": common_config(a, meta, *base)}
    else:
        base = (samples, clusters, skew)
        path, meta = ensure_db(a, *base)
        cfg = common_config(a, meta, *base)
    t_gen = time.perf_counter() - t_begin
    U_full = int(meta["updates"])
    w0 = time.perf_counter()
    secs_full = run_reference_binary(path, cores)          # keeps <db>.ref.csv
    wall_full = time.perf_counter() - w0
    full_run = {"value": U_full / secs_full, "unit": "updates/s", "seconds": round(secs_full, 3), "wall_seconds": round(wall_full, 3),
                "updates": U_full, "database": path.name}
    more_full = max(0, a.warmup - 1) + a.steps             # (the run above is the first warm-up)
    elapsed = time.perf_counter() - t_begin
    # (runs of the whole database only when they fit with a wide margin — 60 % of the budget — so that boxes of different core
    #  counts, and the N = 1 .. 8 runs of one box, do not end up on different sides of the line)
    same_config = elapsed + more_full * wall_full * 1.05 <= 0.6 * a.ref_budget_s
    if same_config:
        for _ in range(max(0, a.warmup - 1)):
            run_reference_binary(path, cores)
        secs = [run_reference_binary(path, cores) for _ in range(a.steps)]
        U = U_full
        sample = scale_note + f"the whole workload: {meta['num_samples']} genomes, U={U:.4g} per step (same .db file as the GPU arm)"
        step_db = path.name
    else:
        gen_full = max(t_gen, 2.0 * wall_full)             # (the database may have come from the cache)
        den = REF_SAMPLE_FRACTIONS[-1]
        for d in REF_SAMPLE_FRACTIONS:
            if elapsed + gen_full / d + (a.warmup + a.steps) * wall_full * 1.3 / d <= a.ref_budget_s:
                den = d
                break
        L = max(min(1000, a.genome_kmers), a.genome_kmers // den)
        spath, smeta = ensure_db(a, *base, genome_kmers=L)
        for _ in range(a.warmup):
            run_reference_binary(spath, cores, keep_csv=False)
        secs = [run_reference_binary(spath, cores, keep_csv=False) for _ in range(a.steps)]
        U = int(smeta["updates"])
        cfg["step_sample"] = dict(common_config(a, smeta, *base, genome_kmers=L), genome_length_fraction=f"1/{a.genome_kmers // L}" if a.genome_kmers % L == 0 else L / a.genome_kmers)
        sample = (scale_note + f"every step = the same generator and shape at {L / 1e6:g} Mbp per genome instead of {a.genome_kmers / 1e6:g} "
                  f"({smeta['num_samples']} genomes, {base[1]} clusters, U={U:.4g} per step): {a.warmup + a.steps} runs of the whole database "
                  f"({wall_full:.1f} s each) do not fit the arm's budget of {a.ref_budget_s:g} s; the whole database ran once in this "
                  f"invocation: {secs_full:.2f} s = {U_full / secs_full:.4g} updates/s (full_config_run)")
        step_db = spath.name
    total = sum(secs)
    v = U * a.steps / total
    cfg["l2_policy"] = "CPU arm: every step is a fresh process reading the .db from the page cache"
    print(json.dumps({
        "impl": "reference", "metric": "k-mer-pair updates/sec on all2all", "value": v, "unit": "updates/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total / a.steps,
        "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": cfg, "same_config": bool(same_config), "full_config_run": full_run,
        "steps_rate_over_full_config_rate": v / full_run["value"],
        "arm": {"binary": "oracle/_ref/kmer-db 2.3.1 all2all (unmodified reference, built by oracle/build_ref.sh)", "threads": cores,
                "step_database": step_db, "updates_per_step": U, "seconds_per_step": [round(x, 3) for x in secs],
                "csv_kept": str(ref_csv_path(path).name), "budget_s": a.ref_budget_s,
                "wall_s": round(time.perf_counter() - t_begin, 1)},
        "cpu_baseline": {"value": v, "unit": "updates/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def _own_cells(rank, block, cells):
    return max(0, min(cells, (rank + 1) * block) - min(cells, rank * block))


def _broadcast_bytes(dist, torch, payload):
    """rank 0's 128-byte NCCL unique id to every rank (torch.distributed is the launcher's plumbing here)."""
    t = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if payload is not None:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1

    if a.impl == "reference":
        reference_arm(a, rank, world)
        return

    import numpy as np
    import torch
    import kdbx

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on stdout at communicator creation; keep stdout for the JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            # (rank 0 generates and cuts the database while the others wait at a barrier: minutes on a box with few cores)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(minutes=40))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return int(x)
        t = torch.tensor([int(x)], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        return int(t.item())

    samples, clusters, skew = db_shape(a, world)
    if rank == 0:
        ensure_db(a, samples, clusters, skew)
    if world > 1:
        barrier()
    path, meta = ensure_db(a, samples, clusters, skew)   # (now only reads the side file)
    N, U_total = int(meta["num_samples"]), int(meta["updates"])
    cells = kdbx.tri_cells(N)
    flags = ((kdbx.FLAG_CHUNKED_LISTS if a.chunked_lists else 0) | kdbx.FLAG_ASYNC_UPLOAD |
             {"auto": 0, "ids": kdbx.FLAG_ID_LISTS, "boundaries": kdbx.FLAG_BOUNDARY_LISTS}[a.list_form])
    ctx = kdbx.Context(device=local_rank, chunk_ids=a.chunk_ids, tile_cols=a.tile_cols, unit_updates=a.unit_updates,
                       tile_rows=a.tile_rows, scatter_threads=a.scatter_threads, flags=flags,
                       upload_chunk_bytes=a.upload_chunk_mb << 20)
    t_shard = 0.0
    window = (0, N)
    if world == 1:
        trie = kdbx.Trie.read_db(path, pinned=True)
    else:
        # rank 0 cuts the database once and writes the parts next to it; every rank reads its own
        tag = f"{path}.{a.scaling}"
        part_path = Path(f"{tag}.part{rank}of{world}.db")
        if rank == 0 and not all(Path(f"{tag}.part{r}of{world}.db.json").exists() for r in range(world)):
            t0 = time.perf_counter()
            whole = kdbx.Trie.read_db(path, pinned=False)
            for r, (owned, win, upd, pats) in enumerate(whole.partition_write_all(world, f"{tag}.part")):   # one host thread per part
                Path(f"{tag}.part{r}of{world}.db.json").write_text(json.dumps({"window": list(win), "owned_updates": owned, "updates": upd,
                                                                              "num_patterns": pats}))
            whole.close()
            t_shard = time.perf_counter() - t0
            print(f"[bench] cut {path.name} into {world} parts in {t_shard:.1f} s", file=sys.stderr)
        barrier()
        trie = kdbx.Trie.read_db(part_path, pinned=True)
        window = tuple(json.loads(Path(str(part_path) + ".json").read_text())["window"])
        ctx.comm_init_rank(world, rank, _broadcast_bytes(dist, torch, kdbx.Context.comm_unique_id() if rank == 0 else None))
    tot = trie.totals()
    P = int(tot.num_patterns)
    # an independent check of the whole result that costs one pass over the trie: every pattern adds num_kmers to each
    # pair of its n samples (the flat form of the reference's sparse path), so the sum of all cells of the matrix is
    # sum_p num_kmers_p * n_p (n_p - 1) / 2 — over the OWNED patterns of this rank's part (ancestors carry num_kmers = 0)
    arr = trie.arrays()
    nn = arr["n"].astype(np.uint64)
    with np.errstate(over="ignore"):   # exact modulo 2^64 (numpy wraps), reduced to 60 bits so that the ranks' shares add without overflow
        expected_sum = sum_over_ranks(int((arr["num_kmers"].astype(np.uint64) * (nn * (nn - np.uint64(1)) // np.uint64(2))).sum(dtype=np.uint64)) % (1 << 60)) % (1 << 60)
    del arr, nn

    def stage():
        ctx.load_patterns(trie)
        if world > 1 and not a.no_window:
            ctx.set_sample_window(*window)
    stage()
    block = ctx.block_cells() if world > 1 else cells
    d_out = torch.zeros(max(1, block), dtype=torch.int32, device="cuda")

    def step():
        """One all2all over this rank's resident (sub-)trie, the reduce-scatter included; returns the library's stats."""
        if world == 1:
            return ctx.all2all_dense_rows_device(0, N, d_out.data_ptr())
        return ctx.all2all_dense_reduce_scatter_device(d_out.data_ptr())[2]

    # ---- device-resident leg -------------------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None   # started now, so that it is settled before the timed steps
    st = step()  # one untimed call establishes this rank's share of U (and warms the allocator)
    for _ in range(max(0, a.warmup - 1)):
        st = step()
    U_rank = int(st.updates)
    assert U_rank == int(tot.updates), "updates executed on this rank != U of its (sub-)trie"
    U_exec = sum_over_ranks(U_rank)
    # sharding replicates only ancestor chains (num_kmers = 0 there, but their rows are still visited)
    assert U_total <= U_exec <= U_total * 1.001, "updates executed by all ranks != U of the database"
    if sampler:
        sampler.wait_first_sample()
    barrier()
    if sampler:
        sampler.begin()
    wall0 = time.perf_counter()
    dev_ms = scat_ms = 0.0
    per_step_ms = []
    launches = scat_launches = 0
    stage_ms = {"prepare": 0.0, "expand": 0.0, "bucket": 0.0, "scatter": 0.0, "collective": 0.0}
    for i in range(a.steps):
        st = step()
        assert st.updates == U_rank
        dev_ms += st.ms_total
        per_step_ms.append(round(st.ms_total, 3))
        scat_ms += st.ms_scatter
        launches += st.kernel_launches
        scat_launches += st.scatter_launches
        for k in stage_ms:
            stage_ms[k] += getattr(st, "ms_" + k)
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    clocks = sampler.stop() if sampler else None
    # every step ends in a collective (N > 1), so the ranks advance in lockstep: the library's device time of a step
    # (CUDA events on its stream, reduce-scatter included) already contains the wait for the slowest rank
    T_ms = max_over_ranks(dev_ms)
    value = U_total * a.steps / (T_ms / 1e3)
    own = _own_cells(rank, block, cells)
    checksum = sum_over_ranks(int((d_out[:own].to(torch.int64) & 0xFFFFFFFF).sum().item()) if own else 0)
    physical = sum_over_ranks(int(st.physical_updates))
    # what every rank did in the last step (the collective's time on a rank is mostly its wait for the slowest one)
    per_rank = None
    if dist is not None:
        mine = torch.tensor([st.ms_total - st.ms_collective, st.ms_scatter, st.ms_collective, float(window[1] - window[0]), float(st.list_form),
                             float(U_rank), float(P)], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [{"compute_ms": round(float(x[0]), 2), "scatter_ms": round(float(x[1]), 2), "collective_and_wait_ms": round(float(x[2]), 2),
                     "window_ids": int(x[3]), "list_form": ["ids", "run boundaries"][int(x[4])], "updates": int(x[5]), "patterns": int(x[6])}
                    for x in allr]
    assert checksum % (1 << 60) == expected_sum, "sum of the matrix != sum over patterns of num_kmers * n (n - 1) / 2"

    # ---- end-to-end leg: host trie -> H2D -> compute (-> reduce-scatter) -> D2H of every block ------------------
    # (right after the device-resident leg: the reference binary of the parity leg below keeps the GPU idle for half a
    #  minute, and the first steps after that ran 15 % slower while the clocks came back)
    e2e = None
    if not a.no_e2e:
        out_host = kdbx.pinned_empty(max(1, block), np.uint32)

        def e2e_step():
            stage()
            if world == 1:
                return ctx.all2all_dense_rows(0, N, out_host[:cells])[1]
            return ctx.all2all_dense_reduce_scatter(out_host)[2]
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            st2 = e2e_step()
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        barrier()
        got = sum_over_ranks(int(out_host[:own].astype(np.int64).sum()))
        assert got == checksum, "e2e result differs from the device-resident result"
        hdr_bytes = 24 if trie.view().parent_id32 else 40   # (32-bit mirrors of parent_id / num_kmers: kdbx_trie_view)
        if trie.view().payload_off:                          # (offsets travel too unless the payload is densely packed)
            hdr_bytes += 8
        h2d = sum_over_ranks(P * hdr_bytes + int(tot.payload_bytes))  # summed over the ranks (every rank copies its own shard)
        e2e = {"value": U_total * a.steps / e2e_s, "unit": "updates/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": cells * 4, "ms_per_step": 1e3 * e2e_s / a.steps,
               "ms_upload": st2.ms_upload, "ms_download": st2.ms_download,
               # device time of the compute call of the last step, from its first kernel to its last: it starts when the
               # headers have arrived and contains the decoder's waits for the payload chunks still on the link
               "ms_compute_after_headers": st2.ms_total, "ms_prepare": st2.ms_prepare}

    # ---- parity: the matrix of the timed configuration against the reference binary's CSV, byte for byte --------
    parity = None
    cpu_baseline = None
    host_full = None
    if world > 1:
        blocks = [torch.zeros_like(d_out) for _ in range(world)] if rank == 0 else None   # (a check, untimed)
        dist.gather(d_out, blocks, dst=0)
        if rank == 0:
            host_full = torch.cat(blocks)[:cells].cpu().numpy().view(np.uint32)
    else:
        host_full = d_out[:cells].cpu().numpy().view(np.uint32)
    if rank == 0 and not a.no_cpu_baseline:
        try:
            ref_csv = ref_csv_path(path)
            if world == 1 or (not ref_csv.exists() and (U_total < 4e11 or a.check_reference)):
                secs = run_reference_binary(path, cores)
                cpu_baseline = {"value": U_total / secs, "unit": "updates/s", "cores": cores, "kind": "reference",
                                "sample": f"the whole workload once: {N} genomes, U={U_total:.4g}, {secs:.2f} s, kmer-db 2.3.1 all2all -t {cores} "
                                          f"on the same .db file"}
            if ref_csv.exists():
                whole = kdbx.Trie.read_db(path, pinned=False) if world > 1 else trie
                ours = Path(str(path) + ".gpu.csv")
                whole.write_all2all_csv(host_full, ours)
                same = files_identical(ours, ref_csv)
                parity = {"parity_checked": "cmp of the CSV written from the GPU matrix with the CSV of oracle/_ref/kmer-db all2all on the same .db",
                          "csv_identical": bool(same), "csv_bytes": os.path.getsize(ours)}
                ours.unlink(missing_ok=True)
                if world > 1:
                    whole.close()
                assert same, "GPU CSV differs from the reference binary's CSV"
            else:
                parity = {"parity_checked": None, "why": "no reference CSV for this database (too large to run the reference here)"}
        except AssertionError:
            raise
        except Exception as e:  # the baseline is a reported number, not a dependency of the GPU path
            cpu_baseline = {"value": None, "unit": "updates/s", "cores": cores, "kind": "reference", "sample": f"failed: {e}"}
    host_full = None

    if world > 1 and not a.keep_cache:
        # the parts (and the N-times-larger weak-scaling database) are tens of GB under /tmp: do not leave them behind
        trie.close()
        barrier()
        for f in (part_path, Path(str(part_path) + ".json")):
            f.unlink(missing_ok=True)
        if rank == 0 and a.scaling == "weak":
            for f in (path, Path(str(path) + ".json"), ref_csv_path(path)):
                f.unlink(missing_ok=True)
    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    peak, peak_src = measured_hbm_peak()
    achieved = ALGO_BYTES_PER_UPDATE * U_rank * a.steps / (scat_ms / 1e3) / 1e9 if scat_ms > 0 else 0.0
    phys_rank = int(st.physical_updates)
    atomic_rate = phys_rank * a.steps / (scat_ms / 1e3) if scat_ms > 0 else 0.0
    kernel = "k_scatter_diff" if int(st.list_form) == 1 else "k_scatter_add"
    cap = ncu_capture(kernel)
    u_rate = U_rank * a.steps / (scat_ms / 1e3) if scat_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": cap.get("dram_bytes_per_launch"), "peak_source": peak_src,
                "note": "HBM-EQUIVALENT figure (12 B x U / kernel time): the accumulators live in shared memory, so it is not a bound; "
                        "the kernel's physical bound is smem_atomic",
                "algorithmic_bytes_per_update": ALGO_BYTES_PER_UPDATE,
                "updates_per_launch": U_rank / max(1, scat_launches / a.steps),
                "avg_launch_ms": scat_ms / max(1, scat_launches), "launches_per_step": scat_launches // max(1, a.steps),
                "kernel_share_of_step": scat_ms / dev_ms if dev_ms else None,
                "smem_atomic": {"achieved": atomic_rate, "peak": SMEM_ATOMIC_PEAK, "unit": "red.shared.add.u32 lane-updates/s",
                                "frac": atomic_rate / SMEM_ATOMIC_PEAK,
                                "peak_source": "conflict-free microbenchmark, ids in registers (profiles/r01_microbench_atomics.txt)",
                                "reductions_per_launch": phys_rank / max(1, scat_launches / a.steps),
                                "note": "lane-updates actually ISSUED (run-boundary lists issue fewer than U: two per run of consecutive ids); "
                                        "effective_frac counts the algorithmic updates U the launch delivers against the same peak; "
                                        "l1tex_busy_pct is the utilisation of the shared-memory pipe in the committed ncu capture "
                                        "(a reduction takes 2.5 bank-conflict wavefronts on sorted-but-gappy columns, so the conflict-free "
                                        "peak is not reachable by any schedule of these updates)",
                                "effective_frac": u_rate / SMEM_ATOMIC_PEAK,
                                "l1tex_busy_pct": cap.get("l1tex_throughput_pct_of_peak"),
                                "wavefronts_per_reduction": (cap.get("shared_atomic_wavefronts_per_launch") / cap.get("shared_atomic_instructions_per_launch"))
                                if cap.get("shared_atomic_instructions_per_launch") else None}}

    cfg = common_config(a, meta, samples, clusters, skew)
    cfg.update({
        "patterns_on_rank0": P,
        "parallelism": ("1 GPU, no collective" if world == 1 else
                        f"trie cut into {world} sub-tries by kdbxh_partitioner (preorder pieces + ancestor chains), one per GPU, "
                        f"sample windows declared, one ncclReduceScatter of the {cells * 4 / 1e6:.0f} MB matrix per step inside libkdbx.so"),
        "shard_seconds_host": t_shard, "sample_window_rank0": list(window),
        "l2_policy": "inputs (trie %.1f GB + lists) exceed the 126 MB L2; no explicit flush" % ((P * 40 + int(tot.payload_bytes)) / 1e9),
        "chunk_ids": a.chunk_ids, "tile_cols": a.tile_cols, "unit_updates": a.unit_updates,
        "tile_rows": a.tile_rows, "scatter_threads": a.scatter_threads, "list_form_requested": a.list_form,
        "upload_chunk_mb": a.upload_chunk_mb})
    line = {
        "metric": "k-mer-pair updates/sec on all2all", "value": value, "unit": "updates/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": T_ms / a.steps, "higher_is_better": True,
        "scaling": a.scaling, "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": cfg,
        "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks,
        "stage_ms_per_step": {k: v / a.steps for k, v in stage_ms.items()}, "wall_ms_per_step": wall_ms / a.steps,
        "library_ms_per_step": per_step_ms,
        "result_checksum": checksum, "checksum_identity": "sum of all cells == sum_p num_kmers_p n_p (n_p - 1) / 2 over the trie: checked",
        "list_form": ["ids", "run boundaries"][int(st.list_form)],
        "physical_updates_per_step": physical,
    }
    if per_rank:
        line["per_rank_last_step"] = per_rank
    if parity:
        line.update(parity)
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
