#!/usr/bin/env python
"""Benchmark of the all2all common-k-mer counting path (BASELINE.json metric: k-mer-pair updates/s).

    python bench.py --gpus N --steps K --warmup W            # our arm (sm_100a kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the UNMODIFIED reference binary on the host cores

Workload (config.workload): BASELINE.json configs[1] — 1,000 synthetic 5 Mbp bacterial genomes,
k=18, f=1.0, dense all2all — produced by the pattern-level generator (kmer-db_b200/host/synth.cpp:
4 clusters x 250, every genome a copy of a random earlier cluster member with 0.5 % substitutions).
A "step" is one full all2all over that database.  The unit of work U (updates) is a property of
the database: U = sum_p l_p(2 n_p - l_p - 1)/2 = the number of `row[col] += w` executions of the
reference's dense path.

  value    K*U / device time of K steps, inputs (the raw trie) resident in HBM, result left in HBM;
           device time = CUDA events on the library's stream around each step, summed, max over ranks.
  e2e      same metric through the public C-ABI calls with HOST buffers: every step copies the trie
           from pinned host memory (kdbx_load_patterns) and reads the matrix back (kdbx_all2all_dense_rows).
  roofline the scatter-add kernel: 12 algorithmic bytes per update (4 B id + 4 B cell read + 4 B cell
           write) x updates per launch / measured average launch duration, against the
           measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the reference binary (oracle/_ref/kmer-db all2all -t <all cores>) on a bounded sample:
           the first cluster (250 genomes x 5 Mbp = exactly a quarter of the workload's updates).
With N>1 (torchrun, one rank per GPU) the database is SHARDED: every rank holds, uploads and processes only its
own sub-trie (kdbxh_partition: a piece of the trie's depth-first preorder plus the ancestors of that piece with
num_kmers = 0; the matrix is linear in num_kmers, so the parts' matrices add up), runs the complete single-GPU
pipeline on it into a partial matrix, and ONE NCCL all-reduce (uint32 sum over NVLink) adds the partial matrices —
the exchange step BASELINE.json's north_star names.  Nothing is replicated but the few ancestor chains.
  --scaling weak  (default)  per-GPU work fixed: N GPUs process a database of N x 1000 genomes (4N clusters, the
                  shape of BASELINE.json configs[2]); rank r's shard is the configs[1] database laid onto the sample
                  ids [1000 r, 1000 (r+1)) (kdbxh_relabel) — what the partitioner yields for a database whose clusters
                  are disjoint subtrees.  U = N x U(configs[1]); the matrix is (1000 N)^2 / 2 cells.
  --scaling strong           total work fixed: the configs[1] database is cut into N parts by kdbxh_partition
                  (host, untimed: it is the layout of the sharded database); rank 0 also computes the unsharded
                  matrix once and the all-reduced result must equal it bit for bit.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "kmer-db_b200"))

ALGO_BYTES_PER_UPDATE = 12


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--samples", type=int, default=1000)
    ap.add_argument("--clusters", type=int, default=4)
    ap.add_argument("--genome-kmers", type=int, default=5_000_000)
    ap.add_argument("--k", type=int, default=18)
    ap.add_argument("--mu", type=float, default=0.005)
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--chunk-ids", type=int, default=0)
    ap.add_argument("--tile-cols", type=int, default=0)
    ap.add_argument("--unit-updates", type=int, default=0)
    ap.add_argument("--tile-rows", type=int, default=0)
    ap.add_argument("--scatter-threads", type=int, default=0)
    ap.add_argument("--chunked-lists", action="store_true", help="force the chunked parent-chain expansion")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak",
                    help="N>1: 'weak' = every rank owns a configs[1]-sized shard of an N-times larger database; "
                         "'strong' = the configs[1] database cut into N sub-tries; both end in one NCCL all-reduce")
    ap.add_argument("--emulate-shard", default="", help="debug, N=1 only: 'r/w' = run the weak-scaling shard of rank r of w "
                    "ranks alone (no collective): what that rank would execute in a w-GPU run")
    ap.add_argument("--list-form", choices=["auto", "ids", "boundaries"], default="auto",
                    help="form of the full sample lists (kdbx.h: KDBX_FLAG_ID_LISTS / KDBX_FLAG_BOUNDARY_LISTS); auto = the library decides")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cache-dir", default=os.environ.get("KDBX_CACHE", "/tmp/kdbx_cache"))
    return ap.parse_args()


def workload_name(a):
    return (f"{a.samples} synthetic {a.genome_kmers / 1e6:g} Mbp genomes, {a.clusters} clusters, mu={a.mu}, "
            f"k={a.k}, f=1.0, dense all2all (BASELINE.json configs[1] shape)")


def cache_path(a, samples, clusters):
    return Path(a.cache_dir) / f"synth_n{samples}_c{clusters}_L{a.genome_kmers}_k{a.k}_mu{a.mu}_s{a.seed}.db"


def get_workload(kdbx, a, samples, clusters, pinned, rank=0, barrier=None):
    """Generate (rank 0) or read the cached database.  Same seed => cluster c is identical whether
    generated alone or as part of the full workload, so the CPU sample is a true subset."""
    path = cache_path(a, samples, clusters)
    if rank == 0 and not path.exists():
        path.parent.mkdir(parents=True, exist_ok=True)
        t0 = time.time()
        t = kdbx.Trie.synth(num_samples=samples, num_clusters=clusters, genome_kmers=a.genome_kmers, k=a.k,
                            mutation_rate=a.mu, seed=a.seed)
        tmp = path.with_suffix(".tmp%d" % os.getpid())
        t.write_db(tmp)
        os.replace(tmp, path)
        t.close()
        print(f"[bench] generated {path.name} in {time.time() - t0:.1f} s", file=sys.stderr)
    if barrier:
        barrier()
    return kdbx.Trie.read_db(path, pinned=pinned), path


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


def run_reference_binary(db_path, threads):
    """Times the reference's own all2all (the span it prints as 'OK (x seconds)' after
    'Calculating matrix of common k-mers...', src/console_all2all.cpp:31-36)."""
    exe = ROOT / "oracle" / "_ref" / "kmer-db"
    if not exe.exists():
        raise RuntimeError("oracle/_ref/kmer-db missing (run oracle/build_ref.sh where /root/reference is mounted)")
    out = Path(db_path).with_suffix(".ref.csv")
    r = subprocess.run([str(exe), "all2all", "-t", str(threads), str(db_path), str(out)], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference all2all failed: " + r.stderr[-300:])
    txt = r.stdout + r.stderr
    m = re.search(r"Calculating matrix of common k-mers\.\.\..*?OK \(([0-9.eE+-]+) seconds\)", txt, re.S)
    if not m:
        raise RuntimeError("could not parse the reference's timing line")
    try:
        out.unlink()
    except OSError:
        pass
    return float(m.group(1))


def ncu_traffic_per_launch():
    """dram bytes per scatter launch from the committed ncu summary, if one exists."""
    p = ROOT / "profiles" / "scatter_traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    sample_n = max(1, a.samples // a.clusters)

    if a.impl == "reference":
        if rank != 0:
            return
        import kdbx
        t, path = get_workload(kdbx, a, sample_n, 1, pinned=False)
        U = int(t.totals().updates)
        t.close()
        for _ in range(a.warmup):
            run_reference_binary(path, cores)
        secs = [run_reference_binary(path, cores) for _ in range(a.steps)]
        total = sum(secs)
        v = U * a.steps / total
        sample = f"first cluster only: {sample_n} genomes x {a.genome_kmers / 1e6:g} Mbp, U={U:.4g} per step"
        print(json.dumps({
            "impl": "reference", "metric": "k-mer-pair updates/sec on all2all", "value": v, "unit": "updates/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total / a.steps,
            "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload_name(a), "sample": sample, "threads": cores},
            "cpu_baseline": {"value": v, "unit": "updates/s", "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
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
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
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

    trie, db_path = get_workload(kdbx, a, a.samples, a.clusters, pinned=True, rank=rank, barrier=barrier if world > 1 else None)
    tot = trie.totals()
    N0, P0, U0 = int(tot.num_samples), int(tot.num_patterns), int(tot.updates)
    chunk_ids = a.chunk_ids
    ctx = kdbx.Context(device=local_rank, chunk_ids=chunk_ids, tile_cols=a.tile_cols, unit_updates=a.unit_updates,
                       tile_rows=a.tile_rows, scatter_threads=a.scatter_threads, flags=(kdbx.FLAG_CHUNKED_LISTS if a.chunked_lists else 0) | kdbx.FLAG_ASYNC_UPLOAD |
                       {"auto": 0, "ids": kdbx.FLAG_ID_LISTS, "boundaries": kdbx.FLAG_BOUNDARY_LISTS}[a.list_form])
    scaling = a.scaling
    full_ref = None
    t_shard = 0.0
    if world == 1 and a.emulate_shard:
        r_, w_ = (int(x) for x in a.emulate_shard.split("/"))
        trie.relabel(r_ * N0, N0 * w_)
        N, U_total = N0 * w_, U0
    elif world == 1:
        N, U_total = N0, U0
    elif scaling == "weak":
        t0 = time.perf_counter()
        trie.relabel(rank * N0, N0 * world)   # this rank's shard of the N0*world-sample database
        t_shard = time.perf_counter() - t0
        N, U_total = N0 * world, U0 * world
    else:
        if rank == 0:  # the unsharded matrix, once, as the bit-exact check of the sharded path
            ctx.load_patterns(trie)
            full_ref = torch.zeros(max(1, kdbx.tri_cells(N0)), dtype=torch.int32, device="cuda")
            ctx.all2all_dense_rows_device(0, N0, full_ref.data_ptr())
        t0 = time.perf_counter()
        part, _owned = trie.partition(world, rank, pinned=True)
        t_shard = time.perf_counter() - t0
        trie.close()
        trie = part
        N, U_total = N0, U0
    tot = trie.totals()
    P = int(tot.num_patterns)
    ctx.load_patterns(trie)
    r0, r1 = 0, N
    cells = kdbx.tri_cells(N)
    d_out = torch.zeros(max(1, cells), dtype=torch.int32, device="cuda")

    def step():
        """One all2all over this rank's resident (sub-)trie (+ the all-reduce); returns the library's stats."""
        st = ctx.all2all_dense_rows_device(r0, r1, d_out.data_ptr())
        if dist is not None:
            dist.all_reduce(d_out)  # uint32 sums wrap like int32 sums: same bits
            # the library works on its own stream: the next step must not start (and zero d_out)
            # while this all-reduce is still in flight on NCCL's stream
            torch.cuda.current_stream().synchronize()
        return st

    def total_updates(st_updates):
        if dist is None:
            return st_updates
        t = torch.tensor([st_updates], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        return int(t.item())

    # ---- device-resident leg -------------------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None   # started now, so that it is settled before the timed steps
    st = step()  # one untimed call establishes this rank's share of U (and warms the allocator)
    for _ in range(max(0, a.warmup - 1)):
        st = step()
    U_rank = int(st.updates)
    assert U_rank == int(tot.updates), "updates executed on this rank != U of its (sub-)trie"
    U_exec = total_updates(U_rank)
    # sharding replicates only ancestor chains (num_kmers = 0 there, but their rows are still visited)
    assert U_total <= U_exec <= U_total * 1.001, "updates executed by all ranks != U of the database"
    if full_ref is not None:
        assert torch.equal(d_out[:cells], full_ref[:cells]), "all-reduced matrix of the sharded run != unsharded matrix"
        full_ref = None
    if dist is not None and scaling == "weak":
        # ranks own disjoint diagonal blocks: this rank's block must hold exactly its own partial result
        mine = torch.zeros_like(d_out)
        ctx.all2all_dense_rows_device(r0, r1, mine.data_ptr())
        lo, hi = kdbx.tri_cells(rank * N0), kdbx.tri_cells((rank + 1) * N0)
        assert torch.equal(mine[lo:hi], d_out[lo:hi]) and int(mine.to(torch.int64).sum().item()) == int(mine[lo:hi].to(torch.int64).sum().item())
        del mine
    if sampler:
        sampler.wait_first_sample()
    barrier()
    if sampler:
        sampler.begin()
    wall0 = time.perf_counter()
    dev_ms = scat_ms = 0.0
    per_step_ms = []
    launches = scat_launches = 0
    stage = {"prepare": 0.0, "expand": 0.0, "bucket": 0.0, "scatter": 0.0}
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    for i in range(a.steps):
        ev[i][0].record()
        st = step()
        ev[i][1].record()
        assert st.updates == U_rank
        dev_ms += st.ms_total
        per_step_ms.append(round(st.ms_total, 3))
        scat_ms += st.ms_scatter
        launches += st.kernel_launches
        scat_launches += st.scatter_launches
        for k in stage:
            stage[k] += getattr(st, "ms_" + k)
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    clocks = sampler.stop() if sampler else None
    if dist is not None:
        # with a collective in the step, time the whole step: CUDA events on torch's stream bracket the
        # (synchronous) library call and the NCCL all-reduce that follows it
        dev_ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    T_ms = max_over_ranks(dev_ms)
    value = U_total * a.steps / (T_ms / 1e3)
    checksum = int(d_out[:cells].to(torch.int64).sum().item()) if cells else 0

    # ---- end-to-end leg: host trie -> H2D -> compute (-> all-reduce) -> D2H host matrix ----------
    e2e = None
    if not a.no_e2e:
        out_host = kdbx.pinned_empty(max(1, cells), np.uint32)
        out_t = torch.from_numpy(out_host.view(np.int32))

        def e2e_step():
            ctx.load_patterns(trie)
            if dist is not None:
                st2 = step()
                out_t.copy_(d_out, non_blocking=False)
            else:
                _, st2 = ctx.all2all_dense_rows(r0, r1, out_host[:cells])
            return st2
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            st2 = e2e_step()
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        barrier()
        assert int(out_host[:cells].astype(np.int64).sum()) == checksum, "e2e result differs from the device-resident result"
        h2d = total_updates(P * 40 + int(tot.payload_bytes))  # summed over the ranks (every rank copies its own shard)
        e2e = {"value": U_total * a.steps / e2e_s, "unit": "updates/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": cells * 4 * world, "ms_per_step": 1e3 * e2e_s / a.steps,
               "ms_upload": st2.ms_upload, "ms_download": st2.ms_download}

    if rank != 0:
        return
    peak, peak_src = measured_hbm_peak()
    achieved = ALGO_BYTES_PER_UPDATE * U_rank * a.steps / (scat_ms / 1e3) / 1e9 if scat_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "k_scatter_add", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic_per_launch(), "peak_source": peak_src,
                "algorithmic_bytes_per_update": ALGO_BYTES_PER_UPDATE,
                "updates_per_launch": U_rank / max(1, scat_launches / a.steps),
                "avg_launch_ms": scat_ms / max(1, scat_launches), "launches_per_step": scat_launches // max(1, a.steps),
                "kernel_share_of_step": scat_ms / dev_ms if dev_ms else None}

    cpu_baseline = None
    if world == 1 and not a.no_cpu_baseline:
        try:
            sub = trie.prefix(sample_n)
            sub_path = cache_path(a, sample_n, 1)
            if not sub_path.exists():
                sub.write_db(sub_path)
            U_s = int(sub.totals().updates)
            secs = run_reference_binary(sub_path, cores)
            cpu_baseline = {"value": U_s / secs, "unit": "updates/s", "cores": cores, "kind": "reference",
                            "sample": f"first cluster only: {sample_n} genomes x {a.genome_kmers / 1e6:g} Mbp, "
                                      f"U={U_s:.4g}, {secs:.2f} s, kmer-db 2.3.1 all2all -t {cores}"}
        except Exception as e:  # the baseline is a reported number, not a dependency of the GPU path
            cpu_baseline = {"value": None, "unit": "updates/s", "cores": cores, "kind": "reference", "sample": f"failed: {e}"}

    print(json.dumps({
        "metric": "k-mer-pair updates/sec on all2all", "value": value, "unit": "updates/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": T_ms / a.steps, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(a) + ("" if world == 1 else
                               f"; x{world} shards laid side by side = {N} genomes ({a.clusters * world} clusters)" if scaling == "weak"
                               else f"; cut into {world} sub-tries"),
                   "num_samples": N, "num_patterns": P if world == 1 else None, "patterns_on_rank0": P, "updates_per_step": U_total,
                   "sum_n": int(tot.sum_n), "sum_l": int(tot.sum_l),
                   "parallelism": ("1 GPU, no collective" if world == 1 else
                                   f"trie sharded into {world} sub-tries (preorder pieces + ancestor chains), one per GPU, "
                                   f"+ one NCCL all-reduce of the {cells * 4 / 1e6:.0f} MB matrix per step"),
                   "shard_seconds_host": t_shard,
                   "l2_policy": "inputs (trie %.1f GB + per-chunk lists) exceed the 126 MB L2; no explicit flush" %
                                ((P * 40 + int(tot.payload_bytes)) / 1e9),
                   "chunk_ids": chunk_ids, "tile_cols": a.tile_cols, "unit_updates": a.unit_updates,
                   "tile_rows": a.tile_rows, "scatter_threads": a.scatter_threads},
        "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks,
        "stage_ms_per_step": {k: v / a.steps for k, v in stage.items()}, "wall_ms_per_step": wall_ms / a.steps,
        "library_ms_per_step": per_step_ms,
        "result_checksum": checksum, "list_form": ["ids", "run boundaries"][int(st.list_form)],
        "physical_updates_per_step": int(st.physical_updates),
    }))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
