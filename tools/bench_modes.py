#!/usr/bin/env python
"""Side benchmarks of the widened rows (all2all-sp, new2all, build) next to the unmodified reference
binary on the same files, with byte comparison of the CSVs.  Not the headline metric (bench.py is);
prints one JSON line per mode.  Inputs are synthetic and seeded:

  all2all-sp   pattern-level database (kmer-db_b200/host/synth.cpp): many small clusters -> sparse matrix
  build        FASTA-level genomes written here (clusters of mutated copies), our host `build` vs the reference's
  new2all      queries = further mutated copies; GPU probe + scatter vs the reference's one2all per query
  all2all-parts (--parts P) the genomes dealt round-robin to P partial databases (every part holds members of every cluster, so
               the cells off the diagonal are as full as the diagonal ones); device db2db + all2all-sp per cell vs the reference

    python tools/bench_modes.py [--out-dir DIR] [--sp-samples N --sp-clusters C --sp-len L] [--db-genomes G --queries Q --len L]
"""
import argparse
import json
import os
import re
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "kmer-db_b200" / "bin" / "kmer-db-b200"
REF = ROOT / "oracle" / "_ref" / "kmer-db"


def run(cmd, **kw):
    t0 = time.perf_counter()
    r = subprocess.run([str(c) for c in cmd], capture_output=True, text=True, **kw)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError(f"{' '.join(map(str, cmd))} failed: {r.stderr[-800:]}")
    return r.stdout + r.stderr, dt


def stats_json(text):
    """the JSON stats line kmer-db-b200 prints on stderr"""
    for line in reversed(text.splitlines()):
        if line.startswith("{\"updates\""):
            return json.loads(line)
    return {}


def phase_seconds(text, after):
    m = re.search(re.escape(after) + r".*?OK \(([0-9.eE+-]+) seconds\)", text, re.S)
    return float(m.group(1)) if m else None


def same(a, b):
    return subprocess.run(["cmp", "-s", str(a), str(b)]).returncode == 0


def write_fasta(path, seq, name):
    with open(path, "w") as f:
        f.write(f">{name}\n")
        s = seq.tobytes().decode()
        for i in range(0, len(s), 100):
            f.write(s[i:i + 100])
            f.write("\n")


def make_genomes(d, genomes, clusters, length, mu, seed, prefix):
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    d.mkdir(parents=True, exist_ok=True)
    members = [[] for _ in range(clusters)]
    names = []
    for g in range(genomes):
        c = g * clusters // genomes
        if not members[c]:
            seq = acgt[rng.integers(0, 4, size=length)]
        else:
            seq = members[c][int(rng.integers(0, len(members[c])))].copy()
            pos = rng.random(length) < mu
            seq[pos] = acgt[rng.integers(0, 4, size=int(pos.sum()))]
        members[c].append(seq)
        name = f"{prefix}{g:05d}"
        write_fasta(d / f"{name}.fasta", seq, name)
        names.append(str(d / name))
    return names, members


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out-dir", default="/tmp/kdbx_modes")
    ap.add_argument("--sp-samples", type=int, default=20000)
    ap.add_argument("--sp-clusters", type=int, default=400)
    ap.add_argument("--sp-len", type=int, default=200000)
    ap.add_argument("--db-genomes", type=int, default=400)
    ap.add_argument("--db-clusters", type=int, default=8)
    ap.add_argument("--queries", type=int, default=100)
    ap.add_argument("--len", type=int, default=1000000)
    ap.add_argument("--parts", type=int, default=0, help="all2all-parts over this many partial databases (0 = skip the mode)")
    ap.add_argument("--parts-genomes", type=int, default=400)
    ap.add_argument("--parts-clusters", type=int, default=8)
    ap.add_argument("--parts-len", type=int, default=500000)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--skip", default="")
    a = ap.parse_args()
    out = Path(a.out_dir)
    out.mkdir(parents=True, exist_ok=True)
    have_ref = REF.exists()

    if "sp" not in a.skip:
        db = out / "sp.db"
        text, _ = run([EXE, "synth", "-n", a.sp_samples, "-clusters", a.sp_clusters, "-len", a.sp_len, "-seed", 4, db])
        U = int(re.search(r"U=(\d+)", text).group(1))
        runs = []
        for _ in range(2):   # one-shot CLI runs include device allocations; the slower one is usually the first
            text, wall = run([EXE, "all2all-sp", db, out / "sp.ours.csv"])
            runs.append((stats_json(text), wall))
        st, wall = min(runs, key=lambda r: r[0].get("seconds", 1e9))
        line = {"mode": "all2all-sp", "workload": f"{a.sp_samples} samples, {a.sp_clusters} clusters, {a.sp_len} k-mers each (pattern-level synthetic)",
                "updates": U, "ours_seconds": st.get("seconds"), "ours_updates_per_s": U / st["seconds"] if st.get("seconds") else None,
                "ours_stage_ms": {k: st.get(k) for k in ("ms_upload", "ms_prepare", "ms_expand", "ms_bucket", "ms_scatter", "ms_compact", "ms_download")},
                "ours_wall_incl_io": wall, "ours_seconds_both_runs": [r[0].get("seconds") for r in runs]}
        if have_ref:
            text, rwall = run([REF, "all2all-sp", "-t", a.threads, db, out / "sp.ref.csv"])
            secs = phase_seconds(text, "Calculating matrix of common k-mers...")
            line.update({"reference_seconds": secs, "reference_threads": a.threads, "reference_wall_incl_io": rwall,
                         "csv_identical": same(out / "sp.ours.csv", out / "sp.ref.csv"),
                         "speedup_compute": secs / st["seconds"] if secs and st.get("seconds") else None})
        print(json.dumps(line), flush=True)

    if "n2a" not in a.skip:
        names, members = make_genomes(out / "fa", a.db_genomes, a.db_clusters, a.len, 0.005, 11, "g")
        (out / "db.list").write_text("\n".join(names) + "\n")
        # queries: mutated copies of random database genomes
        rng = np.random.default_rng(12)
        acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
        qnames = []
        for q in range(a.queries):
            c = int(rng.integers(0, a.db_clusters))
            seq = members[c][int(rng.integers(0, len(members[c])))].copy()
            pos = rng.random(a.len) < 0.01
            seq[pos] = acgt[rng.integers(0, 4, size=int(pos.sum()))]
            write_fasta(out / "fa" / f"q{q:05d}.fasta", seq, f"q{q:05d}")
            qnames.append(str(out / "fa" / f"q{q:05d}"))
        (out / "q.list").write_text("\n".join(qnames) + "\n")
        def update_time(text):
            m = re.search(r"Database update time:\s*([0-9.eE+-]+)", text)
            return float(m.group(1)) if m else None
        text, bwall = run([EXE, "build", "-t", a.threads, out / "db.list", out / "n2a.ours.db"])   # on the device
        dev = {}
        for ln in text.splitlines():
            if ln.startswith("{\"samples\""):
                dev = json.loads(ln)
        line = {"mode": "build", "workload": f"{a.db_genomes} genomes x {a.len} bp, {a.db_clusters} clusters, k=18 (FASTA-level synthetic)",
                "ours_device_wall_seconds": bwall, "ours_device_update_seconds": update_time(text), "ours_device_stats": dev,
                "host_threads": a.threads}
        text, hwall = run([EXE, "build", "-host-build", "-t", a.threads, out / "db.list", out / "n2a.host.db"])
        line.update({"ours_host_builder_wall_seconds": hwall, "ours_host_builder_update_seconds": update_time(text)})
        run([EXE, "all2all", out / "n2a.ours.db", out / "b.dev.csv"])
        run([EXE, "all2all", out / "n2a.host.db", out / "b.host.csv"])
        line["device_and_host_builder_give_identical_all2all"] = same(out / "b.dev.csv", out / "b.host.csv")
        if have_ref:
            text, rbwall = run([REF, "build", "-t", a.threads, out / "db.list", out / "n2a.ref.db"])
            line.update({"reference_wall_seconds": rbwall, "reference_update_seconds": update_time(text)})
            run([EXE, "all2all", out / "n2a.ref.db", out / "b.ref.csv"])
            line["all2all_identical_to_reference_built_db"] = same(out / "b.dev.csv", out / "b.ref.csv")
        print(json.dumps(line), flush=True)
        runs = []
        for _ in range(2):
            text, wall = run([EXE, "new2all", "-t", a.threads, out / "n2a.ours.db", out / "q.list", out / "n2a.ours.csv"])
            runs.append((stats_json(text), wall))
        st, wall = min(runs, key=lambda r: r[0].get("seconds", 1e9))
        line = {"mode": "new2all", "workload": f"{a.queries} queries x {a.len} bp vs {a.db_genomes}-genome database",
                "probes": st.get("probes"), "hits": st.get("hits"), "ours_wall_seconds_incl_db_load_and_fasta": wall,
                "ours_processing_seconds": st.get("seconds"), "ours_processing_seconds_both_runs": [r[0].get("seconds") for r in runs],
                "ours_device_ms": {k: st.get(k) for k in ("ms_prepare", "ms_expand", "ms_probe", "ms_scatter", "ms_total", "ms_download")},
                "ours_probes_per_s_device": st["probes"] / (st["ms_probe"] / 1e3) if st.get("ms_probe") else None}
        if have_ref:
            text, rwall = run([REF, "new2all", "-t", a.threads, out / "n2a.ref.db", out / "q.list", out / "n2a.ref.csv"])
            m = re.search(r"Total: ([0-9.eE+-]+)", text)
            line.update({"reference_wall_seconds_incl_db_load_and_fasta": rwall, "reference_processing_seconds": float(m.group(1)) if m else None,
                         "reference_threads": a.threads, "csv_identical": same(out / "n2a.ours.csv", out / "n2a.ref.csv")})
            # our database must serve the reference too, and vice versa
            run([EXE, "new2all", out / "n2a.ref.db", out / "q.list", out / "n2a.cross.csv"])
            line["csv_identical_on_reference_built_db"] = same(out / "n2a.cross.csv", out / "n2a.ref.csv")
        print(json.dumps(line), flush=True)

    if a.parts > 0:
        names, _ = make_genomes(out / "fa_parts", a.parts_genomes, a.parts_clusters, a.parts_len, 0.005, 21, "p")
        ours_list, ref_list = [], []
        t_build_ours = t_build_ref = 0.0
        for j in range(a.parts):
            lst = out / f"part{j}.list"
            lst.write_text("\n".join(names[j::a.parts]) + "\n")
            _, dt = run([EXE, "build", "-t", a.threads, lst, out / f"part{j}.ours.db"])
            t_build_ours += dt
            ours_list.append(str(out / f"part{j}.ours.db"))
            if have_ref:
                _, dt = run([REF, "build", "-t", a.threads, lst, out / f"part{j}.ref.db"])
                t_build_ref += dt
                ref_list.append(str(out / f"part{j}.ref.db"))
        (out / "parts.ours.list").write_text("\n".join(ours_list) + "\n")
        runs = []
        for _ in range(2):
            text, wall = run([EXE, "all2all-parts", out / "parts.ours.list", out / "parts.ours.csv"])
            runs.append((stats_json(text), wall))
        st, wall = min(runs, key=lambda r: r[1])
        saved = re.search(r"No\. saved pairs: (\d+)", text)
        line = {"mode": "all2all-parts",
                "workload": f"{a.parts_genomes} genomes x {a.parts_len} bp, {a.parts_clusters} clusters, dealt round-robin to {a.parts} partial databases "
                            f"(k=18, FASTA-level synthetic, built on the device)",
                "cell_updates": st.get("updates"), "kmers_probed": st.get("probes"), "kmers_in_both": st.get("hits"),
                "saved_pairs": int(saved.group(1)) if saved else None,
                "ours_wall_seconds_incl_db_load": wall, "ours_wall_both_runs": [r[1] for r in runs],
                "ours_device_ms": {k: st.get(k) for k in ("ms_prepare", "ms_probe", "ms_scatter", "ms_compact", "ms_total", "ms_download")},
                "ours_build_wall_seconds_all_parts": t_build_ours}
        if have_ref:
            (out / "parts.ref.list").write_text("\n".join(ref_list) + "\n")
            text, rwall = run([REF, "all2all-parts", "-t", a.threads, out / "parts.ref.list", out / "parts.ref.csv"])
            m = re.search(r"All2All time\s*:\s*([0-9.eE+-]+)", text)
            ml = re.search(r"Load time\s*:\s*([0-9.eE+-]+)", text)
            line.update({"reference_wall_seconds_incl_db_load": rwall, "reference_all2all_seconds": float(m.group(1)) if m else None,
                         "reference_load_seconds": float(ml.group(1)) if ml else None, "reference_threads": a.threads,
                         "reference_build_wall_seconds_all_parts": t_build_ref,
                         "csv_identical": same(out / "parts.ours.csv", out / "parts.ref.csv")})
            # the grid over the databases the REFERENCE built must give the same table
            run([EXE, "all2all-parts", out / "parts.ref.list", out / "parts.cross.csv"])
            line["csv_identical_on_reference_built_parts"] = same(out / "parts.cross.csv", out / "parts.ref.csv")
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
