#!/usr/bin/env bash
# One gpurun call: GPU parity tests, microbenchmarks, a small and (optionally) the full bench.
# Everything is logged under gpurun_out/<tag>/.
set -u
TAG="${1:-run}"; shift || true
OUT="gpurun_out/$TAG"; mkdir -p "$OUT"
{ nvidia-smi; nproc; free -g; } > "$OUT/box.txt" 2>&1
for step in "$@"; do
  case "$step" in
    tests)  timeout 1500 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/summary.txt";;
    tests_fast) timeout 900 python -m pytest tests -m gpu -x -q -k "not full_size" > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/summary.txt";;
    micro)  timeout 300 kmer-db_b200/bin/microbench > "$OUT/microbench.txt" 2>&1; echo "micro rc=$?" | tee -a "$OUT/summary.txt";;
    smoke)  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/summary.txt";;
    bench_small) timeout 600 python bench.py --samples 400 --genome-kmers 1000000 --steps 3 --warmup 2 > "$OUT/bench_small.json" 2> "$OUT/bench_small.err"; echo "bench_small rc=$?" | tee -a "$OUT/summary.txt";;
    bench)  timeout 1500 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?" | tee -a "$OUT/summary.txt";;
    bench_ref) timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"; echo "bench_ref rc=$?" | tee -a "$OUT/summary.txt";;
    ncu_list) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches.csv" python bench.py --samples 400 --genome-kmers 1000000 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > "$OUT/ncu_list.out" 2>&1; echo "ncu_list rc=$?" | tee -a "$OUT/summary.txt";;
    ncu_full) timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_scatter_add -s 2 -c 2 -f -o "$OUT/scatter" python bench.py --samples 400 --genome-kmers 1000000 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > "$OUT/ncu_full.out" 2>&1; echo "ncu_full rc=$?" | tee -a "$OUT/summary.txt";;
    ncu_list_cfg2) timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file "$OUT/launches_cfg2.csv" python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > "$OUT/ncu_list_cfg2.out" 2>&1; echo "ncu_list_cfg2 rc=$?" | tee -a "$OUT/summary.txt";;
    ncu_full_cfg2) timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_scatter -s 1 -c 1 -f -o "$OUT/scatter_cfg2" python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > "$OUT/ncu_full_cfg2.out" 2>&1; echo "ncu_full_cfg2 rc=$?" | tee -a "$OUT/summary.txt";;
    bench_chunks) for c in 16777216 33554432; do timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --chunk-ids $c > "$OUT/bench_chunk_$c.json" 2> "$OUT/bench_chunk_$c.err"; echo "bench_chunk $c rc=$?" | tee -a "$OUT/summary.txt"; done;;
    bench_tiles) for cfg in "32 1024 1024" "16 1024 512" "16 1024 1024" "8 1024 256" "32 1024 512"; do set -- $cfg; timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --tile-rows $1 --tile-cols $2 --scatter-threads $3 > "$OUT/bench_tile_$1_$2_$3.json" 2> "$OUT/bench_tile_$1_$2_$3.err"; echo "bench_tile $cfg rc=$?" | tee -a "$OUT/summary.txt"; done;;
    ncu_stages) timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:k_job_fill|k_scatter|k_decode_locals|k_key_totals" -s 0 -c 8 -f -o "$OUT/stages_cfg2" python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > "$OUT/ncu_stages.out" 2>&1; echo "ncu_stages rc=$?" | tee -a "$OUT/summary.txt";;
    scale) for n in ${SCALE_NS:-2}; do for mode in ${SCALE_MODES:-weak strong}; do timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 --scaling $mode > "$OUT/scale_${mode}_$n.json" 2> "$OUT/scale_${mode}_$n.err"; echo "scale $mode $n rc=$?" | tee -a "$OUT/summary.txt"; done; done;;
    tests_build) timeout 900 python -m pytest tests/test_gpu_build.py -m gpu -x -q > "$OUT/pytest_build.log" 2>&1; echo "pytest_build rc=$?" | tee -a "$OUT/summary.txt";;
    bench_shard) for sh in ${SHARDS:-1/2 7/8}; do tag=$(echo $sh | tr / _); timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --emulate-shard $sh > "$OUT/bench_shard_$tag.json" 2> "$OUT/bench_shard_$tag.err"; echo "bench_shard $sh rc=$?" | tee -a "$OUT/summary.txt"; done;;
    ncu_stages2) timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:k_job_hist|k_job_fill|k_decode_locals" -s 3 -c 3 -f -o "$OUT/stages2_cfg2" python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > "$OUT/ncu_stages2.out" 2>&1; echo "ncu_stages2 rc=$?" | tee -a "$OUT/summary.txt";;
    bench2) timeout 900 python bench.py --no-cpu-baseline > "$OUT/bench2.json" 2> "$OUT/bench2.err"; echo "bench2 rc=$?" | tee -a "$OUT/summary.txt";;
    cli_trace) D=/tmp/kdbx_modes; ( time kmer-db_b200/bin/kmer-db-b200 new2all $D/n2a.ours.db $D/q.list $D/t.csv ) > "$OUT/cli_new2all.txt" 2>&1; ( time kmer-db_b200/bin/kmer-db-b200 all2all $D/n2a.ours.db $D/t2.csv ) > "$OUT/cli_all2all.txt" 2>&1; ( time kmer-db_b200/bin/kmer-db-b200 build $D/db.list $D/t.db ) > "$OUT/cli_build.txt" 2>&1; ( time oracle/_ref/kmer-db new2all -t 16 $D/n2a.ref.db $D/q.list $D/t3.csv ) > "$OUT/ref_new2all.txt" 2>&1; ( time python -c "import ctypes,time; t=time.time(); l=ctypes.CDLL('libcudart.so.12'); l.cudaFree(0); print('cuda init', time.time()-t)" ) > "$OUT/cuda_init.txt" 2>&1; echo "cli_trace rc=$?" | tee -a "$OUT/summary.txt";;
    bench_2000) timeout 1200 python bench.py --samples 2000 --clusters 8 --steps 3 --warmup 3 --no-cpu-baseline > "$OUT/bench_2000.json" 2> "$OUT/bench_2000.err"; echo "bench_2000 rc=$?" | tee -a "$OUT/summary.txt";;
    test_multi) timeout 600 python -m pytest tests -m gpu -x -q -k "multi_gpu or sharding" > "$OUT/pytest_multi.log" 2>&1; echo "pytest_multi rc=$?" | tee -a "$OUT/summary.txt";;
    modes) timeout 1500 python tools/bench_modes.py --out-dir /tmp/kdbx_modes > "$OUT/modes.jsonl" 2> "$OUT/modes.err"; echo "modes rc=$?" | tee -a "$OUT/summary.txt";;
    bench_ids) timeout 900 python bench.py --no-cpu-baseline --no-e2e --list-form ids --steps 5 --warmup 3 > "$OUT/bench_ids.json" 2> "$OUT/bench_ids.err"; echo "bench_ids rc=$?" | tee -a "$OUT/summary.txt";;
    sanitize) timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_tries or boundary_lists or sample_window or edge_cases" > "$OUT/sanitizer_memcheck.log" 2>&1; echo "memcheck rc=$?" | tee -a "$OUT/summary.txt"
              timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_tries or boundary_lists" > "$OUT/sanitizer_racecheck.log" 2>&1; echo "racecheck rc=$?" | tee -a "$OUT/summary.txt";;
    scale_new) for n in ${SCALE_NS:-2}; do for mode in ${SCALE_MODES:-weak strong}; do timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps ${SCALE_STEPS:-5} --warmup 3 --scaling $mode > "$OUT/scale_${mode}_$n.json" 2> "$OUT/scale_${mode}_$n.err"; echo "scale $mode $n rc=$?" | tee -a "$OUT/summary.txt"; done; done;;
    cfg4) timeout 2400 python tools/run_cfg4.py --gpus ${CFG4_GPUS:-1} > "$OUT/cfg4_g${CFG4_GPUS:-1}.json" 2> "$OUT/cfg4.err"; echo "cfg4 rc=$?" | tee -a "$OUT/summary.txt";;
    cfg5) timeout 2400 python tools/bench_modes.py --out-dir /tmp/kdbx_cfg5 --skip sp --db-genomes ${CFG5_DB:-2000} --db-clusters 40 --queries ${CFG5_Q:-1000} --len 1000000 > "$OUT/cfg5.jsonl" 2> "$OUT/cfg5.err"; echo "cfg5 rc=$?" | tee -a "$OUT/summary.txt";;
    bench_nochunk) timeout 900 python bench.py --no-cpu-baseline --upload-chunk-mb 100000 > "$OUT/bench_nochunk.json" 2> "$OUT/bench_nochunk.err"; echo "bench_nochunk rc=$?" | tee -a "$OUT/summary.txt";;
    parts) timeout 1200 python tools/bench_modes.py --out-dir /tmp/kdbx_parts --skip sp,n2a --parts ${PARTS:-4} --parts-genomes ${PARTS_GENOMES:-400} --parts-len ${PARTS_LEN:-500000} > "$OUT/parts.jsonl" 2> "$OUT/parts.err"; echo "parts rc=$?" | tee -a "$OUT/summary.txt";;
    sanitize_new) timeout 900 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "db2db or asynchronous_chunked or cli_all2all_parts" > "$OUT/sanitizer_memcheck_new.log" 2>&1; echo "memcheck_new rc=$?" | tee -a "$OUT/summary.txt";;
    ncu_final) timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:k_scatter_diff|k_decode_locals|k_job_fill_runs" -s 3 -c 3 -f -o "$OUT/final_cfg2" python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > "$OUT/ncu_final.out" 2>&1; echo "ncu_final rc=$?" | tee -a "$OUT/summary.txt";;
    sanitize_new2) timeout 900 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "db2db_cells or asynchronous_chunked" > "$OUT/sanitizer_memcheck_new.log" 2>&1; echo "memcheck_new rc=$?" | tee -a "$OUT/summary.txt";;
    tests_new) timeout 900 python -m pytest tests -m gpu -x -q -k "parts or db2db or asynchronous or sliding or golden" > "$OUT/pytest_new.log" 2>&1; echo "pytest_new rc=$?" | tee -a "$OUT/summary.txt";;
    tests_cli_new) timeout 900 python -m pytest tests/test_gpu_widen.py -m gpu -x -q -k "sample_rows or one2all or minhash" > "$OUT/pytest_cli_new.log" 2>&1; echo "pytest_cli_new rc=$?" | tee -a "$OUT/summary.txt";;
    tests_2gpu) timeout 900 python -m pytest tests -m gpu -x -q -k "two_gpus or multi_gpu" > "$OUT/pytest_2gpu.log" 2>&1; echo "pytest_2gpu rc=$?" | tee -a "$OUT/summary.txt";;
    *) echo "unknown step $step";;
  esac
done
tail -n 30 "$OUT"/*.log "$OUT"/*.txt "$OUT"/*.json "$OUT"/*.err 2>/dev/null | tail -n 120
