set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15; echo "pytest exit ${PIPESTATUS[0]}" ) > gpurun_out/r1f_pytest_gpu.log 2>&1
timeout 400 python bench.py > gpurun_out/r1f_bench_ecoli100x.json 2> gpurun_out/r1f_bench.err
timeout 300 python tools/stage_times.py chr20_30x 3 > gpurun_out/r1f_stage_times_chr20.log 2>&1
timeout 300 python tools/e2e_times.py ecoli100x 4 > gpurun_out/r1f_e2e_times.log 2>&1
timeout 600 python tools/sweep.py 3.3 6.5 13 26 > gpurun_out/r1f_sweep_1gpu.jsonl 2> gpurun_out/r1f_sweep.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1f.csv python tools/profile_step.py ecoli100x 2 > gpurun_out/r1f_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'kmer_upsert|kmer_partition|probe_kernel|correct_kernel|onesweep|tables_kernel|walk_phase1|dedup_flag|tie_small|table_sweep' -c 14 -f -o gpurun_out/r1f_full python tools/profile_step.py ecoli100x 1 > gpurun_out/r1f_ncu_full.log 2>&1
tail -3 gpurun_out/r1f_pytest_gpu.log; cut -c1-400 gpurun_out/r1f_bench_ecoli100x.json; cat gpurun_out/r1f_sweep_1gpu.jsonl | cut -c1-300
