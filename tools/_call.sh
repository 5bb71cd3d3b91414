set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15; echo "pytest exit ${PIPESTATUS[0]}" ) > gpurun_out/r1g_pytest_gpu.log 2>&1
timeout 400 python bench.py > gpurun_out/r1g_bench_ecoli100x.json 2> gpurun_out/r1g_bench.err
timeout 300 python tools/stage_times.py chr20_30x 3 > gpurun_out/r1g_stage_times_chr20.log 2>&1
timeout 400 python bench.py --workload chr20_30x --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1g_bench_chr20.json 2> gpurun_out/r1g_bench_chr20.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r1g_smoke.log 2>&1
tail -3 gpurun_out/r1g_pytest_gpu.log; cut -c1-300 gpurun_out/r1g_bench_ecoli100x.json; tail -2 gpurun_out/r1g_smoke.log
