set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15; echo "pytest exit ${PIPESTATUS[0]}" ) > gpurun_out/l_pytest.log 2>&1
timeout 300 python tools/e2e_times.py ecoli100x 4 > gpurun_out/l_e2e_sync.log 2>&1
timeout 300 python tools/e2e_times.py ecoli100x 4 overlap > gpurun_out/l_e2e_overlap.log 2>&1
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err
tail -3 gpurun_out/l_pytest.log; tail -3 gpurun_out/l_e2e_sync.log; tail -3 gpurun_out/l_e2e_overlap.log
