set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15; echo "pytest exit ${PIPESTATUS[0]}" ) > gpurun_out/o_pytest.log 2>&1
timeout 300 python tools/stage_times.py ecoli100x 3 2>&1 | grep "^run 2" >> gpurun_out/o_ab.log
timeout 300 python tools/stage_times.py chr20_30x 2 2>&1 | grep "^run 1" >> gpurun_out/o_ab.log
tail -3 gpurun_out/o_pytest.log; cut -c1-330 gpurun_out/o_ab.log
