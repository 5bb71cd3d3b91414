set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15; echo "pytest exit ${PIPESTATUS[0]}" ) > gpurun_out/h_pytest.log 2>&1
for v in 1 0; do
  BGX_UPSERT_CAS_FIRST=$v timeout 300 python tools/stage_times.py ecoli100x 3 2>&1 | grep "^run 2" | sed "s/^/casfirst$v /" >> gpurun_out/h_ab_ecoli.log
done
for f in 1.4 2.0 2.9; do
  BGX_SOLID_FACTOR=$f timeout 300 python tools/stage_times.py ecoli100x 3 2>&1 | grep "^run 2" | sed "s/^/solid$f /" >> gpurun_out/h_ab_ecoli.log
done
BGX_UPSERT_CAS_FIRST=1 timeout 300 python tools/stage_times.py chr20_30x 2 2>&1 | grep "^run 1" | sed "s/^/casfirst1 /" >> gpurun_out/h_ab_chr20.log
BGX_UPSERT_CAS_FIRST=0 BGX_SOLID_FACTOR=2.0 timeout 300 python tools/stage_times.py chr20_30x 2 2>&1 | grep "^run 1" | sed "s/^/casfirst0_solid2.0 /" >> gpurun_out/h_ab_chr20.log
tail -3 gpurun_out/h_pytest.log
