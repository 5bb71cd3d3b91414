set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q 2>&1 | tail -8; echo "pytest exit ${PIPESTATUS[0]}" ) > gpurun_out/p_pytest_dist2.log 2>&1
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 tests/dist_worker.py medium 2>&1 | tail -3 ) > gpurun_out/p_dist2_medium.log 2>&1
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 2 --steps 4 --warmup 3 ) > gpurun_out/p_bench2.json 2> gpurun_out/p_bench2.err
tail -3 gpurun_out/p_pytest_dist2.log; tail -1 gpurun_out/p_dist2_medium.log | cut -c1-200; cut -c1-200 gpurun_out/p_bench2.json
