#!/bin/bash
# third session: the bench line at N GPUs (tiles + sweep + e2e) with the final kernels, and the device-set tests
set -u
mkdir -p gpurun_out
T=${1:-r03n}
N=${2:-8}
echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err; tail -c 3000 gpurun_out/${T}_bench_n$N.json; tail -3 gpurun_out/${T}_bench_n$N.err
echo "== device set tests"; timeout 600 python -m pytest tests/test_multi_device.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${T}_multi_pytest.txt
