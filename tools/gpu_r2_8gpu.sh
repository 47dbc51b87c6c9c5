#!/bin/bash
# Round-2 run on 8 GPUs of one box: host-link ceiling with every GPU copying at once, the bench line at N = 8
# (tiles + sweep + e2e), the device-set tests on all GPUs.
set -u
mkdir -p gpurun_out
T=${1:-r02d}
N=${2:-8}
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1; lscpu | grep -E 'NUMA|Socket|Model name|^CPU\(s\)' >> gpurun_out/${T}_topo.txt
free -g | head -2 >> gpurun_out/${T}_topo.txt
echo "== host link ceiling"; timeout 400 python tools/gpu_pcie_all.py $N 2 2>&1 | tail -1 | tee gpurun_out/${T}_pcie_all.json
echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err; tail -c 5000 gpurun_out/${T}_bench_n$N.json; tail -3 gpurun_out/${T}_bench_n$N.err
echo "== device set tests"; timeout 600 python -m pytest tests/test_multi_device.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${T}_multi_pytest.txt
