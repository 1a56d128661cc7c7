#!/bin/bash
# Developer tool (GPU box, N GPUs): the driver's multi-GPU launch of bench.py + the N-link PCIe ceiling.  Usage: tools/r2_scale.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
lscpu | grep -E "NUMA|Socket|^CPU\(s\)|Model name" > gpurun_out/lscpu_n$N.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench exit $?"; tail -c 400 gpurun_out/bench_n$N.err
python tools/bench_summary.py gpurun_out/bench_n$N.json | head -3
timeout 600 python tools/pcie_probe.py --gpus $N > gpurun_out/pcie_n$N.log 2>&1
echo "pcie exit $?"; tail -15 gpurun_out/pcie_n$N.log
