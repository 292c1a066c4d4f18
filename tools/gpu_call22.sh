#!/bin/bash
# round-2 GPU call 22 (2 GPUs): bench.py --gpus 2 of the final build (batch-sharded value + collective / sharded sections), peer tests
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 3 --warmup 3 > $O/c22_bench_2gpu.json 2> $O/c22_bench_2gpu.err; echo "bench2 rc=$?"; tail -c 1800 $O/c22_bench_2gpu.json; tail -3 $O/c22_bench_2gpu.err
timeout 900 python -m pytest tests/test_gpu_peer.py -q -m gpu > $O/c22_peer_tests.log 2>&1; echo "peer tests rc=$?"; tail -3 $O/c22_peer_tests.log
