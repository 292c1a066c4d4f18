#!/bin/bash
# round-2 GPU call 23 (8 GPUs): bench.py --gpus 8 of the final build with its multi-GPU sections
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29561 bench.py --gpus 8 --steps 3 --warmup 3 > $O/c23_bench_8gpu.json 2> $O/c23_bench_8gpu.err; echo "bench8 rc=$?"; tail -c 1500 $O/c23_bench_8gpu.json; tail -2 $O/c23_bench_8gpu.err
timeout 600 $TR --master-port 29562 tools/sharded_plugin_check.py --qubits 24 > $O/c23_plugin_shard.json 2> $O/c23_plugin_shard.err; echo "plugin shard rc=$?"; cat $O/c23_plugin_shard.json
