#!/bin/bash
# round-2 GPU call 9 (8 GPUs): bench.py --gpus 8 with its multi-GPU sections, the 36-qubit sharded run, shard="qubits" through the plugin
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $O/c9_smi.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29561 bench.py --gpus 8 --steps 3 --warmup 3 > $O/c9_bench_8gpu.json 2> $O/c9_bench_8gpu.err; echo "bench8 rc=$?"; tail -c 2500 $O/c9_bench_8gpu.json; tail -2 $O/c9_bench_8gpu.err
timeout 600 $TR --master-port 29562 tools/sharded_plugin_check.py --qubits 24 > $O/c9_plugin_shard.json 2> $O/c9_plugin_shard.err; echo "plugin shard rc=$?"; cat $O/c9_plugin_shard.json
timeout 600 $TR --master-port 29563 tools/sharded_check.py --qubits 30 --peer --grad --verify --reps 1 > $O/c9_sharded_n30.json 2> $O/c9_sharded_n30.err; echo "n30 rc=$?"; tail -c 700 $O/c9_sharded_n30.json
timeout 900 $TR --master-port 29564 tools/sharded_check.py --qubits 36 --peer --reps 1 > $O/c9_sharded_n36.json 2> $O/c9_sharded_n36.err; echo "n36 rc=$?"; tail -c 1200 $O/c9_sharded_n36.json; tail -3 $O/c9_sharded_n36.err
