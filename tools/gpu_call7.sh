#!/bin/bash
# round-2 GPU call 7 (2 GPUs): sharded register — adjoint + general observables over peer memory; multi-GPU bench sections
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/c7_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_peer.py -q -m gpu > $O/c7_peer_tests.log 2>&1; echo "peer tests rc=$?"; tail -4 $O/c7_peer_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/sharded_check.py --qubits 26 --peer --grad --verify --reps 1 > $O/c7_sharded_n26.json 2> $O/c7_sharded_n26.err; echo "sharded rc=$?"; tail -c 900 $O/c7_sharded_n26.json
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 3 --warmup 3 > $O/c7_bench_2gpu.json 2> $O/c7_bench_2gpu.err; echo "bench2 rc=$?"; tail -c 1500 $O/c7_bench_2gpu.json; tail -3 $O/c7_bench_2gpu.err
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "pauli or hea" > $O/c7_pauli_tests.log 2>&1; echo "pauli tests rc=$?"; tail -3 $O/c7_pauli_tests.log
