set -x
nvidia-smi -L
timeout 60 profiles/micro/tc_alias_probe > gpurun_out/tc_alias_probe.log 2>&1; cat gpurun_out/tc_alias_probe.log
timeout 900 python -m pytest tests/test_nccl_gpu.py -q -m gpu --timeout 600 > gpurun_out/test_nccl.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/test_nccl.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
