set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 600 python -m pytest tests/test_host_path.py tests/test_sharded.py -q -m gpu 2>&1 | tail -5
timeout 600 python tools/inprocess_scaling.py 8192 3 > gpurun_out/r2_inprocess_sharding_2gpu.jsonl 2> gpurun_out/inproc.err; cat gpurun_out/r2_inprocess_sharding_2gpu.jsonl; tail -3 gpurun_out/inproc.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/bench2.err; cat gpurun_out/r2_bench_2gpu.json | cut -c1-1500; tail -5 gpurun_out/bench2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --workload batch --codec bc7_amd --textures 64 > gpurun_out/r2_batch64_bc7amd_2gpu.json 2> gpurun_out/batch2.err; cat gpurun_out/r2_batch64_bc7amd_2gpu.json | cut -c1-1200; tail -5 gpurun_out/batch2.err
