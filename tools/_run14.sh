set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 2> gpurun_out/bench8.err | grep -v "^NCCL" > gpurun_out/r2_bench_8gpu.json; cut -c1-900 gpurun_out/r2_bench_8gpu.json; tail -3 gpurun_out/bench8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 3 --warmup 3 2> gpurun_out/bench4.err | grep -v "^NCCL" > gpurun_out/r2_bench_4gpu.json; cut -c1-300 gpurun_out/r2_bench_4gpu.json
timeout 600 python tools/inprocess_scaling.py 8192 3 > gpurun_out/r2_inprocess_sharding_8gpu.jsonl 2> gpurun_out/inproc8.err; cat gpurun_out/r2_inprocess_sharding_8gpu.jsonl; tail -3 gpurun_out/inproc8.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 1 --warmup 1 --workload batch --codec bc7_amd --textures 1024 2> gpurun_out/batch8.err | grep -v "^NCCL" > gpurun_out/r2_batch1024_bc7amd_8gpu.json; cut -c1-1300 gpurun_out/r2_batch1024_bc7amd_8gpu.json; tail -3 gpurun_out/batch8.err
