python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^$" | tail -8 | cut -c1-250 > gpurun_out/r2_t8_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 4 --warmup 1 --jobs-per-step 4 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r2_bench_2gpu_ref.json 2> gpurun_out/r2_bench_2gpu_ref.err
tail -3 gpurun_out/r2_t8_pytest.log; tail -4 gpurun_out/r2_bench_2gpu.err | cut -c1-250; wc -c gpurun_out/r2_bench_2gpu.json gpurun_out/r2_bench_2gpu_ref.json
