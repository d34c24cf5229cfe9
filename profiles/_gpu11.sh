python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err
tail -3 gpurun_out/r2_bench_8gpu.err | cut -c1-250; wc -c gpurun_out/r2_bench_8gpu.json
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "boundary or two_devices" 2>&1 | tail -2
