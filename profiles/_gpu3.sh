python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^$" | tail -60 | cut -c1-250 > gpurun_out/r2_t3_pytest.log
for v in 1 6 7 0; do echo "tma=$v" >> gpurun_out/r2_burst_torch_v2.log; python profiles/burst_probe.py --requests 128 --rng torch --tma $v 2>&1 | tail -2 | cut -c1-200 >> gpurun_out/r2_burst_torch_v2.log; done
for R in 128 8 1; do python profiles/node_probe.py --requests $R 2>&1 | tail -9 | cut -c1-220 > gpurun_out/r2_node_probe_R$R.log; done
python profiles/node_probe.py --requests 128 --rng philox 2>&1 | tail -9 | cut -c1-220 > gpurun_out/r2_node_probe_R128_philox.log
python profiles/node_probe.py --requests 128 --sampler heun 2>&1 | tail -9 | cut -c1-220 > gpurun_out/r2_node_probe_R128_heun.log
timeout 600 python bench.py --steps 4 --warmup 1 --jobs-per-step 4 --no-frame-shard > gpurun_out/r2_bench_try1.json 2> gpurun_out/r2_bench_try1.err; tail -5 gpurun_out/r2_bench_try1.err | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:substep_torch_tma -s 60 -c 1 -o gpurun_out/r2_torch_tma_v2 python profiles/burst_probe.py --requests 128 --rng torch 2>&1 | tail -2 | cut -c1-200 > gpurun_out/r2_ncu_torch_tma.log
tail -3 gpurun_out/r2_t3_pytest.log
