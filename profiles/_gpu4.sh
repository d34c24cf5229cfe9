python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^$" | tail -40 | cut -c1-250 > gpurun_out/r2_t4_pytest.log
for v in 1 0; do echo "tma=$v" >> gpurun_out/r2_burst_torch_v3.log; python profiles/burst_probe.py --requests 128 --rng torch --tma $v 2>&1 | tail -2 | cut -c1-200 >> gpurun_out/r2_burst_torch_v3.log; done
timeout 900 python bench.py --steps 4 --warmup 1 --jobs-per-step 4 > gpurun_out/r2_bench_try2.json 2> gpurun_out/r2_bench_try2.err; tail -5 gpurun_out/r2_bench_try2.err | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:substep_torch_tma -s 60 -c 1 -o gpurun_out/r2_torch_tma_v3 python profiles/burst_probe.py --requests 128 --rng torch 2>&1 | tail -2 | cut -c1-200 > gpurun_out/r2_ncu_torch_tma.log
tail -3 gpurun_out/r2_t4_pytest.log
