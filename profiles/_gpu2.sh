python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_t2_pytest.log
for R in 128 8 1; do python profiles/node_probe.py --requests $R 2>&1 | tail -12 > gpurun_out/r2_node_probe_R$R.log; done
python profiles/node_probe.py --requests 128 --rng philox 2>&1 | tail -12 > gpurun_out/r2_node_probe_R128_philox.log
python profiles/node_probe.py --requests 128 --sampler heun 2>&1 | tail -12 > gpurun_out/r2_node_probe_R128_heun.log
python profiles/burst_probe.py --requests 128 --rng torch 2>&1 | tail -3 > gpurun_out/r2_burst_torch_tma.log
python profiles/burst_probe.py --requests 128 --rng torch --tma 0 2>&1 | tail -3 > gpurun_out/r2_burst_torch_ldg.log
python profiles/burst_probe.py --requests 128 --rng philox 2>&1 | tail -3 > gpurun_out/r2_burst_philox.log
ncu --set full --clock-control none --import-source on -k regex:substep_torch_tma -s 60 -c 2 -o gpurun_out/r2_torch_tma_v1 python profiles/burst_probe.py --requests 128 --rng torch 2>&1 | tail -5 > gpurun_out/r2_ncu_torch_tma.log
tail -5 gpurun_out/r2_t2_pytest.log
