python -m pytest tests/test_gpu_nodes.py -m gpu -q --tb=short 2>&1 | grep -v "^$" | tail -30 | cut -c1-250 > gpurun_out/r2_t17_pytest.log
for R in 8 128; do python profiles/node_probe.py --requests $R --sampler heun 2>&1 | tail -3 | head -2 | cut -c1-200 >> gpurun_out/r2_heun_sg.log; done
tail -12 gpurun_out/r2_t17_pytest.log; cat gpurun_out/r2_heun_sg.log
