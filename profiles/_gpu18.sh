for R in 128; do python profiles/node_probe.py --requests $R --sampler heun 2>&1 | tail -3 | head -2 | cut -c1-200 >> gpurun_out/r2_heun_sg.log; done
python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^$" | tail -4 | cut -c1-250 > gpurun_out/r2_t18_pytest.log
cat gpurun_out/r2_heun_sg.log; tail -2 gpurun_out/r2_t18_pytest.log
