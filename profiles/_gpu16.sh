python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^$" | tail -8 | cut -c1-250 > gpurun_out/r2_t16_pytest.log
python -m pytest tests/test_gpu_nodes.py -m gpu -q --tb=short 2>&1 | tail -2 >> gpurun_out/r2_t16_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300 > gpurun_out/r2_smoke.log
for R in 8 128; do python profiles/node_probe.py --requests $R --sampler heun 2>&1 | tail -2 | head -1 | cut -c1-200 >> gpurun_out/r2_heun_graph2.log; done
tail -4 gpurun_out/r2_t16_pytest.log; cat gpurun_out/r2_smoke.log; cat gpurun_out/r2_heun_graph2.log
