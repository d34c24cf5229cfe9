python -m pytest tests/test_gpu_nodes.py -m gpu -q --tb=short 2>&1 | grep -v "^$" | tail -14 | cut -c1-250 > gpurun_out/r2_t15_pytest.log
python profiles/node_probe.py --requests 128 --sampler heun 2>&1 | tail -3 | cut -c1-200 > gpurun_out/r2_heun_graph.log
python profiles/node_probe.py --requests 128 --sampler heun --no-graph 2>&1 | tail -3 | cut -c1-200 > gpurun_out/r2_heun_eager.log
python profiles/node_probe.py --requests 8 --sampler heun 2>&1 | tail -3 | cut -c1-200 >> gpurun_out/r2_heun_graph.log
python profiles/node_probe.py --requests 8 --sampler heun --no-graph 2>&1 | tail -3 | cut -c1-200 >> gpurun_out/r2_heun_eager.log
tail -3 gpurun_out/r2_t15_pytest.log; echo graph; cat gpurun_out/r2_heun_graph.log; echo eager; cat gpurun_out/r2_heun_eager.log
