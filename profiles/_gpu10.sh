python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^$" | tail -6 | cut -c1-250 > gpurun_out/r2_t10_pytest.log
for R in 128 256; do for b in 1 0; do echo "R=$R tma_boundary=$b" >> gpurun_out/r2_boundary_ab.log; LANPAINT_B200_TMA_BOUNDARY=$b python profiles/job_probe.py --requests $R 2>&1 | tail -1 >> gpurun_out/r2_boundary_ab.log; LANPAINT_B200_TMA_BOUNDARY=$b python profiles/job_probe.py --requests $R --net cond_uncond --rng torch 2>&1 | tail -1 >> gpurun_out/r2_boundary_ab.log; done; done
for R in 1 8 128; do python profiles/node_probe.py --requests $R 2>&1 | tail -4 | cut -c1-200 > gpurun_out/r2_node_probe_R$R.log; done
tail -2 gpurun_out/r2_t10_pytest.log; cat gpurun_out/r2_boundary_ab.log; for R in 1 8 128; do tail -2 gpurun_out/r2_node_probe_R$R.log | head -1; done
