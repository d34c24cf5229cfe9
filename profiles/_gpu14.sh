python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^$" | tail -6 | cut -c1-250 > gpurun_out/r2_t14_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-300 > gpurun_out/r2_smoke.log
tail -2 gpurun_out/r2_t14_pytest.log; cat gpurun_out/r2_smoke.log
