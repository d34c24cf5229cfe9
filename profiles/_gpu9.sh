python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^$" | tail -12 | cut -c1-250 > gpurun_out/r2_t9_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-300 > gpurun_out/r2_smoke.log
for R in 1 8 128; do python profiles/node_probe.py --requests $R 2>&1 | tail -10 | cut -c1-200 > gpurun_out/r2_node_probe_R$R.log; done
python profiles/timeline_probe.py --requests 1 8 128 > gpurun_out/r2_timeline_default.txt 2>/dev/null
python profiles/timeline_probe.py --requests 1 8 128 --per-step > gpurun_out/r2_timeline_steps.txt 2>/dev/null
python profiles/timeline_probe.py --requests 1 8 128 --no-callback > gpurun_out/r2_timeline_job.txt 2>/dev/null
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2_bench_full2.json 2> gpurun_out/r2_bench_full2.err
tail -3 gpurun_out/r2_t9_pytest.log; cat gpurun_out/r2_smoke.log; tail -4 gpurun_out/r2_bench_full2.err | cut -c1-200; for R in 1 8 128; do tail -3 gpurun_out/r2_node_probe_R$R.log; done
