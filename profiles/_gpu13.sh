python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^$" | tail -6 | cut -c1-250 > gpurun_out/r2_t13_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-300 > gpurun_out/r2_smoke.log
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2_bench_ref_final.json 2> gpurun_out/r2_bench_ref_final.err
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
python profiles/timeline_probe.py --requests 1 8 128 > gpurun_out/r2_timeline_default.txt 2>/dev/null
python profiles/timeline_probe.py --requests 1 8 128 --per-step > gpurun_out/r2_timeline_steps.txt 2>/dev/null
python profiles/timeline_probe.py --requests 1 8 128 --no-callback > gpurun_out/r2_timeline_job.txt 2>/dev/null
tail -2 gpurun_out/r2_t13_pytest.log; cat gpurun_out/r2_smoke.log; tail -4 gpurun_out/r2_bench_final.err | cut -c1-200; grep "median" gpurun_out/r2_timeline_*.txt
