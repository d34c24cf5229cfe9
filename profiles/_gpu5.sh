python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^$" | tail -25 | cut -c1-250 > gpurun_out/r2_t5_pytest.log
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
python profiles/timeline_probe.py --requests 1 8 128 > gpurun_out/r2_timeline_steps.txt 2> gpurun_out/r2_timeline_steps.err
python profiles/timeline_probe.py --requests 1 8 128 --no-callback > gpurun_out/r2_timeline_job.txt 2> gpurun_out/r2_timeline_job.err
ncu --set full --clock-control none --import-source on -k regex:substep_torch_tma -s 60 -c 1 -o gpurun_out/r2_torch_tma_final python profiles/burst_probe.py --requests 128 --rng torch 2>&1 | tail -2 | cut -c1-200 > gpurun_out/r2_ncu_torch_tma.log
tail -3 gpurun_out/r2_t5_pytest.log; tail -4 gpurun_out/r2_bench_full.err | cut -c1-200
