( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2_bench_final2.json 2> gpurun_out/r2_bench_final2.err
tail -4 gpurun_out/r2_bench_final2.err | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_final2.json').read().strip().splitlines()[-1])
print('value %.0f e2e %.0f'%(d['value'], d['e2e']['value']), d['other_sampler'])
print(d['config']['launch'], [ (c['name'], round(c['torch']['ms_per_job_device'],3), c['torch']['launch_mode']) for c in d['config']['configs']])
PY
