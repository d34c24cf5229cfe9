# racecheck, full hazard list (deduplicated), default release and strict per-thread release
for mode in 0 1; do
  LANPAINT_B200_STRICT_RELEASE=$mode timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "torch_stream_tma or tma_staged or boundary_tma or fused_cfg_combine" > /tmp/race_$mode.log 2>&1
  ( echo "LANPAINT_B200_STRICT_RELEASE=$mode compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k 'torch_stream_tma or tma_staged or boundary_tma or fused_cfg_combine'"; grep -E "passed|failed|RACECHECK SUMMARY|Error: Race|Warning: Race" /tmp/race_$mode.log | sed 's/+0x[0-9a-f]*//g' | cut -c1-230 | sort | uniq -c | sort -rn | head -30 ) > gpurun_out/r2_racecheck_mode$mode.log
done
LANPAINT_B200_STRICT_RELEASE=1 python profiles/burst_probe.py --requests 128 --rng torch 2>&1 | tail -1 | cut -c1-200 > gpurun_out/r2_burst_torch_strict.log
ncu --set full --clock-control none --import-source on -k regex:"boundary_tma|substep_torch_tma|synth_denoiser" -s 440 -c 12 -o gpurun_out/r2_job_kernels python profiles/node_probe.py --requests 128 --calls 4 2>&1 | tail -2 | cut -c1-200 > gpurun_out/r2_ncu_job_kernels.log
cat gpurun_out/r2_racecheck_mode0.log | head -8 | cut -c1-200; cat gpurun_out/r2_racecheck_mode1.log | head -8 | cut -c1-200; cat gpurun_out/r2_burst_torch_strict.log
