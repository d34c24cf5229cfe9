for v in 8 9; do python profiles/variant_check.py $v 2>&1 | tail -5 | cut -c1-120 >> gpurun_out/r2_variant_p2.log; done
for v in 1 8 9; do echo "tma=$v" >> gpurun_out/r2_variant_p2.log; python profiles/burst_probe.py --requests 128 --rng torch --tma $v 2>&1 | tail -1 | cut -c1-200 >> gpurun_out/r2_variant_p2.log; done
cat gpurun_out/r2_variant_p2.log
