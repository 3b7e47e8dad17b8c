#!/bin/bash
# ncu captures of the kernels bench.py times in round 2 (one GPU): the launch list of a short default run, full
# captures of the k-mer form of find() on configs[1] and on the 3 Gbp index of configs[3], of the short-range locate
# kernels and of the MEM-style scan.  Everything lands in gpurun_out/ (summaries go to profiles/ from there).
set -u
OUT=gpurun_out
SHORT="--no-locate --no-mem --no-cfg4 --no-cpu-baseline --steps 2 --warmup 1"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r02_launches_bench_final.csv \
    python bench.py --no-cpu-baseline --steps 2 --warmup 1 --cfg4-queries 250000000 --cfg4-steps 1 --mem-steps 1 > $OUT/r02_ncu_launches_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:find_fast_kernel -s 3 -c 1 -o $OUT/r02_prof_fast4_cfg2 -f python bench.py $SHORT > $OUT/r02_ncu_fast4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:find_quad_kernel -s 3 -c 1 -o $OUT/r02_prof_quad_cfg2 -f python bench.py $SHORT > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"find_fast_kernel|find_quad_kernel|find_kernel" -s 6 -c 3 -o $OUT/r02_prof_find_cfg4 -f \
    python scripts/bench_build.py --mbp 3000 --options '[{"walk_table":0,"two_step":false,"fused_table":false}]' > $OUT/r02_ncu_cfg4.log 2>&1
echo done
