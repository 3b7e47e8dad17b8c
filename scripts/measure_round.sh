#!/bin/bash
# One gpurun call that measures everything that was built without a GPU at the end of round 1
# ("Not yet measured"), cheapest and most important first, each step under its own timeout so
# that a hang costs minutes, not the box.  Everything lands in gpurun_out/ (merged back by gpurun).
#
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/measure_round.sh r02'
#
# The index of configs[1] is built once and cached (npz) so that the variants below do not pay for it again.
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
CACHE=/dev/shm/gcsa2_b200_cfg2_index.npz
run() { local limit=$1; shift; echo "=== $* (limit ${limit}s)" | tee -a $OUT/${TAG}_steps.log; timeout $limit "$@"; echo "    exit $?" | tee -a $OUT/${TAG}_steps.log; }

# 1. parity first: a fast kernel with different answers is not done
run 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
tail -3 $OUT/${TAG}_pytest_gpu.log

# 2. the bench line with the defaults (fused table, short-range locate, automatic host packing)
run 900 python bench.py --index-cache $CACHE > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err
tail -c 600 $OUT/${TAG}_bench_default.json

# 3. one switch at a time against the default (find leg only: --no-locate --no-cpu-baseline keeps these short)
SHORT="--index-cache $CACHE --no-locate --no-cpu-baseline --steps 10 --warmup 3"
run 600 python bench.py $SHORT --fused-table 0 > $OUT/${TAG}_bench_unfused.json 2>> $OUT/${TAG}_bench_variants.err
GCSA_B200_HOST_PACK=0 run 600 python bench.py $SHORT > $OUT/${TAG}_bench_nopack.json 2>> $OUT/${TAG}_bench_variants.err
GCSA_B200_HOST_PACK=$(nproc) run 600 python bench.py $SHORT > $OUT/${TAG}_bench_forcepack.json 2>> $OUT/${TAG}_bench_variants.err
# the locate leg with and without the short-range path
GCSA_B200_LOCATE_SMALL=0 run 900 python bench.py --index-cache $CACHE --no-cpu-baseline > $OUT/${TAG}_bench_locate_general.json 2>> $OUT/${TAG}_bench_variants.err

# 4. launch list of the default bench (shares of the step, not absolute times) and one full capture per hot kernel
run 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_bench.csv \
    python bench.py $SHORT --steps 2 --warmup 1 > $OUT/${TAG}_ncu_launches.log 2>&1
run 900 ncu --set full --clock-control none --import-source on -k regex:find_kernel -s 3 -c 2 -o $OUT/${TAG}_prof_find -f \
    python bench.py $SHORT --steps 2 --warmup 1 > $OUT/${TAG}_ncu_find.log 2>&1
run 900 ncu --set full --clock-control none --import-source on -k regex:locate_small -s 4 -c 4 -o $OUT/${TAG}_prof_locate -f \
    python bench.py --index-cache $CACHE --no-cpu-baseline --queries 1000000 --steps 1 --warmup 1 > $OUT/${TAG}_ncu_locate.log 2>&1

# 5. the other operations on configs[2] (find 64-mers, count, locate, parent, MEM scan with and without jumps, k-mers)
run 1500 python scripts/bench_ops.py --kmer-table-k 16 --ops count,locate,parent,mem,kmers,compare --out $OUT/${TAG}_ops_cfg3.json > $OUT/${TAG}_ops_cfg3.log 2>&1
tail -5 $OUT/${TAG}_ops_cfg3.log
echo "measure_round: done" | tee -a $OUT/${TAG}_steps.log
