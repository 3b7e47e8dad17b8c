#!/bin/bash
# compute-sanitizer over the small GPU parity tests (the `engine` tests in their cuda form: every kernel of the library runs at
# least once on inputs of a few hundred thousand path nodes).  memcheck: out-of-bounds and misaligned accesses; leakcheck:
# device memory still allocated by THIS library at exit (the caching allocator of torch keeps its blocks by design: its
# frames are not counted); racecheck: shared-memory hazards (the block-sized locate sort, the builder's block scans, the
# find kernels' tables).  Run on a GPU box: gpurun -- 'bash scripts/sanitize.sh [tools]'; the logs land in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOLS="${*:-memcheck leakcheck racecheck}"
SEL='kat1 or locate_medium or locate_short_range or locate_into_host or all_operations_random_graphs or fused_table'
SEL="$SEL or mem_scan or linear_builder or correct_index or device_compare"
for tool in $TOOLS; do
  args="--tool $tool"
  sel="$SEL"
  [ "$tool" != racecheck ] && sel="$SEL or two_kernel"          # every table shape x pattern length of the k-mer form: too slow under racecheck
  [ "$tool" = leakcheck ] && args="--tool memcheck --leak-check full --print-limit 100000"
  timeout 1500 compute-sanitizer $args --error-exitcode 9 \
    python -m pytest tests/test_gpu_parity.py tests/test_mem.py tests/test_linear_builder.py tests/test_verify_gpu.py tests/test_compare_kmers.py \
      -x -q -m gpu -k "$sel" -p no:cacheprovider > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "$tool: exit $?"
  grep -E "ERROR SUMMARY|passed|failed|LEAK SUMMARY|RACECHECK SUMMARY" gpurun_out/r02_sanitizer_$tool.log | tail -3
  if [ "$tool" = leakcheck ]; then
    echo "leaked allocations made by libgcsa2_b200.so: $(grep -A3 'Leaked' gpurun_out/r02_sanitizer_$tool.log | grep -c 'in libgcsa2_b200.so')"
    grep -v "Host Frame\|Device Frame\|^=========$" gpurun_out/r02_sanitizer_$tool.log | tail -400 > gpurun_out/r02_sanitizer_$tool.short.log
    mv gpurun_out/r02_sanitizer_$tool.short.log gpurun_out/r02_sanitizer_$tool.log
  fi
done
