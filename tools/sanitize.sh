#!/bin/bash
# compute-sanitizer passes over the kernel-level and small-geometry GPU tests plus the conv head at full geometry (run on a B200: `gpurun --timeout 1500 -- bash tools/sanitize.sh`).
# memcheck: out-of-bounds / misaligned global + shared accesses; racecheck: shared-memory hazards; initcheck: reads of
# uninitialised global memory. The full-size tests are skipped (the tools slow kernels down 10-100x).
# Summaries land in gpurun_out/sanitize_<tool>.log; exit status is non-zero if any tool reported an error.
set -u
mkdir -p gpurun_out
SEL='tests/test_gpu_kernels.py tests/test_gpu_entropy.py'
MODEL='tests/test_gpu_model.py tests/test_gpu_batch.py tests/test_gpu_precision.py'
KSEL='small or (tiny69 and (g_s or round_trip))'   # + the conv head (grouped un-patchify epilogue) at full 721x1440 geometry
rc=0
for tool in ${TOOLS:-memcheck racecheck initcheck}; do
  log=gpurun_out/sanitize_${tool}.log
  # racecheck flags tcgen05.alloc.cta_group::2 itself (the instruction writes the TMEM address into the slot of BOTH
  # CTAs of the pair, each CTA's allocator warp issuing it as the two-CTA pattern prescribes: same value, read only
  # after the cluster barrier), so the CTA-pair GEMM test is left out of that pass; memcheck and initcheck cover it
  KSKIP=''; [ "${tool}" = racecheck ] && KSKIP='not cta_pair'
  timeout 1200 compute-sanitizer --tool ${tool} --error-exitcode 7 --print-limit 20 \
      python -m pytest ${SEL} -m gpu -x -q -k "${KSKIP}" > ${log} 2>&1
  st=$?
  timeout 1200 compute-sanitizer --tool ${tool} --error-exitcode 7 --print-limit 20 \
      python -m pytest ${MODEL} -k "${KSEL}" -m gpu -x -q >> ${log} 2>&1
  st2=$?
  [ ${st2} -ne 0 ] && st=${st2}
  echo "${tool}: exit ${st}; $(grep -c 'ERROR SUMMARY' ${log}) summaries; $(grep 'ERROR SUMMARY' ${log} | tail -1)"
  tail -3 ${log}
  [ ${st} -ne 0 ] && rc=1
done
exit ${rc}
