#!/bin/bash
# One ncu --set full capture of every hand-written kernel at its bench shape (second repetition: warm), after the same
# command has run cleanly without ncu.  Output: gpurun_out/r2_prof_all.ncu-rep (read with scripts/ncu_summary.py).
mkdir -p gpurun_out
CMD="python scripts/run_all_kernels.py"
REPS=2 $CMD > gpurun_out/plain_all.log 2>&1 && \
REPS=2 ncu --set full --clock-control none --import-source on \
    -k regex:'stitch|u8_to_f32|conv|pool4|linear|bce|adam|wgrad|colsum|ats_kernel|split|dil' -s ${NCU_SKIP:-40} -c ${NCU_COUNT:-60} \
    -f -o gpurun_out/${NCU_OUT:-r2_prof_all} $CMD > gpurun_out/ncu_all.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/ncu_all.log
