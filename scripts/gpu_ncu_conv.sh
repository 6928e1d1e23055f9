#!/bin/bash
# ncu --set full capture of the c2-shaped tcgen05 conv kernels (fwd, dgrad, wgrad) at the bench shape.
# KREGEX selects kernels (default: all three), NCU_COUNT how many launches after the 3 warm-up ones.
mkdir -p gpurun_out
CMD="python scripts/run_conv_kernels.py"
REPS=2 $CMD > gpurun_out/plain_conv.log 2>&1 && \
REPS=2 ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-'conv3x3_c32_(s1_tc|tc|wgrad_tc|dgrad)'} -s ${NCU_SKIP:-3} -c ${NCU_COUNT:-3} \
    -f -o gpurun_out/${NCU_OUT:-prof_conv} $CMD > gpurun_out/ncu_conv.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/ncu_conv.log
