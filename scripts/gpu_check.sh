#!/bin/bash
# One GPU-box visit: kernel tests group by group (separate processes so one fault does not poison
# the rest), module tests, smoke, short bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/tests.log
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for k in ${KGROUPS:-tcgen05 c1c2_fused stitch conv3x3_fwd_simt dgrad_wgrad conv_c1 conv2d pool4 layout linear bce_threat binarise threat_score ats guard}; do
  echo "=== $k" >> gpurun_out/tests.log
  timeout 240 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "$k" -x 2>&1 | tail -25 >> gpurun_out/tests.log
done
if [ -z "$SKIP_MODULES" ]; then
echo "=== modules" >> gpurun_out/tests.log
timeout 420 python -m pytest tests/test_modules_gpu.py -m gpu -q 2>&1 | grep -v Warning | tail -60 >> gpurun_out/tests.log
echo "=== smoke" >> gpurun_out/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/tests.log 2>&1
fi
echo "=== bench" >> gpurun_out/tests.log
timeout 300 python bench.py --steps ${BENCH_STEPS:-3} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/bench.err >> gpurun_out/tests.log
cat gpurun_out/bench.json >> gpurun_out/tests.log
grep -v "^{" gpurun_out/tests.log | grep -v Warning | tail -120
