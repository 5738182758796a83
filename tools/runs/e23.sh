mkdir -p gpurun_out
for w in cfg4 cfg5; do echo "== $w"; timeout 120 python tools/step_breakdown.py $w 100 2>&1 | grep "blob structure" -A3; done > gpurun_out/e23_breakdown.log
