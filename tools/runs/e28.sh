mkdir -p gpurun_out
for w in cfg4 cfg5 cfg3; do echo "== $w"; timeout 120 python tools/step_breakdown.py $w 100 2>&1 | grep "clocked launch\|CTAs\|general cluster\|graph=True" -A1 | grep -v "^--"; done > gpurun_out/e28_breakdown.log
