#!/bin/bash
# One gpurun call that re-validates a frozen build: GPU tests, smoke, both bench arms, ncu evidence, the config-4 rollout
# measurement, compute-sanitizer.  Text results land in gpurun_out/ (copy the r2_* files to profiles/).
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | cut -c1-300 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > $O/bench_final.json 2> $O/bench_final.err; echo "bench rc=$?"
python bench.py --impl reference > $O/bench_final_ref.json 2>> $O/bench_final.err; echo "ref rc=$?"
bash tools/ncu_r2.sh > $O/ncu_r2.log 2>&1; tail -2 $O/ncu_r2.log
(for w in kernels torch; do python tools/rollout_policy.py --envs 65536 --steps 16 --wrapper $w; done) > $O/r2_rollout_policy.txt 2>&1; tail -4 $O/r2_rollout_policy.txt
python tools/layout_sweep.py > $O/r2_layout_sweep.txt 2>&1; tail -3 $O/r2_layout_sweep.txt
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py > $O/r2_sanitizer_memcheck.log 2>&1; tail -3 $O/r2_sanitizer_memcheck.log
timeout 2000 compute-sanitizer --tool racecheck python tools/sanitize_run.py > $O/r2_sanitizer_racecheck.log 2>&1; tail -3 $O/r2_sanitizer_racecheck.log
