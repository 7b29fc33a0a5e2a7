cd /root/repo
O=gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | cut -c1-300 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > $O/bench_s3_0.json 2> $O/bench_s3_0.err; echo "bench rc=$?"
python bench.py --impl reference > $O/bench_s3_0_ref.json 2>> $O/bench_s3_0.err; echo "ref rc=$?"
timeout 300 python examples/ppo_device_rollout.py --iters 3 > $O/ppo_example.log 2>&1; echo "ppo rc=$?"; tail -3 $O/ppo_example.log
bash tools/ncu_r2.sh > $O/ncu_r2.log 2>&1; tail -2 $O/ncu_r2.log
