cd /root/repo
O=gpurun_out
python -m pytest tests -m gpu -q --tb=short -x 2>&1 | cut -c1-300 | tail -12
python tools/profile_host_step.py 2>&1 | grep -E "per call" | tail -4
python bench.py > $O/bench_s3_2.json 2> $O/bench_s3_2.err; echo "bench rc=$?"
