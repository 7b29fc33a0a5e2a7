cd /root/repo
O=gpurun_out
python -m pytest tests/test_gpu_features.py -m gpu -q --tb=short -k "overlapped or split" 2>&1 | cut -c1-300 | tail -15
python bench.py > $O/bench_s3_1.json 2> $O/bench_s3_1.err; echo "bench rc=$?"
python tools/profile_host_step.py > $O/host_step_profile.txt 2>&1; tail -25 $O/host_step_profile.txt
for v in base tput3 tput5 tput6 unfused2; do python tools/time_lib2.py tools/ubench/exp_$v.so 2>&1 | grep "B=262144"; done
