#!/bin/bash
# Round-2 ncu evidence (run under gpurun; text summaries come back in gpurun_out/, copy them to profiles/):
#   full captures of the step kernel in its four regimes + kernel A + the reset kernel, and the bench launch list
cd "$(dirname "$0")/.."
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
cap() {  # name, kernel regex, skip, args...
  name=$1; kern=$2; skip=$3; shift 3
  timeout 600 $NCU -k regex:$kern --launch-skip $skip --launch-count 1 -f -o /tmp/$name python tools/prof_cfg.py "$@" > $O/$name.log 2>&1
  python tools/ncu_summary.py /tmp/$name.ncu-rep $O/$name.txt > /dev/null 2>&1
  ncu -i /tmp/$name.ncu-rep --page source --csv > /tmp/$name.src.csv 2>/dev/null && python tools/ncu_src_stalls.py /tmp/$name.src.csv >> $O/$name.txt 2>/dev/null
}
python -c "import bench; print(bench.csrc_hash())" > $O/r2_csrc_sha16.txt   # the sources these captures belong to
cap r2_lat_graded_B4096 glg_step_units 2 B=4096 integrator=graded
cp /tmp/r2_lat_graded_B4096.ncu-rep $O/
cap r2_lat_fp32_graded_B4096 glg_step_units 2 B=4096 integrator=graded precision=fp32
cap r2_lat_fixed600_B4096 glg_step_units 2 B=4096 integrator=fixed
cap r2_tput_graded_B262144 glg_step_units 2 B=262144 integrator=graded
cap r2_tput_fp32_unc03_B262144 glg_step_units 2 B=262144 integrator=graded precision=fp32 uncertainty_scale=0.3 tables=19
cap r2_kernelA_fixed600_B16384 glg_step_kernel 1 B=16384 integrator=fixed role_warps=1 steps=2
cap r2_reset_B262144 glg_reset 0 B=262144 reset_only=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launch_list_bench.csv python bench.py --steps 2 --warmup 1 > $O/r2_launch_list_bench.log 2>&1
ls -la $O | tail -20
