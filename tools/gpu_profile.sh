#!/bin/bash
# usage (under gpurun): bash tools/gpu_profile.sh <tag>
# plain bench -> ncu launch list -> ncu --set full captures of the conv, GroupNorm and step kernels.
# The .ncu-rep files are converted to raw CSV on the box and deleted (gpurun_out/ is capped at 64 MiB).
T=${1:-rXX}
set -x
CMD="python bench.py --steps 1 --warmup 3 --denoise-steps 2 --no-extras --no-cpu-baseline"
$CMD > gpurun_out/${T}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu1.log 2>&1; echo ncu1 rc=$?
gzip -f gpurun_out/${T}_launches.csv
ncu --set full --clock-control none -k regex:conv_igemm -s 384 -c 8 -o /tmp/${T}_conv -f $CMD > gpurun_out/${T}_ncu2.log 2>&1; echo ncu2 rc=$?
ncu -i /tmp/${T}_conv.ncu-rep --page raw --csv > gpurun_out/${T}_conv_raw.csv
ncu --set full --clock-control none -k regex:gn_apply -s 279 -c 3 -o /tmp/${T}_gn -f $CMD > gpurun_out/${T}_ncu4.log 2>&1; echo ncu4 rc=$?
ncu -i /tmp/${T}_gn.ncu-rep --page raw --csv > gpurun_out/${T}_gn_raw.csv
CMD2="python tools/step_kernels_bench.py 256 1"
$CMD2 > gpurun_out/${T}_stepk_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:"guided_step|l2reg|extract_noise|map2|apply_mask|to_uint8|color_grad" -c 24 -o /tmp/${T}_stepk -f $CMD2 > gpurun_out/${T}_ncu3.log 2>&1; echo ncu3 rc=$?
ncu -i /tmp/${T}_stepk.ncu-rep --page raw --csv > gpurun_out/${T}_stepk_raw.csv
tail -c 400 gpurun_out/${T}_ncu1.log gpurun_out/${T}_ncu2.log gpurun_out/${T}_ncu3.log gpurun_out/${T}_ncu4.log
du -sh gpurun_out
