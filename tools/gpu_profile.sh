#!/bin/bash
# usage (under gpurun): bash tools/gpu_profile.sh <tag>
# plain bench -> ncu launch list -> ncu --set full captures of the conv, step and GroupNorm kernels
T=${1:-rXX}
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
python bench.py --steps 20 --warmup 3 --profile-out gpurun_out/${T}_profile.json > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo bench rc=$?
$CMD > gpurun_out/${T}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu1.log 2>&1; echo ncu1 rc=$?
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 380 -c 12 -o gpurun_out/${T}_conv -f $CMD > gpurun_out/${T}_ncu2.log 2>&1; echo ncu2 rc=$?
ncu --set full --clock-control none --import-source on -k regex:guided_step -s 18 -c 2 -o gpurun_out/${T}_step -f $CMD > gpurun_out/${T}_ncu3.log 2>&1; echo ncu3 rc=$?
ncu --set full --clock-control none --import-source on -k regex:gn_apply -s 277 -c 6 -o gpurun_out/${T}_gn -f $CMD > gpurun_out/${T}_ncu4.log 2>&1; echo ncu4 rc=$?
