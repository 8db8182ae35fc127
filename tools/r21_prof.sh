set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
python bench.py --steps 20 --warmup 3 --profile-out gpurun_out/r21_profile.json > gpurun_out/r21_bench.json 2> gpurun_out/r21_bench.err; echo bench rc=$?
$CMD > gpurun_out/r21_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r21_launches.csv $CMD > gpurun_out/r21_ncu1.log 2>&1; echo ncu1 rc=$?
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 380 -c 12 -o gpurun_out/r21_conv -f $CMD > gpurun_out/r21_ncu2.log 2>&1; echo ncu2 rc=$?
ncu --set full --clock-control none --import-source on -k regex:guided_step -s 18 -c 2 -o gpurun_out/r21_step -f $CMD > gpurun_out/r21_ncu3.log 2>&1; echo ncu3 rc=$?
ncu --set full --clock-control none --import-source on -k regex:gn_apply -s 300 -c 6 -o gpurun_out/r21_gn -f $CMD > gpurun_out/r21_ncu4.log 2>&1; echo ncu4 rc=$?
ls -la gpurun_out/
