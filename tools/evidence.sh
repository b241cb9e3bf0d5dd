# Round evidence on one B200 (run through gpurun): ncu launch list of a short bench command and `--set full` captures of
# forward and backward layers of the same command.  Outputs under gpurun_out/ (<64 MiB).  Usage: bash tools/evidence.sh r02
R=${1:-r02}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-e2e"
set -x
timeout 300 $CMD > gpurun_out/plain_$R.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_$R.csv $CMD > gpurun_out/ncu1.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mlp_tc|gemm_tc|attn|ln_bwd' -s 514 -c 8 -o gpurun_out/prof_${R}_fwd -f $CMD > gpurun_out/ncu2.log 2>&1; echo "ncu fwd exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mlp_tc|gemm_tc|attn|ln_bwd' -s 570 -c 12 -o gpurun_out/prof_${R}_bwd -f $CMD > gpurun_out/ncu3.log 2>&1; echo "ncu bwd exit $?"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_$R.csv
