cd $GRAFT_REPO_ROOT
timeout 120 python tools/mlp_timing.py > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 0 -c 1 -f -o gpurun_out/prof_mlp_fwd_target python tools/mlp_timing.py > gpurun_out/ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 11 -c 1 -f -o gpurun_out/prof_mlp_fwd_online python tools/mlp_timing.py > gpurun_out/ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 33 -c 1 -f -o gpurun_out/prof_mlp_bwd python tools/mlp_timing.py > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/*.ncu-rep
