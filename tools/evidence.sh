# Round-end evidence on one B200 (run through gpurun): default bench line, ncu launch list of the same command,
# one `--set full` capture of two forward layers + one backward layer.  Outputs under gpurun_out/ (<64 MiB).
set -x
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain.log 2>&1; echo "plain exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 230 --csv --log-file gpurun_out/launches_r01_final2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu1.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc|attn|ln_bwd' -s 633 -c 26 -o gpurun_out/prof_r01_final3 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu2.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out/
