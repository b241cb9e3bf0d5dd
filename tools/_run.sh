cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_f_attn_bwd_persist.json 2> gpurun_out/bench_f.err; tail -c 300 gpurun_out/bench_f.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_r02_f_attn_bwd_persist.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value']); print({k:round(v['ms_per_step'],3) for k,v in d['roofline']['classes'].items()})
PY
