cd $GRAFT_REPO_ROOT
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e > gpurun_out/bench_n8_$name.json 2> gpurun_out/bench_n8_$name.err
  python - <<PY
import json
for l in open('gpurun_out/bench_n8_$name.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$name', d['value'], d['ms_per_step'])
PY
}
run c4_62 V2S_COMM_SMS=4 V2S_SYNC_SPLITS=6,2
run c2_62 V2S_COMM_SMS=2 V2S_SYNC_SPLITS=6,2
run c8_62 V2S_COMM_SMS=8 V2S_SYNC_SPLITS=6,2
run c4_731 V2S_COMM_SMS=4 V2S_SYNC_SPLITS=7,3,1
