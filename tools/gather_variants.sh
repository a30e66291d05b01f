# usage: bash tools/gather_variants.sh <n_gpus> "<variant flags>" ...   (one bench run per variant, gather figures only)
N=$1; shift
for v in "$@"; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-weak --e2e-steps 1 $v > gpurun_out/gv.json 2> gpurun_out/gv.err; tail -2 gpurun_out/gv.err | cut -c1-300; python -c "
import json,sys; d=json.loads(open('gpurun_out/gv.json').read().strip().splitlines()[-1]); g=d.get('gather'); print('N=$N $v | sharded ms', round(d['ms_per_step'],3), '| e2e', '%.3g' % d['e2e']['value'], '%.3g' % d['e2e']['lazy_mueller']['value'], '|', {k:g.get(k) for k in ('transport','chunks','ms_per_step','GBps_received_per_rank','ok','cuda_graph','cuda_graph_error')})"; done
