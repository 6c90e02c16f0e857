#!/bin/bash
# bench.py on N GPUs of one node, launched as the driver does; JSON line into gpurun_out/<tag>_bench_n<N>_final.json
n=$1; tag=${2:-r2b}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 8 --warmup 3 \
  > gpurun_out/${tag}_bench_n${n}_final.json 2> gpurun_out/${tag}_bench_n${n}_final.err
tail -c 300 gpurun_out/${tag}_bench_n${n}_final.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/${tag}_bench_n${n}_final.json') if l.startswith('{')][-1])
print('N=$n value %.0f e2e %.0f ms/step %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']), d.get('reduce_check', {}).get('ok'), 'strong %.4f s' % d['strong_scaling']['seconds'], 'reduce %.4f s' % d.get('e2e_reduce_s', 0), [round(c['value']) for c in d.get('other_configs', [])])"
