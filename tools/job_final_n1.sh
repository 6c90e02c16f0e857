#!/bin/bash
# Final single-GPU evidence of a round: the bench line, the ncu launch list of the same command (short), one ncu --set full
# capture of the kernels of one 720-image chunk, the clocks during the bench, the CLI throughput from page-cached files.
tag=${1:-r2b}
o=gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $o/${tag}_clocks.csv &
smi=$!
python bench.py > $o/${tag}_bench_n1_final.json 2> $o/${tag}_bench_n1_final.err
kill $smi
tail -c 600 $o/${tag}_bench_n1_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 800 --csv --log-file $o/${tag}_launches_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $o/${tag}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_gather_sticks|k_fft_cols_slices|k_fft_rows|k_damped_scatter|k_edge2' -s 21 -c 8 \
    -o $o/${tag}_final_full python bench.py --steps 1 --warmup 3 --batch 720 --no-cpu-baseline --no-e2e --no-extras > $o/${tag}_ncu_full.log 2>&1
python tools/cli_throughput.py -n 98304 --thr 16 --gpus 1 > $o/${tag}_cli_1gpu.txt 2>&1
tail -3 $o/${tag}_cli_1gpu.txt
python -c "
import json
d=json.loads([l for l in open('$o/${tag}_bench_n1_final.json') if l.startswith('{')][-1])
print('value %.0f e2e %.0f ms/step %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']), d['config']['stage_ms_per_step'], d['clocks'], d['strong_scaling']['seconds'], [ (c.get('config'), round(c.get('value',0))) for c in d.get('other_configs',[])])
"
