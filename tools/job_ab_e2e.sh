#!/bin/bash
# A/B of complete bench lines (value + e2e) for the default library and every build_variants/*.so, then the e2e stage trace.
out=${1:-gpurun_out/ab_e2e.txt}
: > $out
for lib in "" build_variants/*.so; do
  [ -n "$lib" ] && [ ! -f "$lib" ] && continue
  RFB200_LIB=${lib:+$PWD/$lib} python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('${lib:-default}', 'value %.0f e2e %.0f ms/step %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']), {k: round(v, 2) for k, v in d['config']['stage_ms_per_step'].items()}, 'e2e_insert_s %.4f' % d.get('e2e_insert_s', 0), {k: round(v, 2) for k, v in d.get('e2e_stage_ms_per_step', {}).items()})" >> $out 2>&1
done
cat $out
