#!/bin/bash
# round-2 GPU session Y (1 GPU): the driver's own commands, timed
cd "$(dirname "$0")/.."
O=gpurun_out
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r2y_driver_reference.json 2> $O/r2y_driver_reference.err ) 2> $O/r2y_driver_reference.time; echo "reference rc=$?"
( time python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2y_driver_bench.json 2> $O/r2y_driver_bench.err ) 2> $O/r2y_driver_bench.time; echo "bench rc=$?"
cat $O/r2y_driver_reference.time $O/r2y_driver_bench.time
python -c "
import json
for f in ('r2y_driver_reference', 'r2y_driver_bench'):
    b = json.load(open('$O/' + f + '.json'))
    print(f, 'value', b['value'], 'ms/step', b['ms_per_step'], 'e2e', b['e2e'], 'extrap', b.get('extrapolated'), b.get('sample_steps'), (b.get('roofline') or {}).get('frac'), (b.get('roofline_pair') or {}).get('frac'), b.get('clocks'))
"
