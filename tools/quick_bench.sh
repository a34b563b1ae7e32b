python "$(dirname "$0")/../bench.py" --steps 300 --warmup 10 --no-cpu --no-e2e "$@" 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value',round(d['value']), 'ms/step',round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()}, 'frac',round(d['roofline']['frac'],3))"
