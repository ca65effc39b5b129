for cl in "4 2" "4 3" "4 4" "8 2" "8 3" "8 4" "16 2" "2 4"; do set -- $cl; python bench.py --steps 5 --warmup 3 --no-cpu-baseline --chunk $1 --lanes $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('chunk lanes $cl', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2), 'value', round(d['value']))
"; done
