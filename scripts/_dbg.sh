timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for i in 1 2 3; do python bench.py --no-families --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['e2e']['value']), d['roofline']['by_class_serial_ms']['elementwise'])"; done
