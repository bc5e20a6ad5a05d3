for i in 1 2 3; do timeout 300 python bench.py --steps 30 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'])"; done
nvidia-smi --query-gpu=name,power.limit,temperature.gpu --format=csv
