set -u
run() { python bench.py --workload $1 --samples $2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline ${3:-} 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$1 ${3:-}', round(d['value']), round(d['ms_per_step'],3), d.get('exact_mode',{}).get('value'))"; }
run cfg2 1073741824 "--precision exact"; run cfg2s 1073741824; run cfg5 1073741824
