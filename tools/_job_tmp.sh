python -m pytest tests -m gpu -x -q 2>&1 | tail -6
QG_TRACE=1 python tools/prof_step.py 256 8 2 2>&1 | grep "facet n=1\|cell n=1\|kd-tree" | tail -3
QG_NO_FACET_ALIAS=1 QG_TRACE=1 python tools/prof_step.py 256 8 2 2>&1 | grep "facet n=1" | tail -1
python bench.py --steps 10 --warmup 3 --e2e-steps 0 --no-cpu | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C3', d['ms_per_step'], d['value'], d['pass_ms']['domain'])"
python bench.py --config C5 --steps 5 --warmup 3 --e2e-steps 0 --no-cpu | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C5', d['ms_per_step'], d['value'], d['pass_ms']['domain'])"
