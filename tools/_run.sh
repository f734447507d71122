python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/kernel_roofline.py --only EqThree 2>/dev/null | cut -c1-170
python tools/kernel_roofline.py --only Envelope 2>/dev/null | cut -c1-170
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --workload audio | python -c "import sys,json; r=json.loads(sys.stdin.readline()); print({k:round(v['avg_launch_ms'],4) for k,v in r['kernels'].items()}, round(r['ms_per_step'],4))"
