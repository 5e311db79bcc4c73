for sps in 32 64 128 256; do
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --scans-per-step $sps 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['scans_per_step'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['avg_launch_ms'], d['roofline']['frac'])"
done
