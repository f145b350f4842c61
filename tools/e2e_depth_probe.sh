for cfg in "STSGPU_COMPUTE_INFLIGHT=0 STS_BENCH_DEPTH=3" "STSGPU_COMPUTE_INFLIGHT=2 STS_BENCH_DEPTH=3" "STSGPU_COMPUTE_INFLIGHT=2 STS_BENCH_DEPTH=4" "STSGPU_COMPUTE_INFLIGHT=0 STS_BENCH_DEPTH=2" "STSGPU_COMPUTE_INFLIGHT=1 STS_BENCH_DEPTH=3"; do
  echo "== $cfg"; env $cfg timeout 300 python bench.py --no-cpu-baseline --no-config4 --steps 30 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('value %.2f (%.3f ms)  e2e %.2f (%.3f ms)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))"
done
