one() { env "$@" timeout 300 python bench.py --no-cpu-baseline --no-config4 --steps 30 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('value %.2f (%.3f ms)  e2e %.2f  pattern %.3f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['all_kernel_groups_ms']['pattern(serial,apen)']))"; }
echo "== old kernel: $(one STSGPU_HIST_WIDE=0)"
echo "== wide 48 regs: $(one STSGPU_HIST_WIDE=1)"
echo "== wide 56 regs: $(one STSGPU_LIB=$PWD/sts_b200/ab_h56.so)"
echo "== wide 64 regs: $(one STSGPU_LIB=$PWD/sts_b200/ab_h64.so)"
echo "== old kernel: $(one STSGPU_HIST_WIDE=0)"
echo "== wide 48 regs: $(one STSGPU_HIST_WIDE=1)"
