# Development aid.  Build the variants first (they are not kept in the tree), e.g.
#   for r in 80 96 104; do touch sts_b200/csrc/k_dft_fast.cu; make -C sts_b200 DFT_MAXREG=$r; cp sts_b200/libstsgpu.so sts_b200/ab_r$r.so; done
#   touch sts_b200/csrc/k_dft_fast.cu; make -C sts_b200
# the engine library is chosen with STSGPU_LIB (sts_b200/binding.py).
# A/B of engine builds and knobs through bench.py (resident value and end-to-end), 30 steps each
one() { env "$@" timeout 300 python bench.py --no-cpu-baseline --no-config4 --steps 30 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('value %.2f (%.3f ms)  e2e %.2f (%.3f ms)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))"; }
for lib in libstsgpu ab_r80 ab_r96 ab_r104; do
  for bg in 2 3; do
    echo "== $lib LC_BG=$bg: $(one STSGPU_LIB=$PWD/sts_b200/$lib.so STSGPU_LC_BG=$bg)"
  done
done
echo "== default RANK_BG=2: $(one STSGPU_RANK_BG=2)"
echo "== default RANK_BG=4: $(one STSGPU_RANK_BG=4)"
echo "== default DFT_CHUNK=64: $(one STSGPU_DFT_CHUNK=64)"
echo "== default DFT_CHUNK=256: $(one STSGPU_DFT_CHUNK=256)"
