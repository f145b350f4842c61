import os, sys, time
sys.path.insert(0, '/root/repo')
import sts_b200
B = 256
eng = sts_b200.Engine(max_streams=B, test_mask=1 << 7)
eng.generate_lcg(0, B)
eng.set_profiling(1)
for r in range(4):
    eng.submit_resident(B, 0); eng.wait()
print(os.environ.get('STSGPU_DFT_DBG', '0'), [(n, round(ms * 4, 3)) for n, ms in eng.kernel_times() if n == 'dft'])
