import sys
sys.path.insert(0,'/root/repo')
import sts_b200
n=int(sys.argv[1]); B=256
eng=sts_b200.Engine(max_streams=B,n=n,test_mask=1<<7)
eng.generate_lcg(0,B); eng.set_profiling(1)
for r in range(3):
    eng.submit_resident(B,0); eng.wait()
print(n,[(k,round(v*1024/B,3)) for k,v in eng.kernel_times() if k=='dft'])
