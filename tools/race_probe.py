import sys, ctypes as C
sys.path.insert(0,'.')
from nextpolish_b200 import engine as E
from tests.synth_cases import CASES
from tests.conftest import run_checker
O=C.CDLL('oracle/libnp_oracle.so'); O.np_oracle_run.argtypes=[C.c_void_p,C.c_int,C.c_void_p,C.c_void_p,C.c_int64,C.c_void_p]
kw=dict(CASES['noisy']); kw['contig_len']=int(sys.argv[1]) if len(sys.argv)>1 else 40000
sh=E.Shard.synthetic(E.synth_params(**kw),0,kw['n_contigs'])
cfg=E.default_config(b"")
want=run_checker(O.np_oracle_run,sh,1,cfg)
eng=E.Engine(0)
bad=0
for it in range(int(sys.argv[2]) if len(sys.argv)>2 else 1):
    got=eng.polish(sh,1,cfg)
    if got!=want:
        bad+=1
        for n in want:
            if got[n]!=want[n]:
                a,b=want[n],got[n]; i=next((i for i in range(min(len(a),len(b))) if a[i]!=b[i]),-1)
                print('MISMATCH it',it,n,len(a),len(b),i,a[max(0,i-10):i+10],b[max(0,i-10):i+10])
print('bad',bad,eng.window_stats())
