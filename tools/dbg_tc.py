import sys, os
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import torch as t
import test_gpu_parity as T
from golden_io import rel_err
class MP:
    def setenv(self,k,v): os.environ[k]=v
    def delenv(self,k,raising=False): os.environ.pop(k,None)
for shape in [(130,3,24,16),(64,5,30,18),(301,2,20,6)]:
    P,Q,sample,ip,data,names=T._movielens_case(*shape, seed=21)
    for qf in ("1","0"):
        if qf=="0": os.environ["ALAN_B200_NO_QFUSE"]="1"
        else: os.environ.pop("ALAN_B200_NO_QFUSE",None)
        out=T._run_paths(P,Q,sample,ip,data,names,MP())
        (lp_tc,g_tc,_,_),(lp_ff,g_ff,_,_)=out[True],out[False]
        print(shape,'qfuse',qf,'lp',rel_err(lp_tc.cpu(),lp_ff.cpu()),{k: float('%.2e'%rel_err(g_tc[k].cpu(),g_ff[k].cpu())) for k in names})
