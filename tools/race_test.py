import sys; sys.path.insert(0,'.')
import numpy as np, torch, os
from pymc_statespace_b200 import _lib
if os.environ.get('KFB_LIB'): _lib.LIB_PATH=os.environ['KFB_LIB']
from tests.test_gpu_parity import run_single
from tests.helpers import random_system
from oracle import kalman_numpy as kn
m,p,r=30,1,3
rng=np.random.default_rng(3)
args=random_system(rng,m,p,r,12,n_missing=1,scale_T=0.1)
res,g,info=run_single("univariate",args,bwd=False)
ref=kn.kalman_filter("univariate",*args)
print("info",info,res[4],ref[4])
