"""Random sweep of explicit compression parameters: kernel source on the host (or on the 32-lane emulator) vs the compiled
reference.  usage: python scripts/cparam_sweep.py [seed] [count] [emu]   (dev container only: needs oracle/_ref)"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import random
import numpy as np
from tests import cases
from tests.oracle_util import ref_compress_params, hostsim_compress_params
from zstd_jni_b200 import corpus
random.seed(int(sys.argv[1]) if len(sys.argv)>1 else 1)
inputs=[("c%d"%i, corpus.chunk(i).tobytes()) for i in (0,1,2,3,4,5)]
inputs+= [("c1-20k", corpus.chunk(1)[:20000].tobytes()), ("c0-10k", corpus.chunk(0)[:10000].tobytes()), ("c3-70k", corpus.chunk(3)[:70000].tobytes()), ("c4-1000", corpus.chunk(4)[:1000].tobytes())]
bad=0; tot=0; unsup=0
N=int(sys.argv[2]) if len(sys.argv)>2 else 150
for t in range(N):
    name,data=random.choice(inputs)
    level=random.choice([1,2,3,4,5,6,7,9,10,12,-3])
    params={}
    for k,rng in (("windowLog",(10,27)),("hashLog",(6,22)),("chainLog",(6,22)),("searchLog",(1,9)),("minMatch",(3,7)),("targetLength",(0,200)),("strategy",(1,6))):
        if random.random()<0.4: params[k]=random.randint(*rng)
    if not params: params={"hashLog":random.randint(6,20)}
    exp=ref_compress_params(data,level,params)
    got=hostsim_compress_params(data,level,params,emu=len(sys.argv)>3)
    tot+=1
    if isinstance(got,int) and got==-40:
        unsup+=1
        # should only be when window < src
        wl=params.get("windowLog")
        if not (wl and (1<<wl)<len(data)) and not (len(data)<=16384 and level>10): print("UNSUP?",name,level,params)
        continue
    if got!=exp:
        bad+=1; print("MISMATCH",name,len(data),level,params, exp if isinstance(exp,int) else len(exp), got if isinstance(got,int) else len(got))
print("total",tot,"bad",bad,"unsupported",unsup)
