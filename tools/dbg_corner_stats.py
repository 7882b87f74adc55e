import sys; sys.path.insert(0,'.')
import torch, numpy as np
from karios_b200 import synth, _native as N
from karios_b200.core.configuration import KLTConfiguration
size=int(sys.argv[1]) if len(sys.argv)>1 else 4000
ref,mon=synth.make_pair(size,size,seed=1234,device='cuda')
conf=KLTConfiguration()
c=N.Context(size,size,20000)
rows=N.RowBuffers(20000,torch.device('cuda'))
kc=N.make_conf(conf,compute_zncc=True)
c.match_tile_async(mon,ref,None,(0,0,size,size),kc,rows)
st=c.read_stats()
print(st.as_dict())
