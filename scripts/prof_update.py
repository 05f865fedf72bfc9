import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributional_rl_navigation_b200.iqn_agent import IQNAgent
B=1024
a=IQNAgent(26,9,device="cuda:0",seed=0,BATCH_SIZE=B)
g=torch.Generator(device="cuda"); g.manual_seed(0)
st=torch.randn(B,26,device="cuda",generator=g)*3; ns=torch.randn(B,26,device="cuda",generator=g)*3
ac=torch.randint(0,9,(B,),device="cuda",generator=g); rw=torch.randn(B,device="cuda",generator=g); dn=(torch.rand(B,device="cuda",generator=g)<0.05).float()
tt=torch.rand(B,8,device="cuda",generator=g); tl=torch.rand(B,8,device="cuda",generator=g)
for _ in range(5): a.train_async((st,ac,rw,ns,dn),(tt,tl))
torch.cuda.synchronize()
