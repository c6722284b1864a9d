import sys, torch
sys.path.insert(0, '/root/repo')
from xfm_b200 import lib
from xfm_b200.encoders import closed_form_rel_index
B,H,ws=int(sys.argv[1]),int(sys.argv[2]),int(sys.argv[3])
L,D=ws*ws+1,H*64
g=torch.Generator().manual_seed(1)
qkv=(torch.randn(B*L,3*D,generator=g)*0.7).bfloat16()
dout=torch.randn(B*L,D,generator=g).bfloat16()
f=qkv.float().view(B,L,3,H,64).permute(2,0,3,1,4)
qf,kf,vf=(t.clone().requires_grad_(True) for t in (f[0],f[1],f[2]))
T=(2*ws-1)**2+3
table=torch.randn(T,H,generator=g).requires_grad_(True)
idx=closed_form_rel_index(ws)
s=(qf*0.125)@kf.transpose(-1,-2)+table[idx.view(-1)].view(L,L,H).permute(2,0,1)
ref=torch.softmax(s,-1)@vf
ref.backward(dout.float().view(B,L,H,64).permute(0,2,1,3))
want=torch.stack([qf.grad,kf.grad,vf.grad]).permute(1,3,0,2,4).reshape(B*L,3*D)
c,do=qkv.cuda(),dout.cuda()
q,k,v=c[:,:D],c[:,D:2*D],c[:,2*D:]
tdev=table.detach().cuda()
out,lse=lib.attention_fwd(q,k,v,B,H,L,L,0.125,rel_table=tdev,rel_window=ws)
dqkv=torch.full_like(c,float('nan'))
dtab=torch.zeros_like(tdev)
lib.attention_bwd(do,q,k,v,out,lse,B,H,L,L,0.125,dqkv[:,:D],dqkv[:,D:2*D],dqkv[:,2*D:],rel_table=tdev,rel_window=ws,rel_dtable=dtab)
torch.cuda.synchronize()
got=dqkv.float().cpu()
for name,sl in (('dq',slice(0,D)),('dk',slice(D,2*D)),('dv',slice(2*D,3*D))):
    gg=got[:,sl]; ww=want[:,sl]
    bad=~torch.isfinite(gg)
    print(name,'nonfinite',int(bad.sum()),'of',gg.numel())
    if bad.any():
        rows=bad.any(1).nonzero().flatten(); cols=bad.any(0).nonzero().flatten()
        print('  rows',rows[:10].tolist(),'...',rows[-5:].tolist(),'n',len(rows),' cols',cols[:6].tolist(),'..',cols[-3:].tolist(),'n',len(cols))
    ok=~bad
    print('  max err on finite', float((gg[ok]-ww[ok]).abs().max()), 'scale', float(ww.abs().max()))
    e=(gg-ww).abs(); e[bad]=0
    r=e.max(1).values
    print('  worst rows', torch.topk(r,5).indices.tolist(), torch.topk(r,5).values.tolist())
print('dtab err', float((dtab.cpu()-table.grad).abs().max()), float(table.grad.abs().max()))
