import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests/golden')
import torch, cases
import stribor_b200 as st
from stribor_b200.spec import layers_from_spec
from oracle import coupling_flow_oracle as O
d,hidden,masks = 63,[192],('ordered_left_half','parity_odd')
case = cases._mk_flow('affine', d, hidden, 3, 0, 700, 700 + d + len(hidden), masks=masks, scale=1.3)()
spec=case['spec']; x=case['inputs']['x'].cuda()
def run():
    layers=[l.cuda() for l in layers_from_spec(spec)]
    flow=st.NormalizingFlow(st.UnitNormal(d), layers)
    with torch.no_grad():
        return flow.log_prob(x), *flow.inverse_and_log_det_jacobian(x)
a=run(); b=a
os.environ['STRIBOR_B200_FORCE_GENERIC']='1'; g=run()
s64=O.spec_to(spec,torch.float64)
w=(O.flow_log_prob(s64,x.cpu().double()), *O.flow_inverse(s64,x.cpu().double(),with_ldj=True))
for i,n in enumerate(('lp','xi','li')):
    print(n,'|val|max',w[i].abs().max().item(),'pipe-vs-64',(a[i].cpu().double()-w[i]).abs().max().item(),'old-vs-64',(b[i].cpu().double()-w[i]).abs().max().item(),'gen-vs-64',(g[i].cpu().double()-w[i]).abs().max().item(),'pipe-gen',(a[i]-g[i]).abs().max().item(),'old-gen',(b[i]-g[i]).abs().max().item())
