"""Would a bf16 residual stream in the decoder keep the 1e-2 waveform tolerance?  CPU simulation on the oracle: the
generator with bf16-rounded conv operands (what the tensor-core path feeds its MMAs) and, for the chosen stages, the
residual stream x rounded to bf16 after the ConvTranspose and after every `x = xt + x` (models/convnext_utils.py:112),
against the fp32 oracle.  Result (profiles/r2_bf16_residual_simulation.txt): the operand rounding alone costs
5.5e-3 (W0) / 6.4e-3 (W1) of the waveform's range; a bf16 residual in ANY stage set brings it to 0.8-1.0e-2, i.e. to
the tolerance itself — rejected.   usage: python scripts/sim_bf16_residual.py"""
import sys, torch, torch.nn.functional as F
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from oracle import restatement as R, weights
from tests.golden.inputs import make_latents
torch.set_num_threads(8)
bf=lambda t: t.to(torch.bfloat16).float()
def gen(sd, z, res_bf16_stages=(), operands_bf16=True, rates=(8,4,2,2,2), ks=(16,12,4,4,4)):
    p='generator.'
    W=lambda q: bf(R.weight_norm_weight(sd,q)) if operands_bf16 else R.weight_norm_weight(sd,q)
    A=lambda t: bf(t) if operands_bf16 else t
    x=F.conv1d(A(z), W(p+'conv_pre.'), sd[p+'conv_pre.bias'], padding=6)
    for i,(u,k) in enumerate(zip(rates,ks)):
        x=F.conv_transpose1d(A(F.silu(x)), W(p+f'ups.{i}.'), sd[p+f'ups.{i}.bias'], stride=u, padding=(k-u)//2)
        rb = i in res_bf16_stages
        if rb: x=bf(x)
        outs=[]
        for b,kk in enumerate((3,7,11)):
            xb=x
            for n,d in enumerate((1,3,5)):
                q=p+f'resblocks.{i}.blocks.{b}.'
                xt=F.conv1d(A(F.silu(xb)), W(q+f'convs1.{n}.'), sd[q+f'convs1.{n}.bias'], dilation=d, padding=(kk*d-d)//2)
                xt=F.conv1d(A(F.silu(xt)), W(q+f'convs2.{n}.'), sd[q+f'convs2.{n}.bias'], padding=(kk-1)//2)
                xb=xt+xb
                if rb: xb=bf(xb)
            outs.append(xb)
        x=torch.stack(outs,0).mean(0)
    x=F.conv1d(A(F.silu(x)), R.weight_norm_weight(sd,p+'conv_post.'), sd[p+'conv_post.bias'], padding=6)
    return torch.tanh(x)
rel=lambda a,b: float((a-b).abs().max()/b.abs().max())
with torch.no_grad():
    for variant in ('W0','W1'):
        sd=weights.make_state_dict(variant, codebook_size=64)
        z=make_latents(2,48,seed=31)*0.5
        ref=R.generator_forward(sd,z)
        for stages in ((),(4,),(3,4),(2,3,4),(1,2,3,4),(0,1,2,3,4)):
            out=gen(sd,z,stages)
            print(variant,'bf16 residual in stages',stages,'rel err vs fp32 oracle %.3e'%rel(out,ref))
