import sys, torch
sys.path.insert(0, '/root/repo')
from easyanimate_b200 import ops
torch.manual_seed(0)
B,H,S,St = 1,4,4096+80,256
for name, scale in (("randn",1.0),("peaked",4.0)):
    q = (torch.randn(B,H,S,64,device="cuda")*scale).to(torch.bfloat16)
    k = torch.randn(B,H,S,64,device="cuda").to(torch.bfloat16)
    v = torch.randn(B,H,S,64,device="cuda").to(torch.bfloat16)
    ref = torch.nn.functional.scaled_dot_product_attention(q.double(),k.double(),v.double()).transpose(1,2).reshape(B,S,H*64)
    for variant in (0x10c, 0x100c, 0x101c, 0x102c, 0x90c, 0x1c):
        ot, ov = ops.attention(q,k,v,St,variant=variant)
        got = torch.cat([ot,ov],1).double()
        d = got-ref
        print(name, hex(variant), "rel_rms_err %.3e" % (d.norm()/ref.norm()).item(), "mean_signed_rel %.3e" % ((d*ref).sum()/(ref*ref).sum()).item(), flush=True)
