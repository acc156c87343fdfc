"""Bring-up helper: run each attention operand-path variant in its own process (a trap kills the CUDA context)."""
import subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, torch
sys.path.insert(0, %r)
from easyanimate_b200 import ops
variant = int(sys.argv[1])
torch.manual_seed(0)
for (B,H,S,St) in [(1,2,128,0),(1,2,256,64),(2,3,1000,77),(1,4,4096+80,256)]:
    q = torch.randn(B,H,S,64,device="cuda").to(torch.bfloat16)
    k = torch.randn(B,H,S,64,device="cuda").to(torch.bfloat16)
    v = torch.randn(B,H,S,64,device="cuda").to(torch.bfloat16)
    ot, ov = ops.attention(q,k,v,St,variant=variant)
    torch.cuda.synchronize()
    ref = torch.nn.functional.scaled_dot_product_attention(q.float(),k.float(),v.float()).transpose(1,2).reshape(B,S,H*64)
    got = torch.cat([ot,ov],1).float()
    err = (got-ref).abs().max().item()
    print("variant",variant,"shape",(B,H,S,St),"max_abs_err",round(err,5), "ref_absmax", round(ref.abs().max().item(),3), flush=True)
''' % ROOT
VARIANTS = [int(v, 0) for v in os.environ.get('EA_VARIANTS', '1,0x1c,0x10c,0x100c').split(',')]
for variant in VARIANTS:
    try:
        r = subprocess.run([sys.executable, "-c", CODE, str(variant)], capture_output=True, text=True, timeout=120)
        print(r.stdout.strip()); 
        if r.returncode != 0: print("variant", variant, "FAILED rc", r.returncode, r.stderr.strip()[-600:])
    except subprocess.TimeoutExpired:
        print("variant", variant, "TIMEOUT")
