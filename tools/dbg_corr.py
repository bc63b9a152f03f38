import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ofr_b200 as ofr, oracle
dev='cuda:0'
shapes = [(2,196,4,4),(2,128,8,8),(2,96,16,16),(1,64,32,32),(1,32,64,64),(2,32,128,128)]
for (B,C,H,W) in shapes:
    torch.manual_seed(0)
    f1=torch.randn(B,C,H,W); f2=torch.randn(B,C,H,W)
    a=f1.to(dev).requires_grad_(True); b=f2.to(dev).requires_grad_(True)
    for stage in ('fwd','g1','g2'):
        try:
            if stage=='fwd':
                out=ofr.Correlation(4,1,4,1,1,1)(a,b); torch.cuda.synchronize()
                ref=oracle.correlation(f1.double(),f2.double(),4,1,4,1,1)
                print(B,C,H,W,'fwd err',(out.cpu().double()-ref).abs().max().item(), flush=True)
            elif stage=='g1':
                go=torch.randn_like(out)
                (g1,)=torch.autograd.grad(out,(a,),go,retain_graph=True); torch.cuda.synchronize()
                print('   g1 ok', flush=True)
            else:
                (g2,)=torch.autograd.grad(out,(b,),go); torch.cuda.synchronize()
                print('   g2 ok', flush=True)
        except Exception as e:
            print(B,C,H,W,stage,'FAILED',str(e)[:100], flush=True); sys.exit(1)
