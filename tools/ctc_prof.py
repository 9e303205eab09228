"""Per-kernel CUDA-event times of the CTC head at the headline shape (B=32, T=1000, V=4337, L<=50): log-sum-exp,
alpha/beta recursion, gradient.  python tools/ctc_prof.py"""
import torch, sys, ctypes as C
sys.path.insert(0,".")
import speech_tranformer_pytorch_b200 as stb
F=stb.functional; lib=stb._lib.load()
B,T,V,L=32,1000,4337,50
x=torch.randn(B,T,V,device="cuda",requires_grad=True)
tg=torch.randint(1,V,(B,L),device="cuda"); il=torch.full((B,),T,device="cuda"); tl=torch.randint(10,L+1,(B,),device="cuda")
for _ in range(2):
    x.grad=None; F.ctc_loss(x,tg,il,tl).backward()
torch.cuda.synchronize()
lib.st_profile_reset(); lib.st_profile_enable(1)
x.grad=None; F.ctc_loss(x,tg,il,tl).backward(); torch.cuda.synchronize(); lib.st_profile_enable(0)
lib.st_profile_dump(b"gpurun_out/ctc_prof.csv")
print(open("gpurun_out/ctc_prof.csv").read())
