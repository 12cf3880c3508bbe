import sys, torch
sys.path.insert(0, '.')
from surya_b200 import ops
B,H,cin,cout=32,512,32,32
x=torch.randn(B,H,H,cin,device='cuda').half()
w=(torch.randn(cout,9*cin,device='cuda')*(9*cin)**-0.5).half()
b=torch.randn(cout,device='cuda')*0.1
for _ in range(2):
    ops.conv2d_nhwc(x,w,b,None,3,1,1,'hardswish')
torch.cuda.synchronize()
