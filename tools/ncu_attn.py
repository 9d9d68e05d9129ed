"""One self-attention launch (for `ncu --set full --import-source on`): batch 16, 5 heads, L = 4096."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tango_b200 import lib as L
dev = torch.device("cuda:0")
B, heads, Lq = 16, 5, 4096
Cc = heads * 64
qkv = torch.randn(B * Lq, 3 * Cc, device=dev).to(torch.bfloat16)
out = torch.empty(B * Lq, Cc, device=dev, dtype=torch.bfloat16)
for _ in range(2):
    L.attention(qkv, qkv, qkv, out, batch=B, heads=heads, Lq=Lq, Lk=Lq, scale=0.125, k_col0=Cc, v_col0=2 * Cc)
torch.cuda.synchronize()
