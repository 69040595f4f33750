"""Timeline of the dK/dV attention kernel (MVF_FA_DBG=1): one backward at S = 3840 through the C ABI."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MVF_FA_DBG"] = "1"
import torch
from video_rep_learning_b200 import _lib as L
B, S, heads, dk = 8, 3840, 8, 32
Hd = heads * dk
g = torch.Generator(device="cuda").manual_seed(1)
qkv = torch.randn(B * S, 3 * Hd, device="cuda", generator=g) * 0.7
ctx = torch.empty(B * S, Hd, device="cuda"); lse = torch.empty(B, heads, S, device="cuda")
nb = L.lib().mvf_attention_ws_bytes(B, S, heads, dk)
ws = torch.empty(nb + 1024, dtype=torch.uint8, device="cuda"); wsp = (ws.data_ptr() + 1023) // 1024 * 1024
st = torch.cuda.current_stream().cuda_stream
L.check(L.lib().mvf_attention_fwd(0, B, S, heads, dk, qkv.data_ptr(), None, ctx.data_ptr(), lse.data_ptr(), wsp, nb, st))
d_ctx = torch.randn(B * S, Hd, device="cuda", generator=g); d_qkv = torch.empty_like(qkv); delta = torch.empty(B, heads, S, device="cuda")
for _ in range(2):
    L.check(L.lib().mvf_attention_bwd(0, B, S, heads, dk, qkv.data_ptr(), None, ctx.data_ptr(), lse.data_ptr(), d_ctx.data_ptr(), d_qkv.data_ptr(), delta.data_ptr(), wsp, nb, st))
torch.cuda.synchronize()
