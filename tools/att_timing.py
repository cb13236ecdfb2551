"""Development: per-stage clock totals of the attention softmax warps (needs a -DSTAD_ATT_TIMING build in STAD_LIB)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from simple_tad_b200 import _lib as L
B, H, S = 64, 12, 1568
qkv = torch.randn(B, S, 3, H, 64, device="cuda").to(torch.bfloat16)
L.attention(qkv); torch.cuda.synchronize()
lib = L.load()
buf = (C.c_ulonglong * 16)()
lib.stad_debug_read_att_timing(buf)
L.attention(qkv); torch.cuda.synchronize()
lib.stad_debug_read_att_timing(buf)
names = ["wait s_full (unpipelined)", "LDTM (unpipelined)", "max+rescale", "exps", "o_full+STTM+LDTM issue", "st wait+p_full+ld wait+s_free", "-", "loop top"]
n_sm = 148
for slot in range(2):
    tot = sum(buf[slot * 8 + i] for i in range(8))
    print(f"slot {slot}: total {tot / n_sm / 1e3:.1f} Kclk per CTA")
    for i in range(8):
        v = buf[slot * 8 + i]
        print(f"   {names[i]:32s} {v / n_sm / 1e3:9.1f} Kclk  {100.0 * v / max(tot, 1):5.1f}%   per iteration ~{v / (B * H * 7 * 13 * (6.0/7 if slot else 1)):7.1f} clk")
