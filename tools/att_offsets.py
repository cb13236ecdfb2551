import ctypes as C, os, sys
sys.path.insert(0, os.getcwd())
import torch
from simple_tad_b200 import _lib as L
B, H, S = 64, 12, 1568
qkv = torch.randn(B, S, 3, H, 64, device="cuda").to(torch.bfloat16)
lib = L.load()
CAP = 2048
buf = (C.c_ulonglong * (4 * CAP))(); cnt = (C.c_int * 4)()
L.attention(qkv); torch.cuda.synchronize(); lib.stad_debug_read_att_trace(buf, cnt)
L.attention(qkv); torch.cuda.synchronize(); lib.stad_debug_read_att_trace(buf, cnt)
tops = {0: [], 1: []}
for role in (0, 1):
    for i in range(cnt[role]):
        v = buf[role * CAP + i]
        if (v & 0xff) == 7: tops[role].append(v >> 8)
t0 = min(tops[0][0], tops[1][0])
n = min(len(tops[0]), len(tops[1]))
print("iters", len(tops[0]), len(tops[1]))
for i in range(0, min(n, 130)):
    a, b = tops[0][i] - t0, tops[1][i] - t0
    da = tops[0][i] - tops[0][i-1] if i else 0
    print(i, a, b, "offset", b - a, "period0", da)
