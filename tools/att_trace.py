"""Development: event trace of CTA 0 of the attention kernel (needs a -DSTAD_ATT_TRACE build in STAD_LIB).
Prints, per role, the events of a few steady-state iterations with clock deltas."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from simple_tad_b200 import _lib as L
B, H, S = (int(x) for x in os.environ.get('ATT_SHAPE', '64,12,1568').split(','))
qkv = torch.randn(B, S, 3, H, 64, device="cuda").to(torch.bfloat16)
lib = L.load()
CAP = 2048
buf = (C.c_ulonglong * (6 * CAP))(); cnt = (C.c_int * 6)()
L.attention(qkv); torch.cuda.synchronize(); lib.stad_debug_read_att_trace(buf, cnt)
L.attention(qkv); torch.cuda.synchronize(); lib.stad_debug_read_att_trace(buf, cnt)
TAGS = {7: "tile top", 0: "s_full ok", 1: "S in regs", 3: "exps+st issued", 5: "p_full arrived",
        20: "mma: k_full ok", 21: "mma: QK issued", 22: "mma: v_full ok", 23: "mma: p_full ok", 24: "mma: PV issued", 30: "S buf0 complete", 31: "S buf1 complete", 39: "epi: unit top", 40: "epi: l_ready ok", 41: "epi: o_done ok", 42: "epi: o_free arrived"}
ev = []
for role in ([2] if len(sys.argv) > 3 else range(6)):
    for i in range(cnt[role]):
        v = buf[role * CAP + i]
        ev.append((v >> 8, role, v & 0xff))
ev.sort()
t0 = ev[0][0]
names = ["softmax0", "softmax1", "mma0", "mma1", "poll", "epi"]
lo, hi = int(sys.argv[1]) if len(sys.argv) > 1 else 60000, int(sys.argv[2]) if len(sys.argv) > 2 else 75000
last = {}
for t, role, tag in ev:
    rel = t - t0
    d = rel - last.get(role, rel)
    last[role] = rel
    if lo <= rel <= hi:
        print(f"{rel:8d}  {'                    ' * role}{names[role]}: {TAGS.get(tag, tag)} (+{d})")
