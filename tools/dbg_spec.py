"""Developer diagnostic: error pattern of the attention kernel on the peaked-softmax case (per row / tile / head)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dit4sr_b200 import ops
torch.manual_seed(0)
B, H, S = 1, 2, 700
D = H * 64
qkv = (torch.randn(B, S, 3 * D, device="cuda") * 3).bfloat16()
q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
o = torch.zeros(B, S, D, device="cuda", dtype=torch.bfloat16)
ops.attention(q, k, v, o, H)
hd = lambda t: t.float().reshape(B, S, H, 64).transpose(1, 2)
ref = torch.nn.functional.scaled_dot_product_attention(hd(q), hd(k), hd(v)).transpose(1, 2).reshape(B, S, D)
e = (o.float() - ref).reshape(S, H, 64)
r = ref.reshape(S, H, 64)
rowerr = e.norm(dim=-1) / (r.norm(dim=-1) + 1e-9)          # [S, H]
tot = (e.norm() / r.norm()).item()
bad = (rowerr > 2e-2)
print(f"env SPEC={os.environ.get('DIT4SR_ATTN_SPEC')} POLY={os.environ.get('DIT4SR_ATTN_POLY')} DEBUG={os.environ.get('DIT4SR_ATTN_DEBUG')}: "
      f"rel {tot:.4e}, bad rows {int(bad.sum())}/{S * H}, finite {bool(torch.isfinite(o.float()).all())}")
for h in range(H):
    idx = bad[:, h].nonzero().flatten().tolist()
    print(f"  head {h}: bad rows (first 40) {idx[:40]}")
    if idx:
        i = idx[0]
        sc = (hd(q)[0, h, i] @ hd(k)[0, h].T) / 8.0
        tm = [round(sc[j * 128:(j + 1) * 128].max().item() * 1.4427, 1) for j in range(6)]
        print(f"    row {i}: per-tile max (log2 units) {tm}; out/ref ratio {(o.float().reshape(S,H,64)[i,h] / r[i,h])[:6].tolist()}")
