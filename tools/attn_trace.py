"""Print the clock64() trace of CTA (0,0,0) of the attention kernel (debug aid; see vap_debug_set_attention_trace)."""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
vap = importlib.import_module("video-as-prompt_b200"); ops = vap.ops
lib = vap._lib.load()
H, J, D = 16, 8192, int(sys.argv[1]) if len(sys.argv) > 1 else 128
qkv = torch.randn((1, J, 3 * H * D), device="cuda").to(torch.bfloat16)
q, k, v = (qkv[..., i * H * D:(i + 1) * H * D].unflatten(2, (H, D)).transpose(1, 2) for i in range(3))
for _ in range(2):
    ops.attention(q, k, v)
buf = torch.zeros(3 * 64 * 8, dtype=torch.int64, device="cuda")
lib.vap_debug_set_attention_trace(buf.data_ptr())
ops.attention(q, k, v)
torch.cuda.synchronize()
lib.vap_debug_set_attention_trace(0)
t = buf.cpu().view(3, 64, 8)
t0 = t[0, 4, 0].item()
names = ["softmax tile0", "softmax tile1", "mma issuer"]
for r in range(3):
    print(names[r], "(cycles relative to tile0 iter4 start; per-iteration stamps)")
    for j in range(4, 14):
        print(f"  j={j:2d} ", " ".join(f"{(x - t0):7d}" for x in t[r, j].tolist() if x > 0))
sm = t[0, 4:60]
d = (sm[1:, 0] - sm[:-1, 0]).float()
print("softmax tile0 iteration period: mean %.0f min %.0f max %.0f cycles" % (d.mean(), d.min(), d.max()))
seg = ["wait S", "tmem ld", "mask+max(+rescale)", "exp+pack+st issue", "(unused)", "st wait", "o_done+arrive"]
cur = t[0, 4:60]
for a, b, nm in [(0, 1, "wait s_full"), (1, 2, "tmem ld"), (2, 3, "max/rescale"), (3, 5, "exp + P st issue"), (5, 6, "st wait + o_done + arrive")]:
    print(f"  {nm:28s} {(cur[:, b] - cur[:, a]).float().mean():8.0f}")
mm = t[2, 4:60]
for a, b, nm in [(0, 1, "wait V/K full"), (1, 2, "wait P0"), (2, 3, "issue PV0+QK0"), (3, 4, "wait P1"), (4, 5, "issue PV1+QK1")]:
    print(f"  mma {nm:24s} {(mm[:, b] - mm[:, a]).float().mean():8.0f}")
