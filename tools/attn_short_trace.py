"""clock64() timeline of CTA (0,0,0) of the short-KV attention kernel on a Wan cross-attention shape (needs a -DVAP_ATTN_TRACE=1 build:
python tools/build_attn_variants.py trace:-DVAP_ATTN_TRACE=1; VAP_B200_LIB=build_variants/libvap_trace.so python tools/attn_short_trace.py [Lkv])."""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
vap = importlib.import_module("video-as-prompt_b200"); ops = vap.ops
lib = vap._lib.load()
H, Lq, D = 40, 20280, 128
Lkv = int(sys.argv[1]) if len(sys.argv) > 1 else 512
q = torch.randn((1, Lq, H * D), device="cuda").to(torch.bfloat16).unflatten(2, (H, D)).transpose(1, 2)
kv = torch.randn((1, Lkv, 2 * H * D), device="cuda").to(torch.bfloat16)
k, v = (kv[..., i * H * D:(i + 1) * H * D].unflatten(2, (H, D)).transpose(1, 2) for i in range(2))
os.environ["VAP_ATTN_SHORT"] = "1"
for _ in range(2):
    ops.attention(q, k, v)
buf = torch.zeros(3 * 64 * 8, dtype=torch.int64, device="cuda")
lib.vap_debug_set_attention_trace(buf.data_ptr())
ops.attention(q, k, v)
torch.cuda.synchronize()
lib.vap_debug_set_attention_trace(0)
t = buf.cpu()
t0 = t[1024 + 6].item()
rel = lambda x: int(x - t0) if x > 0 else None
print(f"short kernel, Lkv={Lkv}: cycles since kernel entry of CTA (0,0,0)")
print("  set-up done (barriers, TMEM alloc, __syncthreads)", rel(t[1024 + 7].item()))
n_kv = (Lkv + 127) // 128
for j in range(n_kv):
    row = t[j * 8:(j + 1) * 8].tolist()
    print(f"  step {j}: start {rel(row[0])}  S seen {rel(row[1])}  scores loaded {rel(row[2])}  half 0 done {rel(row[3])}  half 1 stored {rel(row[5])}  published {rel(row[6])}")
print("  last P published", rel(t[1024 + 14].item()), " O complete", rel(t[1024 + 15].item()), " O stored", rel(t[1024 + 22].item()), " exit", rel(t[1024 + 23].item()))
