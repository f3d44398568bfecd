import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
vap = importlib.import_module("video-as-prompt_b200"); ops = vap.ops
torch.set_printoptions(precision=4, linewidth=200, sci_mode=False)
for D in (128, 64):
    for Lkv in (64, 128, 256):
        Lq = 128
        q = torch.zeros((1, 1, Lq, D), dtype=torch.bfloat16, device="cuda")
        k = torch.randn((1, 1, Lkv, D), device="cuda").to(torch.bfloat16)
        v = torch.zeros((1, 1, Lkv, D), dtype=torch.bfloat16, device="cuda")
        for key in range(Lkv):
            v[0, 0, key, key // 16] = 1.0
        o = ops.attention(q, k, v)
        torch.cuda.synchronize()
        print(f"D={D} Lkv={Lkv} expect {16.0 / Lkv:.4f} on the first {Lkv // 16} dims")
        print(" row0  ", o[0, 0, 0, :20].float().cpu())
        print(" row77 ", o[0, 0, 77, :20].float().cpu())
    # second probe: V[key, d] = key (all dims) with q=0 -> mean key index
    Lkv = 128
    v = torch.arange(Lkv, device="cuda", dtype=torch.float32).view(1, 1, Lkv, 1).expand(1, 1, Lkv, D).to(torch.bfloat16).contiguous()
    o = ops.attention(torch.zeros((1, 1, 128, D), dtype=torch.bfloat16, device="cuda"), torch.randn((1, 1, Lkv, D), device="cuda").to(torch.bfloat16), v)
    print(f"D={D} mean-key probe expect 63.5:", o[0, 0, 0, :4].float().cpu(), o[0, 0, 100, -4:].float().cpu())
