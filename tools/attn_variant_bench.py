"""Time one attention-kernel build (VAP_B200_LIB) on the Wan / CogVideoX joint shapes and check it against torch SDPA."""
import importlib, json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
vap = importlib.import_module("video-as-prompt_b200"); ops = vap.ops
out = {"lib": os.path.basename(os.environ.get("VAP_B200_LIB", "default"))}
for H, J, D in [(40, 40560, 128), (48, 35552, 64)]:
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = torch.randn((1, J, 3 * H * D), generator=g, device="cuda").to(torch.bfloat16)
    q, k, v = (qkv[..., i * H * D:(i + 1) * H * D].unflatten(2, (H, D)).transpose(1, 2) for i in range(3))
    for _ in range(3):
        o = ops.attention(q, k, v)
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); o = ops.attention(q, k, v); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ref = torch.nn.functional.scaled_dot_product_attention(q[:, :2, :4096], k[:, :2], v[:, :2])
    err = (o[:, :2, :4096].float() - ref.float()).abs().max().item() / ref.float().abs().max().item()
    ms = min(ts)
    out[f"D{D}"] = dict(ms=round(ms, 3), tflops=round(4.0 * H * J * J * D / ms / 1e9, 1), err=round(err, 5))
print(json.dumps(out), flush=True)
