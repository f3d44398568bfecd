"""Multi-GPU parity of Ulysses sequence parallelism (torchrun, N ranks): the tiny Wan VAP model (one MoT + one plain block, or
more heads with --heads) run token-sharded over N GPUs — in both transports, "p2p" (exchange fused into the kernels over NVLink
peer memory) and "nccl" — must match the same model run on one GPU.  Prints one JSON line on rank 0; exit code != 0 on mismatch.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/ulysses_multi_gpu_check.py
"""
import argparse, importlib, json, os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
vap = importlib.import_module("video-as-prompt_b200")
ap = argparse.ArgumentParser(); ap.add_argument("--heads", type=int, default=8); ap.add_argument("--frames", type=int, default=4)
ap.add_argument("--family", default="wan", choices=["wan", "cog"], help="cog: tiny CogVideoX VAP model, B = 2 (the CFG batch), rows of each "
                "stream's [text | video] sequence sharded across the ranks (rank 0 holds the text rows)")
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
if a.family == "wan":
    cfg = dict(vap.synth.WAN_TINY, num_attention_heads=a.heads, added_kv_proj_dim=a.heads * 128, num_layers=3, block_idx_with_mot_ref=[0, 2])
    with torch.device("meta"):
        model = vap.WanTransformer3DMOTModel(**cfg)
else:
    cfg = dict(vap.synth.COG_TINY, num_attention_heads=a.heads, num_layers=3, block_idx_with_mot_ref=[0, 2])
    with torch.device("meta"):
        model = vap.CogVideoXTransformer3DMOTModel(**cfg)
model = model.to(torch.bfloat16).to_empty(device=dev)
vap.synth.fill_module_(model, seed=1234, num_layers=cfg["num_layers"])
model.eval()
if a.family == "wan":
    inp = vap.synth.wan_inputs(cfg, a.frames, 16, 8 * world, seed=0, device=dev)  # tokens per stream = frames * 8 * 4 * world
else:  # 226 text + 2 * 3 * 9 = 54 video tokens per stream = 280 rows: divisible by 2, 4 and 8 ranks; rank 0 (and 1 at 8 ranks...) hold text rows
    inp = vap.synth.cog_inputs(cfg, 2, 6, 18, seed=0, device=dev, batch=2)
res = {}
with torch.no_grad():
    ref = model(**inp, return_dict=False)[0].float()
    for mode in ("p2p", "nccl"):
        vap.ulysses.enable(mode=mode)
        outs = [model(**inp, return_dict=False)[0].float() for _ in range(3)]  # 3 forwards: exercises the alternating buffer sets
        vap.ulysses.disable()
        err = max(((o - ref).abs().max() / ref.abs().max()).item() for o in outs)
        res[mode] = err
    if a.family == "wan":  # a B = 2 batch (wan_denoise(batch_cfg=True)): one exchange and one attention launch for both samples in "p2p" mode
        inp2 = vap.synth.wan_inputs(cfg, a.frames, 16, 8 * world, seed=3, device=dev, batch=2)
        ref2 = model(**inp2, return_dict=False)[0].float()
        for mode in ("p2p", "nccl"):
            vap.ulysses.enable(mode=mode)
            outs = [model(**inp2, return_dict=False)[0].float() for _ in range(2)]
            vap.ulysses.disable()
            res[mode + "_batch2"] = max(((o - ref2).abs().max() / ref2.abs().max()).item() for o in outs)
ok = all(e < 2e-2 for e in res.values())
t = torch.tensor([0 if ok else 1], device=dev); dist.all_reduce(t)
if rank == 0:
    print(json.dumps(dict(world=world, heads=a.heads, rel_err_vs_single_gpu=res, ok=bool(t.item() == 0))), flush=True)
dist.destroy_process_group()
sys.exit(0 if t.item() == 0 else 1)
