#!/usr/bin/env python
"""BASELINE.json configs[3]/[4]: stream a synthetic corpus sharded over the ranks into CNN / ViT training batches.

    python scripts/corpus_stream.py --clips 10000 --mode vit --batch 50
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/corpus_stream.py --clips 10000

Rank 0 prints one JSON line: whole-job seconds-of-audio/s (max over ranks of the device time), batches, label stats summed
with the path's single collective (an all-gather of 8 x int64 per rank)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "guitar-tablature-classification_b200"))
import torch
import torch.distributed as dist
from gtc_b200 import shard, streams

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=10000)
ap.add_argument("--clip-seconds", type=float, default=30.0)
ap.add_argument("--mode", default="vit", choices=["vit", "cnn"])
ap.add_argument("--batch", type=int, default=None)
ap.add_argument("--consume", action="store_true", help="touch every batch (one reduction per batch) like a model's first layer would")
ap.add_argument("--model", default=None, choices=["resnet18", "vit_s8"],
                help="train a real consumer on every batch (forward + backward + optimiser step): resnet18 = torchvision ResNet18 -> 256 -> six "
                     "19-way heads, the architecture of bestengine.py:18-48 with random weights; vit_s8 = a DINO ViT-S/8-shaped transformers.ViTModel "
                     "+ six heads (ViT_model.py:6-33)")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
batch = a.batch or (50 if a.mode == "vit" else 128)
acc = torch.zeros(1, device=dev)
def consumer(x, labels):
    acc.add_(x[:, 0, ::16, ::16].sum())

model_info = None
if a.model:
    import torch.nn as nn
    torch.manual_seed(0)
    if a.model == "resnet18":
        import torchvision
        trunk = torchvision.models.resnet18(weights=None)
        trunk.fc = nn.Linear(512, 256)
        feat = 256
    else:
        import transformers
        vit = transformers.ViTModel(transformers.ViTConfig(hidden_size=384, num_hidden_layers=12, num_attention_heads=6,
                                                           intermediate_size=1536, patch_size=8, image_size=224), add_pooling_layer=False)
        class Trunk(nn.Module):
            def __init__(self):
                super().__init__()
                self.vit = vit
            def forward(self, x):
                return self.vit(pixel_values=x).last_hidden_state[:, 0]
        trunk, feat = Trunk(), 384
    class SixHeads(nn.Module):
        def __init__(self):
            super().__init__()
            self.trunk = trunk
            self.heads = nn.ModuleList([nn.Sequential(nn.Linear(feat, 128), nn.ReLU(), nn.Linear(128, 19)) for _ in range(6)])
        def forward(self, x):
            f = self.trunk(x)
            return [h(f) for h in self.heads]
    net = SixHeads().to(dev).train()
    opt = torch.optim.Adam(net.parameters(), lr=5e-4)
    losses = []
    def consumer(x, labels):                          # what bestengine.py:899-960 / ViT_engine.py:277-333 do with a batch
        if len(x) < 2:
            return
        tgt = [labels[:, i] for i in range(6)] if a.mode == "cnn" else [torch.argmax(h, dim=1) for h in labels]
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = net(x)
            loss = sum(nn.functional.cross_entropy(o.float(), t, label_smoothing=0.05) for o, t in zip(out, tgt)) / 6
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        losses.append(loss.detach())
    a.consume = True
rep = streams.stream_corpus(a.clips, a.clip_seconds, rank, world, batch, a.mode, consumer if a.consume else None)
if a.model:
    ls = torch.stack(losses).float().cpu().numpy()
    model_info = {"model": a.model, "steps": int(len(ls)), "loss_first": float(ls[:5].mean()), "loss_last": float(ls[-5:].mean()),
                  "autocast": "bf16"}
vec = torch.tensor([rep.n_clips, rep.n_segments, int(rep.seconds_of_audio * 22050), *rep.label_stats.tolist(), rep.n_batches,
                    int(rep.device_ms * 1e6)], dtype=torch.int64, device=dev)
allv = shard.gather_stats(vec).cpu().numpy()
if rank == 0:
    tot = allv.sum(0)
    ms = allv[:, 7].max() / 1e6
    print(json.dumps({"config": f"{a.clips} clips x {a.clip_seconds:.0f} s, clip % {world} sharding, {a.mode} batches of {batch}",
                      "n_gpus": world, "clips": int(tot[0]), "segments": int(tot[1]), "batches": int(tot[6]),
                      "label_stats": {"total": int(tot[3]), "with_notes": int(tot[4]), "with_first_string": int(tot[5])},
                      "device_ms_max_over_ranks": ms, "s_audio_per_s": float(tot[2] / 22050 / (ms * 1e-3)),
                      "checksum": float(acc.item()) if a.consume else None, "consumer": model_info}))
if world > 1:
    dist.destroy_process_group()
