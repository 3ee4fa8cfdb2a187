"""Is the bench step bound by the host (Python enqueue) or by the GPU?  Times the enqueue of one
step (call returns, queue empty beforehand) against the step's GPU duration."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bevformer_b200 import synthetic as syn
from bevformer_b200.plugin import build_transformer_layer_sequence

dev = torch.device("cuda")
w = syn.WORKLOADS["base"]
enc = build_transformer_layer_sequence(syn.encoder_cfg(w))
enc.load_state_dict(syn.make_state_dict(w))
enc = enc.to(dev, torch.bfloat16).train()
host = syn.make_encoder_inputs(w)
inp = {k: getattr(host, k).to(dev, torch.bfloat16) for k in ("bev_query", "feat", "bev_pos", "prev_bev")}
shift, ss, lsi = host.shift.to(dev), host.spatial_shapes.to(dev), host.level_start_index.to(dev)
proj = torch.randn(1, w.num_query, 256, device=dev, dtype=torch.bfloat16)

def step():
    bq = inp["bev_query"].detach().requires_grad_(True); ft = inp["feat"].detach().requires_grad_(True)
    for p in enc.parameters(): p.grad = None
    out = enc(bq, ft, ft, bev_h=w.bev_h, bev_w=w.bev_w, bev_pos=inp["bev_pos"], spatial_shapes=ss,
              level_start_index=lsi, prev_bev=inp["prev_bev"], shift=shift, img_metas=host.img_metas)
    (out * proj).sum().backward()

for _ in range(3): step()
torch.cuda.synchronize()
enq, tot = [], []
for _ in range(5):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(True), torch.cuda.Event(True)
    t0 = time.perf_counter(); s.record(); step(); e.record(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    enq.append((t1 - t0) * 1e3); tot.append((t2 - t0) * 1e3)
    gpu = s.elapsed_time(e)
print("enqueue ms", [round(x, 1) for x in enq], "wall ms", [round(x, 1) for x in tot], "gpu-span ms", round(gpu, 1))
