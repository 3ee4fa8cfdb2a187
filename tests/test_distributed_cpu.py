"""2-rank gloo test of the multi-GPU story on CPU: the path shards along the batch axis only, ranks
hold replicas of the encoder and exchange nothing but parameter gradients (DDP).  The sampler itself
needs a GPU, so the per-rank compute here is the oracle restatement; what is tested is the sharding
logic bench.py uses: per-rank samples, gradient averaging, max-over-ranks timing."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bevformer_b200 import synthetic as syn
    from oracle import torch_ref
    torch.set_num_threads(1)
    w = syn.WORKLOADS["toy"]
    sd = {k: v.clone().requires_grad_(True) for k, v in syn.make_state_dict(w).items()}
    inp = syn.make_encoder_inputs(w, bs=1, seed=rank)          # a different sample per rank
    o = torch_ref.encoder_forward(sd, w.num_layers, inp.bev_query, inp.feat, **inp.kwargs())
    o.square().mean().backward()
    key = "layers.0.ffns.0.layers.1.weight"
    local = sd[key].grad.clone()
    from bevformer_b200.dist import average_gradients_flat
    average_gradients_flat(list(sd.values()), world)            # bench.py's one-bucket exchange
    ms = torch.tensor([float(rank + 1)])
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)                   # bench.py's max-over-ranks timing
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    ok = torch.allclose(sd[key].grad, sum(gathered) / world, atol=1e-7) and ms.item() == world
    out[rank] = bool(ok) and not torch.equal(gathered[0], gathered[1])
    dist.destroy_process_group()


def test_two_rank_gloo_gradient_averaging():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)), dict(out)


def _arena_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bevformer_b200.arena import GradArena
    torch.manual_seed(0)
    lin_a, lin_b = torch.nn.Linear(8, 12).bfloat16(), torch.nn.Linear(12, 4).bfloat16()
    arena = GradArena([[lin_a.weight, lin_b.weight], [lin_a.bias], [lin_b.bias]])
    arena.defer_conversion = True
    # what a backward pass leaves behind: per-rank fp32 sums in the accumulator views
    g = torch.Generator().manual_seed(100 + rank)
    local = {}
    for name, p in (("aw", lin_a.weight), ("bw", lin_b.weight), ("ab", lin_a.bias), ("bb", lin_b.bias)):
        local[name] = torch.randn(p.shape, generator=g)
        arena.acc_view(p).copy_(local[name])
        arena.touched.add(id(p))
    arena._finalize()                                           # assigns p.grad views, conversion deferred
    arena.all_reduce_mean(world)                                # ONE fp32 all-reduce + one conversion
    ok = True
    for name, p in (("aw", lin_a.weight), ("bw", lin_b.weight), ("ab", lin_a.bias), ("bb", lin_b.bias)):
        gathered = [torch.zeros_like(local[name]) for _ in range(world)]
        dist.all_gather(gathered, local[name])
        want = (sum(gathered) / world).bfloat16()               # averaged in fp32, rounded once
        ok = ok and p.grad.dtype == torch.bfloat16 and torch.equal(p.grad, want)
    span = arena.span_view([lin_a.weight, lin_b.weight], (12 * 8 + 4 * 12,))
    out[rank] = bool(ok) and span is not None and arena.span_view([lin_a.weight, lin_a.bias], (108,)) is None
    dist.destroy_process_group()


def test_two_rank_gloo_arena_fp32_all_reduce():
    """The data-parallel exchange bench.py uses in graph mode: the flat fp32 gradient arena is averaged with
    one all-reduce and converted to the parameter dtype afterwards (fp32 averaging, as the reference's DDP)."""
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_arena_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)), dict(out)
