"""N>1 host logic on CPU: world_size-2 gloo.  The flat LoRA bucket aliases the
parameters and their .grad, autograd accumulates straight into it, and ONE all-reduce sums the
LoRA gradients across ranks.  (The optimiser launch itself is CUDA-only and is covered by the GPU
suite / smoke.)"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from unsloth_b200.ddp import FlatLoRABucket, init_distributed
    r, w, _ = init_distributed()
    assert (r, w) == (rank, world) and dist.get_backend() == "gloo"
    torch.manual_seed(0)                       # identical replicas
    A = torch.nn.Parameter(torch.randn(4, 10))
    B = torch.nn.Parameter(torch.randn(6, 4))
    frozen = torch.nn.Parameter(torch.randn(3), requires_grad=False)
    bucket = FlatLoRABucket([A, B, frozen])
    assert bucket.numel() == 40 + 24
    # parameters and grads are views of the flat buffers
    assert A.data_ptr() == bucket.flat_p.data_ptr() and A.grad.data_ptr() == bucket.flat_g.data_ptr()
    bucket.broadcast_params(0)
    bucket.zero_grad()
    x = torch.full((2, 10), float(rank + 1))   # different data per rank
    loss = ((x @ A.t()) @ B.t()).sum()
    loss.backward()
    assert bucket.flat_g.abs().sum() > 0       # autograd accumulated INTO the bucket
    local = bucket.flat_g.clone()
    bucket.all_reduce_grads()
    # second accumulation keeps adding into the same storage
    if rank == 0:
        torch.save({"local": local, "summed": bucket.flat_g.clone()}, out)
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    assert torch.allclose(bucket.flat_g, sum(gathered))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        bucket.step()
    dist.destroy_process_group()


def test_flat_bucket_allreduce_world2(tmp_path):
    port = _free_port()
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    d = torch.load(out)
    # rank 1 used x = 2 * rank-0's x  => grads are linear in x for A; the sum is 3x the local of rank 0 for dA
    assert torch.allclose(d["summed"][:40], 3 * d["local"][:40], rtol=1e-5, atol=1e-5)


def _dp_worker(rank, world, port, out):
    """Data-parallel equivalence of the hot path itself (through the C-ABI emulator): each rank runs
    the patched QLoRA model on ITS rows of the global batch with the loss normalised by the GLOBAL
    item count; after the one all-reduce of the flat bucket every rank holds the full-batch gradient."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import abi_emulator as EMU
    import unsloth_b200.patch as P
    from test_host_logic_cpu import TINY, attention_double
    from unsloth_b200.ddp import FlatLoRABucket, init_distributed
    mpatch = pytest.MonkeyPatch()
    EMU.install(mpatch)
    mpatch.setattr(P, "_attention", attention_double)
    torch.set_num_threads(2)
    init_distributed()
    model = P.build_qlora_model("llama-3-8b", r=4, lora_alpha=8, device="cpu", dtype=torch.float32, seed=3407,
                                num_hidden_layers=2, init_b_std=0.05, **TINY)     # identical replicas
    bucket = FlatLoRABucket(P.lora_parameters(model))
    bucket.broadcast_params(0)
    g = torch.Generator().manual_seed(99)
    ids = torch.randint(0, TINY["vocab_size"], (4, 12), generator=g)             # the GLOBAL batch
    labels = ids.clone(); labels[1, :5] = -100
    n_global = int((labels[:, 1:] != -100).sum())
    n_t = torch.tensor([int((labels[rank::world, 1:] != -100).sum())])
    dist.all_reduce(n_t)                                                          # the scalar item-count exchange
    assert int(n_t) == n_global
    bucket.zero_grad()
    loss = model(input_ids=ids[rank::world], labels=labels[rank::world], num_items_in_batch=n_global).loss
    loss.backward()
    bucket.all_reduce_grads()
    lt = loss.detach().clone()
    dist.all_reduce(lt)
    if rank == 0:
        summed = bucket.flat_g.clone()
        bucket.zero_grad()
        full = model(input_ids=ids, labels=labels, num_items_in_batch=n_global).loss
        full.backward()
        torch.save({"summed": summed, "full": bucket.flat_g.clone(), "loss_sum": lt, "loss_full": full.detach()}, out)
    dist.barrier()
    dist.destroy_process_group()
    mpatch.undo()


def test_data_parallel_equivalence_of_the_hot_path_world2(tmp_path):
    port = _free_port()
    out = str(tmp_path / "dp.pt")
    mp.spawn(_dp_worker, args=(2, port, out), nprocs=2, join=True)
    d = torch.load(out)
    assert torch.allclose(d["loss_sum"], d["loss_full"], rtol=1e-5, atol=1e-6)
    assert d["full"].abs().max() > 0
    assert torch.allclose(d["summed"], d["full"], rtol=1e-4, atol=1e-7)
