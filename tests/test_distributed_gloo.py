"""CPU, world_size 2 over gloo: the host logic of the one exchange step on the path —
`gather_without_grad` must return the rank-major concatenation (= the reference's all_gather + cat,
criterions/image_text_retrieval_loss.py:29-38), detached, and the per-rank targets must be `i + rank*b`, so that the
sharded oracle loss equals the single-process loss over the concatenated batch."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import restated as R
import synth


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, b, d, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from one_peace_b200.criterions.image_text_retrieval_loss import gather_without_grad
    a_all, t_all = synth.contrastive_pair(b * world, d, seed=77)
    a = a_all[rank * b:(rank + 1) * b].clone().requires_grad_(True)
    t = t_all[rank * b:(rank + 1) * b].clone().requires_grad_(True)
    ga, gt = gather_without_grad(a), gather_without_grad(t)
    assert not ga.requires_grad and ga.shape == (world * b, d)
    assert torch.equal(ga, a_all) and torch.equal(gt, t_all)          # rank-major order
    loss, i2t, t2i = R.itc_loss(a, t, ga, gt, torch.tensor(1 / 0.07), rank=rank)
    loss.backward()
    torch.save(dict(loss=loss.detach(), grad=a.grad.clone(), i2t=i2t, t2i=t2i), os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_and_sharded_targets_gloo(tmp_path):
    world, b, d = 2, 6, 32
    port = _free_port()
    mp.spawn(_worker, args=(world, port, b, d, str(tmp_path)), nprocs=world, join=True)
    a_all, t_all = synth.contrastive_pair(b * world, d, seed=77)
    a_all = a_all.requires_grad_(True)
    full, fi, ft = R.itc_loss(a_all, t_all, a_all.detach(), t_all, torch.tensor(1 / 0.07), rank=0)
    full.backward()
    outs = [torch.load(tmp_path / f"r{r}.pt", weights_only=False) for r in range(world)]
    # mean over ranks of the per-rank mean losses == loss over the concatenated batch (equal shard sizes)
    torch.testing.assert_close(sum(o["loss"] for o in outs) / world, full.detach(), atol=1e-6, rtol=1e-6)
    assert sum(float(o["i2t"]) for o in outs) == float(fi) and sum(float(o["t2i"]) for o in outs) == float(ft)
    # local-rows-only gradient: each rank's gradient is the matching slice of the single-process one scaled by W
    # (mean over b local rows vs mean over W*b rows)
    for r, o in enumerate(outs):
        torch.testing.assert_close(o["grad"] / world, a_all.grad[r * b:(r + 1) * b], atol=1e-6, rtol=1e-5)
