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


# ------------------------------------------------------------------------------------------------------------------
# ZeRO-1 sharded optimizer step (optim/distributed_adam.py): partition arithmetic + the exchange pattern
# (reduce-scatter of the flat gradient, shard update, all-gather of the flat parameters) on CPU over gloo, with the
# oracle's adam_step standing in for the sm_100a kernel.  Must equal plain Adam on the rank-averaged gradients.
# ------------------------------------------------------------------------------------------------------------------
def test_shard_layout_covers_every_element_once():
    from one_peace_b200.optim.distributed_adam import shard_layout, shard_segments
    numels = [7, 64, 1, 1000, 33, 8]
    for world in (1, 2, 3, 8):
        offsets, total, shard = shard_layout(numels, world)
        assert total == shard * world and shard % 8 == 0 and all(o % 8 == 0 for o in offsets)
        seen = [torch.zeros(n, dtype=torch.int32) for n in numels]
        for r in range(world):
            for pi, start, ln, so in shard_segments(offsets, numels, r * shard, (r + 1) * shard):
                assert 0 <= so and so + ln <= shard and (offsets[pi] + start) - r * shard == so
                seen[pi][start:start + ln] += 1
        assert all(bool((s == 1).all()) for s in seen)


def _zero_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from one_peace_b200.optim.distributed_adam import shard_layout, shard_segments
    g = torch.Generator().manual_seed(5)
    shapes = [(13,), (4, 9), (1,), (50, 3)]
    params = [torch.randn(s, generator=g) for s in shapes]
    grads = [torch.randn(s, generator=torch.Generator().manual_seed(100 + rank * 10 + i)) for i, s in enumerate(shapes)]
    numels = [p.numel() for p in params]
    offsets, total, shard = shard_layout(numels, world)
    lo, hi = rank * shard, (rank + 1) * shard
    flat_p, flat_g = torch.zeros(total), torch.zeros(total)
    for p, gr, off in zip(params, grads, offsets):
        flat_p[off:off + p.numel()] = p.reshape(-1)
        flat_g[off:off + p.numel()] = gr.reshape(-1)
    # reduce-scatter (mean) — gloo has no reduce_scatter_tensor: all_reduce + slice is the same exchange result
    dist.all_reduce(flat_g)
    gsh = flat_g[lo:hi] / world
    psh, m, v = flat_p[lo:hi].clone(), torch.zeros(shard), torch.zeros(shard)
    for pi, start, ln, so in shard_segments(offsets, numels, lo, hi):
        sl = slice(so, so + ln)
        R.adam_step(psh[sl], gsh[sl], m[sl], v[sl], 1, 1e-2, 0.9, 0.98, 1e-8, 0.05)      # in place on the shard views
    out = [torch.zeros(shard) for _ in range(world)]
    dist.all_gather(out, psh)
    flat_new = torch.cat(out)
    torch.save([flat_new[off:off + n].view(s) for off, n, s in zip(offsets, numels, shapes)], os.path.join(out_dir, f"z{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_adam_step_equals_plain_adam_gloo(tmp_path):
    world = 2
    mp.spawn(_zero_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g = torch.Generator().manual_seed(5)
    shapes = [(13,), (4, 9), (1,), (50, 3)]
    params = [torch.randn(s, generator=g) for s in shapes]
    outs = [torch.load(tmp_path / f"z{r}.pt", weights_only=False) for r in range(world)]
    for i, (p, s) in enumerate(zip(params, shapes)):
        gavg = sum(torch.randn(s, generator=torch.Generator().manual_seed(100 + r * 10 + i)) for r in range(world)) / world
        want = R.adam_step(p.clone(), gavg, torch.zeros(s), torch.zeros(s), 1, 1e-2, 0.9, 0.98, 1e-8, 0.05)
        for r in range(world):
            torch.testing.assert_close(outs[r][i], want, atol=1e-7, rtol=1e-6)       # every rank holds the same new parameters
