#!/usr/bin/env python
"""Multi-GPU check, run under torchrun on N GPUs of one box:

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dp_check.py

N ranks x (B/N impressions) with synchronised BatchNorm statistics must reproduce one process
on the whole batch: same logits, same averaged gradients, same weights after Adam."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import news_recommendation_model_b200 as nrm          # noqa: E402
from news_recommendation_model_b200.dp import DataParallel, shard_range   # noqa: E402
from news_recommendation_model_b200.synthetic import make_batch, Batch   # noqa: E402
from fixtures import load_weights                      # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    B, H, C, U = 64 * world, 50, 5, 100
    full = make_batch(B, H, C, seed=7, user_num=U)
    delta0 = torch.from_numpy(np.random.default_rng(2).normal(0, 0.3, U + 1).astype(np.float32))

    def fresh():
        m = nrm.UserModel(U)
        m.load_state_dict(load_weights('train'), strict=False)
        with torch.no_grad():
            m.delta.copy_(delta0)
        return m.to(dev).train()

    # single-process reference on the whole batch (every rank computes it redundantly)
    ref = fresh()
    d = full.to(dev)
    out = ref(d.x_history, d.x_target, d.x_global)
    ref.loss(d.user_id, out, d.label).backward()
    ref_grads = {k: p.grad.clone() for k, p in ref.named_parameters()}
    ref_out = out.detach()

    # data parallel: this rank's shard
    lo, hi = shard_range(B, rank, world)
    shard = Batch(*[getattr(full, f)[lo:hi] for f in full.__dataclass_fields__]).to(dev)
    m = fresh()
    DataParallel(m, sync_bn=True)
    o = m(shard.x_history, shard.x_target, shard.x_global)
    m.loss(shard.user_id, o, shard.label).backward()
    torch.cuda.synchronize()
    err_logits = (o.detach() - ref_out[lo:hi]).abs().max().item()
    worst = ('', 0.0)
    for k, p in m.named_parameters():
        if k in ('delta', 'out_mlp.fc2.bias'):
            continue
        scale = ref_grads[k].abs().max().item()
        err = (p.grad - ref_grads[k]).abs().max().item() / max(scale, 1e-12)
        if err > worst[1]:
            worst = (k, err)
    bn_err = (m.bn.running_mean - ref.bn.running_mean).abs().max().item()
    ok = err_logits <= 1e-4 and worst[1] <= 5e-4 and bn_err <= 1e-5
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f'dp_check world={world}: logits err {err_logits:.2e}, worst grad rel err {worst[1]:.2e} ({worst[0]}), '
              f'bn running_mean err {bn_err:.2e} -> {"OK" if flag.item() == 1.0 else "FAILED"}', flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == '__main__':
    main()
