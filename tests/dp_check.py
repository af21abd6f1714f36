#!/usr/bin/env python
"""Multi-GPU numerical check, run under torchrun on N GPUs of one box:

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dp_check.py

`run_check()` is also what `bench.py --gpus N` calls as its pre-flight (result in the JSON line as `dp_parity`) and what
tests/test_gpu_dp.py launches when the box shows >= 2 GPUs.

1. module path, synchronised BatchNorm statistics: N ranks x (B/N impressions) must reproduce one process on the whole
   batch -- same logits, same averaged gradients, same BatchNorm buffers;
2. the `delta` gradient (averaged inside the loss backward, engine._LossFn.backward) is bit-identical on every rank, also when
   it is accumulated into an existing .grad (zero_grad(set_to_none=False) + a second micro-batch);
3. FusedTrainStep (the path bench.py measures): three steps on rank-local shards leave bit-identical weights on every rank, and
   they equal a single-process FusedTrainStep on the whole batch within the fp32 tolerances when sync_bn is on."""
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


def _same_on_all_ranks(t: torch.Tensor) -> bool:
    """True when `t` holds the same bits on every rank."""
    lo, hi = t.detach().clone(), t.detach().clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return bool(torch.equal(lo, hi))


def run_check(dev, rank: int, world: int, precision: str = 'bf16x3', fused: bool = True):
    """-> (ok, report dict).  Needs an initialised NCCL process group with one rank per GPU."""
    B, H, C, U = 64 * world, 50, 5, 100
    full = make_batch(B, H, C, seed=7, user_num=U)
    delta0 = torch.from_numpy(np.random.default_rng(2).normal(0, 0.3, U + 1).astype(np.float32))

    def fresh():
        m = nrm.UserModel(U)
        m.load_state_dict(load_weights('train'), strict=False)
        with torch.no_grad():
            m.delta.copy_(delta0)
        return m.to(dev).train().set_precision(precision)

    rep = {}
    # ---- 1. module path vs one process on the whole batch (every rank computes the reference redundantly)
    ref = fresh()
    d = full.to(dev)
    out = ref(d.x_history, d.x_target, d.x_global)
    ref.loss(d.user_id, out, d.label).backward()
    ref_grads = {k: p.grad.clone() for k, p in ref.named_parameters()}
    ref_out = out.detach()

    lo, hi = shard_range(B, rank, world)
    shard_host = Batch(*[getattr(full, f)[lo:hi] for f in full.__dataclass_fields__])
    shard = shard_host.to(dev)
    m = fresh()
    DataParallel(m, sync_bn=True)
    o = m(shard.x_history, shard.x_target, shard.x_global)
    m.loss(shard.user_id, o, shard.label).backward()
    torch.cuda.synchronize()
    rep['logits_err'] = (o.detach() - ref_out[lo:hi]).abs().max().item()
    worst = ('', 0.0)
    for k, p in m.named_parameters():
        if k in ('delta', 'out_mlp.fc2.bias'):       # pure rounding noise (softmax shift invariance): compared in absolute terms below
            continue
        scale = ref_grads[k].abs().max().item()
        err = (p.grad - ref_grads[k]).abs().max().item() / max(scale, 1e-12)
        if err > worst[1]:
            worst = (k, err)
    rep['worst_grad_rel_err'], rep['worst_grad'] = worst[1], worst[0]
    rep['bn_running_mean_err'] = (m.bn.running_mean - ref.bn.running_mean).abs().max().item()
    rep['delta_grad_abs_err'] = (m.delta.grad - ref_grads['delta']).abs().max().item()
    ok = rep['logits_err'] <= 1e-4 and worst[1] <= 5e-4 and rep['bn_running_mean_err'] <= 1e-5 and rep['delta_grad_abs_err'] <= 1e-6

    # ---- 2. every gradient identical on every rank, also after accumulating a second micro-batch into existing .grad
    same = all(_same_on_all_ranks(p.grad) for p in m.parameters())
    o2 = m(shard.x_history, shard.x_target, shard.x_global)
    (2.0 * m.loss(shard.user_id, o2, shard.label)).backward()            # accumulates into the .grad of the first pass
    torch.cuda.synchronize()
    same_acc = all(_same_on_all_ranks(p.grad) for p in m.parameters())
    rep['grads_identical_on_all_ranks'], rep['accumulated_grads_identical_on_all_ranks'] = same, same_acc
    ok = ok and same and same_acc

    # ---- 3. the fused path bench.py measures
    if fused:
        for sync_bn in ((False, True) if getattr(nrm.FusedTrainStep, 'SYNC_BN', False) else (False,)):
            mf = fresh()
            DataParallel(mf, sync_bn=sync_bn)
            tr = nrm.FusedTrainStep(mf, hi - lo, H, C, lr=1e-3, weight_decay=1e-5)
            pinned = shard_host.pin()
            for _ in range(3):
                h = tr.step(pinned)
            h.item()
            torch.cuda.synchronize()
            rep[f'fused_weights_identical_sync_bn_{int(sync_bn)}'] = _same_on_all_ranks(mf.flat_parameters().buf)
            ok = ok and rep[f'fused_weights_identical_sync_bn_{int(sync_bn)}']
            if sync_bn:
                m1 = fresh()
                tr1 = nrm.FusedTrainStep(m1, B, H, C, lr=1e-3, weight_decay=1e-5)
                fp = full.pin()
                for _ in range(3):
                    h1 = tr1.step(fp)
                h1.item()
                torch.cuda.synchronize()
                a, b = mf.flat_parameters(), m1.flat_parameters()
                # three Adam steps move a weight by up to 3e-3; elements with ~zero gradients follow the sign of rounding noise:
                # parity.assert_weights_follow's two bounds (99.9 % of the elements within 1 % of the travel, none beyond a quarter)
                import parity as P
                ref = {name: b.buf[off:off + n].detach().cpu() for name, off, n, _ in b.slots}
                got = [(name, a.buf[off:off + n].detach().cpu()) for name, off, n, _ in a.slots]
                try:
                    P.assert_weights_follow(got, ref, 3)
                    rep['fused_sync_bn_vs_single_process_weights'] = 'ok'
                except AssertionError as ex:
                    rep['fused_sync_bn_vs_single_process_weights'] = f'FAILED {ex}'
                    ok = False
                # the loss a rank reports is the mean over ITS impressions: the mean over the ranks is the whole-batch loss
                lt = torch.tensor([h.item()], dtype=torch.float64, device=dev)
                dist.all_reduce(lt, op=dist.ReduceOp.SUM)
                rep['fused_sync_bn_vs_single_process_loss_err'] = abs(lt.item() / world - h1.item())
                ok = ok and rep['fused_sync_bn_vs_single_process_loss_err'] <= 1e-5
            rep['transport'] = 'peer memory' if tr.peer is not None else f'nccl ({mf._dp.peer_error})'
    # ---- 4. large user table: delta's gradient as (user id, value) lists (sparse exchange) against the dense exchange
    if fused and getattr(nrm.FusedTrainStep, 'SPARSE_DELTA_MIN', None) is not None:
        UL = 70000
        fullL = make_batch(B, H, C, seed=8, user_num=UL)
        fullL.user_id[1] = fullL.user_id[0]                   # a duplicate user inside one rank's shard
        shardL = Batch(*[getattr(fullL, f)[lo:hi] for f in fullL.__dataclass_fields__]).pin()

        def run(sparse_min):
            mm = nrm.UserModel(UL)
            mm.load_state_dict(load_weights('train'), strict=False)
            mm.to(dev).train().set_precision(precision)
            DataParallel(mm)
            old = nrm.FusedTrainStep.SPARSE_DELTA_MIN
            nrm.FusedTrainStep.SPARSE_DELTA_MIN = sparse_min
            try:
                t = nrm.FusedTrainStep(mm, hi - lo, H, C, lr=1e-3, weight_decay=1e-5)
            finally:
                nrm.FusedTrainStep.SPARSE_DELTA_MIN = old
            for _ in range(2):
                hh = t.step(shardL)
            hh.item()
            torch.cuda.synchronize()
            return mm, t
        m_sp, t_sp = run(65536)
        m_de, t_de = run(1 << 40)
        if t_sp.peer is not None:
            fs, fd = m_sp.flat_parameters(), m_de.flat_parameters()
            rep['sparse_delta_used'] = bool(t_sp.sparse_delta and not t_de.sparse_delta)
            rep['sparse_delta_replicas_identical'] = _same_on_all_ranks(fs.buf)
            # delta cannot influence anything else (softmax is shift invariant): the other weights agree bit for bit
            rep['sparse_vs_dense_other_weights_identical'] = bool(torch.equal(fs.buf[:fs.fixed], fd.buf[:fd.fixed]))
            # the averaged delta gradient itself: first moment after two steps = 0.1 (0.9 g1 + g2), same in both exchanges
            ms, md = t_sp.exp_avg[fs.fixed:fs.fixed + UL + 1], t_de.exp_avg[fd.fixed:fd.fixed + UL + 1]
            rep['sparse_vs_dense_delta_moment_err'] = (ms - md).abs().max().item()
            rep['delta_moment_scale'] = md.abs().max().item()
            ok = ok and rep['sparse_delta_used'] and rep['sparse_delta_replicas_identical'] and rep['sparse_vs_dense_other_weights_identical'] \
                and rep['sparse_vs_dense_delta_moment_err'] <= 1e-12 + 1e-4 * rep['delta_moment_scale']
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item() == 1.0), rep


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    ok, rep = run_check(dev, rank, world)
    if rank == 0:
        print(f'dp_check world={world}: {rep} -> {"OK" if ok else "FAILED"}', flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
