import os, sys, time, cProfile, pstats
import torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import news_recommendation_model_b200 as nrm
from fixtures import load_weights
from news_recommendation_model_b200.synthetic import make_batch
model = nrm.UserModel(1000); model.load_state_dict(load_weights('train'), strict=False)
model.to('cuda').train().set_precision('bf16x3')
opt = nrm.FusedAdam(model.parameters(), lr=1e-3, weight_decay=1e-5)
b = make_batch(1024, 50, 5, seed=1, user_num=1000).to('cuda')
def step():
    out = model(b.x_history, b.x_target, b.x_global)
    loss = model.loss(b.user_id, out, b.label)
    loss.backward()
    opt.step()
    opt.zero_grad()
for i in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(50): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print('cpu issue per step %.3f ms, total per step %.3f ms' % ((t1 - t0) / 50 * 1e3, (t2 - t0) / 50 * 1e3))
pr = cProfile.Profile(); pr.enable()
for i in range(50): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('tottime').print_stats(25)
