import contextlib, importlib, io, os, sys, time
import torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import bench
import dpivae_b200 as dpv
case_mod = importlib.import_module("dpivae_b200.cases.bridge")
dev = torch.device("cuda", 0)
rows = 131072
x, c, y = bench.synth(case_mod, rows, 7, dev)
args = bench.make_args(case_mod, "DPIVAE-A", use_seed=True, n_train=rows, n_batch=rows)
with contextlib.redirect_stdout(io.StringIO()):
    vae = dpv.setup_model(args, case_mod.definition, (x, c, y))
eng = vae.engine(); eng.set_groups(dpv.param_groups(args)); eng.set_math_mode("tc_fp16x3")
w = (1.0, 1.0, 1.0, 1.0)
idx = torch.randperm(rows, device=dev)
for i in range(5):
    eng.loss(x, c, y, 16, w, True, idx=idx, adam_step=i + 1)
torch.cuda.synchronize()
import cProfile, pstats
t0 = time.perf_counter()
for i in range(50):
    eng.loss(x, c, y, 16, w, True, idx=idx, adam_step=i + 6)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host time per call (async, queue filling): %.1f us; total %.1f us/step" % ((t1 - t0) / 50 * 1e6, (t2 - t0) / 50 * 1e6))
# host time when the GPU is idle at call time (like the e2e arm: sync every step)
ts = []
for i in range(30):
    torch.cuda.synchronize()
    a = time.perf_counter(); eng.loss(x, c, y, 16, w, True, idx=idx, adam_step=i + 56); b = time.perf_counter()
    torch.cuda.synchronize(); cc = time.perf_counter()
    ts.append((b - a, cc - a))
print("synced: host call %.1f us, call+wait %.1f us" % (1e6 * sum(t[0] for t in ts) / 30, 1e6 * sum(t[1] for t in ts) / 30))
pr = cProfile.Profile(); pr.enable()
for i in range(200):
    eng.loss(x, c, y, 16, w, True, idx=idx, adam_step=i + 100)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
# the same step as ONE graph launch (device-resident step state), synchronised every step
pool = torch.stack([torch.randperm(rows, device=dev) for _ in range(2)])
sg = eng.step_graph(x, c, y, 16, w, idx_pool=pool)
for i in range(5):
    sg.run(1)
torch.cuda.synchronize()
ts = []
for i in range(30):
    torch.cuda.synchronize()
    a = time.perf_counter(); sg.run(1); b = time.perf_counter()
    torch.cuda.synchronize(); cc = time.perf_counter()
    ts.append((b - a, cc - a))
print("graph, synced: host call %.1f us, call+wait %.1f us" % (1e6 * sum(t[0] for t in ts) / 30, 1e6 * sum(t[1] for t in ts) / 30))
t0 = time.perf_counter(); sg.run(50); torch.cuda.synchronize(); t1 = time.perf_counter()
print("graph, 50 steps back to back: %.1f us/step" % ((t1 - t0) / 50 * 1e6))
