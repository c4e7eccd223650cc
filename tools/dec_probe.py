"""Decoder-kernel duration of the bridge P step under different call shapes (global batch, loss-only vs fused Adam)."""
import contextlib, importlib, io, os, sys
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import dpivae_b200 as dpv  # noqa: E402

wl = dict(bench.WORKLOADS["bridge_p"])
case_mod = importlib.import_module("dpivae_b200.cases.bridge")
dev = torch.device("cuda", 0)
xs, cs, ys = bench.synth(case_mod, 4096, 7, dev)
args = bench.make_args(case_mod, wl["preset"], use_seed=True, seed=123, n_train=4096, n_batch=4096)
with contextlib.redirect_stdout(io.StringIO()):
    vae = dpv.setup_model(args, case_mod.definition, (xs, cs, ys))
rows = 131072
x, c, y = bench.synth(case_mod, rows, 1000, dev)
eng = vae.engine()
eng.set_groups(dpv.param_groups(args))
eng.set_math_mode("tc_fp16x3")
w = (1.0, 1.0, 1.0, 1.0)
torch.manual_seed(99)
step = 0
for name, kw, fused in (("Bg=B fused adam", dict(), True), ("Bg=B loss + adam", dict(), False), ("Bg=2B loss + adam", dict(B_global=2 * rows), False),
                        ("Bg=8B loss + adam", dict(B_global=8 * rows), False), ("Bg=8B fused", dict(B_global=8 * rows), True), ("Bg=B fused adam (again)", dict(), True)):
    ms = []
    for i in range(8):
        step += 1
        if i == 3:
            eng.set_timing(True)
        if fused:
            eng.loss(x, c, y, 16, w, True, adam_step=step, **kw)
        else:
            eng.loss(x, c, y, 16, w, True, **kw)
            eng.adam_step(step)
        if i >= 3:
            ms.append(eng.last_kernel_ms()["dec_fused"])
    eng.set_timing(False)
    print(f"{name:28s} dec_tc {sum(ms) / len(ms):.4f} ms   elbo {float(eng.scalars[0]):.4f}")

# the same after a NCCL communicator exists in this process (single rank) and after one collective
import torch.distributed as dist
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29533")
dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)


def timed(tag, with_ar):
    global step
    ms = []
    for i in range(8):
        step += 1
        if i == 3:
            eng.set_timing(True)
        eng.loss(x, c, y, 16, w, True)
        if with_ar:
            dist.all_reduce(eng.gradbuf)
        eng.adam_step(step)
        if i >= 3:
            ms.append(eng.last_kernel_ms()["dec_fused"])
    eng.set_timing(False)
    print(f"{tag:28s} dec_tc {sum(ms) / len(ms):.4f} ms")


timed("nccl initialised, no collective", False)
timed("nccl allreduce per step", True)
timed("after: no collective", False)
print("stack limit", torch.cuda.cudart().cudaDeviceGetLimit(torch.cuda.cudart().cudaLimit.cudaLimitStackSize) if hasattr(torch.cuda.cudart(), "cudaDeviceGetLimit") else "n/a")
dist.destroy_process_group()
