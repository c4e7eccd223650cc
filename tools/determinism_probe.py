"""Run the same fused step several times and report which outputs differ between runs (race detector of last resort).

    python tools/determinism_probe.py [reps] [math]
"""
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import build_from_golden  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    math = sys.argv[2] if len(sys.argv) > 2 else "tc_fp16x3"
    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden("bridge", "P")
    eng = vae.engine()
    X, C_, Y = (torch.cat([t] * reps, 0).cuda() for t in (x, c, y))
    X = X + 0.01 * torch.randn(X.shape, generator=torch.Generator().manual_seed(1)).cuda()
    eng.set_math_mode(math)
    names = {id(p): k for k, p in vae.named_parameters()}
    runs = []
    for i in range(4):
        torch.manual_seed(123)
        rl, s = eng.loss(X, C_, Y, 16, (1.0, 1.0, 1.0, 1.0), True)
        torch.cuda.synchronize()
        runs.append((rl.clone(), s.clone(), eng.grads.clone()))
    gnan = torch.isnan(runs[0][2])
    print("NaN gradients in run 0:", int(gnan.sum()), "; non-finite row losses:", int((~torch.isfinite(runs[0][0])).sum()))
    for p, o in eng.slots:
        n_ = int(torch.isnan(runs[0][2][o:o + p.numel()]).sum())
        if n_:
            print(f"   NaN in {names[id(p)]}: {n_}/{p.numel()}")
    for i in range(1, 4):
        d_rl = (runs[i][0] != runs[0][0])
        print(f"run {i}: row_loss differs in {int(d_rl.sum())} of {d_rl.numel()} entries; rows by component: {[int(v) for v in d_rl.sum(1)]}; scalars equal {torch.equal(runs[i][1], runs[0][1])}")
        if d_rl.any():
            idx = d_rl.any(0).nonzero().flatten()
            print("   first differing rows:", idx[:10].tolist(), " tiles(8 rows):", sorted(set((idx[:200] // 8).tolist()))[:12])
        for p, o in eng.slots:
            a, b = runs[i][2][o:o + p.numel()], runs[0][2][o:o + p.numel()]
            if not torch.equal(a, b) and not (torch.isnan(a) == torch.isnan(b)).all():
                print('   NaN pattern differs in', names[id(p)])
            if not torch.equal(torch.nan_to_num(a), torch.nan_to_num(b)):
                rel = float((a - b).norm() / (b.norm() + 1e-30))
                print(f"   grad {names[id(p)]:45s} differs in {int((a != b).sum())}/{p.numel()} entries, rel {rel:.2e}")


if __name__ == "__main__":
    main()
