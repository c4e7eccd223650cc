"""Checkpoint / state interchange.

The reference has no save path and its `state_dict` omits what a resumable run needs: the
`StandardScaler` statistics fitted in `setup_model` (utils/transforms.py:64-68, dpivae.py:141-146), the
Adam moments and step counter (dpivae.py:373) and the generator positions.  A checkpoint here is ONE
`torch.save`-able dict:

    {"format": "dpivae_b200.checkpoint/1",
     "model":   DPIVAE.state_dict()            -- the reference's key names and nn.Linear layouts, so the weights
                                                  load into the reference's DPIVAE (and the oracle) unchanged
     "scalers": {"x"|"c"|"y": {"mean": (1,d), "scale": (1,d)}}   (population std)
     "optim":   {"step": int, "exp_avg": {name: tensor}, "exp_avg_sq": {name: tensor}}   torch.optim.Adam's names
     "rng":     {"cuda_seed": int, "cuda_offset": int, "cpu": ByteTensor}}

Per-tensor dicts (not the flat device buffers) are stored on purpose: the flat layout is an implementation detail of
libdpivae_b200, the names are the contract shared with the reference."""
import torch

FORMAT = "dpivae_b200.checkpoint/1"


def _names(vae, eng):
    by_id = {id(p): k for k, p in vae.named_parameters()}
    return [(by_id[id(p)], p, o) for p, o in eng.slots]


def checkpoint_state(vae, include_rng=True):
    """Everything needed to resume `train_model` exactly where it stopped (host tensors)."""
    state = {"format": FORMAT, "model": {k: v.detach().cpu().clone() for k, v in vae.state_dict().items()}}
    sc = {}
    for nm, tr in (("x", vae.transform_x), ("c", vae.transform_c), ("y", vae.transform_y)):
        if tr is not None:
            sc[nm] = {"mean": tr.mean_.detach().cpu().clone(), "scale": tr.scale_.detach().cpu().clone()}
    state["scalers"] = sc
    eng = vae._engine
    if eng is not None:
        m, v = eng.exp_avg.detach().cpu(), eng.exp_avg_sq.detach().cpu()
        state["optim"] = {"step": int(eng.step_count),
                          "exp_avg": {k: m[o:o + p.numel()].view(p.shape).clone() for k, p, o in _names(vae, eng)},
                          "exp_avg_sq": {k: v[o:o + p.numel()].view(p.shape).clone() for k, p, o in _names(vae, eng)}}
        if include_rng:
            gen = torch.cuda.default_generators[eng.dev.index if eng.dev.index is not None else torch.cuda.current_device()]
            state["rng"] = {"cuda_seed": int(gen.initial_seed()), "cuda_offset": int(gen.get_offset()),
                            "cpu": torch.get_rng_state()}
    return state


def save_checkpoint(path, vae, include_rng=True):
    torch.save(checkpoint_state(vae, include_rng), path)


def load_checkpoint_state(vae, state, restore_rng=True):
    """Load a checkpoint dict into a DPIVAE built by `setup_model` with the same architecture."""
    if state.get("format") != FORMAT:
        raise ValueError(f"not a {FORMAT} checkpoint")
    for nm, tr in (("x", vae.transform_x), ("c", vae.transform_c), ("y", vae.transform_y)):
        if nm in state["scalers"]:
            if tr is None:
                raise ValueError(f"checkpoint carries scaler statistics for '{nm}' but the model has no transform_{nm}")
            tr.mean_ = state["scalers"][nm]["mean"].clone()
            tr.scale_ = state["scalers"][nm]["scale"].clone()
    if vae._engine is not None:   # the descriptor bakes the scaler statistics in: rebuild the engine lazily
        vae._engine.close()
        vae._engine = None
    res = vae.load_state_dict(state["model"], strict=False)
    bad = [k for k in res.missing_keys if not k.startswith("decoder_x.model.")] + list(res.unexpected_keys)
    if bad:
        raise ValueError(f"checkpoint / model mismatch: {bad}")
    if "optim" in state and torch.cuda.is_available():
        eng = vae.engine()
        with torch.no_grad():
            for k, p, o in _names(vae, eng):
                eng.exp_avg[o:o + p.numel()].copy_(state["optim"]["exp_avg"][k].reshape(-1))
                eng.exp_avg_sq[o:o + p.numel()].copy_(state["optim"]["exp_avg_sq"][k].reshape(-1))
        eng.step_count = int(state["optim"]["step"])
        if restore_rng and "rng" in state:
            gen = torch.cuda.default_generators[eng.dev.index if eng.dev.index is not None else torch.cuda.current_device()]
            gen.manual_seed(state["rng"]["cuda_seed"])
            gen.set_offset(state["rng"]["cuda_offset"])
            torch.set_rng_state(state["rng"]["cpu"])
    return vae


def load_checkpoint(path, vae, restore_rng=True):
    return load_checkpoint_state(vae, torch.load(path, map_location="cpu", weights_only=True), restore_rng)
