"""Parameter-holding mirrors of the reference's `models/encoders.py`, `models/decoders.py` and
`models/nn.py` modules.

They keep the constructor signatures, attribute names and `state_dict` keys of the reference
(so weights interchange with it and `nn.Linear` default init reproduces it seed-for-seed), but
their arithmetic does NOT run in PyTorch: `DPIVAE` (dpivae_b200/vae.py) re-points every parameter
into one flat device buffer and evaluates all of them inside the fused CUDA kernels.  Calling a
sub-module's `forward` directly therefore raises instead of silently falling back to eager ops.
"""
from torch import nn


class _FusedOnly(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError(
            f"{type(self).__name__}.forward is fused into the DPIVAE CUDA kernels; call DPIVAE.loss / forward / "
            "sample / encode (there is no eager PyTorch path)")


def _mlp_stack(layers, prefix_linear, prefix_nonlinear):
    net = nn.Sequential()
    for i in range(len(layers) - 1):
        net.add_module(f"{prefix_linear}_{i}", nn.Linear(layers[i], layers[i + 1]))
        net.add_module(f"{prefix_nonlinear}_{i}", nn.ReLU())
    return net


class FullCovarianceNN(_FusedOnly):
    """models/encoders.py:6-44: MLP -> loc, sigma, strict-lower L (scale_tril = L + diag(sigma + 1e-8))."""

    def __init__(self, n_latent, n_input, layers):
        super().__init__()
        self.n_latent, self.n_input = n_latent, n_input
        self.layers = [n_input] + list(layers)
        if len(self.layers) != 2:
            raise ValueError("the fused kernels support exactly one hidden layer (reference default)")
        self.mean_output, self.sigma_output, self.cov_output = n_latent, n_latent, n_latent * n_latent
        # creation order matters for seed-for-seed init parity: f_mean, f_sigma, f_cov, then net
        self.f_mean = nn.Linear(self.layers[-1], self.mean_output)
        self.f_sigma = nn.Linear(self.layers[-1], self.sigma_output)
        self.f_cov = nn.Linear(self.layers[-1], self.cov_output)
        self.net = _mlp_stack(self.layers, "encoder_linear", "encoder_nonlinear")


class FactorizedNN(_FusedOnly):
    """models/encoders.py:96-128: MLP -> loc, sigma (diagonal scale_tril)."""

    def __init__(self, n_latent, n_input, layers):
        super().__init__()
        self.n_latent, self.n_input = n_latent, n_input
        self.layers = [n_input] + list(layers)
        if len(self.layers) != 2:
            raise ValueError("the fused kernels support exactly one hidden layer (reference default)")
        self.mean_output, self.sigma_output = n_latent, n_latent
        self.f_mean = nn.Linear(self.layers[-1], self.mean_output)
        self.f_sigma = nn.Linear(self.layers[-1], self.sigma_output)
        self.net = _mlp_stack(self.layers, "encoder_linear", "encoder_nonlinear")


class GaussianEncoder(_FusedOnly):
    """models/encoders.py:46-93."""

    def __init__(self, net, input_transform=None, output_transform=None):
        super().__init__()
        if input_transform is not None:
            raise ValueError("per-encoder input transforms are not part of the fused path (reference never sets one)")
        self.net = net
        self.input_transform = input_transform
        self.output_transform = output_transform


class Decoder(_FusedOnly):
    """models/decoders.py:4-49: ReLU MLP n_input -> layers -> 2*n_output (mean || log sigma).
    state_dict keys net.0 / net.2 as in the reference (Sequential.pop renumbers)."""

    def __init__(self, n_input, n_output, layers, nonlinear_last=None, nonlinearity=nn.ReLU):
        super().__init__()
        if nonlinear_last not in (None, False) or nonlinearity is not nn.ReLU or len(layers) != 1:
            raise ValueError("the fused kernels support the reference configuration: one ReLU hidden layer")
        self.n_input, self.n_output = n_input, n_output
        self.layers = [n_input] + list(layers) + [2 * n_output]
        self.net = nn.Sequential(nn.Linear(self.layers[0], self.layers[1]), nn.ReLU(),
                                 nn.Linear(self.layers[1], self.layers[2]))


class GradRevAdditive(_FusedOnly):
    """models/decoders.py:52-92: physics model term + GRL -> fx0 -> ReLU -> fx1 data-driven term."""

    def __init__(self, model, nz_p, nz_d, n_output, hidden=128, grad_reverse=None):
        super().__init__()
        if hidden != 128:
            raise ValueError("the fused kernels support hidden=128 (reference value, dpivae.py:169)")
        self.model = model
        self.nz_p, self.nz_d, self.n_output = nz_p, nz_d, n_output
        self.grad_reverse = grad_reverse
        self.fx0 = nn.Linear(nz_d, hidden)
        self.fx1 = nn.Linear(hidden, n_output)
        self.nonlinearity = nn.ReLU()
