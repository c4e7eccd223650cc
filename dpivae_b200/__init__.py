"""dpivae_b200 -- B200-native (sm_100a) implementation of the DPI-VAE training step.

Mirrors the reference's Python surface for this path (`setup_model`, `train_model`,
`evaluate_model`, `disentanglement_metric`, `DPIVAE`) over a C-ABI CUDA library
(include/dpivae_b200.h, dpivae_b200/csrc).  No PyTorch-eager or CPU fallback exists.
"""
from .checkpoint import checkpoint_state, load_checkpoint, load_checkpoint_state, save_checkpoint  # noqa: F401
from .datagen import sample_response_device  # noqa: F401
from .dpivae import disentanglement_metric, evaluate_model, param_groups, setup_model, train_model  # noqa: F401
from .utils import (Annealing, EarlyStopping, MarginalDistribution, ScalarLogger, StandardScaler,  # noqa: F401
                    get_logger_training_curve, get_prior_dist, get_shapes_from_dict, make_parser, sample_response)
from .vae import DPIVAE  # noqa: F401
