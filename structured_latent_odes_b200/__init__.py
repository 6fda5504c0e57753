"""structured_latent_odes_b200 -- B200-native (sm_100a) latent-ODE solve for SLODE.

Public surface (mirrors the reference's hot path):

    odeint, odeint_adjoint        torchdiffeq-shaped calls (models/blackbox_ode.py:40-45)
    OdeModel, OdeFunc, Dynamics   module mirrors (models/blackbox_ode.py)
    Decoder, GaussianDecoder,     decoder mirrors with fused heads (models/decoders.py)
    VarianceGaussianDecoder
    CvsMechanistic                the CVS mechanistic RHS (data/cvs/cvs_data.py:52-91) as a forward(t, state) module
    generate_cvs_latents          batched replacement of the LSODA generator loop (data/cvs/cvs_data.py:111-134)
    install_as_torchdiffeq        make `import torchdiffeq` in unmodified reference code resolve here

All compute goes through ``csrc/libslode_b200.so`` (C ABI in ``include/slode_b200.h``); there is no
CPU or eager fallback.
"""
from .torchdiffeq_api import odeint, odeint_adjoint, install_as_torchdiffeq, is_blackbox_func  # noqa: F401
from .blackbox_ode import OdeModel, OdeFunc, Dynamics  # noqa: F401
from .decoders import Decoder, GaussianDecoder, VarianceGaussianDecoder, decoder_heads, multiple_samples  # noqa: F401
from .cvs_mechanistic import CvsMechanistic, generate_cvs_latents, observe as cvs_observe  # noqa: F401

__version__ = "0.1.0"
