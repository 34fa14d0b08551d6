"""hydrodl2_b200 — B200-native HBV recurrence + UH routing behind the hydrodl2 API.

    import hydrodl2_b200 as hydrodl2
    Hbv = hydrodl2.load_model('hbv', ver_name='Hbv')
    model = Hbv(config, device=torch.device('cuda'))
    fluxes = model({'x_phy': x}, parameters)

The arithmetic runs in hand-written sm_100a CUDA kernels (``csrc/``) reached
through the C-ABI library ``lib/libhbv_b200.so`` (``include/hbv_b200.h``).
"""

from .api import available_models, available_modules, load_model, load_module

__version__ = '0.1.0'

__all__ = ['__version__', 'available_models', 'available_modules', 'load_model', 'load_module']
